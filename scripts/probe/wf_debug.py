import sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..')); sys.path.insert(0, os.path.join(HERE, '..', '..', 'tests'))
import numpy as np, oracle_util as ou
from opensbli_b200 import run as R, Simulation
import hostsim
PL = os.path.join(HERE, '..', '..', 'tests', 'golden', 'plans')
over = {'block0np0': 16, 'block0np1': 16, 'block0np2': 16, 'dt': 0.003385 * 64 / 16}
plan_sym, env, plan, cold = R.load_case(os.path.join(PL, 'tgv_wf'), overrides=over)
names = ['rho', 'rhou0', 'rhou1', 'rhou2', 'rhoE']
q = [a.copy() for a in R.initial_state(plan_sym, cold)]
hk = hostsim.HostKernels(plan['user_kernels'], q[0].shape)
oplan = {k: v for k, v in plan.items() if k not in ('user_kernels', 'user_fields')}
qc, _ = ou.oracle_advance(oplan, [a.copy() for a in q], 1)
for n, a in zip(names, qc): hk.fields[n] = a
hk.run('iteration_end')
allf = names + ['u0', 'u1', 'u2', 'p', 'a', 'kappa'] + ['wk%d' % i for i in range(15)]
with Simulation(plan) as sim:
    sim.set_state(q)
    sim.step(1)
    g = {f: sim.download(f) for f in allf}
s = (slice(5, -5),) * 3
for f in allf:
    a, b = g[f], hk.fields[f]
    print('%-6s nan(gpu interior) %5d  nan(cpu interior) %5d  max diff interior %.3e   whole-array nan gpu %d cpu %d' % (
        f, np.isnan(a[s]).sum(), np.isnan(b[s]).sum(), np.nanmax(np.abs(a[s] - b[s])), np.isnan(a).sum(), np.isnan(b).sum()))
