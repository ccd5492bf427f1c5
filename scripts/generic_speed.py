"""What the hand-written kernels buy: the two bench workloads on one B200 through the hand-written path and through the generic
path (every loop of the step printed from the reference's equations and compiled at run time -- the shape of the reference's
own OPS-CUDA program: one kernel per loop, work arrays in HBM).  Same plans, same grid, same GPU; prints one JSON line."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
from opensbli_b200 import run as R, Simulation   # noqa: E402

PLANS = os.path.join(HERE, '..', 'tests', 'golden', 'plans')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = {'grid': [N, N, N], 'steps': 5, 'unit': 'ms/step'}
for workload, fast, slow in (('teno5', 'tgv_teno5', 'tgv_teno5_allprinted'), ('central4', 'tgv_central4', 'tgv_central4_allprinted')):
    over = {'block0np%d' % d: N for d in range(3)}
    over['dt'] = 0.003385 * 64 / N
    res = {}
    states = {}
    for label, name in (('hand_written', fast), ('generic', slow)):
        plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides=over)
        with Simulation(plan) as sim:
            q0 = R.initial_state(plan_sym, cold)
            sim.set_state(q0)
            sim.step(2)
            l0 = sim.launch_count()
            ms = sim.step_timed(5) / 5
            res[label] = {'ms_per_step': ms, 'updates_per_s': N ** 3 / (ms * 1e-3), 'launches_per_step': (sim.launch_count() - l0) / 5,
                          'path': plan['conv']}
            states[label] = np.stack([a[5:-5, 5:-5, 5:-5] for a in sim.get_state()])
    a, b = states['hand_written'], states['generic']
    # per field, relative to the field's L-inf norm; the momentum components share the norm of the momentum vector (tests/common.py)
    norm = [np.abs(a[0]).max()] + [np.sqrt((a[1:4] ** 2).sum(axis=0)).max()] * 3 + [np.abs(a[4]).max()]
    res['max_rel_difference_after_7_steps'] = float(max(np.abs(a[m] - b[m]).max() / norm[m] for m in range(5)))
    res['speed_up'] = res['generic']['ms_per_step'] / res['hand_written']['ms_per_step']
    out[workload] = res
print(json.dumps(out))
