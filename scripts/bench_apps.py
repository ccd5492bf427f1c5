#!/usr/bin/env python
"""Throughput of whole reference apps through the drop-in boundary (plan fixtures distilled by B200(alg), cold kernels by the
runner, time loop on the GPU), beside the reference's own generated C (OpenMP build) where oracle/_ref has it.
One JSON line per app.  Not the headline bench (that is bench.py); these are the wall-bounded / general-path configurations."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

CASES = [
    # plan fixture, size overrides, steps, reference config (oracle/_ref) or None
    ('katzer', {'block0np0': 500, 'block0np1': 250}, 400, 'katzer'),                       # BASELINE config 4 as shipped
    ('tcf_teno6', {'block0np0': 128, 'block0np1': 128, 'block0np2': 128}, 20, 'tcf_teno6'),
    ('tcf_central', {'block0np0': 128, 'block0np1': 128, 'block0np2': 128}, 20, 'tcf_central'),
    ('trans', {'block0np0': 240, 'block0np1': 120, 'block0np2': 64}, 20, 'trans'),
    ('vst', {'block0np0': 500, 'block0np1': 250}, 400, 'vst'),
    ('ewc', {'block0np0': 512, 'block0np1': 512}, 200, 'ewc'),
    # filters (run-time compiled UserDefinedEquations loops after every step of the hand-written kernels)
    ('katzer_sfd', {'block0np0': 500, 'block0np1': 250}, 400, 'katzer_sfd'),
    ('tgv_wf', {'block0np0': 128, 'block0np1': 128, 'block0np2': 128, 'dt': 0.003385 * 64 / 128}, 20, 'tgv_wf'),
    # programs on the generic path (every loop of the step printed from its equations, NVRTC)
    ('tg_isot', {'block0np0': 129, 'block0np1': 129, 'block0np2': 129, 'dt': 0.0015}, 20, 'tg_isot'),
    ('sod_weno7', {'block0np0': 100000, 'dt': 4e-7}, 200, 'sod_weno7'),
]


def main():
    import numpy as np
    from opensbli_b200 import run as R, Simulation
    import oracle_util as ou
    for name, over, nsteps, ref in CASES:
        plan_sym, env, plan, cold = R.load_case(os.path.join(REPO, 'tests', 'golden', 'plans', name), overrides=over)
        q0 = R.initial_state(plan_sym, cold)
        pts = float(np.prod(plan['np']))
        with Simulation(plan) as sim:
            sim.set_state(q0)
            sim.step(3)
            ms = sim.step_timed(nsteps)
            finite = bool(np.isfinite(sim.download('rho')).all())
        line = {'app': name, 'path': 'generic' if plan['conv'] == 'generic' else 'hand-written', 'np': plan['np'], 'steps': nsteps, 'ms_per_step': ms / nsteps, 'updates_per_s': pts * nsteps / (ms * 1e-3), 'finite': finite}
        if ref and ou.have_ref(ref) and '--no-cpu' not in sys.argv:
            n = max(2, nsteps // 10)
            r = ou.run_ref(ref, dict(over, niter=n), [], exe='ref_omp', threads=os.cpu_count())
            sec = r['_wall']
            line['reference_cpu_updates_per_s'] = pts * n / sec
            line['reference_cpu_cores'] = os.cpu_count()
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
