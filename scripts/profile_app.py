#!/usr/bin/env python
"""Per-kernel-family device time of one step of a reference app (plan fixture + size overrides):
    python scripts/profile_app.py tcf_teno6 block0np0=256 block0np1=256 block0np2=256"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import numpy as np
    from opensbli_b200 import run as R, Simulation
    name = sys.argv[1]
    over = {k: int(v) for k, v in (a.split('=') for a in sys.argv[2:])}
    plan_sym, env, plan, cold = R.load_case(os.path.join(REPO, 'tests', 'golden', 'plans', name), overrides=over)
    with Simulation(plan) as sim:
        sim.set_state(R.initial_state(plan_sym, cold))
        sim.step(3)
        ms = sim.step_timed(10) / 10
        prof = sim.profile_step()
    pts = float(np.prod(plan['np']))
    print(json.dumps({'app': name, 'np': plan['np'], 'ms_per_step': ms, 'updates_per_s': pts / (ms * 1e-3),
                      'families_ms': {k: round(v['ms'], 3) for k, v in prof.items()}, 'launches': {k: v['launches'] for k, v in prof.items()}}))


if __name__ == '__main__':
    main()
