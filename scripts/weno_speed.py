"""ms/step of the Taylor-Green workload with the WENO5-JS / WENO5-Z sweeps at 256^3 on one B200 (the bench line itself is TENO5);
OSB_B200_LIB selects an experimental library (e.g. one built with -DOSB_WENO_FOUR_DIVISIONS).  One JSON line."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
import bench                                   # noqa: E402
from opensbli_b200 import Simulation           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {'grid': [N] * 3, 'lib': os.environ.get('OSB_B200_LIB', 'product')}
for form in ('JS', 'Z'):
    plan = bench.tgv_plan([N] * 3, 'teno5')
    plan.update(conv='weno', order=5, weno_formulation=form)
    q = [np.zeros((N + 10,) * 3) for _ in range(5)]
    bench.tgv_state_into(q, plan, 0, N)
    with Simulation(plan) as sim:
        sim.set_state(q)
        sim.step(3)
        ms = sim.step_timed(10) / 10
        prof = sim.profile_step()
        rho = sim.download('rho')[5:-5, 5:-5, 5:-5]
    out[form] = {'ms_per_step': ms, 'updates_per_s': N ** 3 / (ms * 1e-3), 'flux_ms': prof['flux']['ms'], 'checksum_rho': float(rho.sum()), 'finite': bool(np.isfinite(rho).all())}
print(json.dumps(out))
