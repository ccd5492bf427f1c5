import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'tests'))
import numpy as np, torch
import opensbli_b200
from opensbli_b200 import hostpipe
from test_gpu_scale import tgv_case
wl, np3, chunk = sys.argv[1], tuple(int(v) for v in sys.argv[2].split('x')), int(sys.argv[3])
plan, q0 = tgv_case(np3, wl)
pin = lambda: [torch.empty(a.shape, dtype=torch.float64, pin_memory=True) for a in q0]
ti, ta, tb = pin(), pin(), pin()
qi, qa, qb = [[t.numpy() for t in ts] for ts in (ti, ta, tb)]
for a, b in zip(qi, q0): a[...] = b
with opensbli_b200.Simulation(plan) as sim:
    sim.advance_host(qi, qa, 1)
with hostpipe.HostPipeline(plan, chunk=chunk, nsteps=1) as pipe:
    for rep in range(2):
        for a in qb: a[...] = np.nan
        pipe.advance(qi, qb)
        a, b = qa[0], qb[0]
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        print('rep', rep, 'nan planes', [int(np.isnan(b[k]).sum()) for k in range(b.shape[0])])
        print('rep', rep, 'bad planes', [int(bad[k].sum()) for k in range(b.shape[0])])
        k = int(np.argmax([bad[k, 5:-5, 5:-5].sum() for k in range(b.shape[0])]))
        print('worst plane', k, 'interior bad rows', [int(bad[k, j, 5:-5].sum()) for j in range(5, b.shape[1] - 5)][:60])
