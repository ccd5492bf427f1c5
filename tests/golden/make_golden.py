#!/usr/bin/env python
"""Mint the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Each fixture is produced by running oracle/_ref/<config>/ref_seq -- the reference's own generated C
(OpenSBLI front end at /root/reference, OPSC back end, PYTHONHASHSEED=0, SymPy version recorded in
oracle/_ref/<config>/provenance.json) compiled with `g++ -O2 -ffp-contract=off` against the sequential OPS
stand-in oracle/ops_seq.h.  Stored per fixture (npz): the resolved plan (json), the conserved fields after the
reference's own initialisation kernel (niter=0) and after n reference time steps, interior points only.

    python oracle/gen_ref.py            # needs /root/reference
    python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import run_ref, REF_DIR  # noqa: E402

LS3 = dict(rk='ls', rk_a=[0.0, -5.0 / 9.0, -153.0 / 128.0], rk_b=[1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0])
SBLI3 = dict(rk='sbli', rk_a=[1.0 / 4.0, 3.0 / 20.0, 3.0 / 5.0], rk_b=[2.0 / 3.0, 5.0 / 12.0, 3.0 / 5.0])


def sod_plan(N, conv, order, form='JS'):
    g = 1.4
    ql = [1.0, 0.0, 1.0 / (g - 1.0)]
    qr = [0.125, 0.0, 0.1 / (g - 1.0)]
    return dict(ndim=1, np=[N], delta=[1.0 / (N - 1)], conv=conv, order=order, weno_formulation=form, averaging='roe',
                viscous=False, constants=dict(gama=g, dt=0.0002, eps=1e-16, TENO_CT=1e-5),
                bc=[[dict(type='dirichlet', q=ql), dict(type='dirichlet', q=qr)]], **LS3)


def tgv_plan(N, conv, order, rk):
    per = [[dict(type='periodic'), dict(type='periodic')] for _ in range(3)]
    return dict(ndim=3, np=[N, N, N], delta=[2 * math.pi / N] * 3, conv=conv, order=order, averaging='roe', viscous=True,
                constants=dict(gama=1.4, Minf=0.1, Re=1600.0, Pr=0.71, dt=0.003385 * 64 / N, eps=1e-16, TENO_CT=1e-6),
                bc=per, **rk)


# fixture name -> (reference config, plan, [step counts])
FIXTURES = {
    'sod_teno5_n200': ('sod_teno5', sod_plan(200, 'teno', 5), [1, 50]),
    'sod_wenojs5_n800': ('sod_wenojs5', sod_plan(800, 'weno', 5, 'JS'), [1, 100, 1000]),
    'sod_wenoz5_n200': ('sod_wenoz5', sod_plan(200, 'weno', 5, 'Z'), [1, 50]),
    'tgv_central4_16': ('tgv_central4', tgv_plan(16, 'central', 4, SBLI3), [1, 3]),
    'tgv_teno5_16': ('tgv_teno5', tgv_plan(16, 'teno', 5, LS3), [1, 3]),
}


def env_params(plan):
    P = {'dt': plan['constants']['dt']}
    for d in range(plan['ndim']):
        P['block0np%d' % d] = plan['np'][d]
    return P


def main():
    for name, (config, plan, steps) in FIXTURES.items():
        nd = plan['ndim']
        fields = ['rho'] + ['rhou%d' % d for d in range(nd)] + ['rhoE']
        inner = (slice(5, -5),) * nd
        out = {'plan': np.array(json.dumps(plan, sort_keys=True)),
               'provenance': np.array(open(os.path.join(REF_DIR, config, 'provenance.json')).read())}
        r = run_ref(config, dict(env_params(plan), niter=0), fields)
        out['q0'] = np.stack([r[f][inner] for f in fields])
        for n in steps:
            r = run_ref(config, dict(env_params(plan), niter=n), fields)
            out['q%d' % n] = np.stack([r[f][inner] for f in fields])
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith('q')}, '%.1f kB' % (os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
