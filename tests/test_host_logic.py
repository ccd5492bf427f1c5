"""CPU tests (-m "not gpu"): host logic, plan serialisation, C-ABI library exports (no compute calls) and the
device arithmetic (opensbli_b200/csrc/osb_math.cuh compiled for the host) against the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_util as ou
from common import load_fixture

REPO = ou.REPO


def test_header_symbols_exported_by_library():
    """Every function declared in include/osbli_b200.h is exported by the in-tree CUDA library."""
    import opensbli_b200.build as b
    if not os.path.exists(b.LIB):
        b.build()
    hdr = open(os.path.join(REPO, 'include', 'osbli_b200.h')).read()
    declared = set(re.findall(r'\b(osb_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(b.LIB)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from opensbli_b200.runtime import SYMBOLS
    assert set(SYMBOLS) == declared


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product must refuse to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import opensbli_b200
    plan, _ = load_fixture('sod_teno5_n200')
    with pytest.raises(opensbli_b200.BackendError, match='no CPU fallback'):
        opensbli_b200.Simulation(plan)


def test_plan_roundtrip_and_validation():
    from opensbli_b200 import plan as P
    plan, _ = load_fixture('tgv_teno5_16')
    txt = P.to_text(plan)
    assert txt.startswith('osbli_plan 1\n') and 'conv teno' in txt and 'bc 2 1 periodic' in txt
    bad = dict(plan, order=7)
    with pytest.raises(P.PlanError):
        P.to_text(bad)
    bad = dict(plan, bc=[[dict(type='forcing_strip_wall')] * 2] * 3)
    with pytest.raises(P.PlanError, match='not implemented'):
        P.to_text(bad)
    p2 = P.with_size(plan, [32, 32, 32], delta=[0.1] * 3, dt=1e-3)
    assert p2['np'] == [32, 32, 32] and p2['constants']['dt'] == 1e-3 and plan['np'] == [16, 16, 16]


def test_split_face_plan_text_and_validation():
    """SplitBC faces (bc_core.py:200-217): parts serialised in order; a part must be one plane thick at its face, inside the
    padded block, of a boundary class with a plane kernel"""
    import copy
    from opensbli_b200 import plan as P
    plan, _ = load_fixture('isr_split_48x32')
    plan = {k: v for k, v in plan.items() if k not in ('q0_padded', 'fields')}
    txt = P.to_text(plan)
    assert 'bc 1 0 split\nbc_part 1 0 symmetry -3 20 0 1 0\nbc_part 1 0 inviscid_wall 20 52 0 1 0\n' in txt
    for edit in (lambda b: b['parts'][0].update(range=[-3, 20, 0, 2]),            # two planes along the normal
                 lambda b: b['parts'][0].update(range=[-3, 20, 1, 2]),            # not the boundary plane
                 lambda b: b['parts'][1].update(range=[20, 60, 0, 1]),            # outside the padded block
                 lambda b: b['parts'][1].update(type='periodic'),
                 lambda b: b.update(parts=[])):
        bad = copy.deepcopy(plan)
        edit(bad['bc'][1][0])
        with pytest.raises(P.PlanError, match='split bc'):
            P.to_text(bad)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under opensbli_b200/ may reference it."""
    for root, _, files in os.walk(os.path.join(REPO, 'opensbli_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                s = open(os.path.join(root, f)).read()
                assert 'osbli_oracle' not in s and 'oracle_util' not in s, f


@pytest.fixture(scope='module')
def hostcheck():
    src = os.path.join(REPO, 'tests', 'csrc', 'host_math_check.cpp')
    so = os.path.join(REPO, 'tests', 'csrc', 'libhostcheck.so')
    hdrs = [os.path.join(REPO, 'opensbli_b200', 'csrc', h) for h in ('osb_math.cuh', 'osb_flux.cuh', 'osb_flux3.cuh')]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', src, '-o', so])
    return ctypes.CDLL(so)


SCHEMES = [('weno', 5, 'JS'), ('weno', 5, 'Z'), ('teno', 5, 'JS'), ('teno', 6, 'JS')]


@pytest.mark.parametrize('nd', [1, 2, 3])
@pytest.mark.parametrize('recon', range(4))
@pytest.mark.parametrize('avg', [0, 1])
def test_device_math_matches_oracle(hostcheck, nd, recon, avg):
    """interface_flux<> (sparse eigen-projections, division-free TENO cut-off) vs the oracle's dense
    restatement, on smooth, shocked, nearly-constant and constant 6-point windows."""
    lib = ou.oracle_lib()
    P = ctypes.POINTER(ctypes.c_double)
    conv, order, form = SCHEMES[recon]
    rng = np.random.default_rng(7 + nd)
    nv = nd + 2
    for d in range(nd):
        plan = dict(ndim=nd, np=[8] * nd, delta=[0.1] * nd, conv=conv, order=order, weno_formulation=form,
                    averaging='roe' if avg else 'simple', viscous=False, rk='ls', rk_a=[0.], rk_b=[1.],
                    constants=dict(gama=1.4, dt=0.1, eps=1e-16, TENO_CT=1e-5), bc=[[dict(type='periodic')] * 2] * nd)
        cfg = ou.make_cfg(plan)
        worst = 0.0
        for trial in range(120):
            kind = trial % 4
            amp = {0: 1e-2, 1: 0.3, 2: 1e-6, 3: 0.0}[kind]
            q = np.zeros((6, nv))
            for p in range(6):
                rho = abs(1.0 + 0.3 * amp * rng.standard_normal()) + 0.05
                u = np.array([0.3, -0.2, 0.1][:nd]) + amp * rng.standard_normal(nd)
                pr = abs(1.0 + 0.3 * amp * rng.standard_normal()) + 0.05
                if kind == 1 and p >= 3:
                    rho *= 0.3
                    pr *= 0.2
                q[p, 0] = rho
                q[p, 1:1 + nd] = rho * u
                q[p, nd + 1] = pr / 0.4 + 0.5 * rho * (u ** 2).sum()
            f1, f2 = np.zeros(nv), np.zeros(nv)
            lib.osbo_interface_flux(ctypes.byref(cfg), d, q.ctypes.data_as(P), f1.ctypes.data_as(P))
            rc = hostcheck.hostcheck_interface_flux(nd, d, recon, avg, q.ctypes.data_as(P), ctypes.c_double(1.4),
                                                    ctypes.c_double(1e-16), ctypes.c_double(1e-5), f2.ctypes.data_as(P))
            assert rc == 0
            worst = max(worst, np.abs(f1 - f2).max() / np.abs(f1).max())
            # the pass-split form on the staged layout (what the sweep kernels inline)
            for fn in (hostcheck.hostcheck_interface_flux_split,):
                f3 = np.zeros(nv)
                assert fn(nd, d, recon, avg, q.ctypes.data_as(P), ctypes.c_double(1.4), ctypes.c_double(1e-16), ctypes.c_double(1e-5),
                          f3.ctypes.data_as(P)) == 0
                worst = max(worst, np.abs(f1 - f3).max() / np.abs(f1).max())
        assert worst < 1e-13, (nd, d, recon, avg, worst)


def test_teno_cutoff_is_scale_safe(hostcheck):
    """The division-free cut-off must not under/overflow for tiny eps or large flux magnitudes."""
    lib = ou.oracle_lib()
    P = ctypes.POINTER(ctypes.c_double)
    plan = dict(ndim=1, np=[8], delta=[0.1], conv='teno', order=5, averaging='roe', viscous=False, rk='ls', rk_a=[0.],
                rk_b=[1.], constants=dict(gama=1.4, dt=0.1, eps=1e-40, TENO_CT=1e-6), bc=[[dict(type='periodic')] * 2])
    cfg = ou.make_cfg(plan)
    rng = np.random.default_rng(3)
    for scale in (1e-6, 1.0, 1e6):
        for trial in range(50):
            q = np.zeros((6, 3))
            for p in range(6):
                rho = (1.0 + 1e-3 * rng.standard_normal()) * (0.3 if p > 3 else 1.0)
                u = 0.1 + 1e-3 * rng.standard_normal()
                pr = 1.0 * (0.2 if p > 3 else 1.0)
                q[p] = [rho * scale, rho * u * scale, (pr / 0.4 + 0.5 * rho * u * u) * scale]
            f1, f2 = np.zeros(3), np.zeros(3)
            lib.osbo_interface_flux(ctypes.byref(cfg), 0, q.ctypes.data_as(P), f1.ctypes.data_as(P))
            hostcheck.hostcheck_interface_flux(1, 0, 2, 1, q.ctypes.data_as(P), ctypes.c_double(1.4),
                                               ctypes.c_double(1e-40), ctypes.c_double(1e-6), f2.ctypes.data_as(P))
            assert np.all(np.isfinite(f2))
            assert np.abs(f1 - f2).max() / np.abs(f1).max() < 1e-12
