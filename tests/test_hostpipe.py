"""CPU tests of the window pipeline's host logic (opensbli_b200/hostpipe.py): windows, guard planes, periodic wrap and the
planes each window returns, with the ORACLE standing in for the per-window solver through the same five calls the GPU
contexts receive.  The windowed advance must reproduce the whole-block oracle run bit for bit in every cell a kernel
defines.  (The GPU test of the same property through the C ABI is in tests/test_gpu_hostpipe.py.)"""
import numpy as np
import pytest

import oracle_util as ou
from common import load_fixture, pad
from opensbli_b200 import hostpipe
from opensbli_b200.decomp import scheme_halos
from opensbli_b200.plan import PlanError


class OracleWindow(object):
    """Stands in for runtime.Simulation: padded local arrays on the host, the oracle as the solver.  Planes nobody fed are NaN, so a
    window stepped on incomplete data shows up."""

    def __init__(self, plan):
        self.plan = plan
        self.q = [np.full(ou.padded_shape(plan), np.nan) for _ in range(plan['ndim'] + 2)]   # NaN: unset planes show up
        self.log = []

    def stage_fed(self):
        assert not any(np.isnan(a).any() for a in self.q), 'window stepped before all of its planes were fed'

    def step(self, nsteps, sync=True):
        self.q, _ = ou.oracle_advance(self.plan, self.q, nsteps)

    def planes_download(self, arrays, host_plane0, plane0, nplanes):
        for a, b in zip(arrays, self.q):
            a[host_plane0:host_plane0 + nplanes] = b[plane0:plane0 + nplanes]

    def planes_sync(self):
        self.q = [np.full_like(a, np.nan) for a in self.q]

    def launch_count(self):
        return 0

    def close(self):
        pass


class HostStage(object):
    """Stands in for runtime.Stage: the staging copy of the block; planes never uploaded stay NaN."""

    def __init__(self, nv, plane, nplanes):
        self.buf = None
        self.nv, self.nplanes = nv, nplanes
        self.uploaded = 0

    def upload(self, arrays, host_plane0, plane0, nplanes):
        if self.buf is None:
            self.buf = [np.full((self.nplanes,) + a.shape[1:], np.nan) for a in arrays]
        for a, b in zip(arrays, self.buf):
            assert np.isnan(b[plane0:plane0 + nplanes]).all(), 'plane uploaded twice'
            b[plane0:plane0 + nplanes] = a[host_plane0:host_plane0 + nplanes]
        self.uploaded += nplanes

    def feed(self, sim, stage_plane0, plane0, nplanes):
        for a, b in zip(self.buf, sim.q):
            b[plane0:plane0 + nplanes] = a[stage_plane0:stage_plane0 + nplanes]

    def sync(self):
        assert self.uploaded == self.nplanes, 'every plane crosses to the device exactly once'
        self.buf, self.uploaded = None, 0

    def close(self):
        pass


def defined_cells(plan, a):
    """grid points + the halo cells the periodic boundary-condition pass defines: hm planes on either side (periodic.py:42-56)"""
    hm, _ = scheme_halos(plan)
    return a[tuple(slice(5 - hm, 5 + n + hm) for n in reversed(plan['np']))]


@pytest.mark.parametrize('fixture,chunk,nsteps', [('tgv_teno5_16', 8, 1), ('tgv_teno5_16', 5, 1), ('tgv_central4_16', 4, 2),
                                                  ('tgv_central4_16', 6, 1), ('tgv_teno5_16', [2, 3, 6, 5], 1)])
def test_windowed_advance_reproduces_whole_block(fixture, chunk, nsteps):
    plan, states = load_fixture(fixture)
    q0 = pad(plan, states[0])
    whole, _ = ou.oracle_advance(plan, [a.copy() for a in q0], nsteps)
    out = [np.full_like(a, np.nan) for a in q0]
    with hostpipe.HostPipeline(plan, chunk=chunk, nsteps=nsteps, contexts=2, factory=OracleWindow, stage_factory=HostStage) as pipe:
        assert pipe.guard == scheme_halos(plan)[0] * (len(plan['rk_a']) * nsteps - 1)
        pipe.advance(q0, out)
        up, down = pipe.bytes_per_call()
        assert down >= up > 0
    for a, b in zip(whole, out):
        assert np.array_equal(defined_cells(plan, a), defined_cells(plan, b))


@pytest.mark.parametrize('fixture,less', [('tgv_teno5_16', 2), ('tgv_central4_16', 1)])
def test_one_plane_less_guard_is_not_enough(fixture, less):
    """the guard depth is tight: with the stencil reach minus one (TENO5 reads 3 planes away, Central(4) 2) the result differs"""
    plan, states = load_fixture(fixture)
    q0 = pad(plan, states[0])
    whole, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 1)
    out = [np.zeros_like(a) for a in q0]
    real = hostpipe.stencil_depth
    try:
        hostpipe.stencil_depth = lambda p: less
        with hostpipe.HostPipeline(plan, chunk=8, factory=OracleWindow, stage_factory=HostStage) as pipe:
            pipe.advance(q0, out)
    finally:
        hostpipe.stencil_depth = real
    assert not all(np.array_equal(defined_cells(plan, a), defined_cells(plan, b)) for a, b in zip(whole, out))


def test_window_helpers():
    assert hostpipe.windows(16, 8) == [(0, 8), (8, 16)]
    assert hostpipe.windows(16, 6) == [(0, 6), (6, 12), (10, 16)]
    assert hostpipe.windows(8, 8) == [(0, 8)]
    assert hostpipe.windows(16, [2, 4, 6, 4]) == [(0, 2), (2, 6), (6, 12), (12, 16)]
    assert hostpipe.ramp(512) == [32, 64, 128, 128, 96, 64] and sum(hostpipe.ramp(1024)) == 1024
    with pytest.raises(PlanError):
        hostpipe.windows(16, [4, 4])
    with pytest.raises(PlanError):
        hostpipe.ramp(100)
    assert hostpipe.wrapped_runs(-13, 30, 16) == [(3, 0, 13), (0, 13, 16), (0, 29, 1)]
    assert hostpipe.wrapped_runs(2, 5, 16) == [(2, 0, 5)]
    runs = hostpipe.wrapped_runs(500, 90, 512)
    assert runs == [(500, 0, 12), (0, 12, 78)]


def test_out_of_scope_plans_are_refused():
    plan, _ = load_fixture('sod_teno5_n200')                  # not periodic along the slab axis
    with pytest.raises(PlanError):
        hostpipe.window_plan(plan, 50)
    plan, _ = load_fixture('ewc_teno5_32')                    # per-point metric arrays
    with pytest.raises(PlanError):
        hostpipe.window_plan(plan, 8)
    plan, _ = load_fixture('tgv_teno5_16')
    with pytest.raises(PlanError):
        hostpipe.window_plan(plan, 16)                        # one window = the whole block
    p, g = hostpipe.window_plan(plan, 8)
    assert g == 6 and p['np'] == [16, 16, 20] and p['bc'][2][0]['type'] == 'open' and p['bc'][0][0]['type'] == 'periodic'


def test_aliasing_is_refused():
    plan, states = load_fixture('tgv_central4_16')
    q0 = pad(plan, states[0])
    with hostpipe.HostPipeline(plan, chunk=8, factory=OracleWindow, stage_factory=HostStage) as pipe:
        with pytest.raises(ValueError):
            pipe.advance(q0, q0)
