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


# ---- slab-decomposed blocks: the window pipeline per rank, guard planes from the neighbours (gloo, CPU) -------------------------
def _free_port():
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def slab_case(workload, nz):
    """a periodic box long enough along the slab axis for every rank to hold guard + halo planes (Taylor-Green state of the bench)"""
    import math
    import bench
    np3 = (8, 8, nz)
    plan = bench.tgv_plan(list(np3), workload)
    plan['delta'] = [2.0 * math.pi / n for n in np3]
    plan['constants']['dt'] = 0.003385 * 64 / 16
    q = [np.zeros(tuple(n + 10 for n in reversed(np3))) for _ in range(5)]
    bench.tgv_state_into(q, plan, 0, nz)
    rng = np.random.default_rng(3)
    for a in q:
        a[5:-5, 5:-5, 5:-5] *= 1.0 + 0.02 * rng.standard_normal(a[5:-5, 5:-5, 5:-5].shape)
    return plan, q


def _dist_worker(rank, world, port, fixture, chunk, out):
    import os
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from opensbli_b200.decomp import local_extent
    plan, q0 = slab_case(*fixture)
    k0, nk = local_extent(plan, rank, world)
    q_in = [np.ascontiguousarray(a[k0:k0 + nk + 10]).copy() for a in q0]
    for a in q_in:                      # the slab's own halo planes are never read: poison them
        a[:5] = np.nan
        a[-5:] = np.nan
    q_out = [np.full_like(a, np.nan) for a in q_in]

    def exchange(stage, G, n):
        """what osb_staging_pull does over NVLink, with gloo send / recv on the host staging copy"""
        low, high = (rank - 1) % world, (rank + 1) % world
        reqs, recvs = [], []
        for m, b in enumerate(stage.buf):
            assert not np.isnan(b[G:2 * G]).any() and not np.isnan(b[n:n + G]).any(), 'boundary planes must be staged before the exchange'
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(b[G:2 * G])), low, tag=2 * m))          # my bottom planes
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(b[n:n + G])), high, tag=2 * m + 1))     # my top planes
        for m, b in enumerate(stage.buf):
            t_hi, t_lo = torch.empty(b[:G].shape, dtype=torch.float64), torch.empty(b[:G].shape, dtype=torch.float64)
            reqs.append(dist.irecv(t_hi, high, tag=2 * m))       # high neighbour's bottom planes -> my upper guard
            reqs.append(dist.irecv(t_lo, low, tag=2 * m + 1))    # low neighbour's top planes -> my lower guard
            recvs.append((b, t_lo, t_hi))
        for r in reqs:
            r.wait()
        for b, t_lo, t_hi in recvs:
            b[:G] = t_lo.numpy()
            b[G + n:] = t_hi.numpy()

    class SlabStage(HostStage):
        def sync(self):
            pass

        def upload(self, arrays, host_plane0, plane0, nplanes):
            if self.buf is None:
                self.buf = [np.full((self.nplanes,) + a.shape[1:], np.nan) for a in arrays]
            HostStage.upload(self, arrays, host_plane0, plane0, nplanes)

    pipe = hostpipe.DistributedHostPipeline(plan, dist, device=-1, chunk=chunk, contexts=2, factory=OracleWindow,
                                            stage_factory=SlabStage, exchange=exchange)
    pipe.advance(q_in, q_out)
    np.save(os.path.join(out, 'q_%d.npy' % rank), np.stack([a[5:-5] for a in q_out]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('fixture,world,chunk', [(('teno5', 32), 2, 6), (('central4', 24), 2, 5), (('teno5', 40), 3, 8)])
def test_distributed_window_pipeline_reproduces_whole_block(fixture, world, chunk, tmp_path):
    """world ranks, each advancing its slab window by window with guard planes taken from the neighbours' staging copies:
    together they reproduce the whole-block oracle step bit for bit (uneven slabs at world = 3: 14 + 13 + 13 planes)"""
    import torch.multiprocessing as mp
    plan, q0 = slab_case(*fixture)
    whole, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 1)
    mp.spawn(_dist_worker, args=(world, _free_port(), fixture, chunk, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(str(tmp_path / ('q_%d.npy' % r))) for r in range(world)], axis=1)
    hm, _ = scheme_halos(plan)
    s = (slice(None), slice(None)) + tuple(slice(5 - hm, 5 + n + hm) for n in reversed(plan['np'][:-1]))
    for m in range(len(whole)):
        assert np.array_equal(got[m][s[1:]], whole[m][5:-5][s[1:]]), m
