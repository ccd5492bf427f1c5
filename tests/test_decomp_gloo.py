"""CPU tests of the N>1 path (world_size 2 and 4, gloo): the slab decomposition logic of opensbli_b200.decomp
(extents, neighbours, 'exchange' faces, which planes are pushed where) driven with the oracle as the per-rank
solver must reproduce the single-domain oracle run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def host_exchange(q, plan_local, low, high, halo=5):
    """What osb_halo_push does with peer stores, done with gloo send/recv on host arrays."""
    from opensbli_b200.decomp import push_planes
    pp = push_planes(plan_local, halo)
    reqs, recv = [], []
    for m, a in enumerate(q):
        if high is not None:
            (s0, s1), _ = pp['up']
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[s0:s1])), high, tag=2 * m))
        if low is not None:
            (s0, s1), _ = pp['down']
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[s0:s1])), low, tag=2 * m + 1))
    for m, a in enumerate(q):
        if low is not None:      # my low halo <- low neighbour's 'up' planes
            _, (d0, d1) = pp['up']
            t = torch.empty(a[d0:d1].shape, dtype=torch.float64)
            reqs.append(dist.irecv(t, low, tag=2 * m))
            recv.append((a, d0, d1, t))
        if high is not None:     # my high halo <- high neighbour's 'down' planes
            _, (d0, d1) = pp['down']
            t = torch.empty(a[d0:d1].shape, dtype=torch.float64)
            reqs.append(dist.irecv(t, high, tag=2 * m + 1))
            recv.append((a, d0, d1, t))
    for r in reqs:
        r.wait()
    for a, d0, d1, t in recv:
        a[d0:d1] = t.numpy()


def _worker(rank, world, port, fixture, nsteps, out, order='bcs_first'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle_util as ou
    from common import load_fixture, pad
    from opensbli_b200.decomp import local_plan, local_extent, neighbours
    plan, states = load_fixture(fixture)
    lp = local_plan(plan, rank, world)
    k0, nk = local_extent(plan, rank, world)
    low, high = neighbours(plan, rank, world)
    if 'q0_padded' in plan:      # general path: padded initial state, slab incl. its halo planes
        q = [np.ascontiguousarray(a[k0:k0 + nk + 10]) for a in plan['q0_padded']]
    else:
        q = [np.ascontiguousarray(a) for a in pad(lp, states[0][:, k0:k0 + nk])]
    rk = [np.zeros_like(a) for a in q]
    if order == 'exchange_first':
        # the order of the GPU protocol: planes first, rank-local BCs after the neighbours' planes have landed (they also
        # rewrite the halo parts of the received planes).  A copy of the plan with every face an 'exchange' face makes the
        # oracle's stage call skip the BCs, which are then applied separately.
        import copy
        nobc = copy.deepcopy(lp)
        nobc['bc'] = [[{'type': 'exchange'}, {'type': 'exchange'}] for _ in range(lp['ndim'])]
        nobc_cl = [[b.get('closure') for b in pair] for pair in lp['bc']]
        for d, pair in enumerate(nobc['bc']):
            for s_, b in enumerate(pair):
                if nobc_cl[d][s_]:
                    b['closure'] = nobc_cl[d][s_]          # one-sided derivative rows belong to the face, not to its BC kernel
        for _ in range(nsteps):
            host_exchange(q, lp, low, high)
            ou.oracle_apply_bcs(lp, q)
            ou.oracle_stage(nobc, q, rk, -1)
            for s in range(len(lp['rk_a'])):
                ou.oracle_stage(nobc, q, rk, s)
                host_exchange(q, lp, low, high)
                ou.oracle_apply_bcs(lp, q)
        np.save(os.path.join(out, 'q_%d.npy' % rank), np.stack([a[5:-5] for a in q]))
        dist.barrier()
        dist.destroy_process_group()
        return
    for _ in range(nsteps):
        ou.oracle_stage(lp, q, rk, -1)
        host_exchange(q, lp, low, high)
        for s in range(len(lp['rk_a'])):
            ou.oracle_stage(lp, q, rk, s)
            host_exchange(q, lp, low, high)
    np.save(os.path.join(out, 'q_%d.npy' % rank), np.stack([a[5:-5] for a in q]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('fixture,world', [('tgv_teno5_16', 2), ('tgv_teno5_16', 4), ('tgv_central4_16', 2),
                                           # general path: metric fields follow the slab (3-D channel, periodic slab axis) ...
                                           ('tcf_central_16x24x12', 2), ('tcf_teno6_16x24x12', 2),
                                           # ... and physical walls / closures stay with the ranks that own them (2-D, slabs along y)
                                           ('lam2d_16x64', 4), ('vst_60x30', 2),
                                           # blocks that do not divide evenly: slabs of 9 + 8 and 22 + 21 + 21 planes
                                           ('tgv_sym_17', 2), ('lam2d_16x64', 3)])
def test_slab_decomposition_reproduces_single_domain(fixture, world, tmp_path):
    import oracle_util as ou
    from common import load_fixture, pad, inner
    nsteps = 2
    mp.spawn(_worker, args=(world, _free_port(), fixture, nsteps, str(tmp_path)), nprocs=world, join=True)
    plan, states = load_fixture(fixture)
    from common import initial_padded
    q, _ = ou.oracle_advance(plan, initial_padded(plan, states), nsteps)
    ref = inner(plan, q)
    cut = (slice(None), slice(None)) + (slice(5, -5),) * (plan['ndim'] - 1)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), 'q_%d.npy' % r))[cut] for r in range(world)], axis=1)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)        # same arithmetic on every point: bit-exact


@pytest.mark.parametrize('fixture,world', [('tcf_central_16x24x12', 2), ('vst_60x30', 2), ('tgv_teno5_16', 2)])
def test_exchange_before_rank_local_bcs_reproduces_single_domain(fixture, world, tmp_path):
    """The order the GPU driver uses on every path (exchange, then the rank-local BCs) gives the single-domain result bit for bit."""
    import oracle_util as ou
    from common import load_fixture, inner, initial_padded
    nsteps = 2
    mp.spawn(_worker, args=(world, _free_port(), fixture, nsteps, str(tmp_path), 'exchange_first'), nprocs=world, join=True)
    plan, states = load_fixture(fixture)
    q, _ = ou.oracle_advance(plan, initial_padded(plan, states), nsteps)
    ref = inner(plan, q)
    cut = (slice(None), slice(None)) + (slice(5, -5),) * (plan['ndim'] - 1)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), 'q_%d.npy' % r))[cut] for r in range(world)], axis=1)
    assert np.array_equal(got, ref)


def test_decomp_helpers():
    from common import load_fixture
    from opensbli_b200 import decomp
    plan, _ = load_fixture('tgv_teno5_16')
    assert decomp.local_extent(plan, 3, 4) == (12, 4)
    assert [decomp.local_extent({'ndim': 1, 'np': [250]}, r, 4) for r in range(4)] == [(0, 63), (63, 63), (126, 62), (188, 62)]
    assert decomp.neighbours(plan, 0, 4) == (3, 1) and decomp.neighbours(plan, 3, 4) == (2, 0)
    lp = decomp.local_plan(plan, 1, 2)
    assert lp['np'] == [16, 16, 8] and lp['bc'][2][0]['type'] == 'exchange' and lp['bc'][0][0]['type'] == 'periodic'
    pp = decomp.push_planes(lp)
    assert pp['up'] == ((5 + 8 - 3, 5 + 8), (2, 5)) and pp['down'] == ((5, 9), (13, 17))
    with pytest.raises(Exception):
        decomp.local_plan(plan, 0, 5)          # 16 planes over 5 ranks: slabs of 3 planes are thinner than the halo
    # point-wise user kernels (statistics) follow the slab
    withuk = dict(plan, user_kernels=[{'name': 'stats', 'range': [0, 16, 0, 16, 0, 16], 'when': 'iteration_end'}])
    assert decomp.local_plan(withuk, 1, 4)['user_kernels'][0]['range'] == [0, 16, 0, 16, 0, 4]
    assert withuk['user_kernels'][0]['range'] == [0, 16, 0, 16, 0, 16]
    sod, _ = load_fixture('sod_teno5_n200')
    assert decomp.neighbours(sod, 0, 2) == (None, 1) and decomp.neighbours(sod, 1, 2) == (0, None)
    lp = decomp.local_plan(sod, 0, 2)
    assert lp['bc'][0][0]['type'] == 'dirichlet' and lp['bc'][0][1]['type'] == 'exchange'
