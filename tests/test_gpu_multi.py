"""GPU tests (-m gpu, need >= 2 GPUs): slab-decomposed run with peer-store halo exchange (CUDA IPC over NVLink) against
the single-GPU run of the same block."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _load(fixture):
    """fixture, optionally repeated along z ('name*2': the slab of each of two ranks must keep >= 6 planes)"""
    from common import load_fixture
    name, _, rep = fixture.partition('*')
    plan, states = load_fixture(name)
    if rep and 'q0_padded' not in plan:           # periodic box: repeat the interior along z
        r = int(rep)
        plan['np'][2] *= r
        states = {0: np.concatenate([states[0]] * r, axis=1)}
    elif rep:
        r = int(rep)
        tile = lambda a: np.concatenate([a[:5]] + [a[5:-5]] * r + [a[-5:]], axis=0)
        nz, h = plan['np'][2], 5
        plan['np'][2] = nz * r
        plan['q0_padded'] = np.stack([tile(a) for a in plan['q0_padded']])
        plan['fields'] = {n: tile(a) for n, a in plan['fields'].items()}
        for pair in plan['bc'][:2]:
            for b in pair:
                if b.get('table') is not None:
                    tb = np.asarray(b['table'])
                    tb = tb.reshape(tb.shape[0], nz + 2 * h, -1)
                    b['table'] = np.stack([tile(x) for x in tb]).reshape(tb.shape[0], -1)
        states = {0: np.stack([a[5:-5, 5:-5, 5:-5] for a in plan['q0_padded']])}
    return plan, states


def _worker(rank, world, port, fixture, nsteps, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from common import load_fixture, pad, initial_padded
    from opensbli_b200.decomp import DistributedSimulation
    plan, states = _load(fixture)
    ds = DistributedSimulation(plan, dist, device=rank)
    k0, nk = ds.offset, ds.nloc
    if 'q0_padded' in plan:      # general path: the padded initial state (halo values of the walls' directions matter)
        q0 = [np.ascontiguousarray(a[k0:k0 + nk + 10]) for a in initial_padded(plan, states)]
    else:
        q0 = pad(ds.plan, states[0][:, k0:k0 + nk])
    ds.set_state(q0)       # collective: uploads also write the halos the neighbours push into
    ds.step(nsteps)
    ds.barrier()
    q = ds.sim.get_state()
    np.save(os.path.join(out, 'q_%d.npy' % rank), np.stack([a[(slice(5, -5),) * a.ndim] for a in q]))
    ds.barrier()
    ds.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('fixture,world', [('tgv_teno5_16', 2), ('tgv_central4_16', 2), ('tcf_teno6_16x24x12', 2), ('tcf_central_16x24x12', 2),
                                           ('trans_40x30x8*2', 2), ('katzer_60x40', 2), ('vst_60x30', 2),
                                           # uneven slabs (17 = 9 + 8 planes; 40 = 14 + 13 + 13) and more ranks
                                           ('tgv_sym_17', 2), ('katzer_60x40', 3), ('tgv_teno5_16*2', 4)])
def test_slabs_match_single_gpu(fixture, world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    import opensbli_b200
    from common import load_fixture, pad, inner
    nsteps = 3
    plan, states = _load(fixture)
    from common import initial_padded
    with opensbli_b200.Simulation(plan, device=0) as sim:
        sim.set_state(initial_padded(plan, states))
        sim.step(nsteps)
        ref = inner(plan, sim.get_state())
    mp.spawn(_worker, args=(world, _free_port(), fixture, nsteps, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), 'q_%d.npy' % r)) for r in range(world)], axis=1)
    diff = np.abs(got - ref)
    where = np.argwhere(diff > 0)
    assert np.array_equal(got, ref), (float(diff.max()), len(where), where[:5].tolist())       # identical arithmetic per point: bit-exact


@pytest.mark.parametrize('case,world,names', [('katzer', 2, ('rho', 'rhou0', 'rhou1', 'rhoE', 'TENO')), ('katzer', 4, ('rho', 'rhou0', 'rhou1', 'rhoE', 'TENO')),
                                              # run-time compiled boundary kernels on every face, slabs along y
                                              ('isr_invwall_generic', 2, ('rho', 'rhou0', 'rhou1', 'rhoE'))])
def test_apps_decomposed_by_the_runner_match_single_gpu(case, world, names, tmp_path):
    """BASELINE configs[3] at its shipped size through the app-level runner on 2 / 4 GPUs (slabs along y: 250 = 63 + 63 + 62 + 62,
    wall and shock-generator faces stay with the ranks that own them): the dataset file equals the single-GPU run's bit for bit."""
    import subprocess
    import shutil
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    from opensbli_b200 import iodata
    repo = os.path.dirname(HERE)
    outs = []
    for n, d in ((1, tmp_path / 'one'), (world, tmp_path / 'many')):
        d.mkdir()
        for f in ('opensbli_b200.plan.json', 'opensbli.cpp'):
            shutil.copy(os.path.join(HERE, 'golden', 'plans', case, f), str(d))
        cmd = [sys.executable, '-m', 'opensbli_b200.run', str(d), '--niter', '10']
        if n > 1:
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
                   '--master-port', str(_free_port()), '-m', 'opensbli_b200.run', str(d), '--niter', '10']
        subprocess.check_call(cmd, cwd=repo, env=dict(os.environ, PYTHONPATH=repo))
        outs.append(iodata.read_datasets(str(d / 'opensbli_output'))[0])
    for name in names:
        assert np.array_equal(outs[0][name][5:-5, 5:-5], outs[1][name][5:-5, 5:-5]), name


def _pipe_worker(rank, world, port, workload, np3, chunk, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from test_gpu_scale import tgv_case
    from opensbli_b200.hostpipe import DistributedHostPipeline
    plan, q0 = tgv_case(np3, workload)
    rng = np.random.default_rng(5)
    for a in q0:
        a *= 1.0 + 0.02 * rng.standard_normal(a.shape)
    with DistributedHostPipeline(plan, dist, device=rank, chunk=chunk) as dp:
        k0, nk = dp.offset, dp.nloc
        pin = lambda: [torch.empty((nk + 10,) + a.shape[1:], dtype=torch.float64, pin_memory=True) for a in q0]
        ti, to = pin(), pin()
        q_in, q_out = [t.numpy() for t in ti], [t.numpy() for t in to]
        for a, b in zip(q_in, q0):
            a[...] = b[k0:k0 + nk + 10]
            a[:5] = np.nan                       # the slab's own halo planes are never read
            a[-5:] = np.nan
        for rep in range(2):                     # twice: staging copy and contexts are reused
            for a in q_out:
                a[...] = np.nan
            dp.advance(q_in, q_out)
        np.save(os.path.join(out, 'q_%d.npy' % rank), np.stack([a[5:-5, 5:-5, 5:-5] for a in q_out]))
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.parametrize('workload,np3,chunk,world', [('teno5', (64, 48, 96), 16, 2), ('central4', (48, 40, 90), 20, 3)])
def test_distributed_window_pipeline_matches_single_gpu(workload, np3, chunk, world, tmp_path):
    """The end-to-end call on a slab-decomposed host-resident block (hostpipe.DistributedHostPipeline): every rank advances its
    slab window by window, the guard planes of its outer windows pulled out of the neighbours' staging copies over NVLink, no
    per-stage halo exchange -- the grid points equal the single-GPU whole-block step bit for bit."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    import opensbli_b200
    from test_gpu_scale import tgv_case
    plan, q0 = tgv_case(np3, workload)
    rng = np.random.default_rng(5)
    for a in q0:
        a *= 1.0 + 0.02 * rng.standard_normal(a.shape)
    with opensbli_b200.Simulation(plan, device=0) as sim:
        sim.set_state(q0)
        sim.step(1)
        ref = np.stack([a[5:-5, 5:-5, 5:-5] for a in sim.get_state()])
    mp.spawn(_pipe_worker, args=(world, _free_port(), workload, np3, chunk, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), 'q_%d.npy' % r)) for r in range(world)], axis=1)
    assert np.isfinite(got).all()
    assert np.array_equal(got, ref), float(np.abs(got - ref).max())
