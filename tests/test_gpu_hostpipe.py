"""GPU tests (-m gpu): the window pipeline over a host-resident block (opensbli_b200/hostpipe.py + osb_host_planes_* in the
C ABI).  A block advanced window by window -- 'open' cut faces, guard planes, three contexts taking turns so that uploads,
sweeps and downloads overlap -- must equal the whole-block end-to-end call bit for bit in every cell a kernel defines, and
the whole-block call is what the parity tests hold against the oracle."""
import numpy as np
import pytest
import torch

from opensbli_b200 import hostpipe
from opensbli_b200.decomp import scheme_halos
from test_gpu_scale import tgv_case
from common import inner, field_errors
import oracle_util as ou

pytestmark = pytest.mark.gpu


def pinned_like(arrays):
    ts = [torch.empty(a.shape, dtype=torch.float64, pin_memory=True) for a in arrays]
    return ts, [t.numpy() for t in ts]


def defined_cells(plan, a):
    hm, _ = scheme_halos(plan)
    return a[tuple(slice(5 - hm, 5 + n + hm) for n in reversed(plan['np']))]


@pytest.mark.parametrize('workload,np3,chunk,nsteps', [('teno5', (64, 48, 96), 32, 1), ('teno5', (64, 48, 90), 32, 1),
                                                       ('central4', (64, 64, 64), 16, 1), ('teno5', (40, 40, 64), 24, 2),
                                                       ('teno5', (64, 48, 96), [8, 16, 40, 20, 12], 1)])
def test_window_pipeline_equals_whole_block_call(workload, np3, chunk, nsteps):
    import opensbli_b200
    plan, q0 = tgv_case(np3, workload)
    rng = np.random.default_rng(5)
    for a in q0:                                    # perturbed: every TENO branch is taken somewhere
        a *= 1.0 + 0.02 * rng.standard_normal(a.shape)
    keep_in, q_in = pinned_like(q0)
    keep_a, out_whole = pinned_like(q0)
    keep_b, out_pipe = pinned_like(q0)
    for a, b in zip(q_in, q0):
        a[...] = b
    with opensbli_b200.Simulation(plan) as sim:
        sim.advance_host(q_in, out_whole, nsteps)
    for a in out_pipe:
        a[...] = np.nan
    with hostpipe.HostPipeline(plan, chunk=chunk, nsteps=nsteps) as pipe:
        launches = pipe.advance(q_in, out_pipe)
        assert launches > 0
        for a in out_pipe:                          # second call through the same contexts (reuse after a download)
            a[...] = np.nan
        pipe.advance(q_in, out_pipe)
    for a, b in zip(out_whole, out_pipe):
        assert np.isfinite(defined_cells(plan, a)).all()
        assert np.array_equal(defined_cells(plan, a), defined_cells(plan, b))
    if workload == 'teno5' and nsteps == 1 and np3[2] == 96:     # and the whole-block call is the oracle's step
        qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], nsteps)
        err = field_errors(plan, inner(plan, out_pipe), inner(plan, qo))
        assert max(err) < 1e-12, err


def test_plane_copies_straight_from_host():
    """osb_host_planes_upload / _ready / _download without the staging copy: a block uploaded in two runs of planes, stepped
    and read back plane-wise equals upload-step-download of whole arrays"""
    import opensbli_b200
    plan, q0 = tgv_case((32, 24, 40), 'teno5')
    keep_in, q_in = pinned_like(q0)
    keep_a, out = pinned_like(q0)
    for a, b in zip(q_in, q0):
        a[...] = b
    with opensbli_b200.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.step(2)
        want = sim.get_state()
    with opensbli_b200.Simulation(plan) as sim:
        sim.planes_upload(q_in, 0, 0, 17)
        sim.planes_upload(q_in, 17, 17, 33)
        sim.planes_ready()
        sim.step(2, sync=False)
        sim.planes_download(out, 0, 0, 50)
        sim.planes_sync()
    for a, b in zip(want, out):
        assert np.array_equal(defined_cells(plan, a), defined_cells(plan, b))
