"""GPU tests (-m gpu): parity at the sizes BASELINE.json's configs are stated on -- grids that span several 32 x 8 tiles,
several z-chunks of the marching kernels and several 128-point staging blocks, i.e. the code paths the 512^3 headline
times.  The CUDA path (through the C ABI) is compared with the CPU oracle and, where oracle/_ref holds the executable,
with the reference's own generated C run on the same grid from the same initial state."""
import math
import os
import sys

import numpy as np
import pytest

from common import pad, inner, field_errors, tol_for
import oracle_util as ou

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PLANS = os.path.join(HERE, 'golden', 'plans')
sys.path.insert(0, os.path.join(HERE, 'golden'))


def describe(err):
    return ['%.2e' % e for e in err]


def tgv_case(np3, workload):
    import bench
    plan = bench.tgv_plan(list(np3), workload)
    plan['delta'] = [2.0 * math.pi / n for n in np3]
    plan['constants']['dt'] = 0.003385 * 64 / np3[0]
    q = [np.zeros(tuple(n + 10 for n in reversed(np3))) for _ in range(5)]
    bench.tgv_state_into(q, plan, 0, np3[2])
    return plan, q


def run_gpu(plan, q0, nsteps):
    import opensbli_b200
    with opensbli_b200.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.step(nsteps)
        return sim.get_state()


def increment_errors(plan, q0, qg, qo):
    """error of the CHANGE of the state over the steps, relative to the largest change per field: the per-step tolerance of the
    north star is relative to the field norm, which a time step of 1e-3 turns into a loose bound on the residual itself"""
    q0i = inner(plan, q0)
    dg, do = qg - q0i, qo - q0i
    return [float(np.abs(dg[m] - do[m]).max() / max(np.abs(do[m]).max(), 1e-300)) for m in range(len(do))]


def ref_threads():
    return os.cpu_count() or 1


def test_tgv_central4_64cubed_is_config1():
    """BASELINE configs[1] exactly: apps/taylor_green_vortex as shipped, Central(4) + RungeKutta(3), 64^3 -- two x-tiles,
    eight y-tiles, several z-chunks of k_central3d_fused; 3 steps vs the oracle and vs the reference executable."""
    plan, q0 = tgv_case((64, 64, 64), 'central4')
    qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], 3))
    qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 3)
    err = field_errors(plan, qg, inner(plan, qo))
    print('64^3 central4 vs oracle', describe(err))
    assert max(err) < 1e-12, err
    if ou.have_ref('tgv_central4'):
        names = ['rho', 'rhou0', 'rhou1', 'rhou2', 'rhoE']
        r = ou.run_ref('tgv_central4', dict(block0np0=64, block0np1=64, block0np2=64, niter=3, dt=plan['constants']['dt']), names,
                       exe='ref_omp', threads=ref_threads())
        ref = np.stack([r[n][5:-5, 5:-5, 5:-5] for n in names])
        err = field_errors(plan, qg, ref)
        print('64^3 central4 vs reference executable', describe(err))
        assert max(err) < 1e-12, err


@pytest.mark.parametrize('np3,nsteps', [((96, 72, 80), 2), ((128, 128, 128), 1)], ids=['96x72x80', '128cubed'])
def test_tgv_teno5_multi_tile(np3, nsteps):
    """The headline path (k_flux2_x/yz + k_viscous3d_tiled<RK, FROMQ>) on grids wider than one tile in every direction:
    non-cubic 96 x 72 x 80 (ragged z-chunks) against the oracle, 128^3 against the reference executable (oracle if absent)."""
    plan, q0 = tgv_case(np3, 'teno5')
    qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], nsteps))
    assert np.isfinite(qg).all()
    names = ['rho', 'rhou0', 'rhou1', 'rhou2', 'rhoE']
    if np3[0] <= 96 or not ou.have_ref('tgv_teno5'):
        qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], nsteps)
        err = field_errors(plan, qg, inner(plan, qo))
        inc = increment_errors(plan, q0, qg, inner(plan, qo))
        print(np3, 'TENO5 vs oracle', describe(err), 'increments', describe(inc))
        assert max(err) < tol_for(plan, nsteps), err
        assert max(inc[:3] + inc[4:]) < 1e-10, inc          # (rhou2 starts at zero and stays tiny: its increments are round-off)
    if ou.have_ref('tgv_teno5'):
        r = ou.run_ref('tgv_teno5', dict(block0np0=np3[0], block0np1=np3[1], block0np2=np3[2], niter=nsteps, dt=plan['constants']['dt']),
                       names, exe='ref_omp', threads=ref_threads())
        ref = np.stack([r[n][5:-5, 5:-5, 5:-5] for n in names])
        err = field_errors(plan, qg, ref)
        print(np3, 'TENO5 vs reference executable', describe(err))
        assert max(err) < tol_for(plan, nsteps), err


@pytest.mark.parametrize('name', ['tcf_teno6', 'tcf_central'])
def test_channel_48x40x36(name):
    """3-D general path (k_viscous3d_tiled_general, closures, stretched grid) on 48 x 40 x 36: two x-tiles, five y-tiles;
    cold kernels by the runner, 2 steps vs the oracle from the same cold data."""
    from opensbli_b200 import run as R
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides={'block0np0': 48, 'block0np1': 40, 'block0np2': 36})
    q0 = R.initial_state(plan_sym, cold)
    qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], 2))
    qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 2)
    err = field_errors(plan, qg, inner(plan, qo))
    print(name, '48x40x36', describe(err))
    assert max(err) < 1e-12, err


@pytest.mark.parametrize('config', ['katzer', 'katzer_wenoz'])
def test_katzer_500x250_as_shipped(config):
    """BASELINE configs[3] at its shipped size (500 x 250; adaptive TENO5 as shipped and the WENO-Z variant BASELINE names):
    10 steps from the reference's own initial state and metric arrays (dumped by its executable with niter=0) against the
    reference executable's state after 10 steps; without oracle/_ref the cold data come from the runner and the oracle checks."""
    import make_golden as G
    N0, N1, nsteps = 500, 250, 10
    plan = G.katzer_plan(N0, N1) if config == 'katzer' else G.katzer_wenoz_plan(N0, N1)
    names = ['rho', 'rhou0', 'rhou1', 'rhoE']
    if ou.have_ref(config):
        # the same build for the initial state and the final one: the degree-50 polynomial of the initial profile moves by
        # 1e-10 between the -ffp-contract=off and the -O3 build of the reference's own initialisation kernel
        r0 = ou.run_ref(config, dict(block0np0=N0, block0np1=N1, niter=0), names + ['D11', 'SD111'], exe='ref_omp', dump_all=True, threads=ref_threads())
        plan['fields'] = {'D11': r0['D11'], 'SD111': r0['SD111']}
        plan['bc'][1][1]['table'] = G.katzer_dirichlet_table(N0)
        q0 = [np.ascontiguousarray(r0[n]) for n in names]
        qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], nsteps))
        r = ou.run_ref(config, dict(block0np0=N0, block0np1=N1, niter=nsteps), names, exe='ref_omp', threads=ref_threads())
        ref = np.stack([r[n][5:-5, 5:-5] for n in names])
        err = field_errors(plan, qg, ref)
        print(config, '500x250 vs reference executable', describe(err))
        assert max(err) < tol_for(plan, nsteps), err
    else:
        if config != 'katzer':
            pytest.skip('oracle/_ref/%s absent and no plan fixture for this variant' % config)
        from opensbli_b200 import run as R
        plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, 'katzer'))
        assert plan['np'] == [N0, N1]
        q0 = R.initial_state(plan_sym, cold)
        qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], nsteps))
    qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], nsteps)
    err = field_errors(plan, qg, inner(plan, qo))
    print(config, '500x250 vs oracle', describe(err))
    assert max(err) < tol_for(plan, nsteps), err


def test_tgv_64_cubed_100_steps_norms_and_diagnostics():
    """north_star: "final-time L2 norms and diagnostics within 1e-10".  TGV TENO5 64^3 advanced 100 steps (a non-chaotic
    horizon) on the GPU and by the reference executable: L2 norms of the conserved fields, and the volume-averaged kinetic
    energy / enstrophy -- from the device reductions (osb_diagnostics) on one side, computed offline with numpy from the
    reference's dump (how the reference workflow obtains them) on the other."""
    import opensbli_b200
    from test_gpu_diag import numpy_diagnostics
    if not ou.have_ref('tgv_teno5'):
        pytest.skip('oracle/_ref/tgv_teno5 not built')
    n, nsteps = 64, 100
    plan, q0 = tgv_case((n, n, n), 'teno5')
    with opensbli_b200.Simulation(plan) as sim:
        sim.set_state([a.copy() for a in q0])
        sim.step(nsteps)
        d = sim.diagnostics()
        qg = inner(plan, sim.get_state())
    names = ['rho', 'rhou0', 'rhou1', 'rhou2', 'rhoE']
    r = ou.run_ref('tgv_teno5', dict(block0np0=n, block0np1=n, block0np2=n, niter=nsteps, dt=plan['constants']['dt']), names,
                   exe='ref_omp', threads=ref_threads())
    # the reference's dump carries the halos its last exchange left: periodic images, enough for the 4th-order vorticity
    want = numpy_diagnostics(plan, [r[f] for f in names])
    npts = float(n ** 3)
    for key in ('sum_ke', 'sum_enstrophy', 'sum_rho', 'sum_rhoE'):
        rel = abs(d[key] - want[key]) / abs(want[key])
        print(key, d[key] / npts, want[key] / npts, 'rel %.2e' % rel)
        assert rel < 1e-10, (key, d[key], want[key])
    ref = np.stack([r[f][5:-5, 5:-5, 5:-5] for f in names])
    for m, f in enumerate(names):
        l2g, l2r = np.sqrt(np.mean(qg[m] ** 2)), np.sqrt(np.mean(ref[m] ** 2))
        assert abs(l2g - l2r) <= 1e-10 * max(l2r, 1e-30), (f, l2g, l2r)
    err = field_errors(plan, qg, ref)
    print('100 steps, fields', describe(err))
    assert max(err) < 1e-10, err


@pytest.mark.parametrize('conv,order,form,avg', [('weno', 5, 'Z', 'roe'), ('weno', 5, 'JS', 'simple'), ('teno', 5, 'JS', 'simple'), ('teno', 6, 'JS', 'roe')],
                         ids=['wenoZ-roe', 'wenoJS-simple', 'teno5-simple', 'teno6-roe'])
def test_other_reconstructions_on_the_marching_and_tile_sweeps_at_scale(conv, order, form, avg):
    """The marching TMA sweeps are instantiated for WENO5-JS/Z and both averagings as well (TENO6 keeps the tile kernel): the
    same 96 x 72 x 80 block, non-symmetric smooth state, 1 step against the oracle."""
    from test_gpu_parity import synthetic_state
    np3 = (96, 72, 80)
    plan, _ = tgv_case(np3, 'teno5')
    plan.update(conv=conv, order=order, weno_formulation=form, averaging=avg)
    plan['constants'].update(Minf=0.5, Re=200.0, dt=2e-3, TENO_CT=1e-5)
    q0 = pad(plan, synthetic_state(plan))
    qg = inner(plan, run_gpu(plan, [a.copy() for a in q0], 1))
    qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 1)
    err = field_errors(plan, qg, inner(plan, qo))
    inc = increment_errors(plan, q0, qg, inner(plan, qo))
    print(conv, order, form, avg, describe(err), 'increments', describe(inc))
    assert max(err) < tol_for(plan, 1), err
    assert max(inc) < (1e-6 if form == 'Z' else 1e-10), inc      # the residual itself (WENO-Z: the reference's own noise floor, see tests/common.py)
