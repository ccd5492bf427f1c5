"""CPU tests of the kernels the back end PRINTS for run-time compilation (filters, boundary classes without a hand-written
kernel, SplitBC parts): the same CUDA C the GPU compiles with NVRTC is compiled for the host (tests/hostsim.py) and run
between oracle steps, against the reference's own golden states.  Works from the committed plan fixtures, no reference needed."""
import os

import numpy as np
import pytest

import hostsim
import oracle_util as ou
from common import load_fixture, field_errors

PLANS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'plans')
APPS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'apps')


def oracle_plan(plan):
    return {k: v for k, v in plan.items() if k not in ('user_kernels', 'user_fields')}


def test_weno_filter_kernels_reproduce_the_reference():
    """central-4 TGV + WENOFilter (filters/WENO_filter.py): oracle step, then the 15 printed loops on the host, 3 iterations"""
    from opensbli_b200 import run as R
    over = {'block0np0': 16, 'block0np1': 16, 'block0np2': 16, 'dt': 0.003385 * 64 / 16}
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, 'tgv_wf'), overrides=over)
    assert plan['halos'] == [3, 4] and len(plan['user_kernels']) == 15
    z = np.load(os.path.join(APPS, 'tgv_wf_16.npz'))
    q = [a.copy() for a in R.initial_state(plan_sym, cold)]
    hk = hostsim.HostKernels(plan['user_kernels'], q[0].shape)
    rk = None
    for it in range(3):
        q, rk = ou.oracle_advance(oracle_plan(plan), q, 1, rk)
        for n, a in zip(plan_sym['q_names'], q):
            hk.fields[n] = a
        hk.run('iteration_end')
        if it + 1 in (1, 3):
            err = field_errors(plan, np.stack([a[5:-5, 5:-5, 5:-5] for a in q]), z['q%d' % (it + 1)])
            assert max(err) < 1e-12, (it + 1, err)
    kappa = hk.fields['kappa'][5:-5, 5:-5, 5:-5]
    assert np.abs(kappa - z['stat_kappa']).max() < 1e-9 * np.abs(z['stat_kappa']).max()      # a ratio of squared derivatives: cancellation


def test_sfd_filter_kernel_reproduces_the_reference():
    """Katzer + SFD (filters/SFD.py), from the reference's own cold data: oracle step, then `Apply the filter` on the host"""
    from opensbli_b200 import run as R
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, 'katzer_sfd'), overrides={'block0np0': 60, 'block0np1': 40})
    z = np.load(os.path.join(APPS, 'katzer_sfd_60x40.npz'))
    oplan = oracle_plan(plan)
    oplan['fields'] = {'D11': np.ascontiguousarray(z['field_D11']), 'SD111': np.ascontiguousarray(z['field_SD111'])}
    q = [np.ascontiguousarray(a).copy() for a in z['q0_padded']]
    hk = hostsim.HostKernels(plan['user_kernels'], q[0].shape)
    names = plan_sym['q_names']
    for n, a in zip(names, q):
        hk.fields[n + '_filt'] = a.copy()
    rk = None
    for it in range(10):
        q, rk = ou.oracle_advance(oplan, q, 1, rk)
        for n, a in zip(names, q):
            hk.fields[n] = a
        hk.run('iteration_end')
    err = field_errors(plan, np.stack([a[5:-5, 5:-5] for a in q]), z['q10'])
    assert max(err) < 1e-12, err
    for n in names:
        ref = z['stat_%s_filt' % n]
        assert np.abs(hk.fields[n + '_filt'][5:-5, 5:-5] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


CASES = [('sod_zgo_generic', 'sod_zgo_n200', {'block0np0': 200}), ('sod_pout_generic', 'sod_pout_n200', {'block0np0': 200}),
         ('isr_invwall_generic', 'isr_invwall_48x32', {'block0np0': 48, 'block0np1': 32}),
         ('isr_split', 'isr_split_48x32', dict(block0np0=48, block0np1=32, split_range_100=[0, 20, 0, 1], split_halo_range_100=[-3, 0, 0, 0],
                                               split_range_101=[20, 48, 0, 1], split_halo_range_101=[0, 4, 0, 0]))]


@pytest.mark.parametrize('name,fixture,over', CASES, ids=[c[0] for c in CASES])
def test_printed_boundary_kernels_on_perturbed_state(name, fixture, over):
    """one boundary-condition pass of the printed kernels ('generic' faces, SplitBC parts) on a perturbed state equals the
    oracle's hand-written boundary functions under the fixture's hand-written plan"""
    import ctypes
    from opensbli_b200 import run as R
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides=over)
    want, _ = load_fixture(fixture)
    rng = np.random.default_rng(11)
    q0 = [a.copy() for a in R.initial_state(plan_sym, cold)]
    nd = plan['ndim']
    for m, a in enumerate(q0):
        a *= 1.0 + 0.05 * rng.standard_normal(a.shape)
        if 1 <= m <= nd:
            a += 0.05 * rng.standard_normal(a.shape)
    cfg = ou.make_cfg(want)
    P = ctypes.POINTER(ctypes.c_double)
    qo = [a.copy() for a in q0]
    ou.oracle_lib().osbo_apply_bcs(ctypes.byref(cfg), (P * len(qo))(*[a.ctypes.data_as(P) for a in qo]))
    # the printed kernels on the host; faces that kept a hand-written kernel are applied by the oracle under the runner's plan
    hk = hostsim.HostKernels(plan['user_kernels'], q0[0].shape)
    for n, a in zip(plan_sym['q_names'], q0):
        hk.fields[n] = a
    for n, a in plan.get('user_fields', {}).items():
        hk.fields[n] = np.ascontiguousarray(a)
    generic = [(d, s) for d in range(nd) for s in range(2) if plan['bc'][d][s]['type'] == 'generic']
    assert generic
    for d in range(nd):
        for s in range(2):
            if (d, s) in generic:
                hk.run('bc_%d_%d' % (d, s))
            else:
                one = oracle_plan(plan)
                one['bc'] = [[dict(type='exchange'), dict(type='exchange')] for _ in range(nd)]
                one['bc'][d][s] = plan['bc'][d][s]
                c1 = ou.make_cfg(one)
                ou.oracle_lib().osbo_apply_bcs(ctypes.byref(c1), (P * len(q0))(*[a.ctypes.data_as(P) for a in q0]))
    for a, b in zip(q0, qo):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14), name


def run_generic_on_host(plan_sym, plan, cold, nsteps, state=None, fields=None):
    """the whole time loop of a GENERIC-path plan on the host: every loop of the step is a printed kernel, launched in program order.
    `state` / `fields`: start from given padded arrays instead of the runner's cold data (the reference's own, where numpy and the C
    library disagree on an ill-conditioned initial profile)"""
    from opensbli_b200 import run as R
    q = [np.ascontiguousarray(a).copy() for a in (state if state is not None else R.initial_state(plan_sym, cold))]
    hk = hostsim.HostKernels(plan['user_kernels'], q[0].shape)
    for n, a in zip(plan_sym['q_names'], q):
        hk.fields[n] = a
    for n, a in plan.get('user_fields', {}).items():
        hk.fields[n] = np.ascontiguousarray(a).copy()
    for n, a in (fields or {}).items():
        hk.fields[n] = np.ascontiguousarray(a).copy()
    for it in range(nsteps):
        hk.run('iteration_start', it)
        for s in range(plan['generic']['nstages']):
            hk.run('stage_%d' % s, it)
        hk.run('iteration_end', it)
    return hk


@pytest.mark.parametrize('name,golden,over,steps', [('sod_weno7', 'sod_weno7_n200', {'block0np0': 200}, (1, 50)),
                                                    ('sod_weno3', 'sod_weno3_n200', {'block0np0': 200}, (1, 50)),
                                                    ('tg_isot', 'tg_isot_17', {'block0np0': 17, 'block0np1': 17, 'block0np2': 17}, (3,))])
def test_generic_path_reproduces_the_reference(name, golden, over, steps):
    """Programs the hand-written kernels do not cover fall back to the generic path (backend.extract_plan): here WENO7-JS and
    WENO3-Z on the Sod tube.  Every loop of the time step -- boundary kernels, constituent relations, reconstruction, residual, RK
    update with rkA[stage] / rkB[stage] -- is a printed kernel; run on the host in program order against the reference's own run."""
    from opensbli_b200 import run as R
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides=over)
    assert plan['conv'] == 'generic' and plan['generic']['nstages'] == 3 and plan['generic']['reason']
    whens = [k['when'] for k in plan['user_kernels']]
    assert {'stage_0', 'stage_1', 'stage_2'} <= set(whens) and 'iteration_start' in whens      # the RK update is compiled per stage
    z = np.load(os.path.join(APPS, golden + '.npz'))
    for n in steps:
        hk = run_generic_on_host(plan_sym, plan, cold, n)
        ref = z['q%d' % n]
        inner_ = (slice(5, -5),) * plan['ndim']
        got = np.stack([hk.fields[f][inner_] for f in plan_sym['q_names'][:len(ref)]])
        ax = tuple(range(1, ref.ndim))
        err = np.abs(got - ref).max(axis=ax) / np.abs(ref).max(axis=ax)
        tol = 1e-9 if name == 'sod_weno3' else 1e-12          # WENO-Z: the reference's own build-to-build noise floor (common.TOL)
        assert err.max() < tol * max(1, n / 10), (n, err)


@pytest.mark.parametrize('name,fixture', [('tgv_teno5_allprinted', 'tgv_teno5_16'), ('tgv_central4_allprinted', 'tgv_central4_16')])
def test_bench_workloads_through_the_generic_path(name, fixture):
    """The two bench workloads (TENO5 + StoreSome + RK-LS, Central-4 + RK3 Taylor-Green) FORCED through the generic path
    (OSB_FORCE_GENERIC_PATH=1 when the plan was distilled): the printed loops reproduce the goldens the hand-written kernels are
    held to -- the two paths of the back end agree through the reference."""
    from opensbli_b200 import run as R
    want, states = load_fixture(fixture)
    over = {'block0np%d' % d: 16 for d in range(3)}
    over['dt'] = want['constants']['dt']
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides=over)
    assert plan['conv'] == 'generic' and 'forced' in plan['generic']['reason']
    hk = run_generic_on_host(plan_sym, plan, cold, 3)
    got = np.stack([hk.fields[f][5:-5, 5:-5, 5:-5] for f in plan_sym['q_names']])
    err = field_errors(want, got, states[3])
    assert max(err) < 1e-12, err


@pytest.mark.parametrize('name,fixture,over,nsteps', [('katzer_allprinted', 'katzer_60x40', {'block0np0': 60, 'block0np1': 40}, 10),
                                                      ('ewc_allprinted', 'ewc_wenoz5_32', {'block0np0': 32, 'block0np1': 32}, 10),
                                                      ('tcf_teno6_allprinted', 'tcf_teno6_16x24x12', {'block0np0': 16, 'block0np1': 24, 'block0np2': 12}, 5),
                                                      ('trans_allprinted', 'trans_40x30x8', {'block0np0': 40, 'block0np1': 30, 'block0np2': 8}, 5),
                                                      ('vst_allprinted', 'vst_60x30', {'block0np0': 60, 'block0np1': 30}, 10)])
def test_general_path_apps_through_the_generic_path(name, fixture, over, nsteps):
    """Three apps of the general path FORCED through the generic path: Katzer (stretched grid, ReducedAccess closures selected by
    grid index, adaptive TENO with the Ducros sensor, isothermal wall / inflow / outflow / tabulated Dirichlet kernels), the 3-D
    channel (TENO6, Carpenter closures, power-law viscosity, body force) and the fully curvilinear Euler wave (WENO-Z, metric
    cosines) -- the printed loops reproduce the goldens of the reference's own generated C (the curvilinear case bit for bit).
    This is what stands behind 'programs outside the hand-written kernels run on the generic path' for 3-D / viscous curvilinear
    set-ups, for which no reference configuration exists to pin them on directly."""
    from opensbli_b200 import run as R
    want, states = load_fixture(fixture)
    plan_sym, env, plan, cold = R.load_case(os.path.join(PLANS, name), overrides=over)
    assert plan['conv'] == 'generic'
    state = want.get('q0_padded')            # general-path fixtures carry the reference's own cold data (initial state incl. halos, metrics)
    fields = want.get('fields')
    hk = run_generic_on_host(plan_sym, plan, cold, nsteps, state=state, fields=fields)
    nd = plan['ndim']
    got = np.stack([hk.fields[f][(slice(5, -5),) * nd] for f in plan_sym['q_names']])
    err = field_errors(want, got, states[nsteps])
    assert max(err) < 1e-12, err
