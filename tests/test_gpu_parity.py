"""GPU tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden vectors."""
import copy
import math

import numpy as np
import pytest

from common import fixtures, load_fixture, pad, inner, field_errors, tol_for, initial_padded
import oracle_util as ou

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def osb():
    import opensbli_b200
    return opensbli_b200


def describe(err):
    return ['%.2e' % e for e in err]


@pytest.mark.parametrize('name', fixtures())
def test_steps_match_golden(osb, name):
    """n reference time steps from the reference's own initial state (the golden vectors)."""
    plan, states = load_fixture(name)
    for n in sorted(k for k in states if k > 0):
        with osb.Simulation(plan) as sim:
            sim.set_state(initial_padded(plan, states))
            sim.step(n)
            q = sim.get_state()
        err = field_errors(plan, inner(plan, q), states[n])
        print(name, n, describe(err))
        tol = 1e-10 if n >= 1000 else tol_for(plan, n)
        assert max(err) < tol, (name, n, err)


@pytest.mark.parametrize('name', fixtures())
def test_residual_matches_oracle(osb, name):
    """constituent relations + all spatial kernels of one stage (Residual arrays) vs the oracle."""
    plan, states = load_fixture(name)
    q0 = initial_padded(plan, states)
    cfg = ou.make_cfg(plan)
    import ctypes
    P = ctypes.POINTER(ctypes.c_double)
    qo = [a.copy() for a in q0]
    ou.oracle_lib().osbo_apply_bcs(ctypes.byref(cfg), (P * len(qo))(*[a.ctypes.data_as(P) for a in qo]))
    Ro = ou.oracle_residual(plan, qo)
    with osb.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.apply_bcs()
        qb = sim.get_state()
        Rg = sim.residual()
    for a, b in zip(qb, qo):   # boundary conditions: copies are bit-exact, wall/inflow formulas to round-off
        if 'q0_padded' in plan:
            assert np.allclose(a, b, rtol=1e-14, atol=1e-300)
        else:
            assert np.array_equal(a, b)
    Ro_i, Rg_i = inner(plan, Ro), inner(plan, Rg)
    scale = np.abs(Ro_i).max(axis=tuple(range(1, Ro_i.ndim)))
    scale = np.maximum(scale, scale.max() * 1e-6)
    err = [float(np.abs(Rg_i[m] - Ro_i[m]).max() / scale[m]) for m in range(len(scale))]
    print(name, 'residual', describe(err))
    lim = 1e-7 if plan.get('weno_formulation') == 'Z' and plan['conv'] == 'weno' else 1e-11
    assert max(err) < lim, err


@pytest.mark.parametrize('name', [n for n in fixtures() if n.startswith(('katzer', 'vst', 'trans', 'tcf', 'lam2d', 'tgv_sym', 'sod_zgo', 'sod_pout', 'isr'))])
def test_boundary_conditions_on_perturbed_state(osb, name):
    """Wall / inflow / outflow / symmetry / (partial) Dirichlet kernels on a randomly perturbed state -- every branch of the
    formulas carries signal (e.g. the free spanwise momentum at the transitional-SBLI top boundary) -- vs the oracle."""
    import ctypes
    plan, states = load_fixture(name)
    rng = np.random.default_rng(5)
    q0 = initial_padded(plan, states)
    nd = plan['ndim']
    for m, a in enumerate(q0):
        a *= 1.0 + 0.05 * rng.standard_normal(a.shape)
        if 1 <= m <= nd:
            a += 0.05 * rng.standard_normal(a.shape)
    cfg = ou.make_cfg(plan)
    P = ctypes.POINTER(ctypes.c_double)
    qo = [a.copy() for a in q0]
    ou.oracle_lib().osbo_apply_bcs(ctypes.byref(cfg), (P * len(qo))(*[a.ctypes.data_as(P) for a in qo]))
    with osb.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.apply_bcs()
        qb = sim.get_state()
    changed = sum(int(np.count_nonzero(b != a)) for a, b in zip(q0, qo))
    assert changed > 0
    for a, b in zip(qb, qo):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-15), name


@pytest.mark.parametrize('name', ['tgv_teno5_16', 'tgv_central4_16'])
def test_primitive_fields_follow_the_state(osb, name):
    """u_i, p, T handed out after a run are the constituent relations of the CURRENT state, also on the paths whose stage
    kernels derive them on the fly and never write the arrays."""
    plan, states = load_fixture(name)
    g, M = plan['constants']['gama'], plan['constants']['Minf']
    with osb.Simulation(plan) as sim:
        sim.set_state(initial_padded(plan, states))
        sim.step(3)
        q = inner(plan, sim.get_state())
        u0, p, T = (inner(plan, [sim.download(n)])[0] for n in ('u0', 'p', 'T'))
    pw = (g - 1.0) * (q[4] - 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2) / q[0])
    assert np.allclose(u0, q[1] / q[0], rtol=1e-14, atol=1e-16)
    assert np.allclose(p, pw, rtol=1e-12) and np.allclose(T, g * M * M * pw / q[0], rtol=1e-12)


CASES = [
    # nd, N, conv, order, formulation, averaging, viscous, rk
    (1, 64, 'teno', 6, 'JS', 'roe', False, 'ls'),
    (1, 64, 'weno', 5, 'JS', 'simple', False, 'sbli'),
    (2, 24, 'teno', 5, 'JS', 'roe', True, 'ls'),
    (2, 24, 'weno', 5, 'JS', 'simple', False, 'ls'),
    (2, 24, 'central', 4, 'JS', 'roe', True, 'sbli'),
    (2, 23, 'teno', 6, 'JS', 'simple', True, 'ls'),
    (3, 13, 'teno', 6, 'JS', 'roe', True, 'ls'),
    (3, 14, 'weno', 5, 'JS', 'roe', False, 'sbli'),
    (3, 12, 'teno', 5, 'JS', 'simple', True, 'sbli'),
    # smallest grids the halo logic admits (np >= 6 per direction: periodic copies of depth 3/4 must not overlap)
    (1, 6, 'teno', 5, 'JS', 'roe', False, 'ls'),
    (2, 6, 'weno', 5, 'Z', 'roe', True, 'ls'),
    (3, 6, 'teno', 5, 'JS', 'roe', True, 'ls'),
    (3, 6, 'central', 4, 'JS', 'roe', True, 'sbli'),
    # x extent spanning several 128-point staging blocks and 32-wide tiles with a ragged tail
    (2, 300, 'teno', 5, 'JS', 'roe', True, 'ls'),
]


def synthetic_plan(nd, N, conv, order, form, avg, visc, rk):
    per = [[dict(type='periodic'), dict(type='periodic')] for _ in range(nd)]
    coeff = (dict(rk='ls', rk_a=[0.0, -5.0 / 9.0, -153.0 / 128.0], rk_b=[1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0]) if rk == 'ls' else
             dict(rk='sbli', rk_a=[0.25, 3.0 / 20.0, 0.6], rk_b=[2.0 / 3.0, 5.0 / 12.0, 0.6]))
    np_ = [N, N + 3, N + 1][:nd] if N > 6 and N < 100 else ([N, N, N][:nd] if N <= 6 else [N, 37, 1][:nd])
    return dict(ndim=nd, np=np_, delta=[2 * math.pi / n for n in np_], conv=conv, order=order, weno_formulation=form,
                averaging=avg, viscous=visc, constants=dict(gama=1.4, Minf=0.5, Re=200.0, Pr=0.71, dt=2e-3, eps=1e-16, TENO_CT=1e-5),
                bc=per, **coeff)


def synthetic_state(plan, seed=0):
    """smooth periodic state with O(1) variations in every variable and direction (no symmetry)."""
    nd = plan['ndim']
    ax = [np.arange(n) * dl for n, dl in zip(plan['np'], plan['delta'])]
    X = np.meshgrid(*reversed(ax), indexing='ij')[::-1]     # X[d] has numpy shape (k, j, i)
    ph = [0.3, 1.1, 2.0]
    s = sum(np.sin(X[d] + ph[d]) for d in range(nd))
    c = sum(np.cos(2 * X[d] - ph[d]) for d in range(nd))
    rho = 1.0 + 0.2 * s / nd
    u = [0.4 * np.sin(X[d] + ph[(d + 1) % 3]) * (1 + 0.3 * c / nd) + 0.1 * (d + 1) for d in range(nd)]
    p = 1.0 / (1.4 * 0.25) * (1.0 + 0.15 * c / nd)
    E = p / 0.4 + 0.5 * rho * sum(v * v for v in u)
    return np.stack([rho] + [rho * v for v in u] + [E])


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%dd-%s%d-%s-%s-%s' % (c[0], c[2], c[3], c[5], 'ns' if c[6] else 'euler', c[7]))
def test_scheme_matrix_vs_oracle(osb, case):
    """Every scheme / averaging / RK / dimensionality combination, non-cubic grids (x-size not a multiple of 32),
    two steps from a non-symmetric state: CUDA vs oracle."""
    plan = synthetic_plan(*case)
    q0 = pad(plan, synthetic_state(plan))
    qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 2)
    with osb.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.step(2)
        qg = sim.get_state()
    err = field_errors(plan, inner(plan, qg), inner(plan, qo))
    print(case, describe(err))
    assert max(err) < tol_for(plan, 2), err


def test_sod_discontinuity_weno_teno(osb):
    """Dirichlet boundaries + shock/contact/rarefaction: 200 steps, all four reconstructions, vs oracle."""
    base, states = load_fixture('sod_teno5_n200')
    for conv, order, form in (('teno', 5, 'JS'), ('teno', 6, 'JS'), ('weno', 5, 'JS')):
        plan = copy.deepcopy(base)
        plan.update(conv=conv, order=order, weno_formulation=form)
        q0 = pad(plan, states[0])
        qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], 200)
        with osb.Simulation(plan) as sim:
            sim.set_state(q0)
            sim.step(200)
            qg = sim.get_state()
        err = field_errors(plan, inner(plan, qg), inner(plan, qo))
        print(conv, order, describe(err))
        assert max(err) < 1e-11, (conv, order, err)


def test_e2e_host_call_equals_resident_path(osb):
    plan, states = load_fixture('tgv_teno5_16')
    q0 = pad(plan, states[0])
    with osb.Simulation(plan) as sim:
        sim.set_state(q0)
        sim.step(2)
        qa = sim.get_state()
    with osb.Simulation(plan) as sim:
        out = [np.empty_like(a) for a in q0]
        ms = sim.advance_host(q0, out, 2)
        assert ms > 0
        n = sim.launch_count()
        assert n > 0
    for a, b in zip(qa, out):
        assert np.array_equal(a, b)


def test_periodic_conservation_full_size(osb):
    """Size-independent property at a production-like size: on a periodic box the flux-difference form
    conserves the integrals of rho, rhou_i, rhoE to round-off (telescoping sum); TGV 128^3, TENO5, 3 steps."""
    base, _ = load_fixture('tgv_teno5_16')
    N = 128
    plan = copy.deepcopy(base)
    plan['np'] = [N] * 3
    plan['delta'] = [2 * math.pi / N] * 3
    plan['constants']['dt'] = 0.003385 * 64 / N
    x = np.arange(N) * (2 * math.pi / N)
    Z, Y, X = np.meshgrid(x, x, x, indexing='ij')
    u0, u1 = np.sin(X) * np.cos(Y) * np.cos(Z), -np.cos(X) * np.sin(Y) * np.cos(Z)
    p = 1.0 / (1.4 * 0.01) + (1.0 / 16.0) * (np.cos(2 * X) + np.cos(2 * Y)) * (2.0 + np.cos(2 * Z))
    r = 1.4 * 0.01 * p
    q0 = np.stack([r, r * u0, r * u1, 0 * r, p / 0.4 + 0.5 * r * (u0 ** 2 + u1 ** 2)])
    with osb.Simulation(plan) as sim:
        sim.set_state(pad(plan, q0))
        sim.step(3)
        q = inner(plan, sim.get_state())
    assert np.all(np.isfinite(q))
    # mass: exact telescoping sum; energy: the viscous work terms are in non-conservative product form
    for m, lim in ((0, 1e-12), (4, 1e-9)):
        s0, s1 = q0[m].sum(), q[m].sum()
        assert abs(s1 - s0) < lim * abs(s0), (m, s0, s1)
    for m in (1, 2, 3):   # zero net momentum stays zero (relative to |rho u| scale)
        assert abs(q[m].sum()) < 1e-9 * np.abs(q0[1]).sum()
    # TGV symmetry u2(x,y,-z) = -u2(x,y,z) is preserved by the discrete operators
    assert np.abs(q[3][1:, :, :] + q[3][:0:-1, :, :]).max() < 1e-12


def test_cuda_graph_replay_equals_direct_launches(osb):
    """Small grids replay one captured time step as a CUDA graph; the result must be bit-identical to launching
    the kernels directly (step(1) never uses the graph), also after a constant has changed."""
    plan, states = load_fixture('sod_teno5_n200')
    q0 = pad(plan, states[0])
    with osb.Simulation(plan) as a, osb.Simulation(plan) as b:
        a.set_state(q0)
        b.set_state(q0)
        a.step(40)
        for _ in range(40):
            b.step(1)
        for x, y in zip(a.get_state(), b.get_state()):
            assert np.array_equal(x, y)
        a.set_const('dt', 1.0e-4)
        b.set_const('dt', 1.0e-4)
        a.step(10)
        for _ in range(10):
            b.step(1)
        for x, y in zip(a.get_state(), b.get_state()):
            assert np.array_equal(x, y)


@pytest.mark.parametrize('name', ['tgv_teno5_16', 'tgv_central4_16'])
def test_cuda_graph_replay_with_buffer_role_exchange(osb, name):
    """3-D paths whose stage kernels write out of place exchange the roles of the q and Residual buffers every stage: the
    captured unit is two steps (three stages each), replayed only from the parity it was captured at.  Odd and even step
    counts, repeated calls and field access in between must all equal step-by-step direct launches bit for bit."""
    plan, states = load_fixture(name)
    q0 = initial_padded(plan, states)
    with osb.Simulation(plan) as a, osb.Simulation(plan) as b:
        a.set_state(q0)
        b.set_state(q0)
        done = 0
        for n in (5, 4, 1, 3, 2):
            a.step(n)
            for _ in range(n):
                b.step(1)
            done += n
            for x, y in zip(a.get_state(), b.get_state()):
                assert np.array_equal(x, y), (name, done)
            ra, rb = a.residual(), b.residual()          # the parity entry point works on whichever buffers hold the roles now
            for x, y in zip(ra, rb):
                assert np.array_equal(x, y), (name, done, 'residual')


@pytest.mark.parametrize('name', ['tgv_teno5_16', 'tgv_central4_16'])
def test_graph_replay_refreshes_primitives_and_iteration(osb, name):
    """A replayed CUDA graph re-applies the host-side effects of the captured steps: primitives handed out after a replay are
    those of the current state (the fused stage kernels never write u/p/T), and the iteration counter advances."""
    plan, states = load_fixture(name)
    q0 = initial_padded(plan, states)
    with osb.Simulation(plan) as a, osb.Simulation(plan) as b:
        a.set_state(q0)
        b.set_state(q0)
        for n in (4, 4, 6):
            a.step(n)                       # graph replays (two-step units)
            for _ in range(n):
                b.step(1)                   # direct launches
            # (the reference's periodic exchange fills 3 planes on the high side while the constituent relations run over 4:
            # p is 0/0 in the outermost plane of the TENO halo, in both runs alike)
            pa, pb = a.download('p'), b.download('p')
            assert np.array_equal(pa, pb, equal_nan=True), name
            assert np.isfinite(inner(plan, [pa])).all()
            assert a.get_iteration() == b.get_iteration()
        assert a.get_iteration() == 14


def test_user_kernel_reads_current_primitives_and_rejects_unknown_inputs(osb):
    """Point-wise user kernels on the fused 3-D paths: a kernel reading a primitive sees the constituent relations of the
    current state (the stage kernels never write those arrays); a kernel reading a dataset nobody provided is refused."""
    plan, states = load_fixture('tgv_teno5_16')
    src = ('struct UserFields { double *p[96]; };\n'
           'extern "C" __global__ void k(long long off, int n0, int n1, int n2, int lo0, int lo1, int lo2, long long s1, long long s2, UserFields f) {\n'
           '  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, kk = blockIdx.z;\n'
           '  if (i >= n0 || j >= n1 || kk >= n2) return;\n'
           '  const long long X = off + (lo0 + i) + (lo1 + j) * s1 + (lo2 + kk) * s2;\n'
           '  f.p[1][X] = f.p[1][X] + f.p[0][X];\n}\n')
    with osb.Simulation(plan) as sim:
        sim.add_user_kernel(src, 'k', ['p', 'psum'], [0, 16, 0, 16, 0, 16], 'iteration_end', writes=['psum'])
        sim.set_state(initial_padded(plan, states))
        acc = np.zeros((16, 16, 16))
        for _ in range(3):
            sim.step(1)
            q = inner(plan, sim.get_state())
            acc += 0.4 * (q[4] - 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2) / q[0])
        got = sim.download('psum')[5:-5, 5:-5, 5:-5]
        assert np.allclose(got, acc, rtol=1e-13, atol=0), float(np.abs(got - acc).max())
        with pytest.raises(osb.BackendError, match='does not exist'):
            sim.add_user_kernel(src, 'k', ['x0', 'psum2'], [0, 16, 0, 16, 0, 16], 'iteration_end', writes=['psum2'])
