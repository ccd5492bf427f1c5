"""CPU tests (-m "not gpu"): the oracle (oracle/osbli_oracle.c) against the golden vectors minted from the
reference's own generated C, and -- when oracle/_ref is present -- against the reference executables."""
import numpy as np
import pytest

from common import fixtures, load_fixture, pad, inner, field_errors, tol_for, initial_padded
import oracle_util as ou


@pytest.mark.parametrize('name', fixtures())
def test_oracle_matches_golden(name):
    plan, states = load_fixture(name)
    for n in sorted(k for k in states if k > 0):
        if n > 200:
            continue
        q, _ = ou.oracle_advance(plan, initial_padded(plan, states), n)
        err = field_errors(plan, inner(plan, q), states[n])
        assert max(err) < tol_for(plan, n), (name, n, err)


def test_oracle_sod_full_run_final_norms():
    """Config 1 (Sod, WENO-JS5, N=800, 1000 steps): final-time fields and L2 norms within 1e-10."""
    plan, states = load_fixture('sod_wenojs5_n800')
    q, _ = ou.oracle_advance(plan, pad(plan, states[0]), 1000)
    qi = inner(plan, q)
    err = field_errors(plan, qi, states[1000])
    assert max(err) < 1e-10, err
    for m in range(3):
        l2, l2r = np.sqrt(np.mean(qi[m] ** 2)), np.sqrt(np.mean(states[1000][m] ** 2))
        assert abs(l2 - l2r) <= 1e-10 * max(l2r, 1e-300)


def test_oracle_sod_matches_exact_solution():
    """Known-answer anchor: L1(rho) of TENO5 N=200 against the exact Riemann solution at t=0.2 is the
    reference's own 2.51e-3 (apps/Sod_shock_tube/reference.txt)."""
    plan, states = load_fixture('sod_teno5_n200')
    q, _ = ou.oracle_advance(plan, pad(plan, states[0]), 1000)
    rho = inner(plan, q)[0]
    x = np.arange(200) / 199.0
    exact = sod_exact_density(x, 0.2)
    l1 = np.mean(np.abs(rho - exact))
    assert abs(l1 - 2.51e-3) < 1e-4, l1


def sod_exact_density(x, t, g=1.4):
    """Exact Sod solution (rho) for (1,0,1)|(0.125,0,0.1), diaphragm at 0.5."""
    from scipy.optimize import brentq
    rl, pl, rr, pr = 1.0, 1.0, 0.125, 0.1
    al, ar = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p):
        fl = 2 * al / (g - 1) * ((p / pl) ** ((g - 1) / (2 * g)) - 1)
        A, B = 2 / ((g + 1) * rr), (g - 1) / (g + 1) * pr
        fr = (p - pr) * np.sqrt(A / (p + B))
        return fl + fr
    ps = brentq(f, 1e-6, 1.0)
    us = 0.5 * (0 + 0) + 0.5 * ((ps - pr) * np.sqrt((2 / ((g + 1) * rr)) / (ps + (g - 1) / (g + 1) * pr))
                               - 2 * al / (g - 1) * ((ps / pl) ** ((g - 1) / (2 * g)) - 1))
    rsl = rl * (ps / pl) ** (1 / g)
    rsr = rr * ((ps / pr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * ps / pr + 1))
    asl = al * (ps / pl) ** ((g - 1) / (2 * g))
    S = ar * np.sqrt((g + 1) / (2 * g) * ps / pr + (g - 1) / (2 * g))
    xi = (x - 0.5) / t
    rho = np.where(xi < -al, rl, 0.0)
    fan = (xi >= -al) & (xi < us - asl)
    rho = np.where(fan, rl * (2 / (g + 1) + (g - 1) / ((g + 1) * al) * (0 - xi)) ** (2 / (g - 1)), rho)
    rho = np.where((xi >= us - asl) & (xi < us), rsl, rho)
    rho = np.where((xi >= us) & (xi < S), rsr, rho)
    rho = np.where(xi >= S, rr, rho)
    return rho


REF_CASES = [('sod_teno5', 'sod_teno5_n200'), ('sod_wenojs5', 'sod_wenojs5_n800'), ('sod_wenoz5', 'sod_wenoz5_n200'),
             ('tgv_central4', 'tgv_central4_16'), ('tgv_teno5', 'tgv_teno5_16')]


@pytest.mark.parametrize('config,fixture', REF_CASES)
def test_oracle_matches_reference_executable(config, fixture):
    """Pin the oracle against the reference itself run here (other grid size than the fixtures)."""
    if not ou.have_ref(config):
        pytest.skip('oracle/_ref/%s not built (python oracle/gen_ref.py)' % config)
    plan, _ = load_fixture(fixture)
    import copy
    plan = copy.deepcopy(plan)
    nd = plan['ndim']
    N = 20 if nd == 3 else 150
    plan['np'] = [N] * nd
    if nd == 3:
        plan['delta'] = [2 * np.pi / N] * 3
        plan['constants']['dt'] = 0.003385 * 64 / N
    else:
        plan['delta'] = [1.0 / (N - 1)]
    env = {'dt': plan['constants']['dt']}
    for d in range(nd):
        env['block0np%d' % d] = N
    fields = ['rho'] + ['rhou%d' % d for d in range(nd)] + ['rhoE']
    r0 = ou.run_ref(config, dict(env, niter=0), fields)
    r2 = ou.run_ref(config, dict(env, niter=2), fields)
    q, _ = ou.oracle_advance(plan, [r0[f].copy() for f in fields], 2)
    err = field_errors(plan, inner(plan, q), inner(plan, [r2[f] for f in fields]))
    assert max(err) < tol_for(plan, 2), err


def test_reference_noise_floor_weno_z():
    """Why WENO-Z parity is asserted at 1e-9 and not at 1e-12 (tests/common.py): the reference's OWN generated C, built twice
    (ref_seq: g++ -O2 -ffp-contract=off; ref_omp: g++ -O3 -march=x86-64-v3, FMA contraction on), disagrees with itself by
    ~5e-11 after ONE step of the Sod WENO-Z case, because with eps = 1e-14 its Horner-form smoothness indicators lose all
    digits where beta <~ 1e-12.  No independent implementation can sit closer to "the reference" than its two builds sit to
    each other.  The same two builds of the WENO-JS and TENO cases agree to round-off."""
    if not (ou.have_ref('sod_wenoz5') and ou.have_ref('sod_wenojs5')):
        pytest.skip('oracle/_ref not built (python oracle/gen_ref.py)')
    names = ['rho', 'rhou0', 'rhoE']

    def two_builds(config, n):
        a = ou.run_ref(config, dict(block0np0=n, niter=1), names, exe='ref_seq')
        b = ou.run_ref(config, dict(block0np0=n, niter=1), names, exe='ref_omp', threads=1)
        return max(float(np.abs(a[f] - b[f]).max() / np.abs(a[f]).max()) for f in names)
    z, js = two_builds('sod_wenoz5', 200), two_builds('sod_wenojs5', 800)
    print('two builds of the reference, one step: WENO-Z %.2e  WENO-JS %.2e' % (z, js))
    assert 1e-12 < z < 1e-9, z          # the noise floor itself: above the north-star 1e-12, below the asserted 1e-9
    assert js < 1e-13, js
