"""GPU tests (-m gpu): in-loop diagnostics -- device reductions (kinetic energy, enstrophy, mass, total energy, max Mach
number) against the same sums formed with numpy from the downloaded state, and the NaN check (the reference's ops_NaNcheck,
core/diagnostics/simulation_monitors.py:112-113, utilities/helperfunctions.py:172-190)."""
import numpy as np
import pytest

from common import load_fixture, initial_padded, inner

pytestmark = pytest.mark.gpu


def numpy_diagnostics(plan, q):
    """volume sums over the interior from the padded state q (halos valid), 4th-order central vorticity"""
    nd, h = plan['ndim'], 5
    inv = [1.0 / d for d in plan['delta']]
    u = [q[1 + d] / q[0] for d in range(nd)]
    s = (slice(h, -h),) * nd

    def d1(a, b):      # d a / d x_b on the interior; numpy axis of direction b is nd-1-b
        ax = nd - 1 - b
        def sh(o):
            sl = [slice(h, -h)] * nd
            sl[ax] = slice(h + o, a.shape[ax] - h + o)
            return a[tuple(sl)]
        return (1.0 / 12.0) * inv[b] * ((sh(-2) - sh(2)) + 8.0 * (sh(1) - sh(-1)))
    rho = q[0][s]
    u2 = sum(v[s] ** 2 for v in u)
    if nd == 3:
        w2 = (d1(u[2], 1) - d1(u[1], 2)) ** 2 + (d1(u[0], 2) - d1(u[2], 0)) ** 2 + (d1(u[1], 0) - d1(u[0], 1)) ** 2
    elif nd == 2:
        w2 = (d1(u[1], 0) - d1(u[0], 1)) ** 2
    else:
        w2 = 0.0 * rho
    E = q[nd + 1][s]
    p = (plan['constants']['gama'] - 1.0) * (E - 0.5 * rho * u2)
    return {'sum_rho': rho.sum(), 'sum_ke': (0.5 * rho * u2).sum(), 'sum_enstrophy': (0.5 * rho * w2).sum(), 'sum_rhoE': E.sum(),
            'max_mach': np.sqrt(u2 * rho / (plan['constants']['gama'] * p)).max(), 'nonfinite': 0.0}


@pytest.mark.parametrize('name', ['tgv_teno5_16', 'tgv_central4_16', 'katzer_60x40'])
def test_device_reductions_match_numpy(name):
    import opensbli_b200
    plan, states = load_fixture(name)
    with opensbli_b200.Simulation(plan) as sim:
        sim.set_state(initial_padded(plan, states))
        sim.step(3)
        got = sim.diagnostics()
        q = sim.get_state()
        assert sim.nan_check('rho') == 0 and sim.nan_check('p') == 0
    want = numpy_diagnostics(plan, q)
    if plan.get('metric_fields'):       # stretched grid: the vorticity carries the metric factors; compare the other sums
        want.pop('sum_enstrophy')
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-12 * max(abs(v), 1.0), (k, got[k], v)


def test_nan_check_counts_non_finite_values():
    import opensbli_b200
    plan, states = load_fixture('tgv_teno5_16')
    q0 = initial_padded(plan, states)
    q0[0][7, 8, 9] = np.nan
    q0[0][10, 8, 9] = np.inf
    q0[0][0, 0, 0] = np.nan            # in the halo: not counted (interior points only)
    with opensbli_b200.Simulation(plan) as sim:
        sim.set_state(q0)
        assert sim.nan_check('rho') == 2
        assert sim.nan_check('rhoE') == 0
        assert sim.diagnostics()['nonfinite'] == 2
