#!/usr/bin/env python
"""Mint the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Each fixture is produced by running oracle/_ref/<config>/ref_seq -- the reference's own generated C
(OpenSBLI front end at /root/reference, OPSC back end, PYTHONHASHSEED=0, SymPy version recorded in
oracle/_ref/<config>/provenance.json) compiled with `g++ -O2 -ffp-contract=off` against the sequential OPS
stand-in oracle/ops_seq.h.  Stored per fixture (npz): the resolved plan (json), the conserved fields after the
reference's own initialisation kernel (niter=0) and after n reference time steps, interior points only.

    python oracle/gen_ref.py            # needs /root/reference
    python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import run_ref, REF_DIR  # noqa: E402

LS3 = dict(rk='ls', rk_a=[0.0, -5.0 / 9.0, -153.0 / 128.0], rk_b=[1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0])
SBLI3 = dict(rk='sbli', rk_a=[1.0 / 4.0, 3.0 / 20.0, 3.0 / 5.0], rk_b=[2.0 / 3.0, 5.0 / 12.0, 3.0 / 5.0])


def sod_plan(N, conv, order, form='JS'):
    g = 1.4
    ql = [1.0, 0.0, 1.0 / (g - 1.0)]
    qr = [0.125, 0.0, 0.1 / (g - 1.0)]
    return dict(ndim=1, np=[N], delta=[1.0 / (N - 1)], conv=conv, order=order, weno_formulation=form, averaging='roe',
                viscous=False, constants=dict(gama=g, dt=0.0002, eps=1e-16, TENO_CT=1e-5),
                bc=[[dict(type='dirichlet', q=ql), dict(type='dirichlet', q=qr)]], **LS3)


def tgv_plan(N, conv, order, rk):
    per = [[dict(type='periodic'), dict(type='periodic')] for _ in range(3)]
    return dict(ndim=3, np=[N, N, N], delta=[2 * math.pi / N] * 3, conv=conv, order=order, averaging='roe', viscous=True,
                constants=dict(gama=1.4, Minf=0.1, Re=1600.0, Pr=0.71, dt=0.003385 * 64 / N, eps=1e-16, TENO_CT=1e-6),
                bc=per, **rk)


# fixture name -> (reference config, plan, [step counts])
FIXTURES = {
    'sod_teno5_n200': ('sod_teno5', sod_plan(200, 'teno', 5), [1, 50]),
    'sod_wenojs5_n800': ('sod_wenojs5', sod_plan(800, 'weno', 5, 'JS'), [1, 100, 1000]),
    'sod_wenoz5_n200': ('sod_wenoz5', sod_plan(200, 'weno', 5, 'Z'), [1, 50]),
    'tgv_central4_16': ('tgv_central4', tgv_plan(16, 'central', 4, SBLI3), [1, 3]),
    'tgv_teno5_16': ('tgv_teno5', tgv_plan(16, 'teno', 5, LS3), [1, 3]),
}


def ssp_rk3():
    """RungeKuttaLS(3, formulation='SSP') coefficients, rk_LS.py:74-89."""
    from math import sqrt
    c = 0.924574
    z1 = float(sqrt(36*c**4 + 36*c**3 - 135*c**2 + 84*c - 12))
    z2 = float(2*c**2 + c - 2)
    z3 = float(12*c**4 - 18*c**3 + 18*c**2 - 11*c + 2)
    z4 = float(36*c**4 - 36*c**3 + 13*c**2 - 8*c + 4)
    z5 = float(69*c**3 - 62*c**2 + 28*c - 8)
    z6 = float(34*c**4 - 46*c**3 + 34*c**2 - 13*c + 2)
    B = [0.924574, (12*c*(c-1)*(3*z2-z1) - (3*z2-z1)**2)/(144*c*(3*c-2)*(c-1)**2),
         (-24*(3*c-2)*(c-1)**2)/((3*z2-z1)**2 - 12*c*(c-1)*(3*z2-z1))]
    A = [0.0, (-z1*(6*c**2 - 4*c + 1) + 3*z3)/((2*c+1)*z1 - 3*(c+2)*(2*c-1)**2),
         (-z1*z4 + 108*(2*c-1)*c**5 - 3*(2*c-1)*z5)/(24*z1*c*(c-1)**4 + 72*c*z6 + 72*c**6 * (2*c-13))]
    return A, B


def katzer_plan(N0, N1):
    """apps/katzer_SBLI/katzer_SBLI.py as shipped, on an N0 x N1 grid (metric fields and the Dirichlet table are
    filled in from the reference's own cold kernels by main())."""
    A, B = ssp_rk3()
    ra = 'reduced_access'
    return dict(ndim=2, np=[N0, N1], delta=[400.0 / (N0 - 1), 115.0 / (N1 - 1)], conv='teno', order=5, averaging='roe',
                viscous=True, rk='ls', rk_a=A, rk_b=B, viscosity=dict(type='sutherland'), teno_adaptive=True,
                metric_fields=[None, 'D11'],
                constants=dict(gama=1.4, Minf=2.0, Pr=0.72, Re=950.0, Twall=1.67619431, dt=0.04, SuthT=110.4, RefT=288.0,
                               eps=1e-15, TENO_CT=1e-5, teno_a1=10.5, teno_a2=4.5,
                               epsilon=1e-12),   # shock_sensors.py:26 fixes it; the app's '1.0e-30' is never substituted
                bc=[[dict(type='inlet_pressure_extrapolate', closure=ra), dict(type='extrapolation', order=0, closure=ra)],
                    [dict(type='isothermal_wall', closure=ra), dict(type='dirichlet_field', closure=ra)]])


def katzer_dirichlet_table(N0, halo=5):
    x0 = (400.0 / (N0 - 1)) * np.arange(-halo, N0 + halo)
    return np.stack([np.where(x0 > 40.0, 1.129734572, 1.00000596004), np.where(x0 > 40.0, 1.0921171, 1.00000268202),
                     np.where(x0 > 40.0, -0.058866065, 0.00565001630205), np.where(x0 > 40.0, 1.0590824, 0.94644428042)])


FIXTURES['katzer_60x40'] = ('katzer', katzer_plan(60, 40), [1, 10])


def sod_outlet_plan(N, kind):
    """Sod app with its right boundary replaced by ZeroGradientOutletBC / PressureOutletBC(back_pressure=0.1) (boundary classes
    no shipped app uses: zero_gradient_outlet.py:12-23, pressure_outlet.py:33-52)."""
    p = sod_plan(N, 'teno', 5)
    p['bc'][0][1] = dict(type=kind)
    if kind == 'pressure_outlet':
        p['constants']['back_pressure'] = 0.1
    return p


def isr_plan(N0, N1, wall):
    """apps/inviscid_shock_reflection/inviscid_shock.py: 2-D Euler, WENO5-Z with simple averaging, RungeKuttaLS(3); constant
    Dirichlet inflow, order-0 extrapolation outflow, shock-generator Dirichlet state along the top (tabulated: main() reads both
    imposed states off the reference's own dumps), bottom wall = SymmetryBC (shipped) or InviscidWallBC (inviscid_wall.py:24-52)."""
    bottom = dict(type=wall)
    if wall == 'split':      # SplitBC (bc_core.py:200-217) as set up by oracle/gen_ref.py CPP_EDITS: SymmetryBC over x-points [-3, 20), InviscidWallBC beyond
        bottom = dict(type='split', parts=[dict(type='symmetry', range=[-3, 20, 0, 1]), dict(type='inviscid_wall', range=[20, N0 + 4, 0, 1])])
    return dict(ndim=2, np=[N0, N1], delta=[350.0 / (N0 - 1), 115.0 / (N1 - 1)], conv='weno', order=5, weno_formulation='Z',
                averaging='simple', viscous=False, constants=dict(gama=1.4, Minf=2.0, dt=0.1),
                bc=[[dict(type='dirichlet', q=None), dict(type='extrapolation', order=0)], [bottom, dict(type='dirichlet_field')]], **LS3)


FIXTURES['sod_zgo_n200'] = ('sod_zgo', sod_outlet_plan(200, 'zero_gradient_outlet'), [1, 50])
FIXTURES['sod_pout_n200'] = ('sod_pout', sod_outlet_plan(200, 'pressure_outlet'), [1, 50])


def katzer_wenoz_plan(N0, N1):
    """BASELINE.json configs[3] as worded: the Katzer app with LLFWeno(5, formulation='Z') in place of adaptive TENO."""
    p = katzer_plan(N0, N1)
    p.update(conv='weno', order=5, weno_formulation='Z', teno_adaptive=False)
    for k in ('eps', 'TENO_CT', 'teno_a1', 'teno_a2', 'epsilon'):
        p['constants'].pop(k)
    return p


# the Katzer app with the SFD filter (filters/SFD.py) switched on: golden states of an APP run (its user kernels change the state, which
# the oracle does not model), kept under golden/apps/ so that the fixture-driven oracle tests do not pick it up
FIXTURES['apps/katzer_sfd_60x40'] = ('katzer_sfd', katzer_plan(60, 40), [10])
# WENO orders 7 and 3 (the hand-written sweeps cover 5): programs of the GENERIC path -- APP goldens, the oracle does not model them
FIXTURES['apps/sod_weno7_n200'] = ('sod_weno7', sod_plan(200, 'weno', 5, 'JS'), [1, 50])
FIXTURES['apps/sod_weno3_n200'] = ('sod_weno3', sod_plan(200, 'weno', 5, 'Z'), [1, 50])
# the central-4 TGV with the non-linear WENO filter (filters/WENO_filter.py) after every step: an APP run as well
FIXTURES['apps/tgv_wf_16'] = ('tgv_wf', tgv_plan(16, 'central', 4, SBLI3), [1, 3])
FIXTURES['katzer_wenoz_60x40'] = ('katzer_wenoz', katzer_wenoz_plan(60, 40), [1, 10])


def carpenter_tables():
    """First-derivative rows of the reference's Carpenter closure, taken from the scheme object itself
    (Carpenter_scheme.py:78-102, a 4x6 matrix al4^-1 ar4 evaluated by SymPy); second-derivative rows :69-76."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle'))
    import refshim
    refshim.install()
    from opensbli.core.boundary_conditions.Carpenter_scheme import Carpenter
    c = Carpenter()
    d1 = [[float(c.bc4_coefficients[i, j]) for j in range(6)] for i in range(4)]
    d2 = [[float(c.bc4_2_coefficients[i, j]) for j in range(5)] for i in range(2)]
    return {'d1': d1, 'd2': d2}


def katzer_carpenter_plan(N0, N1):
    p = katzer_plan(N0, N1)
    for pair in p['bc']:
        for b in pair:
            b['closure'] = 'carpenter'
    p['closures'] = {'carpenter': carpenter_tables()}
    return p


def tgv_sym_plan(N):
    """apps/taylor_green_vortex/TGsym/TGsym.py: 1/8-domain TGV, Central(4) + RungeKutta(3), SymmetryBC on all faces."""
    sym = [[dict(type='symmetry'), dict(type='symmetry')] for _ in range(3)]
    return dict(ndim=3, np=[N, N, N], delta=[math.pi / (N - 1)] * 3, conv='central', order=4, averaging='roe', viscous=True,
                constants=dict(gama=1.4, Minf=0.1, Re=800.0, Pr=0.71, dt=0.005), bc=sym, **SBLI3)


def tcf_teno6_plan(N0, N1, N2):
    """apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py (statistics off): TENO6, Carpenter closures at the
    isothermal walls, wall-normal stretching, T^0.7 viscosity, body force c0 = -1, SSP-RK3, periodic in x and z."""
    A, B = ssp_rk3()
    per = lambda: dict(type='periodic')
    wall = lambda: dict(type='isothermal_wall', closure='carpenter')
    return dict(ndim=3, np=[N0, N1, N2], delta=[4.0 * math.pi / N0, 2.0 / (N1 - 1), (4.0 * math.pi / 3.0) / N2], conv='teno', order=6,
                averaging='roe', viscous=True, rk='ls', rk_a=A, rk_b=B, viscosity=dict(type='power', exponent=0.7),
                metric_fields=[None, 'D11', None], forcing=True, closures={'carpenter': carpenter_tables()},
                constants=dict(gama=1.4, Minf=0.0955, Pr=0.7, Re=190.71, Twall=1.0, dt=0.0002, eps=1e-15, TENO_CT=1e-7,
                               c0=-1.0, c1=0.0, c2=0.0),
                bc=[[per(), per()], [wall(), wall()], [per(), per()]])


def lam2d_plan(N0, N1):
    """apps/channel_flow/laminar_2D/laminar_channel.py: Central(4) in the Skew() (Blaisdell) split with Carpenter closures at
    isothermal walls, uniform grid, Sutherland viscosity, body force c0 = -1, RungeKuttaLS(3)."""
    per = lambda: dict(type='periodic')
    wall = lambda: dict(type='isothermal_wall', closure='carpenter')
    return dict(ndim=2, np=[N0, N1], delta=[2.0 * math.pi / N0, 2.0 / (N1 - 1)], conv='central', order=4, averaging='roe', viscous=True,
                viscosity=dict(type='sutherland'), forcing=True, central_form='blaisdell', closures={'carpenter': carpenter_tables()},
                constants=dict(gama=1.4, Minf=0.1, Pr=0.72, Re=90.0, Twall=1.0, SuthT=110.4, RefT=273.0, dt=0.0002, c0=-1.0, c1=0.0),
                bc=[[per(), per()], [wall(), wall()]], **LS3)


def tcf_central_plan(N0, N1, N2):
    """apps/channel_flow/compressible_TCF_Central/turbulent_channel.py (statistics off): Central(4) in the Feiereisen
    quadratic split, Carpenter closures at the isothermal walls, wall-normal stretching, T^0.7 viscosity, body force."""
    p = tcf_teno6_plan(N0, N1, N2)
    p.update(conv='central', order=4, central_form='feiereisen', **LS3)
    p['constants'] = dict(gama=1.4, Minf=0.0955, Pr=0.7, Re=190.71, Twall=1.0, dt=0.0002, c0=-1.0, c1=0.0, c2=0.0)
    return p


def vst_plan(N0, N1):
    """apps/viscous_shock_tube/viscous_shock_tube.py: 2-D TENO5 + StoreSome viscous terms with a constant `mu` symbol,
    adiabatic walls on three faces and a symmetry face, ReducedAccess closures, RungeKuttaLS(3)."""
    wall = lambda: dict(type='adiabatic_wall', closure='reduced_access')
    return dict(ndim=2, np=[N0, N1], delta=[1.0 / (N0 - 1), 0.5 / (N1 - 1)], conv='teno', order=5, averaging='roe', viscous=True,
                constants=dict(gama=1.4, Minf=1.0, Pr=0.73, Re=200.0, mu=1.0, dt=0.000005, eps=1e-15, TENO_CT=1e-6),
                bc=[[wall(), wall()], [wall(), dict(type='symmetry')]], **LS3)     # SymmetryBC does not modify the derivatives


def trans_plan(N0, N1, N2):
    """apps/transitional_SBLI/transitional_SBLI.py (statistics off): Katzer's set-up in 3-D -- TENO6 adaptive, wall-normal
    stretching, Sutherland viscosity, ReducedAccess closures, inlet / outlet / isothermal wall / shock-generator Dirichlet
    top (spanwise momentum left free, its kinetic energy added to the imposed energy) / periodic span -- plus the
    time-periodic mass source A exp(-(x-xF)^2-(y-yF)^2) cos(bta z) sin(omega dt iter) in the continuity equation."""
    ra = 'reduced_access'
    return dict(ndim=3, np=[N0, N1, N2], delta=[375.0 / (N0 - 1), 140.0 / (N1 - 1), 27.32 / N2], conv='teno', order=6, averaging='roe',
                viscous=True, viscosity=dict(type='sutherland'), teno_adaptive=True, metric_fields=[None, 'D11', None],
                mass_source=dict(field='BF_amp', rate=0.025 * 0.1011),
                constants=dict(gama=1.4, Minf=1.5, Pr=0.72, Re=750.0, Twall=1.3809973268575328, dt=0.025, SuthT=110.4, RefT=202.17,
                               eps=1e-30, TENO_CT=1e-5, teno_a1=9.5, teno_a2=3.5, epsilon=1e-12),
                bc=[[dict(type='inlet_pressure_extrapolate', closure=ra), dict(type='extrapolation', order=0, closure=ra)],
                    [dict(type='isothermal_wall', closure=ra), dict(type='dirichlet_field', closure=ra, free=[3], ke_free=True)],
                    [dict(type='periodic'), dict(type='periodic')]], **LS3)


def trans_dirichlet_table(x0_padded):
    """transitional_SBLI.py:127-139: pre / post oblique-shock states switched at x0 = 20, printed with %.15f by the app."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle'))
    import refshim
    refshim.install()
    from opensbli.utilities.oblique_shock import ShockConditions
    g, M = 1.4, 1.5
    pre = (1.0, 1.0, 0.00466654053208844, (1.0 / (g * M ** 2)) / (g - 1.0) + 0.5 * (1.0 * 1.0 ** 2 + 0.00466654053208844 ** 2))
    post = ShockConditions(44.6607551, M, g).conservative_post_shock_conditions(1.0)
    r = lambda v: float('%.15f' % float(v))
    x = np.moveaxis(x0_padded, 1, 0)[5].reshape(-1)        # tangential plane (z, x) at an interior j, x fastest (x0 does not depend on j)
    rows = [np.where(x > 20.0, r(post[m]), r(pre[m])) for m in range(3)]
    rows.append(np.zeros_like(x))                           # rhou2 is not imposed
    rows.append(np.where(x > 20.0, r(post[3]), r(pre[3])))
    return np.stack(rows)


def ewc_plan(N, conv='weno'):
    """apps/euler_wave_curvilinear/euler_wave.py: 2-D Euler density wave on a wavy, fully curvilinear periodic grid in
    strong-conservation form (metric arrays D_ij, detJ from the reference's own metric kernels), WENO5-Z (shipped) or TENO5."""
    p = dict(ndim=2, np=[N, N], delta=[2.0 / N, 2.0 / N], conv=conv, order=5, averaging='roe', viscous=False, curvilinear=True,
             constants=dict(gama=1.4, dt=0.0005), bc=[[dict(type='periodic'), dict(type='periodic')] for _ in range(2)], **LS3)
    if conv == 'weno':
        p['weno_formulation'] = 'Z'
    else:
        p['constants'].update(eps=1e-15, TENO_CT=1e-6)
    return p


STATS = ['rhomean', 'E_mean', 'u0mean', 'u1u0mean', 'u2u2mean', 'p_mean', 'pp_mean', 'a_mean', 'T_mean', 'TT_mean', 'mu_mean', 'M_mean',
         'rhou1u0mean', 'rhou2u2mean']

if os.path.isdir('/root/reference'):
    # the channel app as shipped (statistics on): the running sums after 5 steps, divided by niter by the loop after the time loop
    FIXTURES['tcf_teno6_stats_16x24x12'] = ('tcf_teno6_stats', tcf_teno6_plan(16, 24, 12), [5])
    FIXTURES['isr_invwall_48x32'] = ('isr_invwall', isr_plan(48, 32, 'inviscid_wall'), [1, 20])
    FIXTURES['isr_48x32'] = ('isr', isr_plan(48, 32, 'symmetry'), [1, 20])
    FIXTURES['isr_split_48x32'] = ('isr_split', isr_plan(48, 32, 'split'), [1, 20])
    FIXTURES['ewc_wenoz5_32'] = ('ewc', ewc_plan(32), [1, 10])
    FIXTURES['ewc_teno5_32'] = ('ewc_teno5', ewc_plan(32, 'teno'), [1, 10])
    FIXTURES['trans_40x30x8'] = ('trans', trans_plan(40, 30, 8), [1, 5, 20])
    FIXTURES['vst_60x30'] = ('vst', vst_plan(60, 30), [1, 10, 200])
    FIXTURES['lam2d_16x64'] = ('lam2d', lam2d_plan(16, 64), [1, 10])
    FIXTURES['tcf_central_16x24x12'] = ('tcf_central', tcf_central_plan(16, 24, 12), [1, 5])
    FIXTURES['tcf_teno6_16x24x12'] = ('tcf_teno6', tcf_teno6_plan(16, 24, 12), [1, 5])
    FIXTURES['katzer_carpenter_60x40'] = ('katzer_carpenter', katzer_carpenter_plan(60, 40), [1, 10])
FIXTURES['tgv_sym_17'] = ('tgv_sym', tgv_sym_plan(17), [1, 3])
# the isothermal-EOS Taylor-Green app (TGsym/TG_IsoT.py: four conserved variables, no energy equation): generic path, APP golden
FIXTURES['apps/tg_isot_17'] = ('tg_isot', tgv_sym_plan(17), [1, 3])


def env_params(plan):
    P = {'dt': plan['constants']['dt']}
    if 'metric_fields' in plan or plan.get('curvilinear') or plan['bc'][0][0]['type'] == 'symmetry':
        P = {}                        # grid spacings follow from the sizes inside the generated program
    for d in range(plan['ndim']):
        P['block0np%d' % d] = plan['np'][d]
    return P


def main():
    for name, (config, plan, steps) in FIXTURES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        nd = plan['ndim']
        fields = ['rho'] + ['rhou%d' % d for d in range(nd)] + ([] if config == 'tg_isot' else ['rhoE'])
        inner = (slice(5, -5),) * nd
        bc_tables = {}
        if config.startswith('isr'):
            # imposed states read off the reference's own dumps: the uniform initial state is the inflow state; the top row
            # after one step holds the shock-generator state along x (boundary plane j = np1 - 1, tangential range incl. halos)
            r0 = run_ref(config, dict(env_params(plan), niter=0), fields)
            plan['bc'][0][0]['q'] = [float(r0[f][5 + 3, 5 + 3]) for f in fields]
            r1 = run_ref(config, dict(env_params(plan), niter=1), fields)
            bc_tables['bc_table_1_1'] = np.stack([r1[f][5 + plan['np'][1] - 1, :] for f in fields])
        out = {'plan': np.array(json.dumps(plan, sort_keys=True)),
               'provenance': np.array(open(os.path.join(REF_DIR, config, 'provenance.json')).read())}
        out.update(bc_tables)
        extra = []
        if 'metric_fields' in plan:
            extra = [n for d, name in enumerate(plan['metric_fields']) if name for n in ('D%d%d' % (d, d), 'SD%d%d%d' % (d, d, d))]
        if plan.get('curvilinear'):
            extra = ['D%d%d' % (i, j) for i in range(nd) for j in range(nd)] + ['detJ']
        r = run_ref(config, dict(env_params(plan), niter=0), fields + extra, dump_all=bool(extra))
        if extra:
            # full padded arrays: the halo values of the initial state and of the metrics are part of the problem
            # (the reference's init and metric kernels fill them once; some BCs never rewrite them)
            out['q0_padded'] = np.stack([r[f] for f in fields])
            for n in extra:
                out['field_' + n] = r[n]
            if config.startswith('katzer'):
                out['bc_table_1_1'] = katzer_dirichlet_table(plan['np'][0])
            if config == 'trans':
                rx = run_ref(config, dict(env_params(plan), niter=0), ['x0', 'x1', 'x2'], dump_all=True)
                out['bc_table_1_1'] = trans_dirichlet_table(rx['x0'])
                out['field_BF_amp'] = 2.5e-3 * np.exp(-(rx['x0'] - 20.0) ** 2 - (rx['x1'] - 4.0) ** 2) * np.cos(0.23 * rx['x2'])
        out['q0'] = np.stack([r[f][inner] for f in fields])
        for n in steps:
            stats = STATS if config.endswith('_stats') else ['rho_filt', 'rhou0_filt', 'rhou1_filt', 'rhoE_filt'] if config.endswith('_sfd') else ['kappa'] if config.endswith('_wf') else []
            r = run_ref(config, dict(env_params(plan), niter=n), fields + stats, dump_all=bool(stats))
            out['q%d' % n] = np.stack([r[f][inner] for f in fields])
            for s in stats:
                out['stat_' + s] = r[s][inner]
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith('q')}, '%.1f kB' % (os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
