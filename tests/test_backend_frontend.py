"""CPU tests of the drop-in boundary: the reference front end (only importable in the build container, where
/root/reference exists) drives `B200(alg)`; the plans it distils must equal the hand-written plans of the golden
fixtures.  The committed plan fixtures under tests/golden/plans/ make the same check possible without the reference."""
import json
import os
import re
import subprocess
import sys

import pytest

from common import load_fixture

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANS = os.path.join(REPO, 'tests', 'golden', 'plans')
REF = '/root/reference'

APPS = {
    'tgv_teno5': (os.path.join(REPO, 'apps', 'tgv_teno5.py'), [], 'tgv_teno5_16'),
    'tgv_central4': (REF + '/apps/taylor_green_vortex/taylor_green_vortex.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'tgv_central4_16'),
    'katzer': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'katzer_60x40'),
    'tcf_teno6': (REF + '/apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py',
                  [("stats = True", "stats = False"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'tcf_teno6_16x24x12'),
    'lam2d': (REF + '/apps/channel_flow/laminar_2D/laminar_channel.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'lam2d_16x64'),
    'tcf_central': (REF + '/apps/channel_flow/compressible_TCF_Central/turbulent_channel.py',
                    [("stats = True", "stats = False"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'tcf_central_16x24x12'),
    'vst': (REF + '/apps/viscous_shock_tube/viscous_shock_tube.py',
            [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'vst_60x30'),
    'trans': (REF + '/apps/transitional_SBLI/transitional_SBLI.py',
              [("stats = True", "stats = False"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'trans_40x30x8'),
    'ewc': (REF + '/apps/euler_wave_curvilinear/euler_wave.py',
            [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'ewc_wenoz5_32'),
    # the channel app exactly as shipped: statistics gathering (`User kernel` loops of stats.py) switched on
    'tcf_teno6_stats': (REF + '/apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py',
                        [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'tcf_teno6_stats_16x24x12'),
    # 3-D channel in the Feiereisen split on a uniform grid, as shipped: statistics user kernels and a SimulationMonitor
    't3d': (REF + '/apps/channel_flow/turbulent_3D/turbulent_channel.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    'sod_teno5': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'sod_teno5_n200'),
    # boundary classes no shipped app uses: Sod with a zero-gradient / pressure outlet, shock reflection with an InviscidWallBC
    'sod_zgo': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [ZeroGradientOutletBC(direction, 1)]"),
                                                                 ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'sod_zgo_n200'),
    'sod_pout': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [PressureOutletBC(direction, 1, 0.1)]"),
                                                                  ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'sod_pout_n200'),
    'isr_invwall': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py',
                    [("boundaries[direction][side] = SymmetryBC(direction, side)",
                      "from opensbli.core.boundary_conditions.inviscid_wall import InviscidWallBC\nboundaries[direction][side] = InviscidWallBC(direction, side)"),
                     ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # the same three through the run-time compiled boundary-kernel path (OSB_GENERIC_BC forces it for classes that have a hand-written
    # kernel): what classes without one (ForcingStripWall, InletLawal, InletTransfer, InviscidWall2D) go through
    'sod_zgo_generic': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [ZeroGradientOutletBC(direction, 1)]"),
                                                                         ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_GENERIC_BC': 'ZeroGradientOutlet,Dirichlet'}),
    'sod_pout_generic': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [PressureOutletBC(direction, 1, 0.1)]"),
                                                                          ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_GENERIC_BC': 'PressureOutlet'}),
    'isr_invwall_generic': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py',
                            [("boundaries[direction][side] = SymmetryBC(direction, side)",
                              "from opensbli.core.boundary_conditions.inviscid_wall import InviscidWallBC\nboundaries[direction][side] = InviscidWallBC(direction, side)"),
                             ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_GENERIC_BC': 'Symmetry,Extrapolation,Dirichlet'}),
    # SplitBC (bc_core.py:200-217): the bottom wall of the shock reflection shared by a SymmetryBC and an InviscidWallBC; each part
    # becomes a run-time compiled kernel whose range comes from the integer arrays the user fills in in opensbli.cpp
    'isr_split': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py',
                  [("boundaries[direction][side] = SymmetryBC(direction, side)",
                    "from opensbli.core.boundary_conditions.inviscid_wall import InviscidWallBC\nfrom opensbli.core.boundary_conditions.bc_core import SplitBC\n"
                    "boundaries[direction][side] = SplitBC(direction, side, [SymmetryBC(direction, side), InviscidWallBC(direction, side)])"),
                   ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # selective frequency damping (filters/SFD.py) on the Katzer app: a cold user kernel + an in-loop user kernel that writes the state
    'katzer_sfd': (REF + '/apps/katzer_SBLI/katzer_SBLI.py',
                   [("block.set_equations([constituent, simulation_eq, initial, metriceq])",
                     "from opensbli.filters.SFD import SFD\nsfd = SFD(block, chifilt=0.1, omegafilt=1.0/0.75)\n"
                     "block.set_equations([constituent, simulation_eq, initial, metriceq] + sfd.equation_classes)"),
                    ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'katzer_60x40'),
    # the non-linear WENO filter (filters/WENO_filter.py) after every step of the central-4 TGV: 15 UserDefinedEquations loops
    'tgv_wf': (REF + '/apps/taylor_green_vortex/taylor_green_vortex.py',
               [("block.set_equations([copy.deepcopy(constituent), copy.deepcopy(simulation_eq), initial])",
                 "from opensbli.filters.WENO_filter import WENOFilter\nwf = WENOFilter(block, order=5)\n"
                 "block.set_equations([copy.deepcopy(constituent), copy.deepcopy(simulation_eq), initial] + wf.equation_classes)"),
                ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # programs outside the hand-written kernels run on the GENERIC path (every loop printed and compiled at run time): WENO orders 7 and 3
    'sod_weno7': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
                  ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(7, formulation='JS', averaging=Avg)"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    'sod_weno3': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
                  ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(3, formulation='Z', averaging=Avg)"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # isothermal-EOS Taylor-Green as shipped (no energy equation): generic path, 3-D central scheme with ~90 loops per stage
    'tg_isot': (REF + '/apps/taylor_green_vortex/TGsym/TG_IsoT.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # the two bench workloads FORCED through the generic path: the printed loops must reproduce what the hand-written kernels compute
    'tgv_teno5_allprinted': (os.path.join(REPO, 'apps', 'tgv_teno5.py'), [], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    'tgv_central4_allprinted': (REF + '/apps/taylor_green_vortex/taylor_green_vortex.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None,
                                    {'OSB_FORCE_GENERIC_PATH': '1'}),
    # ... and three general-path apps forced through it: stretched grid + ReducedAccess closures + adaptive TENO + wall / inflow / outflow
    # kernels (Katzer), 3-D channel with TENO6, Carpenter closures, power-law viscosity and body force, fully curvilinear WENO-Z
    'katzer_allprinted': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    'tcf_teno6_allprinted': (REF + '/apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py',
                             [("stats = True", "stats = False"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    # the 3-D transitional SBLI app (TENO6-adaptive, time-periodic mass source sin(omega dt iter): the loop counter reaches the printed
    # kernels) and the viscous shock tube (adiabatic walls, symmetry plane, a dataset that is read but never written)
    'trans_allprinted': (REF + '/apps/transitional_SBLI/transitional_SBLI.py',
                         [("stats = True", "stats = False"), ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    'vst_allprinted': (REF + '/apps/viscous_shock_tube/viscous_shock_tube.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    'ewc_allprinted': (REF + '/apps/euler_wave_curvilinear/euler_wave.py', [("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None, {'OSB_FORCE_GENERIC_PATH': '1'}),
    # InletTransferBC has no hand-written kernel: generic by itself (Sod with the left boundary copied from its first halo point)
    'sod_inlet_transfer': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 0, left_eqns)]", "boundaries += [InletTransferBC(direction, 0)]"),
                                                                            ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], None),
    # BASELINE configs[3] as worded: the Katzer app with WENO-Z instead of adaptive TENO (same edits as oracle/gen_ref.py)
    'katzer_wenoz': (REF + '/apps/katzer_SBLI/katzer_SBLI.py',
                     [("sc1 = \"**{\\'scheme\\':\\'Teno\\'}\"", "sc1 = \"**{\\'scheme\\':\\'Weno\\'}\""), ("constituent.add_equations(shock_sensor)", "pass"),
                      ("LLF = LLFTeno(teno_order, formulation='adaptive', averaging=Avg, sensor=sensor_array, store_sensor=True)",
                       "LLF = LLFWeno(5, formulation='Z', averaging=Avg)"),
                      (", DataObject('D11'), DataObject('TENO')])", ", DataObject('D11')])"),
                      ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")], 'katzer_wenoz_60x40'),
}

DRIVER = r'''
import sys, os
sys.path.insert(0, %(oracle)r); sys.path.insert(0, %(repo)r); sys.path.insert(0, os.path.dirname(%(app)r))   # apps import their neighbours
import refshim; refshim.install(%(ref)r)
os.environ['OSBLI_BACKEND'] = 'b200'
src = open(%(app)r).read()
for old, new in %(edits)r:
    assert old in src
    src = src.replace(old, new)
exec(compile(src, %(app)r, 'exec'), {'__name__': '__main__'})
'''


def start_app(name, workdir):
    app, edits = APPS[name][0], APPS[name][1]
    code = DRIVER % dict(oracle=os.path.join(REPO, 'oracle'), repo=REPO, ref=REF, app=app, edits=edits)
    env = dict(os.environ, PYTHONHASHSEED='0', **(APPS[name][3] if len(APPS[name]) > 3 else {}))
    return subprocess.Popen([sys.executable, '-W', 'ignore', '-c', code], cwd=workdir, env=env, stdout=subprocess.DEVNULL)


@pytest.fixture(scope='module')
def app_runs(tmp_path_factory):
    """All app scripts run through the reference front end concurrently (each is a minute of single-threaded SymPy)."""
    if not os.path.isdir(REF):
        return {}
    procs = {}
    for name in APPS:
        d = str(tmp_path_factory.mktemp(name))
        procs[name] = (d, start_app(name, d))
    return {name: (d, p.wait()) for name, (d, p) in procs.items()}


def comparable(plan):
    keys = ('ndim', 'np', 'conv', 'order', 'averaging', 'viscous', 'rk', 'rk_a', 'rk_b')
    p = {k: plan[k] for k in keys}
    p['bc'] = [[{k: v for k, v in b.items() if k != 'table'} for b in pair] for pair in plan['bc']]
    p['viscosity'] = plan.get('viscosity', {'type': 'constant'})
    p['teno_adaptive'] = bool(plan.get('teno_adaptive'))
    p['metric_fields'] = plan.get('metric_fields') or [None] * plan['ndim']
    p['forcing'] = bool(plan.get('forcing'))
    p['mass_source'] = plan.get('mass_source')
    p['curvilinear'] = bool(plan.get('curvilinear'))
    p['central_form'] = plan.get('central_form', 'blaisdell') if plan['conv'] == 'central' else None
    if plan['conv'] == 'weno':
        p['weno_formulation'] = plan.get('weno_formulation', 'JS')
    return p


@pytest.mark.parametrize('name', sorted(APPS))
def test_b200_backend_distils_expected_plan(name, app_runs):
    from opensbli_b200 import run as R
    if os.path.isdir(REF):
        workdir, rc = app_runs[name]
        assert rc == 0, 'B200(alg) failed on %s' % APPS[name][0]
        committed = os.path.join(PLANS, name)
        if os.environ.get('OSB_REFRESH_PLAN_FIXTURES') or not os.path.exists(os.path.join(committed, 'opensbli.cpp')):
            os.makedirs(committed, exist_ok=True)                  # opt-in: OSB_REFRESH_PLAN_FIXTURES=1 rewrites the committed fixtures
            for f in ('opensbli_b200.plan.json', 'opensbli.cpp'):
                open(os.path.join(committed, f), 'w').write(open(os.path.join(workdir, f)).read())
        else:                                                      # otherwise they are the regression reference of the back end
            fresh = json.load(open(os.path.join(workdir, 'opensbli_b200.plan.json')))
            kept = json.load(open(os.path.join(committed, 'opensbli_b200.plan.json')))
            assert fresh == kept, 'B200(alg) no longer distils the committed plan of %s: %s' % (
                name, sorted(k for k in set(fresh) | set(kept) if fresh.get(k) != kept.get(k)))
            assert open(os.path.join(workdir, 'opensbli.cpp')).read() == open(os.path.join(committed, 'opensbli.cpp')).read()
    else:
        workdir = os.path.join(PLANS, name)
        if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
            pytest.skip('no committed plan fixture for %s' % name)
    if APPS[name][2] is None:          # no golden run of this app: the plan fixture is refreshed, dedicated tests below read it
        sym = json.load(open(os.path.join(workdir, 'opensbli_b200.plan.json')))
        assert sym['ndim'] in (1, 2, 3)
        if name == 'isr_invwall':      # kernel named 'Symmetry' by the reference, recognised by what it assigns
            assert sym['bc'][1][0]['type'] == 'inviscid_wall' and sym['bc'][0][1] == {'type': 'extrapolation', 'order': 0}
        if name.endswith('_generic') or name == 'sod_inlet_transfer':
            gen = [(d, sd) for d, pair in enumerate(sym['bc']) for sd, b in enumerate(pair) if b['type'] == 'generic']
            assert gen and sorted(k['when'] for k in sym['user_kernels']) == sorted('bc_%d_%d' % f for f in gen)
            plan_sym, env, plan_num, _cold = R.load_case(workdir)
            for k in plan_num['user_kernels']:
                assert 'extern "C" __global__' in k['source'] and k['when'].startswith('bc_')
        return
    plan_sym, env, plan_num, _cold = R.load_case(workdir)
    want, _ = load_fixture(APPS[name][2])
    got = comparable(plan_num)
    exp = comparable(want)
    # the app's own grid size may differ from the fixture's
    got['np'] = exp['np']
    assert json.loads(json.dumps(got)) == json.loads(json.dumps(exp))
    for k in want['constants']:
        if k not in ('dt',) and k in plan_num['constants']:
            assert plan_num['constants'][k] == want['constants'][k], k
    assert 'gama' in plan_num['constants']
    # one-sided closure rows read off the IR equal the tables of the reference's scheme objects
    import numpy as np
    for cname, tab in want.get('closures', {}).items():
        for key in ('d1', 'd2'):
            assert np.allclose(np.array(plan_num['closures'][cname][key]), np.array(tab[key]), rtol=1e-12, atol=1e-14), (cname, key)
    # the stub keeps the reference's contract: every parameter was substituted, `int iter=0;` is present
    stub = open(os.path.join(workdir, 'opensbli.cpp')).read()
    assert '=Input;' not in stub and 'int iter=0;' in stub
    # print_iteration_ops (helperfunctions.py:172-190) works unchanged on the stub and is honoured by the runner
    calls = re.findall(r"^print_iteration_ops\((.*)\)", open(APPS[name][0]).read(), flags=re.M) if os.path.exists(APPS[name][0]) else []
    every, nan_name = R.read_iteration_ops(stub)
    if calls:
        want_every = int(re.search(r'every\s*=\s*(\d+)', calls[-1]).group(1)) if 'every' in calls[-1] else 250
        want_nan = re.search(r"NaN_check\s*=\s*'(\w+)'", calls[-1])
        assert every == want_every and nan_name == (want_nan.group(1)[:-3] if want_nan else None), (calls, every, nan_name)
    else:
        assert every is None and nan_name is None


def test_initial_state_from_cold_kernel_matches_reference_init():
    """numpy evaluation of the Grid_based_initialisation statements == the reference's own init kernel output."""
    import numpy as np
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 'tgv_teno5')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan_num, _cold = R.load_case(workdir)
    want, states = load_fixture('tgv_teno5_16')
    env = dict(env, block0np0=16, block0np1=16, block0np2=16, Delta0block0=want['delta'][0], Delta1block0=want['delta'][1], Delta2block0=want['delta'][2])
    plan_num, cold = R.resolve(plan_sym, env)
    q0 = R.initial_state(plan_sym, cold)
    s = (slice(5, -5),) * 3
    for m in range(5):
        assert np.abs(q0[m][s] - states[0][m]).max() <= 1e-13 * max(1.0, np.abs(states[0][m]).max())


@pytest.mark.parametrize('name,fixture,sizes', [('lam2d', 'lam2d_16x64', (16, 64)), ('vst', 'vst_60x30', (60, 30)), ('tcf_central', 'tcf_central_16x24x12', (16, 24, 12)),
                                                ('tcf_teno6', 'tcf_teno6_16x24x12', (16, 24, 12))])
def test_channel_cold_kernels_match_reference(name, fixture, sizes):
    """Channel apps: initial condition, stretched-grid metrics and their boundary kernels evaluated by the runner from
    the plan fixture == the reference's own cold kernels (golden), at the fixture's grid size."""
    import numpy as np
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, name)
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    over = {'block0np%d' % d: n for d, n in enumerate(sizes)}
    plan_sym, env, plan_num, cold = R.load_case(workdir, overrides=over)
    want, states = load_fixture(fixture)
    assert plan_num['np'] == list(sizes)
    assert np.allclose(plan_num['delta'], want['delta'], rtol=1e-14)
    q0 = R.initial_state(plan_sym, cold)
    s = (slice(5, -5),) * len(sizes)
    for m in range(len(q0)):
        assert np.abs(q0[m][s] - states[0][m]).max() <= 1e-12 * max(1.0, np.abs(states[0][m]).max())
    # metric arrays: interior points (the hot loops read D_dd / SD_ddd at the point itself only; the reference additionally
    # copies them into the periodic halos, which nothing reads)
    for f, a in want.get('fields', {}).items():
        assert np.abs(plan_num['fields'][f][s] - a[s]).max() <= 1e-11 * np.abs(a).max(), f


def test_simulation_monitor_is_driven_by_the_runner(tmp_path):
    """turbulent_3D as shipped: the SimulationMonitor (probes of rho, u0, p, T, u1 every 250 iterations into output.log) is
    distilled into the plan; the runner cuts the time loop at the printing iterations and writes the reference's format."""
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 't3d')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan, cold = R.load_case(workdir, overrides={'block0np0': 16, 'block0np1': 130, 'block0np2': 16, 'niter': 600})
    mon = plan['monitor']
    assert mon['arrays'] == ['rho', 'u0', 'p', 'T', 'u1'] and mon['frequency'] == 250 and mon['output_file'] == 'output.log'
    assert mon['probes'][3] == [64, 15, 64] and plan['central_form'] == 'feiereisen' and len(plan['user_kernels']) == 2

    class FakeSim(object):
        calls, it = [], 0

        def step_timed(self, n):
            self.calls.append(n); self.it += n
            return 1.0 * n

        def read_point(self, name, i, j=0, k=0):
            return self.it + 0.5
    sim = FakeSim()
    ms = R.time_loop(sim, plan, 600, str(tmp_path))
    assert sim.calls == [1, 249, 250, 100] and ms == 600.0
    lines = open(os.path.join(str(tmp_path), 'output.log')).read().splitlines()
    assert lines[0].startswith('Iteration, Time, rho_B0(0, 10, 20)') and len(lines) == 4
    assert lines[1].split(', ')[0] == '1' and lines[3].split(', ')[0] == '500'
    assert lines[2].split(', ')[2] == '250.500000000000'          # fp_precision = 12


def test_user_kernels_become_cuda_source():
    """Statistics loops of the shipped channel app: distilled as point-wise user kernels and printed as CUDA C by the runner."""
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 'tcf_teno6_stats')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan, cold = R.load_case(workdir, overrides={'block0np0': 16, 'block0np1': 24, 'block0np2': 12, 'niter': 5})
    uk = plan['user_kernels']
    assert [k['when'] for k in uk] == ['iteration_end', 'after_loop']
    acc, norm = uk
    assert acc['range'] == [0, 16, 0, 24, 0, 12] and 'u0u0mean' in acc['writes'] and 'rho' in acc['fields']
    assert 'extern "C" __global__ void osb_user_kernel_0(' in acc['source'] and 'u0u0mean[X] = ' in acc['source']
    assert '#define niter (5.0)' in norm['source'] and 'rhomean[X] = rhomean[X]/niter;' in norm['source']


def test_curvilinear_cold_data_match_reference():
    """apps/euler_wave_curvilinear: the full metric tensor and detJ -- metric kernels plus the periodic exchanges of the cold
    phase -- evaluated by the runner from the plan fixture equal the reference's arrays in every point incl. the halos."""
    import numpy as np
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 'ewc')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan, cold = R.load_case(workdir, overrides={'block0np0': 32, 'block0np1': 32})
    want, states = load_fixture('ewc_wenoz5_32')
    assert plan['curvilinear'] and plan['conv'] == 'weno' and plan['weno_formulation'] == 'Z'
    for f in ('D00', 'D01', 'D10', 'D11', 'detJ'):
        assert np.abs(plan['fields'][f] - want['fields'][f]).max() <= 1e-14, f
    q0 = R.initial_state(plan_sym, cold)
    for m in range(4):
        assert np.abs(q0[m][5:-5, 5:-5] - states[0][m]).max() <= 1e-15


def test_transitional_sbli_cold_data_match_reference():
    """apps/transitional_SBLI: mass-source amplitude, metrics, shock-generator table (imposed variables, used range) and the
    polynomial boundary-layer initial condition evaluated by the runner == the reference's own cold kernels (golden)."""
    import numpy as np
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 'trans')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan, cold = R.load_case(workdir, overrides={'block0np0': 40, 'block0np1': 30, 'block0np2': 8})
    want, states = load_fixture('trans_40x30x8')
    s = (slice(5, -5),) * 3
    assert plan['mass_source'] == want['mass_source']
    for f in ('BF_amp', 'D11', 'SD111'):
        assert np.abs(plan['fields'][f][s] - want['fields'][f][s]).max() <= 1e-12 * np.abs(want['fields'][f]).max(), f
    top = plan['bc'][1][1]
    assert top['free'] == [3] and top['ke_free'] and top['closure'] == 'reduced_access'
    used = np.zeros((8 + 10, 40 + 10), bool)
    used[2:-1, 2:-1] = True                                   # the BC kernel's tangential range [-3, np + 4)
    for m in (0, 1, 2, 4):
        assert np.abs(top['table'][m].reshape(18, 50)[used] - want['bc'][1][1]['table'][m].reshape(18, 50)[used]).max() <= 1e-15
    q0 = R.initial_state(plan_sym, cold)
    for m in range(5):
        assert np.abs(q0[m][s] - states[0][m]).max() <= 1e-9      # degree-50 polynomial fit: ill-conditioned (see Katzer)


def test_sod_initial_state_and_dirichlet_states():
    """Piecewise initial condition and the Dirichlet boundary states, evaluated from the plan fixture."""
    import numpy as np
    from opensbli_b200 import run as R
    workdir = os.path.join(PLANS, 'sod_teno5')
    if not os.path.exists(os.path.join(workdir, 'opensbli.cpp')):
        pytest.skip('plan fixture missing')
    plan_sym, env, plan_num, cold = R.load_case(workdir)
    want, states = load_fixture('sod_teno5_n200')
    assert plan_num['niter'] == 1000 and plan_num['np'] == [200]
    q0 = R.initial_state(plan_sym, cold)
    for m in range(3):
        assert np.abs(q0[m][5:-5] - states[0][m]).max() <= 1e-15
    for s in range(2):
        assert np.allclose(plan_num['bc'][0][s]['q'], want['bc'][0][s]['q'], rtol=1e-15, atol=0)


def test_c_expression_semantics():
    from opensbli_b200.run import c_eval
    assert c_eval('ceil(0.2/0.0002)', {}) == 1000.0
    assert c_eval('2*M_PI/block0np0', {'block0np0': 64}) == 2 * 3.141592653589793 / 64
    assert c_eval('7/2', {}) == 3 and c_eval('7/2.0', {}) == 3.5 and c_eval('1.0/(n-1)', {'n': 200}) == 1.0 / 199
    assert c_eval('1.0e-16', {}) == 1e-16


TGV = REF + '/apps/taylor_green_vortex/taylor_green_vortex.py'
B200_LINE = ("OPSC(alg)", "from opensbli_b200 import B200\nB200(alg)")
UNSUPPORTED = {
    # another equation of state (isothermal variant of the TGV app)
    'other_eos': (REF + '/apps/taylor_green_vortex/TGsym/TG_IsoT.py', [B200_LINE], 'constituent relation'),
    # a stress tensor with another bulk-viscosity factor: the viscous loops no longer equal the implemented terms
    'other_stress_tensor': (TGV, [B200_LINE, ("- (2/3)* KD(_i,_j)* Der(u_k,x_k)", "- (1/3)* KD(_i,_j)* Der(u_k,x_k)")], 'viscous terms added to'),
    # plain conservative central convective terms instead of the skew-symmetric split
    'other_convective_split': (TGV, [B200_LINE, ("- Skew(rho*u_j,x_j)", "- Conservative(rho*u_j,x_j)")], 'central convective'),
}


@pytest.fixture(scope='module')
def unsupported_runs(tmp_path_factory):
    """every case twice: as a user runs it (falls back to the generic path) and with OSB_NO_GENERIC_PATH=1 (must raise)"""
    if not os.path.isdir(REF):
        return {}
    procs = {}
    for name, (app, edits, _) in UNSUPPORTED.items():
        code = DRIVER % dict(oracle=os.path.join(REPO, 'oracle'), repo=REPO, ref=REF, app=app, edits=edits)
        for strict in (False, True):
            d = str(tmp_path_factory.mktemp(name + ('_strict' if strict else '')))
            env = dict(os.environ, PYTHONHASHSEED='0', **({'OSB_NO_GENERIC_PATH': '1'} if strict else {}))
            procs[name, strict] = (d, subprocess.Popen([sys.executable, '-W', 'ignore', '-c', code], cwd=d, env=env, stdout=subprocess.PIPE,
                                                       stderr=subprocess.PIPE, text=True))
    return {key: (d,) + p.communicate() + (p.returncode,) for key, (d, p) in procs.items()}


@pytest.mark.parametrize('name', sorted(UNSUPPORTED))
def test_programs_outside_the_hand_written_kernels(name, unsupported_runs):
    """An app whose equations differ from what the hand-written kernels compute must never silently compute something else:
    it is recognised (the reason names what differs) and runs on the GENERIC path -- its own loops, printed and compiled at run
    time -- or, with OSB_NO_GENERIC_PATH=1, raises."""
    if not os.path.isdir(REF):
        pytest.skip('needs the reference front end')
    d, out, err, rc = unsupported_runs[name, True]
    assert rc != 0 and 'UnsupportedByB200' in err and UNSUPPORTED[name][2] in err, err[-600:]
    d, out, err, rc = unsupported_runs[name, False]
    assert rc == 0 and 'GENERIC path' in out and UNSUPPORTED[name][2] in out, (out[-600:], err[-600:])
    sym = json.load(open(os.path.join(d, 'opensbli_b200.plan.json')))
    assert sym['conv'] == 'generic' and UNSUPPORTED[name][2] in sym['generic']['reason'] and sym['generic']['nstages'] == 3
    whens = set(k['when'] for k in sym['user_kernels'])
    assert 'stage' in whens and 'iteration_start' in whens
