#!/usr/bin/env python
"""Build the *reference* CPU executables (oracle/_ref/) from the reference's own generated C.

TEST INFRASTRUCTURE ONLY.  For each named config this script
  1. runs the unmodified reference front end (/root/reference, through oracle/refshim.py) on the
     app script with PYTHONHASHSEED=0 in oracle/_ref/<config>/  -> opensbli.cpp, *_kernels.h,
     defdec_data_set.h, bc_exchanges.h  (the reference's OPSC output);
  2. makes every `name=Input;` simulation parameter overridable from the environment
     (`name = ops_env_int("name", <app value>)`) so one binary serves all grid sizes / step counts;
  3. compiles it unmodified against the OPS stand-in oracle/ops_seq.h:
        ref_seq : g++ -O2 -ffp-contract=off                      (parity oracle, "OPS seq")
        ref_omp : g++ -O3 -march=x86-64-v3 -fopenmp -DOPS_OMP   (CPU baseline, "OPS openmp" stand-in)
Outputs go only to oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
Nothing from /root/reference is copied into the repository.

    python oracle/gen_ref.py [config ...]        # default: all configs
"""
import os
import sys
import subprocess
import hashlib
import json

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get('OSBLI_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')

# name -> (app path, [(old, new) source edits applied to the app text before exec])
CONFIGS = {
    # config 1: shipped Sod app (TENO5, N=200) and its WENO-JS5 / WENO-Z5 variants (N via env)
    'sod_teno5': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', []),
    'sod_wenojs5': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [
        ("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
        ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(5, formulation='JS', averaging=Avg)"),
        ("'200'", "'800'")]),
    'sod_wenoz5': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [
        ("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
        ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(5, formulation='Z', averaging=Avg)")]),
    # WENO orders the hand-written sweeps do not cover (they are built around the 6-point window of order 5): run on the generic path
    'sod_weno7': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [
        ("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
        ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(7, formulation='JS', averaging=Avg)")]),
    'sod_weno3': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [
        ("'scheme\\':\\'Teno\\'", "'scheme\\':\\'Weno\\'"),
        ("LLFTeno(teno_order, averaging=Avg)", "LLFWeno(3, formulation='Z', averaging=Avg)")]),
    # the isothermal-equation-of-state Taylor-Green app (no energy equation, p = rho / (gama Minf^2)): other constituent relations
    # than the hand-written kernels implement -> generic path
    'tg_isot': (REF + '/apps/taylor_green_vortex/TGsym/TG_IsoT.py', []),
    # boundary classes no shipped app uses: the Sod app with a zero-gradient / a pressure outlet on the right,
    # the inviscid shock reflection with its bottom wall as InviscidWallBC instead of SymmetryBC
    'sod_zgo': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [ZeroGradientOutletBC(direction, 1)]")]),
    'sod_pout': (REF + '/apps/Sod_shock_tube/Sod_shock_tube.py', [("boundaries += [DirichletBC(direction, 1, right_eqns)]", "boundaries += [PressureOutletBC(direction, 1, 0.1)]")]),
    'isr_invwall': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py', [("boundaries[direction][side] = SymmetryBC(direction, side)",
                     "from opensbli.core.boundary_conditions.inviscid_wall import InviscidWallBC\nboundaries[direction][side] = InviscidWallBC(direction, side)")]),
    'isr': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py', []),
    # the same with its bottom wall split in two (SplitBC, bc_core.py:200-217): SymmetryBC over x-points [0, 20), InviscidWallBC
    # over the rest; the part ranges are run-time arrays the reference leaves as `Input` for the user to edit (CPP_EDITS below)
    'isr_split': (REF + '/apps/inviscid_shock_reflection/inviscid_shock.py', [("boundaries[direction][side] = SymmetryBC(direction, side)",
                  "from opensbli.core.boundary_conditions.inviscid_wall import InviscidWallBC\nfrom opensbli.core.boundary_conditions.bc_core import SplitBC\n"
                  "boundaries[direction][side] = SplitBC(direction, side, [SymmetryBC(direction, side), InviscidWallBC(direction, side)])")]),
    # config 2: shipped TGV app (central-4 + RK3)
    'tgv_central4': (REF + '/apps/taylor_green_vortex/taylor_green_vortex.py', []),
    # the same with the non-linear WENO filter (filters/WENO_filter.py) applied after every step of the central scheme: characteristic
    # WENO5 reconstruction of the dissipative flux part in the three directions, Ducros sensor, filter application -- all
    # UserDefinedEquations loops at the end of the iteration; the exchanges become 3/4 planes deep
    'tgv_wf': (REF + '/apps/taylor_green_vortex/taylor_green_vortex.py', [("block.set_equations([copy.deepcopy(constituent), copy.deepcopy(simulation_eq), initial])",
               "from opensbli.filters.WENO_filter import WENOFilter\nwf = WENOFilter(block, order=5)\n"
               "block.set_equations([copy.deepcopy(constituent), copy.deepcopy(simulation_eq), initial] + wf.equation_classes)")]),
    # config 3/5: TGV TENO5 + StoreSome + RK-LS (our app script, reference front end + OPSC back end)
    'tgv_teno5': (REPO + '/apps/tgv_teno5.py', []),
    # config 4: shipped Katzer SBLI app (ReducedAccess closures) and a variant with the default Carpenter closures
    'katzer': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', []),
    'katzer_carpenter': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', [(", scheme=ReducedAccess())", ")")]),
    # config 4 with the selective-frequency-damping filter (filters/SFD.py): a `User kernel` before the loop (filtered state <- state)
    # and one at the end of every iteration that relaxes the state towards its filtered copy
    'katzer_sfd': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', [("block.set_equations([constituent, simulation_eq, initial, metriceq])",
                   "from opensbli.filters.SFD import SFD\nsfd = SFD(block, chifilt=0.1, omegafilt=1.0/0.75)\n"
                   "block.set_equations([constituent, simulation_eq, initial, metriceq] + sfd.equation_classes)")]),
    # config 4 as BASELINE.json words it: the same app with WENO-Z instead of adaptive TENO (no shock sensor)
    'katzer_wenoz': (REF + '/apps/katzer_SBLI/katzer_SBLI.py', [
        ("sc1 = \"**{\\'scheme\\':\\'Teno\\'}\"", "sc1 = \"**{\\'scheme\\':\\'Weno\\'}\""),
        ("constituent.add_equations(shock_sensor)", "pass"),
        ("LLF = LLFTeno(teno_order, formulation='adaptive', averaging=Avg, sensor=sensor_array, store_sensor=True)",
         "LLF = LLFWeno(5, formulation='Z', averaging=Avg)"),
        (", DataObject('D11'), DataObject('TENO')])", ", DataObject('D11')])")]),
    # 3-D turbulent channel (TENO6, Carpenter closures, stretched wall-normal grid, isothermal walls, power-law viscosity,
    # body force, SSP-RK3), statistics gathering switched off (it is outside the hot path)
    'tcf_teno6': (REF + '/apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py', [("stats = True", "stats = False")]),
    # Central(4) next to walls: 2-D laminar channel (Blaisdell skew form, StoreSome, Carpenter closures, isothermal walls,
    # Sutherland viscosity, body force) and the 3-D turbulent channel in the Feiereisen split on a stretched grid
    'lam2d': (REF + '/apps/channel_flow/laminar_2D/laminar_channel.py', []),
    'tcf_central': (REF + '/apps/channel_flow/compressible_TCF_Central/turbulent_channel.py', [("stats = True", "stats = False")]),
    # 2-D viscous shock tube: TENO5 + StoreSome viscous terms with a constant `mu` symbol, adiabatic walls and a symmetry
    # face, all with ReducedAccess closures
    'vst': (REF + '/apps/viscous_shock_tube/viscous_shock_tube.py', []),
    # 3-D transitional SBLI (statistics off): Katzer's set-up in 3-D with TENO6-adaptive and a time-periodic mass source
    'trans': (REF + '/apps/transitional_SBLI/transitional_SBLI.py', [("stats = True", "stats = False")]),
    # 2-D Euler wave on a fully curvilinear (wavy) grid: WENO-Z with metric direction cosines in the eigensystem
    'ewc': (REF + '/apps/euler_wave_curvilinear/euler_wave.py', []),
    'ewc_teno5': (REF + '/apps/euler_wave_curvilinear/euler_wave.py', [
        ("sc1 = \"**{\\'scheme\\':\\'Weno\\'}\"", "sc1 = \"**{\\'scheme\\':\\'Teno\\'}\""),
        ("LLFWeno(weno_order, formulation='Z', averaging=Avg)", "LLFTeno(5, averaging=Avg)"),
        ("'Delta0block0', 'Delta1block0']", "'Delta0block0', 'Delta1block0', 'eps', 'TENO_CT']"),
        ("'2.0/(block0np0)', '2.0/(block0np1)']", "'2.0/(block0np0)', '2.0/(block0np1)', '1e-15', '1e-6']")]),
    # the same channel app exactly as shipped, i.e. with its statistics-gathering user kernels (stats.py) switched on
    'tcf_teno6_stats': (REF + '/apps/channel_flow/compressible_TCF_TENO/turbulent_channel.py', []),
    # symmetry boundaries: shipped 1/8-domain TGV (central-4 + RK3, SymmetryBC on all six faces)
    'tgv_sym': (REF + '/apps/taylor_green_vortex/TGsym/TGsym.py', []),
}

# hand edits of the generated opensbli.cpp that the reference expects from its user (values it prints as `Input` and
# substitute_simulation_parameters does not reach): the SplitBC part ranges; the tangential extensions (-3, +4) are those of
# the full-plane kernel ({-3, block0np0 + 4, 0, 1} in oracle/_ref/isr/opensbli.cpp)
CPP_EDITS = {
    'isr_split': [("int split_range_100[] = {Input, Input, Input, Input};", "int split_range_100[] = {0, 20, 0, 1};"),
                  ("int split_halo_range_100[] = {Input, Input, Input, Input};", "int split_halo_range_100[] = {-3, 0, 0, 0};"),
                  ("int split_range_101[] = {Input, Input, Input, Input};", "int split_range_101[] = {20, block0np0, 0, 1};"),
                  ("int split_halo_range_101[] = {Input, Input, Input, Input};", "int split_halo_range_101[] = {0, 4, 0, 0};")],
}

INT_PARAMS = ('niter',)


def acc_header():
    lines = ['// generated by oracle/gen_ref.py -- OPS_ACCn accessor macros for oracle/ops_seq.h']
    for ar, cond in (('x', 'OPS_1D'), ('x,y', 'OPS_2D'), ('x,y,z', 'OPS_3D')):
        lines.append('#ifdef %s' % cond)
        for n in range(192):
            if ar == 'x':
                e = '(x)'
            elif ar == 'x,y':
                e = '((x)+ops_xdim[%d]*(y))' % n
            else:
                e = '((x)+ops_xdim[%d]*(y)+ops_xdim[%d]*ops_ydim[%d]*(z))' % (n, n, n)
            lines.append('#define OPS_ACC%d(%s) %s' % (n, ar, e))
        lines.append('#endif')
    return '\n'.join(lines) + '\n'


def env_substitute(constants, values, simulation_name='opensbli'):
    """Stand-in for helperfunctions.substitute_simulation_parameters (helperfunctions.py:130-149):
    same text replacement of `name=Input;`, but the app's value becomes the *default* of an
    environment lookup."""
    path = './%s.cpp' % simulation_name
    s = open(path).read()
    for const, value in zip(constants, values):
        old = const + '=Input;'
        if old in s:
            is_int = const in INT_PARAMS or (const.startswith('block') and 'np' in const)
            fn = 'ops_env_int' if is_int else 'ops_env_double'
            s = s.replace(old, '%s = %s("%s", %s);' % (const, fn, const, value))
    open(path, 'w').write(s)
    json.dump(dict(zip(constants, values)), open('params.json', 'w'), indent=1)


def generate(name):
    app, edits = CONFIGS[name]
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    os.chdir(d)
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(app))
    import refshim
    refshim.install(REF)
    import opensbli.utilities.helperfunctions as H
    original = H.substitute_simulation_parameters
    H.substitute_simulation_parameters = env_substitute
    for mod in list(sys.modules.values()):      # star-imports (e.g. `from opensbli import *` in an app's stats.py) re-export the original
        if getattr(mod, 'substitute_simulation_parameters', None) is original:
            mod.substitute_simulation_parameters = env_substitute
    src = open(app).read()
    for old, new in edits:
        assert old in src, (name, old)
        src = src.replace(old, new)
    os.environ['OSBLI_BACKEND'] = 'opsc'
    g = {'__name__': '__main__', '__file__': app}
    exec(compile(src, app, 'exec'), g)
    if name in CPP_EDITS:
        cpp = open('opensbli.cpp').read()
        for old, new in CPP_EDITS[name]:
            assert old in cpp, (name, old)
            cpp = cpp.replace(old, new)
        open('opensbli.cpp', 'w').write(cpp)
    sha = {}
    for f in sorted(os.listdir('.')):
        if f.endswith(('.cpp', '.h')):
            sha[f] = hashlib.sha256(open(f, 'rb').read()).hexdigest()
    import sympy
    json.dump({'config': name, 'app': app, 'edits': edits, 'sympy': sympy.__version__,
               'python': sys.version.split()[0], 'PYTHONHASHSEED': os.environ.get('PYTHONHASHSEED'),
               'sha256': sha}, open('provenance.json', 'w'), indent=1)


def compile_ref(name):
    d = os.path.join(OUT, name)
    open(os.path.join(OUT, 'ops_seq_acc.h'), 'w').write(acc_header())
    base = ['g++', '-std=c++17', '-w', '-I', HERE, '-I', OUT, '-I', d, os.path.join(d, 'opensbli.cpp')]
    subprocess.check_call(base + ['-O2', '-ffp-contract=off', '-o', os.path.join(d, 'ref_seq')])
    subprocess.check_call(base + ['-O3', '-march=x86-64-v3', '-fopenmp', '-DOPS_OMP', '-o', os.path.join(d, 'ref_omp')])


def main():
    names = sys.argv[1:] or list(CONFIGS)
    if os.environ.get('_OSBLI_GEN_CHILD'):
        generate(names[0])
        return
    if not os.path.isdir(REF):
        print('reference not present at %s: keeping prebuilt oracle/_ref' % REF)
        return
    for n in names:
        env = dict(os.environ, PYTHONHASHSEED='0', _OSBLI_GEN_CHILD='1')
        print('== generating reference C for', n, flush=True)
        subprocess.check_call([sys.executable, '-W', 'ignore', os.path.abspath(__file__), n], env=env,
                              stdout=open(os.devnull, 'w'))
        print('== compiling', n, flush=True)
        compile_ref(n)


if __name__ == '__main__':
    main()
