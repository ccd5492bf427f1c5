"""Test-side helpers: ctypes binding of the CPU oracle (oracle/osbli_oracle.c), runner for the
reference executables in oracle/_ref/ and reader for their raw dumps.  Test infrastructure only."""
import ctypes
import os
import subprocess
import tempfile
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, 'oracle')
REF_DIR = os.path.join(ORACLE_DIR, '_ref')

CONV = {'central': 0, 'weno': 1, 'teno': 2}
BC = {'periodic': 0, 'dirichlet': 1, 'exchange': 2, 'open': 2, 'split': 14, 'isothermal_wall': 3, 'extrapolation': 4,
      'inlet_pressure_extrapolate': 5, 'symmetry': 6, 'dirichlet_field': 7, 'adiabatic_wall': 8,
      'zero_gradient_outlet': 9, 'pressure_outlet': 10, 'inviscid_wall': 11}
MU = {'constant': 0, 'sutherland': 1, 'power': 2}
CLOSURES = {
    # rows idx = 0.. next to the face x weights of the boundary-absolute points 0..np-1
    # reduced_access_scheme.py:36-43, 76-83 ; Carpenter second derivative Carpenter_scheme.py:69-76
    'reduced_access': dict(d1=[[-25.0 / 12, 48.0 / 12, -36.0 / 12, 16.0 / 12, -3.0 / 12], [-3.0 / 12, -10.0 / 12, 18.0 / 12, -6.0 / 12, 1.0 / 12]],
                           d2=[[35.0 / 12, -104.0 / 12, 114.0 / 12, -56.0 / 12, 11.0 / 12], [11.0 / 12, -20.0 / 12, 6.0 / 12, 4.0 / 12, -1.0 / 12]]),
}


class OsboCfg(ctypes.Structure):
    _fields_ = [('ndim', ctypes.c_int), ('np', ctypes.c_int * 3), ('halo', ctypes.c_int),
                ('conv', ctypes.c_int), ('order', ctypes.c_int), ('weno_z', ctypes.c_int),
                ('averaging', ctypes.c_int), ('viscous', ctypes.c_int), ('rk', ctypes.c_int),
                ('nstages', ctypes.c_int), ('rk_a', ctypes.c_double * 8), ('rk_b', ctypes.c_double * 8),
                ('gama', ctypes.c_double), ('Minf', ctypes.c_double), ('Re', ctypes.c_double),
                ('Pr', ctypes.c_double), ('dt', ctypes.c_double), ('eps', ctypes.c_double),
                ('teno_ct', ctypes.c_double), ('delta', ctypes.c_double * 3),
                ('bc', (ctypes.c_int * 2) * 3), ('bc_q', ((ctypes.c_double * 5) * 2) * 3),
                # general path (keep in sync with oracle/osbli_oracle.h)
                ('visc_law', ctypes.c_int), ('SuthT', ctypes.c_double), ('RefT', ctypes.c_double), ('mu_exp', ctypes.c_double),
                ('D', ctypes.POINTER(ctypes.c_double) * 3), ('SD', ctypes.POINTER(ctypes.c_double) * 3),
                ('closure', (ctypes.c_int * 2) * 3),
                ('c_nr1', ctypes.c_int), ('c_np1', ctypes.c_int), ('c_nr2', ctypes.c_int), ('c_np2', ctypes.c_int),
                ('c_d1', ctypes.c_double * 24), ('c_d2', ctypes.c_double * 12),
                ('teno_adaptive', ctypes.c_int), ('teno_a1', ctypes.c_double), ('teno_a2', ctypes.c_double),
                ('sensor_eps', ctypes.c_double), ('theta', ctypes.POINTER(ctypes.c_double)),
                ('teno_store', ctypes.POINTER(ctypes.c_double)), ('Twall', ctypes.c_double),
                ('extrap_order', (ctypes.c_int * 2) * 3), ('bc_face', (ctypes.POINTER(ctypes.c_double) * 2) * 3),
                ('force', ctypes.c_double * 3), ('bc_free', (ctypes.c_int * 2) * 3), ('src_amp', ctypes.POINTER(ctypes.c_double)), ('src_rate', ctypes.c_double),
                ('src_iter0', ctypes.c_int), ('curv_D', (ctypes.POINTER(ctypes.c_double) * 3) * 3),
                ('curv_detJ', ctypes.POINTER(ctypes.c_double)), ('back_pressure', ctypes.c_double),
                ('split_n', (ctypes.c_int * 2) * 3), ('split_kind', ((ctypes.c_int * 8) * 2) * 3), ('split_lo', (((ctypes.c_int * 3) * 8) * 2) * 3),
                ('split_hi', (((ctypes.c_int * 3) * 8) * 2) * 3), ('split_order', ((ctypes.c_int * 8) * 2) * 3), ('split_q', (((ctypes.c_double * 5) * 8) * 2) * 3),
                ('halo_m', ctypes.c_int), ('halo_p', ctypes.c_int), ('central_form', ctypes.c_int)]


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, 'libosbli_oracle.so')
        src = os.path.join(ORACLE_DIR, 'osbli_oracle.c')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['make', '-C', ORACLE_DIR, 'libosbli_oracle.so'], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(so)
        _lib.osbo_padded_size.restype = ctypes.c_long
        for f in ('osbo_recon_teno5', 'osbo_recon_teno6', 'osbo_recon_weno5'):
            getattr(_lib, f).restype = ctypes.c_double
    return _lib


def make_cfg(plan):
    """plan: dict as produced by opensbli_b200.plan (numeric, resolved)."""
    c = OsboCfg()
    P = ctypes.POINTER(ctypes.c_double)
    keep = []
    c.ndim = plan['ndim']
    for d in range(3):
        c.np[d] = plan['np'][d] if d < plan['ndim'] else 1
        c.delta[d] = plan['delta'][d] if d < plan['ndim'] else 1.0
    c.halo = 5
    c.conv = CONV[plan['conv']]
    c.order = plan['order']
    c.weno_z = 1 if plan.get('weno_formulation', 'JS') == 'Z' else 0
    c.averaging = 1 if plan.get('averaging', 'roe') == 'roe' else 0
    c.viscous = 1 if plan.get('viscous') else 0
    c.rk = 0 if plan['rk'] == 'sbli' else 1
    c.nstages = len(plan['rk_a'])
    for s in range(c.nstages):
        c.rk_a[s] = plan['rk_a'][s]
        c.rk_b[s] = plan['rk_b'][s]
    k = plan['constants']
    c.gama = k['gama']
    c.Minf = k.get('Minf', 1.0)
    c.Re = k.get('Re', 1.0) / (k.get('mu', 1.0) if plan.get('viscosity', {'type': 'constant'})['type'] == 'constant' else 1.0)
    c.Pr = k.get('Pr', 1.0)
    c.dt = k['dt']
    c.eps = k.get('eps', 1e-16)
    c.teno_ct = k.get('TENO_CT', 1e-6)
    for d in range(plan['ndim']):
        for s in range(2):
            b = plan['bc'][d][s]
            c.bc[d][s] = BC[b['type']]
            if b['type'] == 'dirichlet':
                for m, v in enumerate(b['q']):
                    c.bc_q[d][s][m] = v
            if b['type'] == 'dirichlet_field':
                a = np.ascontiguousarray(b['table'], dtype=np.float64)
                keep.append(a)
                c.bc_face[d][s] = a.ctypes.data_as(P)
                c.bc_free[d][s] = sum(1 << m for m in b.get('free', [])) | (256 if b.get('ke_free') else 0)
            if b['type'] == 'extrapolation':
                c.extrap_order[d][s] = int(b.get('order', 0))
            if b['type'] == 'split':
                c.split_n[d][s] = len(b['parts'])
                for n, part in enumerate(b['parts']):
                    c.split_kind[d][s][n] = BC[part['type']]
                    c.split_order[d][s][n] = int(part.get('order', 0))
                    for e in range(3):
                        c.split_lo[d][s][n][e] = part['range'][2 * e] if e < plan['ndim'] else 0
                        c.split_hi[d][s][n][e] = part['range'][2 * e + 1] if e < plan['ndim'] else 1
                    for m, v in enumerate(part.get('q', ())):
                        c.split_q[d][s][n][m] = v
            if b.get('closure'):
                c.closure[d][s] = 1
                cl = plan['closures'][b['closure']] if 'closures' in plan else CLOSURES[b['closure']]
                d1, d2 = np.asarray(cl['d1'], dtype=np.float64), np.asarray(cl['d2'], dtype=np.float64)
                c.c_nr1, c.c_np1 = d1.shape
                c.c_nr2, c.c_np2 = d2.shape
                for i, v in enumerate(d1.ravel()):
                    c.c_d1[i] = v
                for i, v in enumerate(d2.ravel()):
                    c.c_d2[i] = v
    visc = plan.get('viscosity', {'type': 'constant'})
    c.visc_law = MU[visc['type']]
    c.SuthT, c.RefT, c.mu_exp = k.get('SuthT', 0.0), k.get('RefT', 1.0), visc.get('exponent', 0.0)
    c.Twall = k.get('Twall', 1.0)
    c.back_pressure = k.get('back_pressure', 0.0)
    c.central_form = {'blaisdell': 0, 'feiereisen': 1}[plan.get('central_form', 'blaisdell')]
    if plan.get('halos'):
        c.halo_m, c.halo_p = plan['halos']
    if plan.get('forcing'):
        for d in range(plan['ndim']):
            c.force[d] = k.get('c%d' % d, 0.0)
    shape = padded_shape(plan)
    for d, name in enumerate(plan.get('metric_fields', [None] * plan['ndim'])):
        if name:
            for arr, fld in ((c.D, 'D%d%d' % (d, d)), (c.SD, 'SD%d%d%d' % (d, d, d))):
                a = np.ascontiguousarray(plan['fields'][fld], dtype=np.float64)
                assert a.shape == shape
                keep.append(a)
                arr[d] = a.ctypes.data_as(P)
    if plan.get('curvilinear'):
        for i in range(plan['ndim']):
            for j in range(plan['ndim']):
                a = np.ascontiguousarray(plan['fields']['D%d%d' % (i, j)], dtype=np.float64)
                assert a.shape == shape
                keep.append(a)
                c.curv_D[i][j] = a.ctypes.data_as(P)
        a = np.ascontiguousarray(plan['fields']['detJ'], dtype=np.float64)
        keep.append(a)
        c.curv_detJ = a.ctypes.data_as(P)
    ms = plan.get('mass_source')
    if ms:
        a = np.ascontiguousarray(plan['fields'][ms['field']], dtype=np.float64)
        assert a.shape == shape
        keep.append(a)
        c.src_amp, c.src_rate, c.src_iter0 = a.ctypes.data_as(P), ms['rate'], int(plan.get('iteration0', 0))
    ad = plan.get('teno_adaptive')
    if ad:
        c.teno_adaptive = 1
        c.teno_a1, c.teno_a2, c.sensor_eps = k['teno_a1'], k['teno_a2'], k.get('epsilon', 1e-12)
        th, ts = np.zeros(shape), np.zeros(shape)
        keep += [th, ts]
        c.theta, c.teno_store = th.ctypes.data_as(P), ts.ctypes.data_as(P)
        c._theta, c._teno_store = th, ts
    c._keep = keep
    return c


def padded_shape(plan, halo=5):
    nd = plan['ndim']
    return tuple(plan['np'][d] + 2 * halo for d in reversed(range(nd)))   # numpy order (k,j,i)


def oracle_advance(plan, q, nsteps, rk_reg=None):
    """q: list of padded numpy arrays (C-order, x fastest), advanced in place. Returns rk registers."""
    lib = oracle_lib()
    cfg = make_cfg(plan)
    nv = plan['ndim'] + 2
    assert len(q) == nv
    q = [np.ascontiguousarray(a, dtype=np.float64) for a in q]
    if rk_reg is None:
        rk_reg = [np.zeros_like(a) for a in q]
    P = ctypes.POINTER(ctypes.c_double)
    qa = (P * nv)(*[a.ctypes.data_as(P) for a in q])
    ra = (P * nv)(*[a.ctypes.data_as(P) for a in rk_reg])
    rc = lib.osbo_advance(ctypes.byref(cfg), qa, ra, ctypes.c_int(nsteps))
    assert rc == 0
    return q, rk_reg


def oracle_stage(plan, q, rk_reg, stage):
    """In-place: stage < 0 = iteration start, else one RK stage (arrays must be C-contiguous float64)."""
    lib = oracle_lib()
    cfg = make_cfg(plan)
    nv = plan['ndim'] + 2
    P = ctypes.POINTER(ctypes.c_double)
    qa = (P * nv)(*[a.ctypes.data_as(P) for a in q])
    ra = (P * nv)(*[a.ctypes.data_as(P) for a in rk_reg])
    assert lib.osbo_stage(ctypes.byref(cfg), qa, ra, ctypes.c_int(stage)) == 0


def oracle_apply_bcs(plan, q):
    """In-place boundary conditions of the rank-local faces ('exchange' faces are left alone)."""
    lib = oracle_lib()
    cfg = make_cfg(plan)
    nv = plan['ndim'] + 2
    P = ctypes.POINTER(ctypes.c_double)
    lib.osbo_apply_bcs(ctypes.byref(cfg), (P * nv)(*[a.ctypes.data_as(P) for a in q]))


def oracle_residual(plan, q):
    lib = oracle_lib()
    cfg = make_cfg(plan)
    nv = plan['ndim'] + 2
    q = [np.ascontiguousarray(a, dtype=np.float64) for a in q]
    R = [np.zeros_like(a) for a in q]
    P = ctypes.POINTER(ctypes.c_double)
    qa = (P * nv)(*[a.ctypes.data_as(P) for a in q])
    ra = (P * nv)(*[a.ctypes.data_as(P) for a in R])
    lib.osbo_residual(ctypes.byref(cfg), qa, ra)
    return R


# ---------------------------------------------------------------- reference executables
def have_ref(config):
    return os.path.exists(os.path.join(REF_DIR, config, 'ref_seq'))


def read_dump(path):
    raw = open(path, 'rb').read()
    hdr = np.frombuffer(raw[:40], dtype=np.int32)
    nd = int(hdr[0])
    size, d_m, d_p = hdr[1:4], hdr[4:7], hdr[7:10]
    pdim = [int(size[d] - d_m[d] + d_p[d]) for d in range(nd)]
    data = np.frombuffer(raw[40:], dtype=np.float64).reshape(tuple(reversed(pdim))).copy()
    return data


def run_ref(config, params, fields, exe='ref_seq', dump_all=False, threads=None):
    """Run oracle/_ref/<config>/<exe> with parameter overrides (dict of env names, e.g. block0np0, niter, dt)
    and return {field: padded ndarray} of the final-time dump (+ wall time of the reference's own timer)."""
    d = os.path.join(REF_DIR, config)
    with tempfile.TemporaryDirectory() as out:
        env = dict(os.environ, OSBLI_OUT=out)
        if dump_all:
            env['OSBLI_DUMP_ALL'] = '1'
        if threads:
            env['OMP_NUM_THREADS'] = str(threads)
        for k, v in params.items():
            env[k] = repr(v) if isinstance(v, float) else str(v)
        res = subprocess.run([os.path.join(d, exe)], env=env, cwd=out, stdout=subprocess.PIPE, check=True, text=True)
        wall = None
        for line in res.stdout.splitlines():
            if 'Total Wall time' in line:
                wall = float(line.split()[-1])
        outd = {}
        for f in fields:
            cand = [p for p in os.listdir(out) if p.endswith('.%s_B0.bin' % f) and (dump_all == p.startswith('all.'))]
            assert len(cand) == 1, (f, os.listdir(out))
            outd[f] = read_dump(os.path.join(out, cand[0]))
        outd['_wall'] = wall
    return outd
