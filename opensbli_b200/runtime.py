"""ctypes binding of the C ABI (include/osbli_b200.h) and the `Simulation` driver object.

There is no CPU fallback: if the CUDA library is missing or no GPU is present, construction fails loudly.
"""
import ctypes
import os
import numpy as np

from . import plan as _plan
from .build import LIB

_P = ctypes.POINTER(ctypes.c_double)
_lib = None

FAMILIES = ('prim', 'flux', 'central', 'viscous', 'rk', 'bc', 'sync', 'user')

SYMBOLS = ('osb_create', 'osb_destroy', 'osb_last_error', 'osb_set_const_f64', 'osb_get_const_f64', 'osb_set_iteration', 'osb_get_iteration', 'osb_create_field', 'osb_add_user_kernel', 'osb_run_user_kernels', 'osb_read_point',
           'osb_num_fields', 'osb_field_name', 'osb_field_info', 'osb_upload', 'osb_download', 'osb_device_ptr', 'osb_upload_face',
           'osb_step', 'osb_step_begin', 'osb_stage', 'osb_sync', 'osb_apply_bcs', 'osb_residual', 'osb_step_timed', 'osb_timer_start', 'osb_timer_stop', 'osb_advance_host',
           'osb_launch_count', 'osb_slow_path_count', 'osb_profile_step', 'osb_nan_check', 'osb_diagnostics', 'osb_ipc_export', 'osb_ipc_import', 'osb_halo_push',
           'osb_host_planes_upload', 'osb_host_planes_ready', 'osb_host_planes_download', 'osb_host_planes_sync',
           'osb_staging_create', 'osb_staging_destroy', 'osb_staging_last_error', 'osb_staging_upload', 'osb_staging_feed', 'osb_staging_fed', 'osb_staging_sync',
           'osb_staging_ipc_export', 'osb_staging_ipc_import', 'osb_staging_pull',
           'osb_measure_fp64_peak')


class BackendError(RuntimeError):
    pass


def load_library(path=None):
    """Load libosbli_b200.so (built in-tree by opensbli_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get('OSB_B200_LIB') or LIB
    if not os.path.exists(path):
        raise BackendError('%s not found: build it with `python -m opensbli_b200.build` '
                           '(the B200 back end has no CPU fallback)' % path)
    lib = ctypes.CDLL(path)
    lib.osb_last_error.restype = ctypes.c_char_p
    lib.osb_last_error.argtypes = [ctypes.c_void_p]
    lib.osb_field_name.restype = ctypes.c_char_p
    lib.osb_field_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.osb_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.osb_destroy.argtypes = [ctypes.c_void_p]
    lib.osb_set_const_f64.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
    lib.osb_get_const_f64.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _P]
    lib.osb_set_iteration.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
    lib.osb_get_iteration.argtypes = [ctypes.c_void_p]
    lib.osb_get_iteration.restype = ctypes.c_longlong
    lib.osb_add_user_kernel.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.osb_run_user_kernels.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.osb_create_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.osb_read_point.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]
    lib.osb_num_fields.argtypes = [ctypes.c_void_p]
    lib.osb_field_info.argtypes = [ctypes.c_void_p, ctypes.c_char_p] + [ctypes.POINTER(ctypes.c_int)] * 3
    lib.osb_upload.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    lib.osb_download.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    lib.osb_device_ptr.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.osb_upload_face.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.osb_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.osb_sync.argtypes = [ctypes.c_void_p]
    lib.osb_step_begin.argtypes = [ctypes.c_void_p]
    lib.osb_stage.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.osb_apply_bcs.argtypes = [ctypes.c_void_p]
    lib.osb_residual.argtypes = [ctypes.c_void_p]
    lib.osb_step_timed.argtypes = [ctypes.c_void_p, ctypes.c_int, _P]
    lib.osb_timer_start.argtypes = [ctypes.c_void_p]
    lib.osb_timer_stop.argtypes = [ctypes.c_void_p, _P]
    lib.osb_advance_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _P]
    lib.osb_launch_count.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    lib.osb_profile_step.argtypes = [ctypes.c_void_p, _P, ctypes.POINTER(ctypes.c_longlong)]
    lib.osb_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    lib.osb_ipc_import.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.osb_halo_push.argtypes = [ctypes.c_void_p]
    lib.osb_measure_fp64_peak.argtypes = [ctypes.c_int, _P]
    lib.osb_slow_path_count.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    lib.osb_nan_check.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]
    lib.osb_diagnostics.argtypes = [ctypes.c_void_p, _P]
    lib.osb_host_planes_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.osb_host_planes_download.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.osb_host_planes_ready.argtypes = [ctypes.c_void_p]
    lib.osb_host_planes_sync.argtypes = [ctypes.c_void_p]
    lib.osb_staging_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.osb_staging_destroy.argtypes = [ctypes.c_void_p]
    lib.osb_staging_last_error.restype = ctypes.c_char_p
    lib.osb_staging_last_error.argtypes = [ctypes.c_void_p]
    lib.osb_staging_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.osb_staging_feed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.osb_staging_fed.argtypes = [ctypes.c_void_p]
    lib.osb_staging_sync.argtypes = [ctypes.c_void_p]
    lib.osb_staging_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    lib.osb_staging_ipc_import.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.osb_staging_pull.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    _lib = lib
    return lib


class Stage(object):
    """Device copy of a whole host-resident block (planes along the slowest axis) that windows are fed from (hostpipe.py)."""

    def __init__(self, nv, plane_doubles, nplanes, device=-1):
        self.lib = load_library()
        self.nv = nv
        h = ctypes.c_void_p()
        rc = self.lib.osb_staging_create(int(device), int(nv), int(plane_doubles), int(nplanes), ctypes.byref(h))
        if rc:
            raise BackendError('osb_staging_create failed (%d): %s' % (rc, 'out of device memory' if rc == 3 else 'no CUDA device'))
        self.h = h

    def _check(self, rc, what):
        if rc:
            raise BackendError('%s: %s' % (what, self.lib.osb_staging_last_error(self.h).decode()))

    def upload(self, arrays, host_plane0, plane0, nplanes):
        stride = arrays[0].strides[0]
        ptrs = (ctypes.c_void_p * self.nv)(*[a.ctypes.data + host_plane0 * stride for a in arrays])
        self._check(self.lib.osb_staging_upload(self.h, ptrs, int(plane0), int(nplanes)), 'osb_staging_upload')

    def feed(self, sim, stage_plane0, plane0, nplanes):
        sim._check(self.lib.osb_staging_feed(self.h, sim.ctx, int(stage_plane0), int(plane0), int(nplanes)), 'osb_staging_feed')

    def sync(self):
        self._check(self.lib.osb_staging_sync(self.h), 'osb_staging_sync')

    def ipc_export(self):
        buf = ctypes.create_string_buffer(64 * 8)
        n = ctypes.c_int()
        self._check(self.lib.osb_staging_ipc_export(self.h, buf, ctypes.byref(n)), 'osb_staging_ipc_export')
        return buf.raw[:n.value]

    def ipc_import(self, side, handles):
        self._check(self.lib.osb_staging_ipc_import(self.h, int(side), handles, len(handles)), 'osb_staging_ipc_import')

    def pull(self, side, src_plane0, dst_plane0, nplanes):
        self._check(self.lib.osb_staging_pull(self.h, int(side), int(src_plane0), int(dst_plane0), int(nplanes)), 'osb_staging_pull')

    def close(self):
        if self.h:
            self.lib.osb_staging_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_fp64_peak(device=-1):
    v = ctypes.c_double()
    rc = load_library().osb_measure_fp64_peak(device, ctypes.byref(v))
    if rc:
        raise BackendError('osb_measure_fp64_peak failed (%d)' % rc)
    return v.value


class Simulation(object):
    """One solver context on one GPU.  Field arrays are numpy arrays in the reference layout:
    shape (np2+10, np1+10, np0+10)[-ndim:], C order (x fastest), halo 5."""

    HALO = 5

    def __init__(self, plan, device=-1):
        self.plan = _plan.validate(plan)
        self.lib = load_library()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.osb_create(_plan.to_text(plan).encode(), int(device), ctypes.byref(self.ctx))
        if rc:
            raise BackendError('osb_create failed (%d): %s' % (rc, self.lib.osb_last_error(None).decode()))
        self.ndim = plan['ndim']
        self.nv = self.ndim + 2
        self.shape = tuple(int(plan['np'][d]) + 2 * self.HALO for d in reversed(range(self.ndim)))
        self.q_names = ['rho'] + ['rhou%d' % d for d in range(self.ndim)] + ['rhoE']
        # cold-path inputs carried by the plan: metric fields and tabulated Dirichlet states
        for name, arr in plan.get('fields', {}).items():
            self.upload(name, arr)
        for d in range(self.ndim):
            for s in range(2):
                b = plan['bc'][d][s]
                if b['type'] == 'dirichlet_field':
                    t = np.ascontiguousarray(b['table'], dtype=np.float64)
                    self._check(self.lib.osb_upload_face(self.ctx, d, s, t.ctypes.data), 'osb_upload_face')
        # point-wise user kernels (statistics): CUDA source written by opensbli_b200.run.resolve, compiled by NVRTC
        # datasets that only user kernels read (e.g. coordinates evaluated by the cold path): declared and uploaded first
        known = set(self.field_names())
        for name, arr in plan.get('user_fields', {}).items():
            if name not in known:
                self._check(self.lib.osb_create_field(self.ctx, name.encode()), 'osb_create_field')
                self.upload(name, arr)
        for k in plan.get('user_kernels', []):
            self.add_user_kernel(k['source'], k['entry'], k['fields'], k['range'], k['when'], k.get('writes'))

    # -- plumbing
    def _check(self, rc, what):
        if rc:
            raise BackendError('%s failed (%d): %s' % (what, rc, self.lib.osb_last_error(self.ctx).decode()))

    def close(self):
        if self.ctx:
            self.lib.osb_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- fields
    def field_names(self):
        return [self.lib.osb_field_name(self.ctx, i).decode() for i in range(self.lib.osb_num_fields(self.ctx))]

    def upload(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        if a.shape != self.shape:
            raise ValueError('field %s: expected padded shape %s, got %s' % (name, self.shape, a.shape))
        self._check(self.lib.osb_upload(self.ctx, name.encode(), a.ctypes.data), 'osb_upload')

    def download(self, name):
        a = np.empty(self.shape, dtype=np.float64)
        self._check(self.lib.osb_download(self.ctx, name.encode(), a.ctypes.data), 'osb_download')
        return a

    def download_into(self, name, array):
        """device -> caller-owned host array (e.g. pinned memory), no intermediate allocation."""
        if array.shape != self.shape or array.dtype != np.float64 or not array.flags['C_CONTIGUOUS']:
            raise ValueError('download_into needs a C-contiguous float64 array of shape %s' % (self.shape,))
        self._check(self.lib.osb_download(self.ctx, name.encode(), array.ctypes.data), 'osb_download')

    def set_state(self, q):
        for n, a in zip(self.q_names, q):
            self.upload(n, a)

    def get_state(self):
        return [self.download(n) for n in self.q_names]

    def device_ptr(self, name):
        p = ctypes.c_void_p()
        self._check(self.lib.osb_device_ptr(self.ctx, name.encode(), ctypes.byref(p)), 'osb_device_ptr')
        return p.value

    def set_const(self, name, value):
        self._check(self.lib.osb_set_const_f64(self.ctx, name.encode(), float(value)), 'osb_set_const_f64')

    def add_user_kernel(self, source, entry, fields, rng, when, writes=None):
        """when: 'iteration_end' (launched at the end of every time step) | 'after_loop' (run_user_kernels('after_loop')).
        `writes`: the names among `fields` the kernel assigns (created zero-initialised when new); a name it only reads
        must already exist on the device.  writes=None treats every name as written (the pre-existing call shape)."""
        r = (ctypes.c_int * 6)(*(list(rng) + [0, 1] * 3)[:6])
        w = self._when_code(when)
        names = [('+' + f) if (writes is None or f in writes) else f for f in fields]
        self._check(self.lib.osb_add_user_kernel(self.ctx, source.encode(), entry.encode(), ','.join(names).encode(), r, w), 'osb_add_user_kernel')

    @staticmethod
    def _when_code(when):
        """'iteration_end' | 'after_loop' | 'bc_<dir>_<side>' (the boundary kernel of a face whose plan entry is 'generic') |
        generic path: 'iteration_start' | 'stage' (every RK stage) | 'stage_<s>' (stage s only)"""
        if when.startswith('bc_'):
            _, d, s = when.split('_')
            return 100 + 2 * int(d) + int(s)
        if when.startswith('stage_'):
            return 210 + int(when.split('_')[1])
        return {'iteration_end': 0, 'after_loop': 1, 'iteration_start': 200, 'stage': 209}[when]

    def run_user_kernels(self, when='after_loop'):
        self._check(self.lib.osb_run_user_kernels(self.ctx, self._when_code(when)), 'osb_run_user_kernels')
        self.sync()

    def read_point(self, name, i, j=0, k=0):
        v = ctypes.c_double()
        self._check(self.lib.osb_read_point(self.ctx, name.encode(), int(i), int(j), int(k), ctypes.byref(v)), 'osb_read_point')
        return v.value

    def set_iteration(self, iteration):
        """iteration number seen by time-dependent source terms (restart)"""
        self._check(self.lib.osb_set_iteration(self.ctx, int(iteration)), 'osb_set_iteration')

    def get_iteration(self):
        return int(self.lib.osb_get_iteration(self.ctx))

    # -- the time loop
    def step(self, nsteps=1, sync=True):
        self._check(self.lib.osb_step(self.ctx, int(nsteps)), 'osb_step')
        if sync:
            self.sync()

    def step_begin(self):
        self._check(self.lib.osb_step_begin(self.ctx), 'osb_step_begin')

    def stage(self, s):
        self._check(self.lib.osb_stage(self.ctx, int(s)), 'osb_stage')

    def halo_push(self):
        self._check(self.lib.osb_halo_push(self.ctx), 'osb_halo_push')

    def ipc_export(self):
        buf = ctypes.create_string_buffer(64 * 16)
        n = ctypes.c_int()
        self._check(self.lib.osb_ipc_export(self.ctx, buf, ctypes.byref(n)), 'osb_ipc_export')
        return buf.raw[:n.value]

    def ipc_import(self, side, handles):
        self._check(self.lib.osb_ipc_import(self.ctx, int(side), handles, len(handles)), 'osb_ipc_import')

    def sync(self):
        self._check(self.lib.osb_sync(self.ctx), 'osb_sync')

    def apply_bcs(self):
        self._check(self.lib.osb_apply_bcs(self.ctx), 'osb_apply_bcs')
        self.sync()

    def residual(self):
        self._check(self.lib.osb_residual(self.ctx), 'osb_residual')
        self.sync()
        return [self.download('Residual%d' % m) for m in range(self.nv)]

    def step_timed(self, nsteps=1):
        ms = ctypes.c_double()
        self._check(self.lib.osb_step_timed(self.ctx, int(nsteps), ctypes.byref(ms)), 'osb_step_timed')
        return ms.value

    def timer_start(self):
        self._check(self.lib.osb_timer_start(self.ctx), 'osb_timer_start')

    def timer_stop(self):
        ms = ctypes.c_double()
        self._check(self.lib.osb_timer_stop(self.ctx, ctypes.byref(ms)), 'osb_timer_stop')
        return ms.value

    def advance_host(self, q_in, q_out, nsteps=1):
        """End-to-end call with HOST buffers (lists of nv padded float64 arrays, ideally pinned)."""
        pin = (ctypes.c_void_p * self.nv)(*[a.ctypes.data for a in q_in])
        pout = (ctypes.c_void_p * self.nv)(*[a.ctypes.data for a in q_out])
        ms = ctypes.c_double()
        self._check(self.lib.osb_advance_host(self.ctx, pin, pout, int(nsteps), ctypes.byref(ms)), 'osb_advance_host')
        return ms.value

    # -- window pipeline over a host-resident block (hostpipe.py)
    def planes_upload(self, arrays, host_plane0, plane0, nplanes):
        """Asynchronous copy of planes [host_plane0, +nplanes) of the padded host arrays into local planes [plane0, +nplanes)."""
        stride = arrays[0].strides[0]
        ptrs = (ctypes.c_void_p * self.nv)(*[a.ctypes.data + host_plane0 * stride for a in arrays])
        self._check(self.lib.osb_host_planes_upload(self.ctx, ptrs, int(plane0), int(nplanes)), 'osb_host_planes_upload')

    def planes_ready(self):
        self._check(self.lib.osb_host_planes_ready(self.ctx), 'osb_host_planes_ready')

    def planes_download(self, arrays, host_plane0, plane0, nplanes):
        stride = arrays[0].strides[0]
        ptrs = (ctypes.c_void_p * self.nv)(*[a.ctypes.data + host_plane0 * stride for a in arrays])
        self._check(self.lib.osb_host_planes_download(self.ctx, ptrs, int(plane0), int(nplanes)), 'osb_host_planes_download')

    def planes_sync(self):
        self._check(self.lib.osb_host_planes_sync(self.ctx), 'osb_host_planes_sync')

    def stage_fed(self):
        self._check(self.lib.osb_staging_fed(self.ctx), 'osb_staging_fed')

    # -- in-loop diagnostics
    def nan_check(self, name='rho'):
        """number of non-finite values of a dataset over the interior (the reference's ops_NaNcheck)"""
        n = ctypes.c_longlong()
        self._check(self.lib.osb_nan_check(self.ctx, name.encode(), ctypes.byref(n)), 'osb_nan_check')
        return n.value

    DIAG = ('sum_rho', 'sum_ke', 'sum_enstrophy', 'sum_rhoE', 'max_mach', 'nonfinite')

    def diagnostics(self):
        """interior sums of this block: mass, kinetic energy 1/2 rho |u|^2, enstrophy 1/2 rho |curl u|^2, total energy;
        max Mach number; number of points with a non-finite rho / rhoE (halos as the last step left them)"""
        v = (ctypes.c_double * len(self.DIAG))()
        self._check(self.lib.osb_diagnostics(self.ctx, v), 'osb_diagnostics')
        return dict(zip(self.DIAG, list(v)))

    def slow_path_count(self, enable=True):
        """arm (or disarm) the counter of TENO5 waves that take the full cut-off path; returns the count so far"""
        n = ctypes.c_longlong()
        self._check(self.lib.osb_slow_path_count(self.ctx, 1 if enable else 0, ctypes.byref(n)), 'osb_slow_path_count')
        return n.value

    def launch_count(self):
        n = ctypes.c_longlong()
        self._check(self.lib.osb_launch_count(self.ctx, ctypes.byref(n)), 'osb_launch_count')
        return n.value

    def profile_step(self):
        ms = (ctypes.c_double * len(FAMILIES))()
        n = (ctypes.c_longlong * len(FAMILIES))()
        self._check(self.lib.osb_profile_step(self.ctx, ms, n), 'osb_profile_step')
        return {f: {'ms': ms[i], 'launches': n[i]} for i, f in enumerate(FAMILIES)}
