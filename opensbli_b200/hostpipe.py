"""Window pipeline: advance a block that lives in HOST memory with the three engines of the GPU busy at the same time.

The reference moves whole blocks between host and device (ops_fetch / HDF5, opensbli/core/io_hdf5.py:99-127) and has nothing
to cite here; this module is how the end-to-end call of the B200 back end (`Simulation.advance_host`: upload, step, download,
one after the other) is made to overlap its copies with the sweeps.

The block is cut into WINDOWS along its slowest axis.  A window carries `guard` extra planes on each cut side, its cut faces
have the 'open' boundary type (nothing is filled there), and it is advanced as an independent small block:

    upload   the host arrays cross PCIe ONCE, in plane order, into a staging copy of the block on the device; the window takes
             planes [z0 - guard - halo, z1 + guard + halo) (periodic wrap) from there with device-to-device copies, as soon as
             the upload has got that far
    step     nsteps iterations on its compute stream -- every RK stage invalidates `depth` more planes next to an open face
             (depth = the reach of the scheme's stencil along the axis), so with guard = depth * (stages * nsteps - 1) the
             planes [z0, z1) end up exactly as in a run of the whole block (same kernels, same arithmetic per point)
    download planes [z0, z1) into the host arrays, on the context's download stream

A few contexts take the windows in turn, so window k+1 is uploaded and window k-1 downloaded while window k is swept.  The
price is the redundant sweep of the guard planes; the gain is that a step costs max(copy, sweep) instead of their sum.

Scope: periodic along the slab axis, no per-point plan arrays (metrics, tables) and no user kernels -- the periodic-box
configurations the bench line is quoted on.  Anything else raises PlanError; callers fall back to `advance_host`.
"""
import copy

from . import plan as _plan
from .decomp import scheme_halos, slab_axis

HALO = 5


def stencil_depth(plan):
    """Planes next to an open face that one RK stage invalidates = the farthest plane a residual reads along the axis: 2 for
    Central(4) (first, second and mixed derivatives), 3 for WENO5 / TENO5 / TENO6 (interfaces i -+ 1/2 read points i-3 .. i+3;
    the fourth halo plane the reference exchanges on one side, weno.py:17-32, is not read by a residual)."""
    return scheme_halos(plan)[0]


def guard_planes(plan, nsteps=1):
    return stencil_depth(plan) * (len(plan['rk_a']) * int(nsteps) - 1)


def ramp(n):
    """Window sizes that ramp up and down: the pipeline cannot sweep before its first window is uploaded nor finish before its
    last one is downloaded, so the ends are small (n/16, n/8 ... 3n/16, n/8) and the middle large (n/4: less guard-plane work)."""
    if n % 16 or n < 256:
        raise _plan.PlanError('ramp windows need a plane count that is a multiple of 16 and at least 256')
    u = n // 16
    return [u, 2 * u, 4 * u, 4 * u, 3 * u, 2 * u]


def windows(n, chunk):
    """[(z0, z1)] covering [0, n): `chunk` is a list of window sizes that add up to n, or one size -- then the windows have exactly
    `chunk` planes and the last one is shifted back to end at n (its overlap with the one before is computed twice, to the same
    values)."""
    if isinstance(chunk, (list, tuple)):
        if sum(chunk) != n or min(chunk) < 1:
            raise _plan.PlanError('window sizes %s do not add up to %d planes' % (list(chunk), n))
        out, z = [], 0
        for c in chunk:
            out.append((z, z + c))
            z += c
        return out
    if chunk >= n:
        return [(0, n)]
    out = []
    z = 0
    while z + chunk < n:
        out.append((z, z + chunk))
        z += chunk
    out.append((n - chunk, n))
    return out


def window_plan(plan, chunk, nsteps=1):
    """Plan of one window (all windows share it): `chunk` + 2 guard planes along the slab axis, 'open' cut faces."""
    plan = _plan.validate(copy.deepcopy(plan))
    ax = slab_axis(plan)
    n = plan['np'][ax]
    if plan['bc'][ax][0]['type'] != 'periodic' or plan['bc'][ax][1]['type'] != 'periodic':
        raise _plan.PlanError('window pipeline: the slab axis must be periodic')
    if plan.get('fields') or plan.get('user_fields') or plan.get('user_kernels'):
        raise _plan.PlanError('window pipeline: per-point plan arrays and user kernels are not windowed')
    for d in range(plan['ndim']):
        for s in range(2):
            if plan['bc'][d][s].get('table') is not None:
                raise _plan.PlanError('window pipeline: tabulated boundary states are not windowed')
    g = guard_planes(plan, nsteps)
    if chunk >= n:
        raise _plan.PlanError('window pipeline: one window would hold the whole block (use advance_host)')
    if chunk < 1:
        raise _plan.PlanError('window pipeline: empty window')
    p = copy.deepcopy(plan)
    p['np'][ax] = chunk + 2 * g
    p['bc'][ax][0] = {'type': 'open'}
    p['bc'][ax][1] = {'type': 'open'}
    p.pop('io', None)
    return _plan.validate(p), g


def wrapped_runs(first, count, n):
    """Split global interior planes [first, first + count) (may lie outside [0, n): periodic images) into runs of
    consecutive planes inside [0, n): [(global plane of the run, offset within the request, length)]."""
    runs = []
    done = 0
    while done < count:
        g = (first + done) % n
        length = min(count - done, n - g)
        runs.append((g, done, length))
        done += length
    return runs


class HostPipeline(object):
    """advance(q_in, q_out): q_out <- nsteps iterations of q_in, both lists of padded host arrays (PINNED for the copies to
    overlap; reference layout, slab axis first).  Same result, bit for bit, as Simulation.advance_host on the whole block in
    every cell a kernel defines: the grid points and the halo cells within the scheme's halo depth (the last boundary-condition
    pass).  Halo cells beyond that depth are read by nothing and are left unspecified, as they are by the whole-block call."""

    def __init__(self, plan, chunk=64, nsteps=1, device=-1, contexts=3, factory=None, stage_factory=None, exchange=None, slab=None):
        """factory(window_plan) -> window solver, stage_factory(nv, plane_doubles, nplanes) -> staging copy; defaults: a
        Simulation and a Stage on `device` (the CPU tests of the windowing drive the oracle through the same calls).
        Slab-decomposed blocks (one rank per GPU): `slab` = (first plane, number of planes) of this rank and
        `exchange(stage, G, n)` fills the staging copy's G = guard + halo planes below and above the slab with the neighbours'
        planes (their staged bottom / top planes; see DistributedHostPipeline) -- after that the rank's windows are as
        independent of the other ranks as they are of each other: no per-stage halo exchange at all."""
        if factory is None:
            from .runtime import Simulation, Stage
            factory = lambda p: Simulation(p, device=device)
            stage_factory = stage_factory or (lambda nv, plane, n: Stage(nv, plane, n, device=device))
        self.plan_global = _plan.validate(copy.deepcopy(plan))
        self.ax = slab_axis(self.plan_global)
        self.exchange = exchange
        self.n = self.plan_global['np'][self.ax] if slab is None else int(slab[1])
        self.nsteps = int(nsteps)
        if chunk == 'ramp':
            chunk = ramp(self.n)
        self.chunk = list(chunk) if isinstance(chunk, (list, tuple)) else int(chunk)
        self.windows = windows(self.n, self.chunk)
        # one window plan per window size; windows of a size take turns on that size's contexts (at most `contexts` each)
        self.plans, self.pools, self.guard = {}, {}, None
        for z0, z1 in self.windows:
            size = z1 - z0
            if size not in self.plans:
                self.plans[size], self.guard = window_plan(self.plan_global, size, self.nsteps)
                self.pools[size] = []
        count = {}
        for z0, z1 in self.windows:
            count[z1 - z0] = count.get(z1 - z0, 0) + 1
        for size, cnt in count.items():
            self.pools[size] = [factory(self.plans[size]) for _ in range(min(int(contexts), cnt))]
        self.sims = [sim for pool in self.pools.values() for sim in pool]
        self.plan = self.plans[self.windows[0][1] - self.windows[0][0]]
        self.hm, self.hp = scheme_halos(self.plan_global)
        self.nv = self.plan_global['ndim'] + 2
        self.plane = 1
        for d in range(self.plan_global['ndim']):
            if d != self.ax:
                self.plane *= self.plan_global['np'][d] + 2 * HALO
        # single block: the staging copy holds the n planes, guard planes are periodic images; slab of a decomposed block: it
        # holds G more planes on either side, which the neighbours fill
        self.G = 0 if exchange is None else self.guard + HALO
        if exchange is not None and self.n < self.G:
            raise _plan.PlanError('window pipeline: slab of %d planes is thinner than guard + halo' % self.n)
        self.stage = stage_factory(self.nv, self.plane, self.n + 2 * self.G)
        self.launches = 0

    def close(self):
        for s in self.sims:
            s.close()
        self.sims = []
        if self.stage is not None:
            self.stage.close()
            self.stage = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def bytes_per_call(self):
        """(H2D, D2H) bytes of one advance() call, counted from the plane copies it enqueues."""
        down = sum(z1 - z0 for z0, z1 in self.windows) + (2 * self.hm if self.exchange is None else 0)
        return self.n * self.plane * 8 * self.nv, down * self.plane * 8 * self.nv

    def advance(self, q_in, q_out):
        n, g, h = self.n, self.guard, HALO
        if any(a is b or a.ctypes.data == b.ctypes.data for a, b in zip(q_in, q_out)):
            raise ValueError('window pipeline: q_out must not alias q_in (windows read guard planes other windows write)')
        l0 = sum(s.launch_count() for s in self.sims)
        staged = [False] * n                         # planes of the block already on their way to the staging copy
        G = self.G
        if self.exchange is not None:
            # the planes the neighbours need go first; once they have landed the ranks swap them device to device, then the
            # rest of the slab follows in window order while the first windows are already swept
            for a, b in ((0, min(G, n)), (max(n - G, min(G, n)), n)):
                if b > a:
                    self.stage.upload(q_in, h + a, G + a, b - a)
                    staged[a:b] = [True] * (b - a)
            self.stage.sync()
            self.exchange(self.stage, G, n)
        turn = {}
        for w, (z0, z1) in enumerate(self.windows):
            size = z1 - z0
            pool = self.pools[size]
            sim = pool[turn.get(size, 0) % len(pool)]
            turn[size] = turn.get(size, 0) + 1
            # padded local plane j of the window holds plane z0 - g - h + j of the block (periodic image)
            if self.exchange is None:
                runs = wrapped_runs(z0 - g - h, size + 2 * g + 2 * h, n)
            else:                                    # planes below 0 / above n are the neighbours', already in the staging copy
                lo_, hi_ = z0 - g - h, z1 + g + h
                runs = [(p0, p0 - lo_, p1 - p0) for p0, p1 in ((lo_, min(hi_, 0)), (max(lo_, 0), min(hi_, n)), (max(lo_, n), hi_)) if p1 > p0]
            for gp, _, length in runs:
                a = gp
                while 0 <= a < min(gp + length, n):  # upload the planes of this run that no earlier window asked for
                    if staged[a]:
                        a += 1
                        continue
                    b = a
                    while b < min(gp + length, n) and not staged[b]:
                        staged[b] = True
                        b += 1
                    self.stage.upload(q_in, h + a, G + a, b - a)
                    a = b
            for gp, off, length in runs:
                self.stage.feed(sim, G + gp, off, length)
            sim.stage_fed()
            sim.step(self.nsteps, sync=False)
            sim.planes_download(q_out, h + z0, h + g, z1 - z0)
            # halo planes of the whole block as its periodic BC leaves them after the last stage: [-hm, 0) <- [n-hm, n),
            # [n, n+hm) <- [0, hm)  (periodic.py:42-56: side 0 copies hm planes up; side 1 copies hp planes starting at n-hm
            # down to -hm, the last of which lands on plane 0 and is plane 0); they come from the windows that own those planes
            if self.exchange is not None:
                continue                             # slab of a decomposed block: its halo planes belong to the neighbours
            lo, hi = max(z0, n - self.hm), min(z1, n)
            if lo < hi:
                sim.planes_download(q_out, h + lo - n, h + g + lo - z0, hi - lo)
            lo, hi = max(z0, 0), min(z1, self.hm)
            if lo < hi:
                sim.planes_download(q_out, h + n + lo, h + g + lo - z0, hi - lo)
        for sim in self.sims:
            sim.planes_sync()
        self.stage.sync()
        self.launches = sum(s.launch_count() for s in self.sims) - l0
        return self.launches


class DistributedHostPipeline(object):
    """The window pipeline on a slab-decomposed block, one rank per GPU (`dist` = torch.distributed, initialised).  Every rank
    advances ITS slab of the host-resident block window by window; the guard + halo planes its first and last windows need
    from the neighbouring slabs are pulled out of the neighbours' staging copies over NVLink right after those planes were
    uploaded.  The windows then run without any per-stage halo exchange -- the guard planes absorb it -- so the ranks do not
    wait for each other during the sweeps.  q_in / q_out: this rank's padded slab arrays (decomp.local_extent); the grid points
    of q_out equal those of the decomposed in-HBM run bit for bit; its halo planes along the slab axis are not written."""

    def __init__(self, plan_global, dist, device, chunk=64, nsteps=1, contexts=3, factory=None, stage_factory=None, exchange=None):
        from .decomp import local_extent, neighbours
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        plan_global = _plan.validate(copy.deepcopy(plan_global))
        self.offset, self.nloc = local_extent(plan_global, self.rank, self.world)
        self.low, self.high = neighbours(plan_global, self.rank, self.world)
        if self.world > 1 and (self.low is None or self.high is None):
            raise _plan.PlanError('window pipeline: the slab axis must be periodic')
        self.n_low = local_extent(plan_global, self.low, self.world)[1] if self.world > 1 else 0
        self.pipe = HostPipeline(plan_global, chunk=chunk, nsteps=nsteps, device=device, contexts=contexts, factory=factory,
                                 stage_factory=stage_factory, slab=(self.offset, self.nloc) if self.world > 1 else None,
                                 exchange=(exchange or self._exchange) if self.world > 1 else None)
        if self.world > 1 and exchange is None:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.pipe.stage.ipc_export())
            self.pipe.stage.ipc_import(0, handles[self.low])
            self.pipe.stage.ipc_import(1, handles[self.high])
            self._barrier()

    def _barrier(self):
        if self.dist.get_backend() == 'nccl':
            import torch
            self.dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            self.dist.barrier()

    def _exchange(self, stage, G, n):
        self._barrier()                                   # every rank's boundary planes are in its staging copy
        stage.pull(0, self.n_low, 0, G)                   # low neighbour's top G planes (its staging planes [n_low, n_low + G))
        stage.pull(1, G, G + n, G)                        # high neighbour's bottom G planes (its staging planes [G, 2G))
        stage.sync()
        self._barrier()                                   # nobody re-uploads while a neighbour is still reading

    def advance(self, q_in, q_out):
        return self.pipe.advance(q_in, q_out)

    def bytes_per_call(self):
        return self.pipe.bytes_per_call()

    def close(self):
        self.pipe.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
