"""Slab decomposition of the single structured block over the GPUs of one node.

The reference delegates decomposition to OPS (`ops_partition("")`, opensbli/code_generation/opsc.py:540-545); here
the block is cut into equal slabs along its slowest axis, one rank (process) per GPU.  Faces along that axis are
contiguous planes, so a halo exchange is a plain plane copy.  Only the conserved arrays are exchanged -- primitives are
recomputed in the halos, as the reference does (constituent-relation kernels run over grid + halos).

Exchange protocol per RK stage (the place where the reference has its `ops_halo_transfer`s, algorithm.py:440-442), all
ordered on each rank's CUDA stream by flag words the neighbours write through peer pointers -- no host synchronisation:
    out-of-place stage kernels (3-D periodic-box paths): the kernel that finishes the stage also stores the new boundary
    planes into the neighbours' Residual-role buffers over NVLink (CUDA IPC peer pointers) -> "pushed" handshake -> the
    buffers exchange roles -> rank-local BCs;
    in-place paths (2-D, general path with walls / metrics): kernels that read the halos -> "read done" handshake -> RK
    update -> plane copies into the neighbours' halos -> "pushed" handshake -> rank-local BCs (they also rewrite the x/y-halo
    parts of the received planes, so they must come after the neighbours' copies).
Per-point arrays of the general path (metrics, source amplitudes) and tabulated Dirichlet states are cut with the slab;
physical boundary conditions and their one-sided closures stay with the ranks that own the face.
The pure functions in this module (extents, neighbours, plane indices) are shared by the GPU driver and by the
CPU (gloo) tests of the N>1 path.
"""
import copy

from . import plan as _plan


def scheme_halos(plan):
    """(halo_m, halo_p) of the spatial scheme: WENO/TENO 3/4 (weno.py:17-32, teno.py:18-36), central 2/2."""
    if plan.get('halos'):
        return tuple(plan['halos'])
    return (2, 2) if plan['conv'] == 'central' else (3, 4)


def slab_axis(plan):
    return plan['ndim'] - 1


def local_extent(plan, rank, world):
    """(first plane, number of planes) of a rank's slab; when the block does not divide evenly the first n %% world ranks hold
    one plane more"""
    ax = slab_axis(plan)
    n = plan['np'][ax]
    base, rem = divmod(n, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def neighbours(plan, rank, world):
    """(low, high) neighbour ranks along the slab axis, None at a physical (non-periodic) boundary."""
    ax = slab_axis(plan)
    periodic = plan['bc'][ax][0]['type'] == 'periodic'
    low = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    high = rank + 1 if rank < world - 1 else (0 if periodic else None)
    if world == 1:
        return None, None
    return low, high


def local_plan(plan, rank, world):
    """Plan of one rank: its slab of the block; faces shared with a neighbour become 'exchange' BCs."""
    if world == 1:
        return copy.deepcopy(plan)
    if plan.get('conv') == 'generic':
        raise _plan.PlanError('programs on the generic path (every loop a run-time compiled kernel) run on one GPU: their boundary '
                              'kernels and periodic copies are part of the kernel lists and are not cut into slabs')
    ax = slab_axis(plan)
    _, loc = local_extent(plan, rank, world)
    hm, hp = scheme_halos(plan)
    if min(local_extent(plan, r, world)[1] for r in range(world)) < max(hm, hp):
        raise _plan.PlanError('slab of %d planes is thinner than the halo depth' % loc)
    p = copy.deepcopy(plan)
    p['np'][ax] = loc
    low, high = neighbours(plan, rank, world)
    if low is not None:
        p['bc'][ax][0] = {'type': 'exchange'}
    if high is not None:
        p['bc'][ax][1] = {'type': 'exchange'}
    # general path: per-point arrays (metrics, mass-source amplitude) and tabulated Dirichlet states follow the slab
    import numpy as np
    nd, h = plan['ndim'], 5
    k0, _ = local_extent(plan, rank, world)
    if plan.get('fields'):
        p['fields'] = {n: np.ascontiguousarray(np.asarray(a)[k0:k0 + loc + 2 * h]) for n, a in plan['fields'].items()}
    if plan.get('user_fields'):
        p['user_fields'] = {n: np.ascontiguousarray(np.asarray(a)[k0:k0 + loc + 2 * h]) for n, a in plan['user_fields'].items()}
    for d in range(nd):
        for s in range(2):
            b = plan['bc'][d][s]
            if b.get('table') is not None and d != ax:
                pd = [plan['np'][e] + 2 * h for e in range(nd) if e != d]          # tangential padded extents, x fastest
                t = np.asarray(b['table']).reshape((-1,) + tuple(reversed(pd)))
                p['bc'][d][s]['table'] = np.ascontiguousarray(t[:, k0:k0 + loc + 2 * h]).reshape(t.shape[0], -1)
    # point-wise user kernels cover the rank's own slab; run-time compiled boundary kernels ('bc_<dir>_<side>') follow their face
    kept = []
    n_ax = plan['np'][ax]
    for uk in p.get('user_kernels', []):
        r = list(uk['range']) + [0, 1] * (3 - nd)
        if uk['when'].startswith('bc_'):
            _, d, sd = uk['when'].split('_')
            d, sd = int(d), int(sd)
            if d == ax:
                if p['bc'][ax][sd]['type'] == 'exchange':        # an interior cut: the neighbour's planes take the face's place
                    continue
                if sd == 1:
                    r[2 * ax] += loc - n_ax
                    r[2 * ax + 1] += loc - n_ax
            else:                                                # tangential range [-halo, np + halo) -> [-halo, nloc + halo)
                if r[2 * ax] > 0 or r[2 * ax + 1] < n_ax:
                    raise _plan.PlanError('boundary kernel %s does not cover the whole slab axis: cannot be decomposed' % uk.get('name'))
                r[2 * ax + 1] += loc - n_ax
            uk['source'] = '#define OSB_GOFF%d %d\n' % (ax, k0) + uk['source']
        else:
            if r[2 * ax] != 0 or r[2 * ax + 1] != n_ax:
                raise _plan.PlanError('user kernel %s does not cover the whole slab axis: cannot be decomposed' % uk.get('name'))
            r[2 * ax + 1] = loc
        uk['range'] = r[:2 * nd]
        kept.append(uk)
    if 'user_kernels' in p:
        p['user_kernels'] = kept
    return _plan.validate(p)


def push_planes(plan_local, halo=5):
    """Plane index ranges (in the padded array, along the slab axis) of the two pushes of one rank:
    'up'   : my top hm planes    [np-hm, np) -> high neighbour's low halo  [-hm, 0)
    'down' : my bottom hp planes [0, hp)     -> low neighbour's high halo  [np_lo, np_lo+hp)   (np_lo: the low neighbour's thickness)"""
    ax = slab_axis(plan_local)
    n = plan_local['np'][ax]
    hm, hp = scheme_halos(plan_local)
    return {'up': ((halo + n - hm, halo + n), (halo - hm, halo)),
            'down': ((halo, halo + hp), (halo + n, halo + n + hp))}


class DistributedSimulation(object):
    """One rank of a slab-decomposed run on GPUs: a Simulation plus the per-stage halo exchange.
    `dist` is torch.distributed (already initialised, one rank per GPU)."""

    def __init__(self, plan_global, dist, device):
        from .runtime import Simulation
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.plan_global = _plan.validate(plan_global)
        self.plan = local_plan(plan_global, self.rank, self.world)
        self.offset, self.nloc = local_extent(plan_global, self.rank, self.world)
        self.sim = Simulation(self.plan, device=device)
        self.device = device
        self.nstages = len(self.plan['rk_a'])
        self.low, self.high = neighbours(plan_global, self.rank, self.world)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.sim.ipc_export())
            if self.low is not None:
                self.sim.ipc_import(0, handles[self.low])
            if self.high is not None:
                self.sim.ipc_import(1, handles[self.high])
            self.barrier()

    def barrier(self):
        import torch
        if self.world > 1:
            self.sim.sync()
            if self.dist.get_backend() == 'nccl':
                self.dist.barrier(device_ids=[torch.cuda.current_device()])
            else:
                self.dist.barrier()

    def set_state(self, q):
        """Upload this rank's slab (padded arrays).  Collective: no rank may start stepping -- and storing boundary planes
        into its neighbours' halos -- before every rank's upload, which also writes those halos, has landed."""
        self.sim.set_state(q)
        self.barrier()

    def get_state(self):
        return self.sim.get_state()

    def step(self, nsteps=1, sync=True):
        """Enqueue nsteps iterations; the neighbour exchange is part of the stream-ordered stage sequence."""
        self.sim.step(nsteps, sync=sync)

    def close(self):
        self.sim.close()
