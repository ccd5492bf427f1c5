#!/usr/bin/env python
"""bench.py -- grid-point updates/s of the OpenSBLI solver hot path (fp64, TENO5 Taylor-Green vortex).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

One "step" = one full RK3 time step (3 stages) of all 5 conserved variables on every grid point of the block,
i.e. exactly the body of the reference's timed loop (algorithm.py:301-327, 440-474): BCs/exchanges, constituent
relations, TENO5 characteristic flux sweeps in 3 directions, viscous terms, RK update.
Workload (N=1): BASELINE.json configs[2] = TGV Re=1600, TENO5 + StoreSome(4) viscous + RungeKuttaLS(3), 512^3 fp64.
N>1: weak scaling, 512^3 points per GPU, slab decomposition (N=8 -> 1024^3), halo exchange by peer stores over NVLink.

Prints ONE JSON line (see the contract in the task statement): value = whole-job updates/s with the state resident in
HBM; e2e = the same through the C-ABI calls with HOST buffers (pinned H2D of the state + step + D2H every step; at N=1 the
block is advanced window by window so that copies and sweeps overlap, opensbli_b200/hostpipe.py, and the plain
upload-step-download call is reported beside it as e2e.unpipelined);
roofline = dominant kernel family (flux sweeps) against the measured FP64-pipe peak (and HBM for context);
cpu_baseline = the reference's own generated C (oracle/_ref/tgv_teno5/ref_omp) on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'grid-point updates/s (fp64, TENO5 TGV)'
UNIT = 'updates/s'
ALG_BYTES_PER_UPDATE = 480.0          # 3 stages x (5 q + 5 RK regs) x (read + write) x 8 B
ALG_FLOP_PER_POINT_FLUX_SWEEP = 2836.0 + 50.0 / 3.0   # reference count_ops: one LLFTeno_reconstruction_d loop + its share of the Residual loop
ALG_FLOP_PER_UPDATE = 27.1e3          # 3 x 9038 (reference's own operation count, opsc.py:411-418)
# dram__bytes_read.sum + dram__bytes_write.sum per flux-sweep launch at 512^3 from the `ncu --set full` capture
# profiles/r02_final_ncu_top_kernels_512.md: z 11.04 GB, x 16.34 GB, y 16.44 GB -> mean of the three launches of a stage
# (algorithmic: 5 q read + 5 residual read + 5 residual write = 16.1 GB for the accumulating sweeps, 10.7 GB for the first)
NCU_DRAM_BYTES_PER_FLUX_LAUNCH_512 = 14.61e9
NCU_DRAM_BYTES_PER_CENTRAL_LAUNCH_512 = 22.58e9     # k_central3d_fused<.., TMA> at 512^3: 11.81 GB read + 10.77 GB written (algorithmic 21.5 GB)
# sm__pipe_fp64_cycles_active (fraction of peak) of the z, x, y sweep kernels in the same capture: what the hardware says beside
# the algorithmic fraction (the kernels execute ~820 FP64 instructions per interface where the reference's count_ops has 2836)
NCU_PIPE_FP64_BUSY_512 = {'z': 0.634, 'x': 0.642, 'y': 0.599}
LS3 = dict(rk='ls', rk_a=[0.0, -5.0 / 9.0, -153.0 / 128.0], rk_b=[1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0])


SBLI3 = dict(rk='sbli', rk_a=[1.0 / 4.0, 3.0 / 20.0, 3.0 / 5.0], rk_b=[2.0 / 3.0, 5.0 / 12.0, 3.0 / 5.0])


def tgv_plan(np3, workload='teno5'):
    dl = 2 * math.pi / 512 if max(np3) > 512 else 2 * math.pi / np3[0]
    per = [[dict(type='periodic'), dict(type='periodic')] for _ in range(3)]
    consts = dict(gama=1.4, Minf=0.1, Re=1600.0, Pr=0.71, dt=0.003385 * 64 * dl / (2 * math.pi), eps=1e-16, TENO_CT=1e-6)
    if workload == 'central4':      # BASELINE configs[1]: the shipped taylor_green_vortex app, Central(4) + RungeKutta(3)
        return dict(ndim=3, np=list(np3), delta=[dl] * 3, conv='central', order=4, averaging='roe', viscous=True,
                    constants=consts, bc=per, **SBLI3)
    return dict(ndim=3, np=list(np3), delta=[dl] * 3, conv='teno', order=5, averaging='roe', viscous=True,
                constants=consts, bc=per, **LS3)


def global_grid(ngpus, size):
    """weak scaling: size^3 points per GPU; N=2: (s,s,2s) N=4: (s,2s,2s) N=8: (2s,2s,2s)."""
    f = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[ngpus]
    return [size * f[0], size * f[1], size * f[2]]


def tgv_state_into(bufs, plan, k0, nk, halo=5):
    """Analytic TGV initial condition (apps/taylor_green_vortex/taylor_green_vortex.py:87-101) for planes
    [k0, k0+nk) of the block, written into the padded numpy arrays bufs (interior only; BCs fill the halos)."""
    import numpy as np
    n0, n1 = plan['np'][0], plan['np'][1]
    d = plan['delta']
    x = (np.arange(n0) * d[0])[None, None, :]
    y = (np.arange(n1) * d[1])[None, :, None]
    z = ((k0 + np.arange(nk)) * d[2])[:, None, None]
    g, M = 1.4, 0.1
    s = (slice(halo, halo + nk), slice(halo, halo + n1), slice(halo, halo + n0))
    u0 = np.sin(x) * np.cos(y) * np.cos(z)
    u1 = -np.cos(x) * np.sin(y) * np.cos(z)
    p = 1.0 / (g * M * M) + (1.0 / 16.0) * (np.cos(2.0 * x) + np.cos(2.0 * y)) * (2.0 + np.cos(2.0 * z))
    r = g * M * M * p
    for b in bufs:
        b[...] = 0.0
    bufs[0][s] = r
    bufs[1][s] = r * u0
    bufs[2][s] = r * u1
    bufs[3][s] = 0.0
    bufs[4][s] = p / (g - 1.0) + 0.5 * r * (u0 * u0 + u1 * u1)


def pipeline_fits(plan, nloc, chunk, contexts, torch):
    """the window contexts (about 26 arrays each) and the staging copy of the pipelined end-to-end leg next to what is already
    on the device: skip the leg rather than run the device out of memory (large strong-scaling slabs)"""
    plane = 8.0
    for n in plan['np'][:-1]:
        plane *= n + 10
    need = (contexts * (chunk + 2 * 8 + 10) * 26 + (nloc + 32) * 5) * plane
    free, _ = torch.cuda.mem_get_info()
    return need < 0.8 * free, need, free


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_sample_size(cores):
    """Bounded sample of the same workload: TGV TENO5 on n^3 points for 2 steps, ~10-20 s of CPU work."""
    rate = 4.0e4 * max(cores, 1)              # reference generated C: ~4-6e4 updates/s per core
    n = int(round((rate * 12.0 / 2.0) ** (1.0 / 3.0) / 8.0)) * 8
    return max(32, min(n, 256))


def run_reference_cpu(steps, n, threads):
    """Times the reference's own generated C (OPS-OpenMP stand-in) with its own timer (time loop only)."""
    exe = os.path.join(REPO, 'oracle', '_ref', 'tgv_teno5', 'ref_omp')
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), block0np0=str(n), block0np1=str(n), block0np2=str(n),
               niter=str(steps), dt=repr(0.003385 * 64 / n))
    env.pop('OSBLI_OUT', None)
    out = subprocess.run([exe], env=env, stdout=subprocess.PIPE, text=True, check=True).stdout
    for line in out.splitlines():
        if 'Total Wall time' in line:
            return float(line.split()[-1])
    return None


def run_oracle_port_cpu(steps, n):
    """Fallback when oracle/_ref is absent: the plain-C oracle port (1 core) on a small sample."""
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import numpy as np
    import oracle_util as ou
    plan = tgv_plan([n, n, n])
    q = [np.zeros((n + 10,) * 3) for _ in range(5)]
    tgv_state_into(q, plan, 0, n)
    t0 = time.perf_counter()
    ou.oracle_advance(plan, q, steps)
    return time.perf_counter() - t0


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_size or cpu_sample_size(cores)
    exe = os.path.join(REPO, 'oracle', '_ref', 'tgv_teno5', 'ref_omp')
    kind = 'reference'
    if os.path.exists(exe):
        if args.warmup > 0:
            run_reference_cpu(args.warmup, n, cores)
        wall = run_reference_cpu(args.steps, n, cores)
        sample = 'TGV TENO5 %d^3 x %d steps (reference generated C, g++ -O3 -march=x86-64-v3 -fopenmp, OPS stand-in), %s' % (n, args.steps, cpu_model())
    else:   # the oracle always exists: fall back to the plain-C port (scalar, 1 core)
        kind, cores, n = 'port', 1, 40
        wall = run_oracle_port_cpu(args.steps, n)
        sample = 'TGV TENO5 %d^3 x %d steps (oracle/osbli_oracle.c port, 1 core), %s' % (n, args.steps, cpu_model())
    value = n ** 3 * args.steps / wall
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'TGV Re=1600 TENO5+StoreSome(4)+RK-LS3 512^3 fp64 (BASELINE configs[2]); CPU arm timed on a bounded %d^3 sample' % n},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ parity beside the number
def parity_check(args, world, rank, local_rank, dist):
    """Correctness carried by the bench line itself (same library, same context type, right after the timed region):
    a small TGV block (64^3, the workload's scheme) is advanced
      (a) on this rank's GPU alone and compared with the reference's own generated C (oracle/_ref/<config>/ref_seq, run on
          the host for the same grid and steps; the plain-C oracle port when the executable is absent) -> max_rel_err;
      (b) for N > 1, slab-decomposed over all N ranks with the peer-store halo exchange and compared BIT FOR BIT with (a).
    The reference / oracle is the checker here, never the thing measured."""
    import numpy as np
    import opensbli_b200
    from opensbli_b200.decomp import DistributedSimulation, local_extent
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    n, nsteps = 64, 2
    plan = tgv_plan([n, n, n], args.workload)
    plan['delta'] = [2 * math.pi / n] * 3
    plan['constants']['dt'] = 0.003385 * 64 / n
    names = ['rho', 'rhou0', 'rhou1', 'rhou2', 'rhoE']
    inner = (slice(5, -5),) * 3
    out = {'grid': [n, n, n], 'steps': nsteps, 'max_rel_err': None, 'vs': None, 'multi_gpu_bit_identical': None}
    single = None
    if rank == 0:
        q0 = [np.zeros((n + 10,) * 3) for _ in range(5)]
        tgv_state_into(q0, plan, 0, n)
        with opensbli_b200.Simulation(plan, device=local_rank) as sim:
            sim.set_state(q0)
            sim.step(nsteps)
            single = np.stack([a[inner] for a in sim.get_state()])
        config = 'tgv_teno5' if args.workload == 'teno5' else 'tgv_central4'
        try:
            import oracle_util as ou
            if ou.have_ref(config):
                r = ou.run_ref(config, dict(block0np0=n, block0np1=n, block0np2=n, niter=nsteps, dt=plan['constants']['dt']), names, exe='ref_seq')
                ref = np.stack([r[f][inner] for f in names])
                out['vs'] = "reference's generated C oracle/_ref/%s/ref_seq (g++ -O2 -ffp-contract=off, OPS stand-in), %d^3, %d steps" % (config, n, nsteps)
            else:
                qo, _ = ou.oracle_advance(plan, [a.copy() for a in q0], nsteps)
                ref = np.stack([a[inner] for a in qo])
                out['vs'] = 'oracle/osbli_oracle.c port (oracle/_ref absent), %d^3, %d steps' % (n, nsteps)
            norms = [np.abs(ref[0]).max(), np.abs(ref[1:4]).max(), np.abs(ref[1:4]).max(), np.abs(ref[1:4]).max(), np.abs(ref[4]).max()]
            out['max_rel_err'] = float(max(np.abs(single[m] - ref[m]).max() / norms[m] for m in range(5)))
            out['tolerance'] = 1e-12 * nsteps
            out['ok'] = bool(out['max_rel_err'] <= out['tolerance'])
        except Exception as e:       # the checker must never take the measurement down with it
            out['vs'] = 'checker failed: %r' % (e,)
    if world > 1:
        import torch
        ds = DistributedSimulation(plan, dist, device=local_rank)
        k0, nk = local_extent(plan, rank, world)
        q0 = [np.zeros((nk + 10, n + 10, n + 10)) for _ in range(5)]
        tgv_state_into(q0, ds.plan, k0, nk)
        ds.set_state(q0)
        ds.step(nsteps)
        ds.barrier()
        mine = np.stack([a[inner] for a in ds.sim.get_state()])
        ds.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            got = np.concatenate(parts, axis=1)
            out['multi_gpu_bit_identical'] = bool(np.array_equal(got, single))
            out['multi_gpu_max_abs_diff'] = float(np.abs(got - single).max())
            out['multi_gpu'] = '%d slabs of %d planes, peer-store halo exchange, vs the single-GPU run of the same block' % (world, nk)
    return out


def central4_secondary(size, device, hbm_peak, steps=5):
    """BASELINE configs[1]'s scheme (shipped taylor_green_vortex app: Central(4) skew-symmetric + RungeKutta(3)) at the headline
    size on the same GPU, carried in the one JSON line: the bandwidth-bound stage kernel against the measured HBM copy bandwidth."""
    import numpy as np
    import opensbli_b200
    n = size
    plan = tgv_plan([n, n, n], 'central4')
    q = [np.zeros((n + 10,) * 3) for _ in range(5)]
    tgv_state_into(q, plan, 0, n)
    with opensbli_b200.Simulation(plan, device=device) as sim:
        sim.set_state(q)
        del q
        sim.step(3)
        ms = sim.step_timed(steps)
        prof = sim.profile_step()
        finite = bool(np.isfinite(sim.download('rho')).all())
    launch_ms = (prof['central']['ms'] + prof['viscous']['ms'] + prof['prim']['ms']) / max(prof['central']['launches'], 1)
    ach = (ALG_BYTES_PER_UPDATE / 3.0) * n ** 3 / (launch_ms * 1e-3) / 1e9
    return {'workload': 'TGV Re=1600 Central(4) skew-symmetric + RungeKutta(3), %d^3 fp64 (BASELINE configs[1] scheme at the headline size)' % n,
            'value': n ** 3 * steps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps, 'steps': steps, 'finite': finite,
            'roofline': {'bound': 'hbm', 'kernel': 'k_central3d_fused', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                         'launch_ms': launch_ms, 'bytes_model': 'algorithmic 160 B per point per stage'}}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--size', type=int, default=512, help='points per direction per GPU (default 512: the headline case)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='teno5', choices=['teno5', 'central4'], help='teno5 = headline (BASELINE configs[2]); central4 = configs[1] at --size')
    ap.add_argument('--cpu-size', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-chunk', default='64', help="planes per window of the pipelined end-to-end leg, or 'ramp' (n/16, n/8, n/4, n/4, 3n/16, n/8; measured slower: 0.76 vs 0.83 G)")
    ap.add_argument('--e2e-contexts', type=int, default=3, help='window contexts taking turns in the pipelined end-to-end leg')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --size^3 points per GPU (default); strong: one --grid^3 block cut over all GPUs (BASELINE configs[4])')
    ap.add_argument('--grid', type=int, default=1024, help='strong scaling: points per direction of the whole block')
    ap.add_argument('--develop-time', type=float, default=4.0, help='--state developed: time units the half-size box is advanced for')
    ap.add_argument('--state', default='smooth', choices=['smooth', 'perturbed', 'developed'],
                    help='smooth: analytic TGV field (every TENO stencil passes the cut-off: the best case of the all-pass shortcut); '
                         'perturbed: the same field with 5 %% random noise (nearly every wave takes the full cut-off path: the worst case)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the Central-4 (BASELINE configs[1] scheme) line carried as `secondary`')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
        return

    import numpy as np
    import torch
    import opensbli_b200
    from opensbli_b200.decomp import DistributedSimulation, local_extent

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 back end has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if world != args.gpus and rank == 0:
        print('bench.py: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE' % (args.gpus, world), file=sys.stderr)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    if args.scaling == 'strong':
        plan = tgv_plan([args.grid] * 3, args.workload)
        plan['delta'] = [2 * math.pi / args.grid] * 3
        plan['constants']['dt'] = 0.003385 * 64 / args.grid
    else:
        plan = tgv_plan(global_grid(world, args.size), args.workload)

    class _Solo(object):
        def get_rank(self): return 0
        def get_world_size(self): return 1
    dsim = DistributedSimulation(plan, dist if world > 1 else _Solo(), device=local_rank)
    sim = dsim.sim
    lplan = dsim.plan
    k0, nk = local_extent(plan, rank, world)
    shape = sim.shape
    nbytes_state = 5 * int(np.prod(shape)) * 8

    # pinned host buffers (torch only provides the pinned allocation)
    pin = not args.no_e2e                      # the end-to-end leg copies from / to pinned host memory
    hin = [torch.empty(shape, dtype=torch.float64, pin_memory=pin) for _ in range(5)]
    hout = [torch.empty(shape, dtype=torch.float64, pin_memory=True) for _ in range(5)] if pin else []
    q_in = [t.numpy() for t in hin]
    q_out = [t.numpy() for t in hout]
    developed = None
    if args.state == 'developed':
        # a flow with small scales in it: the vortex on a box of half the size per direction (same Mach and Reynolds numbers)
        # advanced to t = --develop-time on this GPU with the same kernels, then tiled 2 x 2 x 2 -- what TENO sees are the
        # differences between neighbouring cells, which the tiling keeps.  Single GPU, cubic block.
        if world != 1 or args.scaling != 'weak' or args.size % 2:
            raise SystemExit('--state developed: one GPU, even --size')
        h = args.size // 2
        hp = tgv_plan([h] * 3, args.workload)
        nst = int(math.ceil(args.develop_time / hp['constants']['dt']))
        from opensbli_b200 import Simulation
        with Simulation(hp, device=local_rank) as small:
            b = [np.zeros((h + 10,) * 3) for _ in range(5)]
            tgv_state_into(b, hp, 0, h)
            small.set_state(b)
            small.step(nst)
            developed = [np.tile(a[5:-5, 5:-5, 5:-5], (2, 2, 2)) for a in small.get_state()]
        if not all(np.isfinite(a).all() for a in developed):
            raise SystemExit('--state developed: the precursor run is not finite')
        developed_note = 'TGV on %d^3 advanced %d steps to t = %.2f with the same kernels, tiled 2 x 2 x 2' % (h, nst, nst * hp['constants']['dt'])

    def fill_state(bufs):
        tgv_state_into(bufs, lplan, k0, nk)
        if developed is not None:
            for b, a in zip(bufs, developed):
                b[5:-5, 5:-5, 5:-5] = a
        if args.state == 'perturbed':           # 5 % multiplicative noise on every conserved variable, seeded per rank
            rng = np.random.default_rng(1234 + rank)
            for b in bufs:
                b *= 1.0 + 0.05 * rng.standard_normal(b.shape)
    fill_state(q_in)
    dsim.set_state(q_in)           # collective: the upload also writes the halo planes the neighbours store into

    def barrier():
        sim.sync()
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    # ---- resident-state timing: W warm-up steps, then EXACTLY K steps between barriers + syncs
    dsim.step(W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sim.launch_count()
    torch.cuda.synchronize()
    barrier()
    sim.timer_start()                              # CUDA events on the solver's own stream (torch events would not see it)
    dsim.step(K)
    ms = sim.timer_stop()
    barrier()
    launches = sim.launch_count() - launches0
    clocks = sampler.summary()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    points = float(np.prod(plan['np']))
    value = points * K / (ms * 1e-3)

    # ---- state sanity (no NaN) after the timed region
    chk = sim.download('rho')
    finite = bool(np.isfinite(chk).all())

    # ---- per-kernel-family device time (events around every launch of one step) -> roofline of the dominant family
    prof = sim.profile_step()          # every rank takes part (the stage sequence contains the neighbour handshakes)
    roofline = roofline_hbm = None
    fp64_peak = opensbli_b200.measure_fp64_peak(local_rank)
    # share of TENO5 characteristic waves that left the all-pass shortcut during one more step (instrumented: not timed)
    slow_fraction = None
    if args.workload == 'teno5':
        sim.slow_path_count(True)
        dsim.step(1)
        barrier()
        nslow = sim.slow_path_count(False)
        pl = lplan['np']
        waves = 3 * 5 * float((pl[0] + 1) * pl[1] * pl[2] + pl[0] * (pl[1] + 1) * pl[2] + pl[0] * pl[1] * (pl[2] + 1))
        slow_fraction = nslow / waves
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    if prof and args.workload == 'central4':
        # dominant kernel: the fused central stage kernel (decomposed runs without peers: k_central + k_viscous3d_tiled)
        tot = sum(v['ms'] for v in prof.values())
        ce = prof['central']
        launch_ms = (ce['ms'] + prof['viscous']['ms'] + prof['prim']['ms']) / max(ce['launches'], 1)
        pts_local = float(np.prod(lplan['np']))
        ach = (ALG_BYTES_PER_UPDATE / 3.0) * pts_local / (launch_ms * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': 'k_central3d_fused (constituent relations + Central(4) convective + viscous terms + RK stage update)',
                    'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                    'traffic': NCU_DRAM_BYTES_PER_CENTRAL_LAUNCH_512 if (world == 1 and args.size == 512) else None,
                    'traffic_source': 'profiles/r02_final_ncu_central_fused_512.md (bytes per launch, ncu --set full)',
                    'launch_ms': launch_ms, 'share_of_step': (ce['ms'] + prof['viscous']['ms'] + prof['prim']['ms']) / tot if tot else None,
                    'bytes_model': 'algorithmic 160 B per point per stage: q read once, q and RK register written once, RK register read once',
                    'families_ms': {k: v['ms'] for k, v in prof.items()},
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'}
    elif prof:
        fl = prof['flux']
        per_launch_ms = fl['ms'] / max(fl['launches'], 1)
        pts_local = float(np.prod(lplan['np']))
        ach = ALG_FLOP_PER_POINT_FLUX_SWEEP * pts_local / (per_launch_ms * 1e-3) / 1e12
        tot = sum(v['ms'] for v in prof.values())
        roofline = {'bound': 'fp64', 'kernel': 'k_flux3_x / k_flux3_march<y|z> (TENO5 characteristic flux sweep + flux difference)',
                    'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': ach / fp64_peak if fp64_peak else None,
                    'traffic': NCU_DRAM_BYTES_PER_FLUX_LAUNCH_512 if (world == 1 and args.size == 512) else None,
                    'traffic_source': 'profiles/r02_final_ncu_top_kernels_512.md (bytes per launch, ncu --set full)', 'launch_ms': per_launch_ms,
                    'pipe_fp64_busy': NCU_PIPE_FP64_BUSY_512 if (world == 1 and args.size == 512 and args.state == 'smooth') else None,
                    'pipe_fp64_busy_source': 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active, profiles/r02_final_ncu_top_kernels_512.md',
                    'frac_note': 'frac uses the ALGORITHMIC operation count of the reference (SURVEY 8d); it exceeds 1 because the kernels need ~3.4x fewer FP64 instructions than that count -- pipe_fp64_busy is the hardware-side fraction', 'share_of_step': fl['ms'] / tot if tot else None,
                    'peak_source': 'measured in this run: DFMA micro-benchmark osb_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)',
                    'flop_model': "reference's own count_ops: 2836 per point per LLFTeno_reconstruction loop + 50/3 Residual",
                    'families_ms': {k: v['ms'] for k, v in prof.items()}}
        roofline_hbm = {'bound': 'hbm', 'achieved': ALG_BYTES_PER_UPDATE * value / world / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': ALG_BYTES_PER_UPDATE * value / world / 1e9 / hbm_peak,
                        'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                        'note': 'whole step, algorithmic 480 B/update; the step is FP64-bound, shown for context'}

    # ---- end to end through the C-ABI call with HOST buffers: every step = H2D state + 1 step + D2H state
    e2e = None
    if not args.no_e2e:
        Ke = min(K, 5)
        fill_state(q_in)
        if world == 1:
            sim.advance_host(q_in, q_out, 1)      # warm-up of the path
            tot_ms = 0.0
            src, dst = q_out, q_in
            for _ in range(Ke):
                tot_ms += sim.advance_host(src, dst, 1)
                src, dst = dst, src
            whole = {'value': points * Ke / (tot_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': nbytes_state,
                     'd2h_bytes_per_step': nbytes_state, 'steps': Ke,
                     'call': 'osb_advance_host(ctx, q_in, q_out, 1) per step, pinned host buffers in the reference layout'}
            e2e = whole
            # the same call pipelined: windows of the block along z on three contexts, so that the upload of window k+1, the
            # sweep of window k and the download of window k-1 overlap (hostpipe.py); host wall clock around the call
            try:
                from opensbli_b200.hostpipe import HostPipeline
                ok_, need_, free_ = pipeline_fits(plan, plan['np'][2], 64 if args.e2e_chunk == 'ramp' else int(args.e2e_chunk), args.e2e_contexts, torch)
                if not ok_:
                    raise MemoryError('pipelined leg skipped: needs about %.0f GB, %.0f GB free on the device' % (need_ / 1e9, free_ / 1e9))
                with HostPipeline(plan, chunk=(args.e2e_chunk if args.e2e_chunk == 'ramp' else int(args.e2e_chunk)), nsteps=1, device=local_rank, contexts=args.e2e_contexts) as pipe:
                    pipe.advance(src, dst)        # warm-up (module load, first-touch of the window contexts)
                    src, dst = dst, src
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(Ke):
                        pipe.advance(src, dst)
                        src, dst = dst, src
                    wall = time.perf_counter() - t0
                    up, down = pipe.bytes_per_call()
                    e2e = {'value': points * Ke / wall, 'unit': UNIT, 'h2d_bytes_per_step': up, 'd2h_bytes_per_step': down, 'steps': Ke,
                           'call': 'HostPipeline.advance(q_in, q_out): osb_staging_upload (each plane once) + per window osb_staging_feed / osb_step / osb_host_planes_download, '
                                   'pinned host buffers in the reference layout, windows of %s planes + 2 x %d guard planes on %d contexts'
                                   % ([z1 - z0 for z0, z1 in pipe.windows], pipe.guard, len(pipe.sims)),
                           'timing': 'host wall clock around the calls (each call ends with a synchronisation of all its streams)',
                           'launches_per_step': pipe.launches, 'unpipelined': whole}
            except Exception as ex:               # a failure of the pipelined leg must not cost the bench line
                e2e = dict(whole, pipelined_error=repr(ex))
        else:
            barrier()
            sim.timer_start()
            for _ in range(Ke):
                dsim.set_state(q_in)
                dsim.step(1)
                for m, nme in enumerate(sim.q_names):
                    sim.download_into(nme, q_out[m])
            t = torch.tensor([sim.timer_stop()], dtype=torch.float64, device='cuda')
            barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            whole = {'value': points * Ke / (float(t.item()) * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': nbytes_state * world,
                     'd2h_bytes_per_step': nbytes_state * world, 'steps': Ke,
                     'call': 'per step: osb_upload x5 + stage loop with halo pushes + osb_download x5 on every rank'}
            e2e = whole
            # pipelined: every rank advances its slab window by window; the guard planes of its outer windows come out of the
            # neighbours' staging copies over NVLink, so there is no per-stage halo exchange at all (hostpipe.DistributedHostPipeline)
            try:
                from opensbli_b200.hostpipe import DistributedHostPipeline
                ok_, need_, free_ = pipeline_fits(plan, nk, int(args.e2e_chunk), args.e2e_contexts, torch)
                fits = torch.tensor([1.0 if ok_ else 0.0], dtype=torch.float64, device='cuda')
                dist.all_reduce(fits, op=dist.ReduceOp.MIN)          # all ranks take the same branch
                if float(fits.item()) < 1.0:
                    raise MemoryError('pipelined leg skipped: needs about %.0f GB per rank, %.0f GB free on the device' % (need_ / 1e9, free_ / 1e9))
                with DistributedHostPipeline(plan, dist, local_rank, chunk=int(args.e2e_chunk), nsteps=1, contexts=args.e2e_contexts) as dp:
                    src, dst = q_in, q_out
                    dp.advance(src, dst)          # warm-up
                    src, dst = dst, src
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(Ke):
                        dp.advance(src, dst)
                        src, dst = dst, src
                    barrier()
                    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    up, down = dp.bytes_per_call()
                    e2e = {'value': points * Ke / float(t.item()), 'unit': UNIT, 'h2d_bytes_per_step': up * world, 'd2h_bytes_per_step': down * world,
                           'steps': Ke, 'call': 'DistributedHostPipeline.advance(q_in, q_out) on every rank: slab staged once, neighbours\' guard planes '
                           'pulled over NVLink (osb_staging_pull), windows of %d + 2 x %d guard planes, no per-stage halo exchange'
                           % (int(args.e2e_chunk), dp.pipe.guard),
                           'timing': 'host wall clock between barriers, max over ranks', 'unpipelined': whole}
            except Exception as ex:
                e2e = dict(whole, pipelined_error=repr(ex))
                barrier()

    # ---- parity beside the number (small block: vs the reference executable; N>1: decomposed vs single GPU, bit for bit)
    parity = None
    if not args.no_parity:
        barrier()
        parity = parity_check(args, world, rank, local_rank, dist)

    # ---- CPU baseline beside it (rank 0, N=1): the reference's generated C on this box's host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_size or cpu_sample_size(cores)
        wall = run_reference_cpu(2, n, cores)
        if wall:
            cpu = {'value': n ** 3 * 2 / wall, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                   'sample': 'TGV TENO5 %d^3 x 2 steps, reference generated C (OPS-OpenMP stand-in, g++ -O3 -march=x86-64-v3 -fopenmp), %s' % (n, cpu_model())}
        else:
            cpu = {'value': None, 'unit': UNIT, 'cores': cores, 'kind': 'reference', 'sample': 'oracle/_ref/tgv_teno5/ref_omp missing'}

    secondary = None
    if rank == 0 and world == 1 and args.workload == 'teno5' and not args.no_secondary:
        try:
            secondary = central4_secondary(args.size, local_rank, hbm_peak)
        except Exception as e:
            secondary = {'error': repr(e)}

    if rank == 0:
        line = {'metric': METRIC if args.workload == 'teno5' else 'grid-point updates/s (fp64, Central-4 TGV)', 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic',
                'config': {'workload': ('TGV Re=1600 TENO5(Roe,LLF)+StoreSome(4) viscous+RK-LS3, %s grid fp64 (BASELINE configs[%d]), %s'
                                        % ('x'.join(str(n) for n in plan['np']), 2 if world == 1 else 4,
                                           ('%d^3 points per GPU' % args.size) if args.scaling == 'weak' else 'strong scaling: the block is cut into %d slabs' % world)) if args.workload == 'teno5' else
                                       'TGV Re=1600 Central(4) skew-symmetric + RungeKutta(3), %s grid fp64 (BASELINE configs[1] at this size)' % 'x'.join(str(n) for n in plan['np']),
                           'grid': plan['np'], 'parallelism': 'slab%d' % world,
                           'state': args.state,
                           'l2_policy': 'working set %.1f GB per GPU >> 126 MB L2 (no flush needed)' % (19 * np.prod(shape) * 8 / 1e9),
                           'finite': finite},
                'roofline': roofline, 'roofline_hbm': roofline_hbm, 'cpu_baseline': cpu, 'clocks': clocks, 'e2e': e2e, 'parity': parity, 'secondary': secondary,
                'gpu_launches': int(launches), 'fp64_peak_tflops_measured': fp64_peak,
                'fp64_peak': {'measured_tflops': fp64_peak, 'how': 'osb_measure_fp64_peak: 8 independent DFMA chains per thread, 148 x 8 blocks x 256 threads, 16384 iterations, best of 4',
                              'sm_mhz_during_run': clocks.get('sm_mhz'), 'theoretical_tflops_at_max_clock': 148 * 64 * 2 * (clocks.get('sm_max_mhz') or 1965.0) * 1e6 / 1e12,
                              'note': '64 FP64 FMA lanes per SM; MEASURED_PEAKS.json carries no FP64 entry'},
                'state': {'kind': args.state if developed is None else 'developed: ' + developed_note, 'waves_on_full_cutoff_path': slow_fraction,
                          'note': 'share of TENO5 characteristic waves (5 per interface per sweep) that needed the full cut-off evaluation in one step after the timed region'},
                'families_ms_rank0': {k: v['ms'] for k, v in prof.items()} if prof else None,
                'alg_flop_per_update': ALG_FLOP_PER_UPDATE, 'achieved_alg_tflops': ALG_FLOP_PER_UPDATE * value / world / 1e12}
        if args.workload != 'teno5':      # the reference operation count quoted above is the TENO5 one
            line.pop('alg_flop_per_update'); line.pop('achieved_alg_tflops')
        print(json.dumps(line))
    dsim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
