#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the fused Ant-HeavenHell step on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the restated CPU oracle (C step, all host cores; same metric)

One "step" = one fused env.step over the whole batch (physics x10 substeps + task logic + obs +
episode/autoreset). Envs shard across GPUs with no data-path collective (weak scaling: envs per GPU
fixed; `strong` in the line / --scaling strong: 1 Mi envs in total); the only collective is an NCCL all-reduce of
the 8-double episode-metric vector every min(100, K) steps. `value` is the STATIONARY regime (stationary_state).
Extra keys of the one JSON line: early_phase, strong, per_config (BASELINE configs 1-4 at their sizes),
roofline.peak_nominal, e2e.roof_gbs / e2e.frac.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_ENV_STEP = {'ant': 4.30e4, 'ant_heavenhell': 4.30e4, 'ant_tag': 4.30e4, 'ant_gather': 4.34e4}
BYTES_PER_ENV_STEP = {'ant': 1332, 'ant_heavenhell': 1444, 'ant_tag': 1428, 'ant_gather': 2212}
METRIC = 'env-steps/sec'
# dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch, from the committed ncu --set full captures
# of the stationary regime (profiles/ncu_step_{hh,ant,ant_tag,ant_gather}_r2.txt; HeavenHell: 596.7 MB read + 982.7 MB
# write at 1 Mi envs = 1506 B per env-step against 1444 algorithmic: no re-reads)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {('ant_heavenhell', 1 << 20): 1.5793e9, ('ant', 1 << 20): 1.4909e9,
                                ('ant_tag', 1 << 20): 1.5779e9, ('ant_gather', 1 << 20): 2.2180e9}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='graft', choices=['graft', 'reference'])
    ap.add_argument('--env', default='ant_heavenhell', choices=sorted(FLOPS_PER_ENV_STEP))
    ap.add_argument('--envs-per-gpu', type=int, default=1 << 20)
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ref-envs-per-core', type=int, default=2048)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --envs-per-gpu on every GPU; strong: --envs-per-gpu is the TOTAL, split over the GPUs')
    ap.add_argument('--no-per-config', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU (oracle) arm
CPU_KIND = ('restated CPU oracle, not reference JAX: oracle/brax_step.c (scalar C restatement of the brax v1 step, '
            'gcc -O3, one process per core) under the NumPy task / episode / autoreset logic of oracle/envs.py')


def _ref_worker(conn, env_name, m, seed):
    import numpy as np
    from oracle import cstep, envs as oenvs, threefry as tf
    env = oenvs.create(env_name, episode_length=1000, auto_reset=True)
    cstep.attach(env.env.sys, threads=1)
    keys = tf.split(tf.prng_key(seed), m + 1)[1:]
    s = env.reset(keys)
    rng = np.random.default_rng(seed)
    conn.send('ready')
    while True:
        msg = conn.recv()
        if msg == 'stop':
            break
        a = rng.uniform(-1, 1, (m, 8)).astype(np.float32)
        s = env.step(s, a)
        conn.send(float(s.reward.sum()))


def run_cpu_oracle(env_name, steps, warmup, m, procs):
    """Times the restated CPU oracle (oracle/: C float32 step + NumPy env logic) on `procs` processes x `m` envs."""
    import multiprocessing as mp
    from oracle import cstep
    cstep.build()   # before the fork: the workers only dlopen it
    ctx = mp.get_context('fork')
    pipes, ps = [], []
    for i in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, env_name, m, 1000 + i), daemon=True)
        p.start()
        pipes.append(a); ps.append(p)
    for a in pipes:
        assert a.recv() == 'ready'

    def one_step():
        for a in pipes:
            a.send('step')
        for a in pipes:
            a.recv()
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    for a in pipes:
        a.send('stop')
    for p in ps:
        p.join(timeout=5)
    return procs * m * steps / dt, dt


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    cores = os.cpu_count() or 1
    m = args.ref_envs_per_core
    # bounded sample: keep the whole run within a few minutes whatever K the driver passes
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 5))
    val, dt = run_cpu_oracle(args.env, steps, warmup, m, cores)
    sample = f'{cores} processes x {m} envs x {steps} steps of {args.env}, {dt:.1f} s ({CPU_KIND})'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.env}, {cores * m} envs per step on the host CPU (bounded sample of the '
                               f'{args.envs_per_gpu}-env-per-GPU workload), episode_length 1000, cached autoreset'},
        'cpu_baseline': {'value': val, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for (t, ln) in self.lines if t0 <= t <= t1] or [ln for (_, ln) in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        for ln in rows:
            f = [x.strip() for x in ln.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def fp32_peak_tflops(torch, lib, device):
    """Non-tensor FP32 roof measured on this GPU: a pure dependent-chain FMA kernel, best of 5."""
    import ctypes as C
    blocks, iters = 148 * 16, 4096
    out = torch.empty(blocks * 256, dtype=torch.float32, device=device)
    flops = C.c_double()
    st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.pobrax_fp32_probe(out.data_ptr(), blocks, iters, st, C.byref(flops))
        assert rc == 0
        e1.record()
        e1.synchronize()
        if i:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 FP32 lanes x 2 (FMA) x 1.965 GHz = 74.4
EPISODE_LENGTH = 1000
PREROLL = 1000          # untimed steps before the timed region (stationary regime, see stationary_state)
ACTION_PERIOD = 64      # pre-generated i.i.d. U(-1,1) actions reused cyclically, period >= 64
CONFIG_OF = {'ant_heavenhell': 'BASELINE configs[4]', 'ant': 'BASELINE configs[1] family (plain Ant)',
             'ant_gather': 'BASELINE configs[2] family (Ant-Gather)', 'ant_tag': 'BASELINE configs[3] family (Ant-Tag)'}


def make_actions(torch, n, device, seed, period=ACTION_PERIOD):
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.rand((period, n, 8), device=device, generator=g) * 2 - 1


def stationary_state(torch, env, keys, actions, preroll=PREROLL, seed=99):
    """Reset, then bring the batch into the STATIONARY regime before anything is timed: the cost of a step grows with
    the time since reset (random-action ants gather along the walls), and a batch reset all at
    once would also truncate all at once at step 1000. So every env gets its own episode age -- info['steps'] drawn
    uniformly from [0, episode_length) -- and the batch is rolled `preroll` untimed steps: from then on the cached
    autoreset re-starts ~1/1000 of the envs per step and the mix of episode ages no longer changes."""
    state = env.reset(keys)
    g = torch.Generator(device=env.device).manual_seed(seed)
    state.buf['steps'].copy_(torch.randint(0, EPISODE_LENGTH, (env.batch_size,), device=env.device, generator=g).float())
    for i in range(preroll):
        state = env.step(state, actions[i % actions.shape[0]])
    return state


def time_steps(torch, env, state, actions, k, start=0, reduce_every=0, world=1, reduce_fn=None):
    """k steps on the current stream between two CUDA events (no host sync inside). Returns (state, ms, acc)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = None
    e0.record()
    for i in range(k):
        state = env.step(state, actions[(start + i) % actions.shape[0]])
        if reduce_every and (i + 1) % reduce_every == 0:
            acc = reduce_fn(state, world)   # NCCL all-reduce(sum) of 8 doubles, async on this stream
    e1.record()
    e1.synchronize()
    return state, e0.elapsed_time(e1), acc


def bench_small_config(torch, envs, name, n, peak_tf, device, steps=200, reps=5):
    """One BASELINE small-batch config (cached autoreset, stationary episode ages): `steps` env steps captured into a
    CUDA graph (the launch-bound way to run a small batch) and replayed `reps` times, best replay reported, next to
    the plain per-step launch path."""
    env = envs.create(name, batch_size=n, episode_length=EPISODE_LENGTH, auto_reset=True)
    keys = env.split_keys((0, 0), n + 1, first=1, count=n)
    acts = make_actions(torch, n, device, 4321)
    state = stationary_state(torch, env, keys, acts)
    torch.cuda.synchronize()
    state, ms_plain, _ = time_steps(torch, env, state, acts, steps)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        state = env.step(state, acts[0])
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for i in range(steps):
                state = env.step(state, acts[i % acts.shape[0]])
        best = float('inf')
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            g.replay()
            e1.record(side)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
    torch.cuda.current_stream(device).wait_stream(side)
    us = 1e3 * best / steps
    rate = n / (us * 1e-6)
    return {'env': name, 'envs': n, 'us_per_step': us, 'us_per_step_plain_launches': 1e3 * ms_plain / steps,
            'env_steps_per_s': rate, 'frac': rate * FLOPS_PER_ENV_STEP[name] / 1e12 / peak_tf,
            'how': f'{steps} steps as one CUDA graph, best of {reps} replays; cached autoreset; stationary episode ages; '
                   f'state fits L2 at this size'}


def d2h_roof_gbs(torch, dist, device, world, nbytes, reps=6):
    """Pinned device->host bandwidth of THIS box for the e2e roof: every rank copies a buffer of the step's D2H size
    at the same time (they share the host's PCIe / memory complex), `reps` times after a warm-up. GB/s per rank,
    min over ranks."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    e1.synchronize()
    gbs = torch.tensor([nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(gbs, op=dist.ReduceOp.MIN)
    return float(gbs.item())


def main_graft(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=graft) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    from po_brax_b200 import envs, _lib
    from po_brax_b200.host import HostStepper, bind_to_gpu_numa_node
    from po_brax_b200.parallel import shard_keys, reduce_metrics
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)   # CPU affinity of this rank -> its GPU's NUMA node (pinned buffers land there)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    lib = _lib.load()
    strong = args.scaling == 'strong'
    n = args.envs_per_gpu // world if strong else args.envs_per_gpu
    total = n * world
    K, W = args.steps, max(args.warmup, 3)
    cadence = min(100, K)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(n_envs, early=False):
        """The contract's measurement at n_envs per GPU: W warm-up steps, then EXACTLY K timed steps between barriers,
        max over ranks. early=True times the K steps right behind a reset of all envs instead of the stationary
        regime. Returns (ms of the K steps, accumulators, best-of-5 ms or None, gpu launches)."""
        env = envs.create(args.env, batch_size=n_envs, episode_length=EPISODE_LENGTH, auto_reset=True, eval_metrics=True)
        keys = shard_keys(env, seed=0, total=n_envs * world, rank=rank, world=world)
        acts = make_actions(torch, n_envs, device, 1234 + rank)
        state = env.reset(keys) if early else stationary_state(torch, env, keys, acts)
        for i in range(W):
            state = env.step(state, acts[i % ACTION_PERIOD])
        barrier()
        state, ms, acc = time_steps(torch, env, state, acts, K, start=W, reduce_every=cadence, world=world,
                                    reduce_fn=reduce_metrics)
        barrier()
        best5 = None
        if not early and K <= 250:   # best of 5 further repetitions of the same K steps (SURVEY 8(d) protocol)
            reps = []
            for r in range(5):
                state, m, _ = time_steps(torch, env, state, acts, K, start=W + (r + 1) * K)
                reps.append(m)
            best5 = min(reps + [ms])
        t = torch.tensor([ms, best5 if best5 is not None else 0.0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, b5 = [float(x) for x in t.tolist()]
        if acc is None:
            acc = reduce_metrics(state, world)
        return ms, acc.cpu().tolist(), (b5 if best5 is not None else None)

    peak_tf = fp32_peak_tflops(torch, lib, device)
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    ms, acc, best5 = measure(n)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    value = total * K / (ms * 1e-3)
    per_gpu = n * K / (ms * 1e-3)
    # kernels of this library inside the timed region: one step_kernel per step and rank (+ tag_rng_kernel for Tag)
    launches = K * world * (2 if args.env == 'ant_tag' else 1)

    ms_early, _, _ = measure(n, early=True)
    early = {'value': total * K / (ms_early * 1e-3), 'ms_per_step': ms_early / K,
             'what': f'the same K steps timed {W} steps behind a reset of ALL envs (round 1\'s headline regime)'}
    # Config 5 as BASELINE states it: 1 Mi envs TOTAL, N/G per GPU (strong scaling), beside the weak-scaling headline
    strong_line = None
    if not strong:
        tot = 1 << 20
        if world == 1 and n == tot:
            strong_line = {'total_envs': tot, 'envs_per_gpu': n, 'value': value, 'ms_per_step': ms / K, 'same_as': 'value'}
        elif tot % world == 0:
            ms_s, _, _ = measure(tot // world)
            strong_line = {'total_envs': tot, 'envs_per_gpu': tot // world, 'value': tot * K / (ms_s * 1e-3),
                           'ms_per_step': ms_s / K, 'scaling': 'strong'}

    per_config = None
    if world == 1 and not args.no_per_config:
        per_config = [bench_small_config(torch, envs, name, m, peak_tf, device)
                      for name, m in (('ant_heavenhell', 128), ('ant', 4096), ('ant_gather', 16384), ('ant_tag', 65536))]
        for c, label in zip(per_config, ('configs[0] batch on the GPU', 'configs[1]', 'configs[2]', 'configs[3]')):
            c['baseline_config'] = label

    # ---- e2e: host buffers in, host buffers out (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        torch.cuda.empty_cache()
        hs = HostStepper(args.env, n, chunks=8, device=device, episode_length=EPISODE_LENGTH, auto_reset=True)
        keys = hs.envs[0].split_keys((0, 0), total + 1, first=1 + rank * n, count=n)   # = shard_keys(seed 0)
        hs.reset(keys)
        hs.action_host.copy_(make_actions(torch, n, device, 1234 + rank, period=1)[0].cpu())
        ke = max(3, min(args.e2e_steps, K))
        for _ in range(3):
            hs.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        for _ in range(ke):
            hs.step()  # enqueue H2D + step + D2H per chunk (after the current stream), then wait for the host buffers
        cur = torch.cuda.current_stream(device)
        for st in hs.streams:
            cur.wait_stream(st)
        e1.record()
        barrier()
        tw1 = time.perf_counter()
        ems = torch.tensor([max(e0.elapsed_time(e1), 0.0), (tw1 - tw0) * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems_dev, ems_wall = [float(x) for x in ems.tolist()]
        d2h = hs.d2h_bytes_per_step
        del hs
        torch.cuda.empty_cache()
        roof = d2h_roof_gbs(torch, dist, device, world, d2h)
        e2e_val = total * ke / (max(ems_dev, ems_wall) * 1e-3)
        e2e = {'value': e2e_val, 'unit': 'env-steps/s',
               'h2d_bytes_per_step': n * 8 * 4 * world, 'd2h_bytes_per_step': d2h * world,
               'steps': ke, 'ms_per_step_device': ems_dev / ke, 'ms_per_step_wall': ems_wall / ke,
               'roof_gbs': roof, 'roof_env_steps_per_s': roof * 1e9 / (d2h / n) * world,
               'frac': e2e_val / (roof * 1e9 / (d2h / n) * world),
               'roof_how': f'pinned D2H copy of the step\'s {d2h / 1e6:.0f} MB per rank, all {world} ranks at once, '
                           'GB/s per rank (min over ranks); roof = that bandwidth / D2H bytes per env-step',
               'numa': numa,
               'what': 'HostStepper.step(): pinned action[N,8] H2D, fused step, obs[N,D]+reward+done D2H, 8 chunks on 8 '
                       'streams; the regime is the early one (a few steps behind a reset): the link, not the kernel, bounds it'}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, hbm_src = measured_peaks()
    flops, byts = FLOPS_PER_ENV_STEP[args.env], BYTES_PER_ENV_STEP[args.env]
    ach_tf = per_gpu * flops / 1e12
    ach_gb = per_gpu * byts / 1e9
    roofline = {
        'bound': 'fp32', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach_tf / peak_tf,
        'peak_nominal': FP32_NOMINAL_TFLOPS, 'frac_nominal': ach_tf / FP32_NOMINAL_TFLOPS,
        'traffic': NCU_TRAFFIC_BYTES_PER_LAUNCH.get((args.env, n)),
        'peak_source': 'measured live: pobrax_fp32_probe (dependent-chain FFMA kernel, burst, best of 5; '
                       'MEASURED_PEAKS.json has no FP32 figure); peak_nominal = 148 SM x 128 lanes x 2 x 1.965 GHz',
        'kernel': f'step_kernel<{args.env}>', 'launch_ms': ms / K,
        'algorithmic_flops_per_env_step': flops, 'units_per_launch': n,
        'hbm': {'bound': 'hbm', 'achieved': ach_gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gb / hbm_peak,
                'algorithmic_bytes_per_env_step': byts, 'peak_source': hbm_src},
        'step_roofline_env_steps_per_s': min(peak_tf * 1e12 / flops, hbm_peak * 1e9 / byts),
        'frac_of_step_roofline': per_gpu / min(peak_tf * 1e12 / flops, hbm_peak * 1e9 / byts),
    }
    if best5 is not None:
        roofline['best_of_6_launch_ms'] = best5 / K
        roofline['best_of_6_frac'] = (n * K / (best5 * 1e-3)) * flops / 1e12 / peak_tf
    cpu_baseline = None
    os.sched_setaffinity(0, all_cpus)   # the CPU leg uses every host core again
    if world == 1 and not args.no_cpu_baseline:
        cores, m, st_ = os.cpu_count() or 1, args.ref_envs_per_core, 100
        v, dt = run_cpu_oracle(args.env, st_, 3, m, cores)
        cpu_baseline = {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                        'sample': f'{cores} processes x {m} envs x {st_} steps of {args.env}, {dt:.1f} s ({CPU_KIND})'}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.env} ({CONFIG_OF[args.env]}), {n} envs per GPU, episode_length {EPISODE_LENGTH}, cached '
                               f'autoreset, i.i.d. U(-1,1) actions (device-resident, period {ACTION_PERIOD})',
                   'phase': f'stationary: every env starts at its own episode age (info.steps ~ U[0,{EPISODE_LENGTH})), then '
                            f'{PREROLL} untimed pre-roll steps + {W} warm-up steps before the {K} timed ones, so ~1/1000 of '
                            'the envs re-start per step and the share of ants along the walls has settled; `early_phase` '
                            'keeps the cheaper just-after-reset figure',
                   'envs_per_gpu': n, 'total_envs': total, 'parallelism': f'env-sharded x{world}, no per-step collective',
                   'l2': f'state+obs per GPU = {n * (512 + 4 * env_obs(args.env)) / 1e6:.0f} MB >> 126 MB L2 (inputs larger than L2)',
                   'metric_allreduce_every': cadence},
        'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': launches,
        'clocks': clocks, 'early_phase': early, 'strong': strong_line, 'per_config': per_config,
        'episode_metrics': dict(zip(('episodes', 'sum_return', 'sum_length', 'truncations', 'hits', 'heavens', 'hells',
                                     'dead_steps'), acc)),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def env_obs(name):
    return {'ant': 87, 'ant_heavenhell': 114, 'ant_tag': 103, 'ant_gather': 211}[name]


def emit(line):
    """The ONE JSON line of the contract, written to the real stdout (see __main__)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1

if __name__ == '__main__':
    a = parse()
    # Libraries write to stdout on their own (NCCL prints "NCCL version ..." there when NCCL_DEBUG is set on the box):
    # fd 1 is pointed at stderr for the whole run and the JSON line goes to a saved copy of the real stdout.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main_reference(a) if a.impl == 'reference' else main_graft(a))
