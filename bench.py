#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the fused Ant-HeavenHell step on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the restated CPU oracle (C step, all host cores; same metric)

One "step" = one fused env.step over the whole batch (physics x10 substeps + task logic + obs +
episode/autoreset). Envs shard across GPUs with no data-path collective (weak scaling: envs per GPU
fixed); the only collective is an NCCL all-reduce of the 8-double episode-metric vector every 100 steps.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch, from the committed ncu --set full capture
# (profiles/ncu_step_hh_r1.txt: 593.4 MB read + 984.2 MB write at 1 Mi HeavenHell envs = 1505 B per env-step)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {('ant_heavenhell', 1 << 20): 1.578e9}


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


def main_graft(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=graft) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    from po_brax_b200 import envs, _lib
    from po_brax_b200.host import HostStepper
    from po_brax_b200.parallel import shard_keys, reduce_metrics
    lib = _lib.load()
    n = args.envs_per_gpu
    total = n * world
    env = envs.create(args.env, batch_size=n, episode_length=1000, auto_reset=True, eval_metrics=True)
    keys = shard_keys(env, seed=0, total=total, rank=rank, world=world)
    state = env.reset(keys)
    period = 8
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    actions = torch.rand((period, n, 8), device=device, generator=g) * 2 - 1
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf = fp32_peak_tflops(torch, lib, device)
    for i in range(W):
        state = env.step(state, actions[i % period])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    launches = 0
    acc_total = None
    for i in range(K):
        state = env.step(state, actions[i % period])
        launches += 1
        if (i + 1) % 100 == 0:
            acc_total = reduce_metrics(state, world)  # NCCL all-reduce(sum) of 8 doubles, async on this stream
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    value = total * K / (ms * 1e-3)
    per_gpu = n * K / (ms * 1e-3)
    if acc_total is None:
        acc_total = reduce_metrics(state, world)
    acc = acc_total.cpu().tolist()

    # ---- e2e: host buffers in, host buffers out (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        del state, env
        torch.cuda.empty_cache()
        hs = HostStepper(args.env, n, chunks=8, device=device, episode_length=1000, auto_reset=True)
        hs.reset(keys)
        hs.action_host.copy_(actions[0].cpu())
        ke = max(3, min(args.e2e_steps, K))
        for _ in range(3):
            hs.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        for st in hs.streams:
            st.wait_event(e0)
        for _ in range(ke):
            hs.step()  # enqueue H2D + step + D2H per chunk, then wait for the host buffers
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
        e2e = {'value': total * ke / (max(ems_dev, ems_wall) * 1e-3), 'unit': 'env-steps/s',
               'h2d_bytes_per_step': hs.h2d_bytes_per_step * world, 'd2h_bytes_per_step': hs.d2h_bytes_per_step * world,
               'steps': ke, 'ms_per_step_device': ems_dev / ke, 'ms_per_step_wall': ems_wall / ke,
               'what': 'HostStepper.step(): pinned action[N,8] H2D, fused step, obs[N,D]+reward+done D2H, 8 chunks on 8 streams'}
        del hs
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
        'traffic': NCU_TRAFFIC_BYTES_PER_LAUNCH.get((args.env, n)),
        'peak_source': 'measured live: pobrax_fp32_probe (dependent-chain FFMA kernel, burst, best of 5); '
                       'nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4',
        'kernel': f'step_kernel<{args.env}>', 'launch_ms': ms / K,
        'algorithmic_flops_per_env_step': flops, 'units_per_launch': n,
        'hbm': {'bound': 'hbm', 'achieved': ach_gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gb / hbm_peak,
                'algorithmic_bytes_per_env_step': byts, 'peak_source': hbm_src},
        'step_roofline_env_steps_per_s': min(peak_tf * 1e12 / flops, hbm_peak * 1e9 / byts),
        'frac_of_step_roofline': per_gpu / min(peak_tf * 1e12 / flops, hbm_peak * 1e9 / byts),
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores, m, st_ = os.cpu_count() or 1, args.ref_envs_per_core, 100
        v, dt = run_cpu_oracle(args.env, st_, 3, m, cores)
        cpu_baseline = {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                        'sample': f'{cores} processes x {m} envs x {st_} steps of {args.env}, {dt:.1f} s ({CPU_KIND})'}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f'{args.env} (BASELINE configs[4]), {n} envs per GPU, episode_length 1000, cached autoreset, '
                               f'i.i.d. U(-1,1) actions (device-resident, period {period})',
                   'phase': f'timed steps are steps {W}..{W + K} after a reset of all envs; the cost of a step grows with the '
                            'time since reset as random-action ants gather along the walls (DESIGN.md section 6)',
                   'envs_per_gpu': n, 'total_envs': total, 'parallelism': f'env-sharded x{world}, no per-step collective',
                   'l2': f'state+obs per GPU = {n * (512 + 4 * env_obs(args.env)) / 1e6:.0f} MB >> 126 MB L2 (inputs larger than L2)',
                   'metric_allreduce_every': 100},
        'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': launches * world,
        'clocks': clocks,
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
