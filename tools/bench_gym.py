"""Gym-path timing: AutoresetVmapGymWrapper.step per step -- sync_free=False: step kernel + done.any() host round
trip + key split + reset_where_done; sync_free=True (default): step kernel + reset_where_done_chain, no host sync;
cuda_graph=True: the same launches replayed as one CUDA graph; copy=True (default): outputs cloned like the reference's
fresh arrays, copy=False: the live buffers.
    python tools/bench_gym.py [n_envs]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for sync_free, graph, copy in ((False, False, True), (True, False, True), (True, True, True), (True, False, False),
                               (True, True, False)):
    e = envs.create_gym_env('ant_heavenhell', batch_size=n, seed=0, cuda_graph=graph, copy=copy)
    e.sync_free = sync_free
    e.reset()
    a = torch.rand((64, n, 8), device='cuda', generator=torch.Generator(device='cuda').manual_seed(1)) * 2 - 1
    for i in range(50): e.step(a[i % 64])          # i.i.d. actions like bench.py (a constant action is a gait)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100): e.step(a[(50 + i) % 64])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 100
    print('gym step, sync_free', sync_free, 'cuda_graph', graph, 'copy', copy, 'ms/step', round(dt * 1e3, 3), 'env-steps/s', f'{n / dt:.3e}')
