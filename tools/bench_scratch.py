import sys, time, torch
sys.path.insert(0, '.')
from po_brax_b200 import envs
for n in (16, 1 << 20):
    e = envs.create_gym_env('ant', batch_size=n, seed=0, episode_length=20, eval_metrics=True, discount=0.99)
    e.reset()
    a = torch.rand((64, n, 8), device='cuda') * 2 - 1
    for i in range(20): e.step(a[i % 64])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100): e.step(a[i % 64])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 100
    print(f'scratch.py-style gym env (ant, eval_metrics, episode_length 20), {n} envs: {dt * 1e6:.1f} us per step', e.get_stats())
