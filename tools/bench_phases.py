"""Step time by episode phase at 1 Mi envs, all envs reset together: the first 5 steps after the reset, steps 5-45,
45-245 (ants have wandered to the walls by then), 245-445, 445-645, 645-845, 845-995 (the batch truncates at 1000).
    python tools/bench_phases.py [env ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
n = 1 << 20
g = torch.Generator(device='cuda').manual_seed(1)
a = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
for name in sys.argv[1:] or ['ant', 'ant_heavenhell', 'ant_tag', 'ant_gather']:
    env = envs.create(name, batch_size=n)
    s = env.reset(shard_keys(env, 0, n, 0, 1))
    out = []
    for steps in (5, 40, 200, 200, 200, 200, 150):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): s = env.step(s, a[i % 4])
        e1.record(); e1.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    print(f'{name}: steps 0-5 {out[0]:.4f}, 5-45 {out[1]:.4f}, 45-245 {out[2]:.4f}, 245-445 {out[3]:.4f}, 445-645 {out[4]:.4f}, '
          f'645-845 {out[5]:.4f}, 845-995 {out[6]:.4f} ms/step', flush=True)
    del env, s
