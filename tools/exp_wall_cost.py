import os, sys, torch
sys.path.insert(0, '/root/repo')
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
n = 1 << 20
for label, kw in (('spawn at the wall (default)', {}), ('spawn mid-corridor (no leg near a wall)', {'init_ant_pos': ((-0.2, 3.0), (0.2, 4.0))}),
                  ('no walls', {'walls': False})):
    env = envs.create('ant_heavenhell', batch_size=n, **kw)
    keys = shard_keys(env, 0, n, 0, 1)
    g = torch.Generator(device='cuda').manual_seed(1)
    a = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
    s = env.reset(keys)
    for i in range(10): s = env.step(s, a[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): s = env.step(s, a[i % 4])
    e1.record(); e1.synchronize()
    print(label, 'ms/step', e0.elapsed_time(e1) / 40)
    del env, s
