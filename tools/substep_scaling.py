"""Fixed (prologue + epilogue) vs per-substep cost of the step kernel: time at 1, 5, 10, 20 substeps of the same h."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
n = 1 << 20
for name in ('ant', 'ant_heavenhell'):
    res = {}
    for sub in (1, 5, 10, 20):
        env = envs.create(name, batch_size=n, sys_dt=0.005 * sub, sys_substeps=sub)
        s = env.reset(shard_keys(env, 0, n, 0, 1))
        g = torch.Generator(device='cuda').manual_seed(1)
        a = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
        for i in range(10): s = env.step(s, a[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30): s = env.step(s, a[i % 4])
        e1.record(); e1.synchronize()
        res[sub] = e0.elapsed_time(e1) / 30
        del env, s
    per = (res[20] - res[10]) / 10
    print(name, {k: round(v, 4) for k, v in res.items()}, 'per-substep ms', round(per, 4), 'fixed ms', round(res[10] - 10 * per, 4))
