"""Small ragged run of every entry point for compute-sanitizer --tool memcheck (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from po_brax_b200 import envs
from po_brax_b200.envs.wrappers import RandomizedAutoResetWrapperNaive
for name in ('ant', 'ant_heavenhell', 'ant_gather', 'ant_tag'):
    for n in (1, 37):
        env = envs.create(name, batch_size=n, episode_length=3, eval_metrics=True)
        keys = env.split_keys((0, 5), n + 1, first=1, count=n)
        s = env.reset(keys)
        for t in range(5):
            s = env.step(s, torch.rand((n, 8), device='cuda') * 2 - 1)
        q = s.qp
        env._pack(q)
        e2 = envs.create(name, batch_size=n, episode_length=2, auto_reset=False)
        s2 = e2.reset(keys)
        for t in range(3):
            s2 = e2.step(s2, torch.zeros((n, 8), device='cuda'))
        s2 = e2.reset_where_done(s2, keys)
        if name != 'ant':
            w = RandomizedAutoResetWrapperNaive(e2)
            s2 = w.step(s2, torch.zeros((n, 8), device='cuda'))
            e2.split_pairs(s2.buf['rng'])
        torch.cuda.synchronize()
print('sanitize run ok')
