"""Rolls every env family a few steps and saves obs / packed qp / reward (bitwise comparison of two library builds).
    POBRAX_LIB=... python tools/dump_rollout.py out.pt [n_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from po_brax_b200 import envs
out, n, T = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2048, int(sys.argv[3]) if len(sys.argv) > 3 else 60
res = {}
g = torch.Generator(device='cuda').manual_seed(3)
acts = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
for kind in ('ant', 'ant_heavenhell', 'ant_tag', 'ant_gather'):
    env = envs.create(kind, batch_size=n, episode_length=25)
    s = env.reset(env.split_keys((0, 7), n + 1, first=1, count=n))
    for t in range(T):
        s = env.step(s, acts[t % 4])
    res[kind] = {k: s.buf[k].cpu() for k in ('obs', 'qp', 'reward', 'done', 'steps')}
torch.save(res, out)
print('saved', out)
