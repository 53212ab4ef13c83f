"""Small ragged run of every entry point (for compute-sanitizer --tool memcheck where the pool allows it; also a plain
smoke of the corner / wall-group paths and the gym adapters)."""
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
# ants pushed against the walls (HeavenHell spawn box moved into a corner of the T junction): exercises the wall groups,
# the multi-candidate cull and the bisection; plus the gym adapters (device key chain, CUDA graph, unbatched)
env = envs.create('ant_heavenhell', batch_size=61, init_ant_pos=((0.6, 4.2), (1.0, 4.9)))
s = env.reset(env.split_keys((0, 9), 62, first=1, count=61))
for t in range(30):
    s = env.step(s, torch.rand((61, 8), device='cuda') * 2 - 1)
for graph in (False, True):
    g = envs.create_gym_env('ant_tag', batch_size=19, seed=1, episode_length=4, cuda_graph=graph)
    g.reset()
    for t in range(9):
        g.step(torch.zeros((19, 8), device='cuda'))
u = envs.create_gym_env('ant_gather', seed=2, episode_length=3)
u.reset()
for t in range(7):
    u.step(np.zeros(8, np.float32))
torch.cuda.synchronize()
print('sanitize run ok')
