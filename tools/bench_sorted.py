"""Experiment: how fast would the step be if near-wall envs shared warps? Sort the reset keys by the spawn's distance
to the nearest wall (torso y for HeavenHell) so that env order itself is clustered, then time the step.
    python tools/bench_sorted.py [env]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
name = sys.argv[1] if len(sys.argv) > 1 else 'ant_heavenhell'
n = 1 << 20
env = envs.create(name, batch_size=n)
keys = shard_keys(env, 0, n, 0, 1)
g = torch.Generator(device='cuda').manual_seed(1)
a = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
def run(keys, label, steps=40):
    s = env.reset(keys)
    for i in range(10): s = env.step(s, a[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): s = env.step(s, a[i % 4])
    e1.record(); e1.synchronize()
    print(label, 'ms/step', e0.elapsed_time(e1) / steps)
    return s
s = env.reset(keys)
pos = s.qp.pos[:, 0]
if name == 'ant_heavenhell':
    score = pos[:, 1]                                   # distance to the y = 0 wall
else:
    score = torch.minimum(4.5 - pos[:, 0].abs(), 4.5 - pos[:, 1].abs())   # distance to the cage
order = torch.argsort(score)
run(keys, 'unsorted')
run(keys[order], 'sorted by spawn wall distance')
run(keys[order], 'sorted, 200 steps', steps=200)
run(keys, 'unsorted, 200 steps', steps=200)

def first_steps(keys, label, k=5, reps=5):
    best = 1e9
    for _ in range(reps):
        s = env.reset(keys)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k): s = env.step(s, a[i % 4])
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / k)
    print(label, 'first', k, 'steps after reset, ms/step', best)
first_steps(keys, 'unsorted')
first_steps(keys[order], 'sorted (ordering still fresh)')
