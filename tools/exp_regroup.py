"""Experiment (upper bound, no kernel change): how much faster is the step when the envs that enter the rare contact
region (a body within reach of a wall, a torso on the ground) share warps? Takes bench.py's stationary state, flags
envs by a geometric proxy, physically permutes every env-indexed buffer so that flagged envs come first, and times the
step right behind the permutation and again later (how fast the grouping decays).
    python tools/exp_regroup.py [env] [margin]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
name = sys.argv[1] if len(sys.argv) > 1 else 'ant_heavenhell'
margin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.42
n = 1 << 20
env = envs.create(name, batch_size=n, episode_length=1000, auto_reset=True)
g = torch.Generator(device='cuda').manual_seed(1)
a = torch.rand((64, n, 8), device='cuda', generator=g) * 2 - 1
s = bench.stationary_state(torch, env, shard_keys(env, 0, n, 0, 1), a)

def timed(s, k, t0=0):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k): s = env.step(s, a[(t0 + i) % 64])
    e1.record(); e1.synchronize()
    return s, e0.elapsed_time(e1) / k

s, ms = timed(s, 50); print(f'{name}: stationary, env order {ms:.4f} ms/step')
p = env.params
nw = p.num_walls
lo = torch.tensor([[p.wall_lo[w][0], p.wall_lo[w][1]] for w in range(nw)], device='cuda')
hi = torch.tensor([[p.wall_hi[w][0], p.wall_hi[w][1]] for w in range(nw)], device='cuda')

def flags(s):
    pos = s.qp.pos                                   # [N, nb, 3]
    legs = pos[:, [2, 4, 6, 8], :2]                  # lower-leg centres
    d = torch.maximum(torch.maximum(lo[None, None] - legs[:, :, None], legs[:, :, None] - hi[None, None]), torch.zeros((), device='cuda'))
    dist = d.norm(dim=-1).amin(dim=(1, 2)) if nw else torch.full((n,), 1e9, device='cuda')
    return (dist < margin) | (pos[:, 0, 2] < 0.27)

def permute(s, order):
    for k, v in s.buf.items():
        if v is None or k == 'acc': continue
        if v.shape[0] == n: v.copy_(v[order].clone())
        elif v.dim() >= 2 and v.shape[1] == n: v.copy_(v[:, order].clone())
        else: raise RuntimeError(k)
    return s

f = flags(s)
print(f'flagged envs {f.float().mean().item():.3f}; warps with a flagged env, env order {f.view(-1, 8).any(1).float().mean().item():.3f}')
order = torch.argsort((~f).to(torch.uint8), stable=True)
s = permute(s, order)
f2 = flags(s)
print(f'warps with a flagged env after grouping {f2.view(-1, 8).any(1).float().mean().item():.3f}')
t = 50
for k in (10, 10, 20, 40, 80, 160):
    s, ms = timed(s, k, t); t += k
    f3 = flags(s)
    print(f'grouped, steps +{t - 50 - k}..+{t - 50}: {ms:.4f} ms/step; warps with a flagged env now {f3.view(-1, 8).any(1).float().mean().item():.3f}')
# interleaved placement: flagged warps spread evenly instead of first
