"""Where the per-step cost growth comes from: after K steps, how many envs are non-finite, have the torso on the
ground, or have a body within reach of a wall.   python tools/diag_phases.py [env ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
n = 1 << 18
g = torch.Generator(device='cuda').manual_seed(1)
a = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
for name in sys.argv[1:] or ['ant', 'ant_heavenhell', 'ant_tag', 'ant_gather']:
    env = envs.create(name, batch_size=n, eval_metrics=True)
    s = env.reset(shard_keys(env, 0, n, 0, 1))
    p = env.params
    for t in range(446):
        if t in (0, 5, 45, 245, 445):
            q = s.qp
            pos = q.pos[:, :9]
            finite = torch.isfinite(pos).all(-1).all(-1) & torch.isfinite(q.vel[:, :9]).all(-1).all(-1)
            tz = pos[:, 0, 2]
            msg = f'{name} t={t}: non-finite {1 - finite.float().mean():.5f}, torso z<0.25 {(tz < 0.25).float().mean():.4f}, z in [0.2,0.25) {((tz < 0.25) & (tz >= 0.2)).float().mean():.4f}'
            if name != 'ant' and p.num_walls:
                lo = torch.tensor([[p.wall_lo[w][0], p.wall_lo[w][1]] for w in range(p.num_walls)], device='cuda')
                hi = torch.tensor([[p.wall_hi[w][0], p.wall_hi[w][1]] for w in range(p.num_walls)], device='cuda')
                xy = pos[:, :, None, :2]
                d = torch.maximum(torch.maximum(lo[None, None] - xy, xy - hi[None, None]), torch.zeros((), device='cuda'))
                dist = d.norm(dim=-1).min(dim=-1).values                 # [n, 9]
                msg += (f', torso<0.25 of wall {(dist[:, 0] < 0.25).float().mean():.4f}, some Aux<0.22 {(dist[:, 1::2] < 0.22).any(-1).float().mean():.4f}'
                        f', some foot<0.37 {(dist[:, 2::2] < 0.37).any(-1).float().mean():.4f}, |torso xy| max {pos[:, 0, :2].abs().max():.1f}')
            sp = q.vel[:, :9].norm(dim=-1).max(-1).values
            msg += f', max body speed median {sp.median():.2f} p99 {sp.quantile(0.99):.1f} max {sp.max():.1f}'
            print(msg, flush=True)
        s = env.step(s, a[t % 4])
    print(name, 'acc', dict(zip(env.ACC_NAMES, s.buf['acc'].tolist())), flush=True)
    del env, s
