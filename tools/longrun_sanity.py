import sys, torch
sys.path.insert(0, '.')
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
for name in ('ant_heavenhell', 'ant_tag', 'ant_gather', 'ant'):
    n = 1 << 18
    env = envs.create(name, batch_size=n, episode_length=1000, auto_reset=True, eval_metrics=True)
    s = env.reset(shard_keys(env, 0, n, 0, 1))
    g = torch.Generator(device='cuda').manual_seed(3)
    a = torch.rand((64, n, 8), device='cuda', generator=g) * 2 - 1
    for t in range(4000):
        s = env.step(s, a[t % 64])
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(s.obs).all()) and bool(torch.isfinite(s.buf['qp']).all()) and bool(torch.isfinite(s.reward).all())
    acc = s.buf['acc'].tolist()
    print(name, 'finite', ok, 'episodes', acc[0], 'mean return', acc[1] / max(acc[0], 1), 'mean length', acc[2] / max(acc[0], 1), 'steps range', float(s.buf['steps'].min()), float(s.buf['steps'].max()), flush=True)
