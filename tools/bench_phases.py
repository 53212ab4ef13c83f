"""Step time by episode phase at 1 Mi envs, all envs reset together: the first 5 steps after the reset, steps 5-45,
45-245, 245-445, 445-645, 645-845, 845-995 (the batch truncates at 1000); then the stationary regime of bench.py
(per-env episode ages + 1000-step pre-roll).
    python tools/bench_phases.py [--period P] [--quick] [env ...]
--period: length of the cyclically reused i.i.d. U(-1,1) action sequence. SURVEY 8(d) prescribes >= 64 (the default).
A SHORT cycle (round 1 used 4) is a periodic gait: the ants then walk metres and pile up along the walls, which is
where round 1's "cost grows with the time since reset" curve came from; i.i.d. actions barely move them."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
args = sys.argv[1:]
period, quick = 64, False
if '--period' in args:
    i = args.index('--period'); period = int(args[i + 1]); del args[i:i + 2]
if '--quick' in args:
    quick = True; args.remove('--quick')
n = 1 << 20
g = torch.Generator(device='cuda').manual_seed(1)
a = torch.rand((period, n, 8), device='cuda', generator=g) * 2 - 1


def timed(env, s, steps, t0):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): s = env.step(s, a[(t0 + i) % period])
    e1.record(); e1.synchronize()
    return s, e0.elapsed_time(e1) / steps


for name in args or ['ant', 'ant_heavenhell', 'ant_tag', 'ant_gather']:
    env = envs.create(name, batch_size=n)
    s = env.reset(shard_keys(env, 0, n, 0, 1))
    out, t = [], 0
    for steps in ((5, 40, 200) if quick else (5, 40, 200, 200, 200, 200, 150)):
        s, ms = timed(env, s, steps, t); t += steps
        out.append(ms)
    # stationary regime (bench.py stationary_state): own episode age per env, 1000-step pre-roll
    s = env.reset(shard_keys(env, 0, n, 0, 1))
    s.buf['steps'].copy_(torch.randint(0, 1000, (n,), device='cuda', generator=g).float())
    for i in range(1000): s = env.step(s, a[i % period])
    s, st = timed(env, s, 200, 1000)
    names = ('0-5', '5-45', '45-245', '245-445', '445-645', '645-845', '845-995')
    print(f'{name} (action period {period}): ' + ', '.join(f'{k} {v:.4f}' for k, v in zip(names, out)) +
          f' | stationary {st:.4f} ms/step', flush=True)
    del env, s
