import sys, time, torch
sys.path.insert(0, '.')
from po_brax_b200.host import HostStepper
from po_brax_b200.parallel import shard_keys
n = 1 << 20
for chunks in (4, 8, 16, 32, 64):
    hs = HostStepper('ant_heavenhell', n, chunks=chunks, episode_length=1000, auto_reset=True)
    keys = torch.cat([shard_keys(e, c * hs.m, hs.m, 0, 1) if False else torch.randint(0, 2**31 - 1, (hs.m, 2), dtype=torch.int32, device='cuda') for c, e in enumerate(hs.envs)])
    hs.reset(keys)
    hs.action_host.uniform_(-1, 1)
    for _ in range(3): hs.step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 20
    for _ in range(K): hs.step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f'chunks {chunks}: {dt * 1e3:.3f} ms/step, {n / dt:.4e} env-steps/s', flush=True)
    del hs
