#!/usr/bin/env python
"""Minimal driver for ncu: reset + a few fused steps of one env family (no timing, no extras).
    python tools/profile_step.py --env ant_heavenhell --envs 1048576 --steps 6 [--spread 60]
--spread K: run K un-profiled... no: steps before the profiled window so envs are spread over episode phases."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from po_brax_b200 import envs  # noqa: E402
from po_brax_b200.parallel import shard_keys  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--env', default='ant_heavenhell')
ap.add_argument('--envs', type=int, default=1 << 20)
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--period', type=int, default=64, help='length of the cyclic i.i.d. action sequence')
ap.add_argument('--stationary', action='store_true',
                help="bench.py's regime: own episode age per env, then --steps steps (use >= 1000 and profile the last)")
a = ap.parse_args()
env = envs.create(a.env, batch_size=a.envs, episode_length=1000, auto_reset=True, eval_metrics=True)
state = env.reset(shard_keys(env, 0, a.envs, 0, 1))
g = torch.Generator(device='cuda').manual_seed(1234)
acts = torch.rand((a.period, a.envs, 8), device='cuda', generator=g) * 2 - 1
if a.stationary:
    state.buf['steps'].copy_(torch.randint(0, 1000, (a.envs,), device='cuda', generator=g).float())
for i in range(a.steps):
    state = env.step(state, acts[i % a.period])
torch.cuda.synchronize()
print('ok', float(state.reward.sum()))
