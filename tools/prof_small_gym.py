import sys, cProfile, pstats, torch
sys.path.insert(0, '.')
from po_brax_b200 import envs
n = 16
e = envs.create_gym_env('ant', batch_size=n, seed=0, episode_length=20, eval_metrics=True, discount=0.99)
e.reset()
a = torch.rand((n, 8), device='cuda') * 2 - 1
for i in range(200): e.step(a)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(2000): e.step(a)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
