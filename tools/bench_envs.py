"""Per-env step time (1 Mi envs, 40 timed steps after 10 warm-up steps), walls on / off (test hook):
    python tools/bench_envs.py [env ...]          (POBRAX_LIB=<tuning build> to time another build)
"""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
n=1<<20
cases=(('ant_tag',{}),('ant_tag',{'walls':False}),('ant_heavenhell',{}),('ant_heavenhell',{'walls':False}),('ant_gather',{}),('ant_gather',{'walls':False}),('ant',{}))
if len(sys.argv)>1: cases=[(a,{}) for a in sys.argv[1:]]
for name,kw in cases:
    env=envs.create(name,batch_size=n,**kw)
    s=env.reset(shard_keys(env,0,n,0,1))
    g=torch.Generator(device='cuda').manual_seed(1)
    a=torch.rand((4,n,8),device='cuda',generator=g)*2-1
    for i in range(10): s=env.step(s,a[i%4])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): s=env.step(s,a[i%4])
    e1.record(); e1.synchronize()
    print(name,kw,'ms/step',e0.elapsed_time(e1)/40)
    del env,s
