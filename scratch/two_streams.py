import sys, torch
sys.path.insert(0,'.')
from po_brax_b200 import envs
from po_brax_b200.parallel import shard_keys
def run(name, parts, n=1<<20, kw={}):
    m=n//parts
    es=[envs.create(name,batch_size=m,**kw) for _ in range(parts)]
    ss=[e.reset(shard_keys(e,0,m,0,1)) for e in es]
    g=torch.Generator(device='cuda').manual_seed(1)
    a=torch.rand((4,m,8),device='cuda',generator=g)*2-1
    st=[torch.cuda.Stream() for _ in range(parts)]
    torch.cuda.synchronize()
    def step(i):
        for k in range(parts):
            with torch.cuda.stream(st[k]):
                ss[k]=es[k].step(ss[k],a[i%4])
    for i in range(10): step(i)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in st: s.wait_event(e0)
    for i in range(40): step(i)
    for s in st: torch.cuda.current_stream().wait_stream(s)
    e1.record(); e1.synchronize()
    print(name,kw,'parts',parts,'ms/step',e0.elapsed_time(e1)/40, flush=True)
for name,kw in (('ant',{}),('ant_gather',{}),('ant_heavenhell',{})):
    for parts in (1,2,4):
        run(name,parts,kw=kw)
