import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import envs as oenvs, threefry as tf
from tests import _parity as P
from po_brax_b200 import envs
np.set_printoptions(precision=6, suppress=True, linewidth=200)
kind='ant'; n=128
keys=P.keys_for(n,0)
oenv=oenvs.ENVS[kind](); s=oenv.reset(keys)
env=envs.create(kind,batch_size=n,auto_reset=False)
rng=tf.prng_key(1)
oenv.sys.track_margin=True
for t in range(14):
    rng,a=P.actions_for(rng,n)
    cs=env.state_from_qp(P.qp_to_torch(s.qp))
    oenv.sys.margin=None
    nxt=oenv.step(oenvs.State(s.qp.copy(),s.obs,s.reward,s.done,dict(s.metrics),dict(s.info)),a)
    got=env.step(cs,torch.as_tensor(a,device='cuda'))
    gp=P.t2n(got.qp.pos); err=np.abs(gp-nxt.qp.pos).reshape(n,-1).max(1)
    bad=np.nonzero(err>1e-4)[0]
    for e in bad:
        print('t',t,'env',e,'err',err[e],'margin',oenv.sys.margin[e])
        print(' torso before',s.qp.pos[e,0],'rot',s.qp.rot[e,0])
        print(' torso after oracle',nxt.qp.pos[e,0],'cuda',gp[e,0])
        print(' pos diff per body', np.abs(gp[e]-nxt.qp.pos[e]).max(1))
        print(' vel diff per body', np.abs(P.t2n(got.qp.vel)[e]-nxt.qp.vel[e]).max(1))
        go=P.t2n(got.obs)[e]; print(' cfrc vel cuda', go[27:57].reshape(10,3)[:9].round(4).tolist()); print(' cfrc vel orac', nxt.obs[e,27:57].reshape(10,3)[:9].round(4).tolist())
        print(' foot z before', s.qp.pos[e,[2,4,6,8],2], 'lower rot', s.qp.rot[e,2])
    s=nxt
