import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import envs as oenvs, threefry as tf, brax_v1 as bx
from tests import _parity as P
from po_brax_b200 import envs
np.set_printoptions(precision=6, suppress=True, linewidth=200)
kind='ant'; n=128
keys=P.keys_for(n,0)
oenv=oenvs.ENVS[kind](); s=oenv.reset(keys)
env1=envs.create(kind,batch_size=n,auto_reset=False,sys_dt=0.005,sys_substeps=1)
rng=tf.prng_key(1)
for t in range(13):
    rng,a=P.actions_for(rng,n)
    if t<12:
        s=oenv.step(s,a); continue
    qp=s.qp.copy()
    e=43
    for sub in range(10):
        cs=env1.state_from_qp(P.qp_to_torch(qp))
        got=env1.step(cs,torch.as_tensor(a,device='cuda'))
        qn,_,_=oenv.sys.substep(qp,a)
        ja,jv=oenv.sys.angle_vel(qn)
        print('sub',sub,'psi',ja[e],'ang torso',qn.ang[e,0])
        for nm in ('pos','rot','vel','ang'):
            d=np.abs(P.t2n(getattr(got.qp,nm))[e]-getattr(qn,nm)[e]).max(1)
            print('   ',nm,d[:9])
        qp=qn
