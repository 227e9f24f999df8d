timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, cgmres_cpp_b200 as cg
from cgmres_cpp_b200.workloads import synthetic_batch
import time
for model,n in ((cg.MSD,1000),(cg.SEMIACTIVE,777),(cg.ARM,333),(cg.MSD,65536)):
    x0,p,u0=synthetic_batch(model,n,seed=3)
    outs={}
    for mode in (cg.MODE_FAST, cg.MODE_ONCHIP_EXACT):
        c=cg.BatchedCgmres(model,n=n,device=0,mode=mode)
        if p is not None: c.set_ptau_repeat(p)
        c.init_u0(u0); c.init_u0_newton(u0,x0,p,10)
        c.set_x(x0); t=time.time(); c.step_closed_loop(200); x=c.get_x(); dt=time.time()-t
        outs[mode]=(x,c.get_u(),c.get_status()); print(model,n,mode,'%.3f s'%dt, flush=True)
    d=np.abs(outs[cg.MODE_FAST][0]-outs[cg.MODE_ONCHIP_EXACT][0]).max()
    print('model',model,'n',n,'max |x_fast-x_exact| after 200 steps',d, 'status eq', np.array_equal(outs[1][2],outs[2][2]), flush=True)
PY
