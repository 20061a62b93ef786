import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mllp_b200 as M
from oracle import pdhg_oracle as O
import torch
for name,K in (('afiro',500),('sc50a',300),('25fv47',300),('pilot87',200),('ken-18',100),('osa-60',100),('pds-20',100)):
    A,b,c=M.load_csr(name); m,n=A.shape
    lp=M.DeviceLP(A, A.data, m, n)
    print(name, lp.info())
    v=np.random.default_rng(0).standard_normal(n); w=np.random.default_rng(1).standard_normal(m)
    o1=lp.spmv(torch.tensor(v,device='cuda')).cpu().numpy(); r1=A@v
    o2=lp.spmv(torch.tensor(w,device='cuda'),trans=True).cpu().numpy(); r2=A.T@w
    print('  spmv rel err', np.linalg.norm(o1-r1)/np.linalg.norm(r1), np.linalg.norm(o2-r2)/np.linalg.norm(r2))
    s_gpu=lp.sigma_max(); s_cpu=O.power_iteration(A,50)
    print('  sigma', s_gpu, s_cpu)
    eta=0.9/s_cpu
    xo,yo=O.pdhg_run(A,b,c,np.zeros(n),np.zeros(m),eta,eta,K)
    obj,x,y,info=M.pdhg_linear_program(A,A.data,b,c,num_iters=K,tau=eta,sigma=eta,handle=lp)
    print('  parity x %.2e y %.2e'%(np.linalg.norm(x-xo)/max(np.linalg.norm(xo),1e-300), np.linalg.norm(y-yo)/max(np.linalg.norm(yo),1e-300)), 'obj',obj, c@xo)
    kk=O.kkt(A,b,c,xo,yo); print('  kkt diff', np.abs(kk-np.array([info[k] for k in M.linear_program_methods.SCALAR_NAMES[:10]])).max())
    # timing
    bt=torch.tensor(b,device='cuda'); ct=torch.tensor(c,device='cuda')
    for KK in (100,1000):
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        M.pdhg_linear_program(A,A.data,bt,ct,num_iters=KK,tau=eta,sigma=eta,handle=lp)
        e0.record(); M.pdhg_linear_program(A,A.data,bt,ct,num_iters=KK,tau=eta,sigma=eta,handle=lp); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1); print('  K=%d: %.3f ms  %.2f us/iter  %.1f GB/s algorithmic'%(KK,ms,ms*1e3/KK, lp.info()['bytes_per_iter']*KK/ms/1e6))
    # graph mode
    lpg=M.DeviceLP(A,A.data,m,n,flags=2)
    obj2,x2,y2,_=M.pdhg_linear_program(A,A.data,b,c,num_iters=K,tau=eta,sigma=eta,handle=lpg)
    print('  graph-mode bitwise equal:', np.array_equal(x2,x), np.array_equal(y2,y))
    torch.cuda.synchronize(); e0.record(); M.pdhg_linear_program(A,A.data,bt,ct,num_iters=960,tau=eta,sigma=eta,handle=lpg); e1.record(); torch.cuda.synchronize()
    print('  graph mode: %.2f us/iter'%(e0.elapsed_time(e1)*1e3/960))
    if name in ('afiro','sc50a'):
        t=time.time(); obj,x,y,info=M.solve_linear_program(A,A.data,b,c,handle=lp,tol=1e-6); 
        print('  solve', obj, {k:info[k] for k in ('iters','restarts','converged','rel_kkt')}, 'time',time.time()-t)
        xs,ys,ks,is_=O.pdhg_solve(A,b,c,np.zeros(n),np.zeros(m),info['eta']); print('  oracle solve', ks[0], is_)
