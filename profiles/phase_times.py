import ctypes, numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
from frankenpaxos_b200 import Engine, traces as T
import bench
cfg=bench.CFG; n=1<<20
eng=Engine(slot_capacity=8*n, max_batch=3*n, **cfg)
L=eng._L
L.fpx_debug_phase_times.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
dev=torch.device('cuda')
def td(x): return torch.from_numpy(x.view(np.int32).reshape(len(x),-1)).to(dev)
outp=torch.empty((3*n,4),dtype=torch.int32,device=dev); outn=torch.empty((3*n,2),dtype=torch.int32,device=dev); outc=torch.empty((3*n,2),dtype=torch.int32,device=dev)
for s in range(6):
    a,p,b=T.workload(s,cfg,n,slot0=s*n)
    da,dp,db=td(a),td(p),td(b)
    eng.proxyleader_arm_dev(da.data_ptr(),n)
    if '--no-acceptor' not in sys.argv: eng.acceptor_phase2a_dev(dp.data_ptr(),3*n,outp.data_ptr(),outn.data_ptr())
    eng.proxyleader_phase2b_dev(db.data_ptr(),3*n,outc.data_ptr())
    eng.replica_chosen_last_dev(outc.data_ptr()); eng.chosen_watermark_dev()
    r=eng.sync(check=False)
    ta=(ctypes.c_ulonglong*8)(); tt=(ctypes.c_ulonglong*8)()
    L.fpx_debug_phase_times(eng.h, ta, tt)
    ta=np.array(ta[:6],dtype=np.int64); tt=np.array(tt[:6],dtype=np.int64)
    print('acceptor phases us (pass1, bar1, carry, pass2, bar2):', np.diff(ta)/1e3, ' tally (A, bar1, B, bar2, C):', np.diff(tt)/1e3, r.n_chosen)
