import ctypes, numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
from frankenpaxos_b200 import Engine, traces as T
import bench
cfg=bench.CFG; n=1<<20
eng=Engine(slot_capacity=8*n, max_batch=3*n, **cfg)
L=eng._L
dev=torch.device('cuda')
def td(x): return torch.from_numpy(x.view(np.int32).reshape(len(x),-1)).to(dev)
outp=torch.empty((3*n,4),dtype=torch.int32,device=dev); outn=torch.empty((3*n,2),dtype=torch.int32,device=dev); outc=torch.empty((3*n,2),dtype=torch.int32,device=dev)
for s in range(5):
    a,p,b=T.workload(s,cfg,n,slot0=s*n)
    da,dp,db=td(a),td(p),td(b)
    eng.proxyleader_arm_dev(da.data_ptr(),n)
    eng.acceptor_phase2a_dev(dp.data_ptr(),3*n,outp.data_ptr(),outn.data_ptr())
    eng.proxyleader_phase2b_dev(db.data_ptr(),3*n,outc.data_ptr())
    eng.replica_chosen_last_dev(outc.data_ptr()); eng.chosen_watermark_dev()
    eng.sync()
    buf=(ctypes.c_ulonglong*(148*8*4))(); grid=ctypes.c_int()
    L.fpx_debug_cta_trace.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.fpx_debug_cta_trace(eng.h, buf, ctypes.byref(grid))
    G=grid.value
    t=np.array(buf[:G*4],dtype=np.uint64).reshape(G,4)
    sm=(t[:,0]&np.uint64(1023)).astype(int); ts=(t>>np.uint64(10)).astype(np.int64)
    t0=ts[:,1].min()
    p2=(ts[:,2]-ts[:,1])/1e3   # pass2 duration per CTA
    p1end=(ts[:,0]-ts[:,0].min())/1e3
    print('grid',G,'pass2 us: min %.1f med %.1f p90 %.1f max %.1f'%(p2.min(),np.median(p2),np.percentile(p2,90),p2.max()), ' pass1-end spread %.1f'%p1end.max())
    if s==4:
        order=np.argsort(p2)
        print('slowest 12 CTAs (block, sm, us):',[(int(i),int(sm[i]),round(float(p2[i]),1)) for i in order[-12:]])
        print('fastest 12 CTAs:',[(int(i),int(sm[i]),round(float(p2[i]),1)) for i in order[:12]])
        # by SM
        bysm={}
        for i in range(G): bysm.setdefault(sm[i],[]).append(p2[i])
        avg=np.array([np.mean(v) for k,v in sorted(bysm.items())]); cnt=np.array([len(v) for k,v in sorted(bysm.items())])
        print('CTAs per SM: min %d max %d; per-SM mean pass2: min %.1f max %.1f'%(cnt.min(),cnt.max(),avg.min(),avg.max()))
        print('corr(block index, pass2) = %.3f'%np.corrcoef(np.arange(G),p2)[0,1])
        # quartiles by block index
        q=np.array_split(p2,8); print('mean pass2 by block-index octile:',[round(float(x.mean()),1) for x in q])
