import sys, time; sys.path.insert(0,'.')
import numpy as np, ctypes as C
from motcpp_b200 import _lib, api, synth
OC=dict(det_thresh=0.2,max_age=30,max_obs=50,min_hits=3,iou_threshold=0.3,min_conf=0.1,delta_t=3,inertia=0.2,use_byte=0,q_xy_scaling=0.01,q_s_scaling=0.0001)
for (name,S,cap,dm,gen,T) in (("C4 2048x2048",148,3072,2048,lambda s:synth.ocsort_stream(s%4,n_frames=40),40),("C2-shape 256x512",296,1536,512,lambda s:synth.bytetrack_stream(s%8,n_frames=80,n_clutter=24,n_low=40,config=4),80)):
    base=[gen(s) for s in range(8)]
    dets=np.stack([base[s%len(base)] for s in range(S)],1)
    cnt=np.full((T,S),dets.shape[2],np.int32)
    eng=api.Engine(_lib.TRACKER_OCSORT,S,cap,dm,**OC)
    t0=time.time(); out,no=eng.update(dets[:T//2],cnt[:T//2],ld_out=cap)
    try: eng.check()
    except Exception as ex: print(name, ex, eng.header(0)[:14], no[:4,0]); continue
    t1=time.time()
    out,no=eng.update(dets[T//2:],cnt[T//2:],ld_out=cap); eng.check(); t2=time.time()
    print(name,"info",eng.info(),"first half %.3fs second half %.3fs -> %.0f frames/s e2e"%(t1-t0,t2-t1,S*(T-T//2)/(t2-t1)), "rows",no[-1,:4], eng.header(0)[:14])
    eng.close()
