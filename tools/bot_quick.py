import sys, time; sys.path.insert(0,'.')
import numpy as np
from motcpp_b200 import _lib, api, synth
CLI = dict(track_high_thresh=0.6, track_low_thresh=0.1, new_track_thresh=0.7, track_buffer=30, match_thresh=0.8,
           proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30, fuse_first_associate=0, with_reid=1)
S,T=32,12
base=[synth.embeddings_stream(s,n_frames=T) for s in range(2)]
dets=np.stack([base[s%2][0] for s in range(S)],1); embs=np.stack([base[s%2][1] for s in range(S)],1)
cnt=np.full((T,S),dets.shape[2],np.int32)
eng=api.Engine(_lib.TRACKER_BOTSORT,S,2048,1024,emb_dim=512,**CLI)
print(eng.info())
h=T//2
t0=time.time(); out,no=eng.update(dets[:h],cnt[:h],ld_out=2048,embs=embs[:h]); eng.check(); t1=time.time()
out,no=eng.update(dets[h:],cnt[h:],ld_out=2048,embs=embs[h:]); eng.check(); t2=time.time()
print("C3 1024x1024x512: %d streams, second half %.3fs -> %.0f frames/s e2e (pageable host memory)"%(S,t2-t1,S*(T-h)/(t2-t1)), no[-1,:4], eng.header(0)[:14])
