"""Timing probe for the DeepOC-SORT engine (development aid): one stable 1024-object scene, variants of the appearance path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from motcpp_b200 import _lib, api, synth, build

build.build(); _lib.require_gpu()
lib = _lib.load(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
DOC = dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, delta_t=3, inertia=0.2, w_association_emb=0.5,
           alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=0, aw_off=0, q_xy_scaling=0.01, q_s_scaling=0.0001)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 148
for dim, over in ((512, {}), (512, {"aw_off": 1}), (512, {"embedding_off": 1}), (64, {}), (64, {"w_association_emb": 0.0})):
    d, e = synth.embeddings_stream(1, n_frames=12, dim=dim)
    dets = torch.from_numpy(d).to(dev)[:, None].expand(-1, S, -1, -1).contiguous()
    embs = torch.from_numpy(e).to(dev)[:, None].expand(-1, S, -1, -1).contiguous()
    counts = torch.full((12, S), 1024, dtype=torch.int32, device=dev)
    off = over.get("embedding_off", 0)
    eng = api.Engine(_lib.TRACKER_DEEPOCSORT, S, 3072, 2048, emb_dim=0 if off else dim, **{**DOC, **over})
    out = torch.empty((4, S, 3072, 8), device=dev); n_out = torch.empty((4, S), dtype=torch.int32, device=dev)
    def run(f0, nf):
        api.check(lib.mot_engine_update_device_embs(eng._h, nf, dets[f0:].data_ptr(), counts[f0:].data_ptr(), 1024,
                                                    None if off else embs[f0:].data_ptr(), out.data_ptr(), n_out.data_ptr(), 3072, st))
    run(0, 4); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(4, 4); b.record(); torch.cuda.synchronize()
    print(dim, over, "ms/frame", a.elapsed_time(b) / 4, "hdr", eng.header(0)[:15].tolist(), flush=True)
    eng.close()
