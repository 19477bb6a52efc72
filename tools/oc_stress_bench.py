"""OC-SORT small-shape throughput on the parity-stress streams (lots of twin ties -> exact LAPJV re-solves)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from motcpp_b200 import _lib, api, synth
OC = dict(det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3, inertia=0.2,
          use_byte=0, q_xy_scaling=0.01, q_s_scaling=0.0001)
S, T = 592, 300
base = [synth.stress_stream(s, n_frames=T) for s in range(16)]
dets = np.stack([base[s % 16][0] for s in range(S)], 1)
cnt = np.stack([base[s % 16][1] for s in range(S)], 1).astype(np.int32)
lib = _lib.load()
dev = torch.device("cuda", 0)
d, c = torch.from_numpy(dets).to(dev), torch.from_numpy(cnt).to(dev)
out = torch.empty((T, S, 256, 8), device=dev); no = torch.empty((T, S), dtype=torch.int32, device=dev)
eng = api.Engine(_lib.TRACKER_OCSORT, S, 256, 64, **OC)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    eng.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    api.check(lib.mot_engine_update_device(eng._h, T, d.data_ptr(), c.data_ptr(), dets.shape[2], out.data_ptr(), no.data_ptr(), 256, st))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
eng.check()
print("OC-SORT stress streams: %d streams x %d frames in %.1f ms -> %.2f M frames/s" % (S, T, dt * 1e3, S * T / dt / 1e6))
