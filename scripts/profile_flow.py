"""Phase timing of the camera-motion RANSAC kernel on the bench scene (debug hook fm_klt_set_debug) and wall /
device time of the tracker stages of a KLT-only frame."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from fastmot_b200 import MOT, _lib

lib = _lib.require_device()
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.fm_klt_set_debug.argtypes = [C.c_void_p]
lib.fm_klt_set_debug(C.c_void_p(dbg.data_ptr()))
scene, frames = bench.make_frames(0, 26)
mot = MOT(scene.size, detections_override=bench.det_override(scene), **bench._cfg())
mot.reset(1 / 30.)
mot.extractors[0]._engine(bench.N_OBJECTS)
dev_frames = [torch.as_tensor(f).cuda() for f in frames]
names = ["compact", "ransac loop", "inlier compaction", "normal eq", "jacobi+denorm", "LM", ]
rows = []
for i, f in enumerate(dev_frames):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mot.step(f)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    t = dbg.cpu().numpy()
    if i >= 6 and t[0] > 0:
        ph = [(t[k + 1] - t[k]) / 1e3 for k in range(6)]
        rows.append(ph + [float(t[8]), float(t[9]), float(t[10]), wall, float(mot.tracker.flow.rounds_last)])
    dbg.zero_()
r = np.array(rows)
print("homography phases (us, mean over frames):")
for k, nme in enumerate(names):
    print(f"  {nme:20s} {r[:, k].mean():8.1f}")
print(f"  ransac iterations {r[:, 6].mean():.1f}   matches {r[:, 7].mean():.0f}   inliers {r[:, 8].mean():.0f}")
print(f"  frame wall ms: mean {r[:, 9].mean():.3f} min {r[:, 9].min():.3f} max {r[:, 9].max():.3f}; affine rounds {r[:, 10].mean():.1f}")
print("per-frame wall ms:", np.round(r[:, 9], 2).tolist())
