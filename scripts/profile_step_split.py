"""cProfile of MOT.step on the bench scene, split into detector frames and tracking-only frames (host-side cost per
frame; GPU waits show up as synchronize / fetch).  Prints wall ms per frame kind and the top host functions of each."""
import cProfile, os, pstats, sys, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fastmot_b200 import MOT

from types import SimpleNamespace
N = 111
c = bench.CONFIGS[int(os.environ.get("FM_CONFIG", 3))]
scene = bench.make_scene(c, 0)
frames = [scene.frame(t) for t in range(N)]
mot = MOT(scene.size, detections_override=bench.det_override(scene), **bench._cfg(c, SimpleNamespace(p5_input=896)))
mot.reset(1 / 30.)
mot.extractors[0]._engine(c["n"])
dev_frames = [torch.as_tensor(f).cuda() for f in frames]
for f in dev_frames[:11]:
    mot.step(f)
torch.cuda.synchronize()
prof = {True: cProfile.Profile(), False: cProfile.Profile()}
wall = {True: [], False: []}
for f in dev_frames[11:N]:
    det = mot.frame_count % mot.detector_frame_skip == 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prof[det].enable()
    mot.step(f)
    prof[det].disable()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    wall[det].append((t1 - t0, t2 - t0))
for det in (False, True):
    w = sorted(wall[det])
    med = w[len(w) // 2]
    print(f"==== {'detector' if det else 'tracking-only'} frames: n={len(w)} median host-return {med[0]*1e3:.3f} ms, "
          f"drained {med[1]*1e3:.3f} ms")
    for key in ("tottime", "cumulative"):
        s = io.StringIO()
        pstats.Stats(prof[det], stream=s).sort_stats(key).print_stats(30)
        print("\n".join(l[:160] for l in s.getvalue().splitlines()[4:]))
