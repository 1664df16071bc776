"""cProfile of MOT.step on the bench scene (host-side cost per frame; GPU waits show up as synchronize/fetch)."""
import cProfile, os, pstats, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fastmot_b200 import MOT

scene, frames = bench.make_frames(0, 61)
mot = MOT(scene.size, detections_override=bench.det_override(scene), **bench._cfg())
mot.reset(1 / 30.)
mot.extractors[0]._engine(bench.N_OBJECTS)
dev_frames = [torch.as_tensor(f).cuda() for f in frames]
for f in dev_frames[:11]:
    mot.step(f)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for f in dev_frames[11:61]:
    mot.step(f)
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(38)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
