"""Times one OSNet forward (CUDA events around the graph replay, L2 flushed between replays).
usage: python scripts/time_osnet.py [batch] [width]   (FM_OSB_FUSED / FM_OSB_WARPS select the kernel variants)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastmot_b200.engine import OSNetEngine  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
width = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
eager = "--eager" in sys.argv
eng = OSNetEngine(width, max_batch=batch, use_graph=not eager)
eng.load_nhwc8(torch.randn(batch, 256, 128, 8, device='cuda').half() * 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2 if eager else 5):
    eng.forward()
torch.cuda.synchronize()
ts = []
for _ in range(2 if eager else 20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.forward()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"osnet x{width} batch {batch} fused={os.environ.get('FM_OSB_FUSED', '1')} warps={os.environ.get('FM_OSB_WARPS', '8')}"
      f" n_osb={eng.n_osb} kernels={eng.kernels_per_replay()}: median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}"
      f"  layer_bytes {eng.layer_bytes / 1e9:.2f} GB")
