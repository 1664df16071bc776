"""Runs the OSNet x1.0 engine on 224 synthetic crops a few times (ncu target for the ReID stack)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmot_b200.engine import OSNetEngine
eng = OSNetEngine(1.0, max_batch=224, use_graph=False)
eng.load_nhwc8(torch.randn(224, 256, 128, 8, device='cuda').half())
for _ in range(3):
    eng.forward()
torch.cuda.synchronize()
print("done", eng.n_tc, eng.n_simt)
