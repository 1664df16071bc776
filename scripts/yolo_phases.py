"""Timeline of the TMA conv launches inside one detector forward (graph replay, PDL chain as in production):
CTA (0,0,0) of every conv_tma launch stamps %globaltimer at entry / after griddepcontrol.wait / accumulator complete /
exit.  Prints per launch: grid, K slices, time spent waiting for the previous layer, main loop, epilogue, and the gap
between the previous launch's exit and this launch's release.
usage: python scripts/yolo_phases.py [model]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastmot_b200 import _lib, models  # noqa: E402
from fastmot_b200.engine import build_yolo_engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'YOLOv4CSP'
lib = _lib.require_device()
lib.fm_conv_tma_set_debug.argtypes = [C.c_void_p]
eng = build_yolo_engine(models.YOLO.get_model(name), use_graph=True)
x = torch.rand(*eng.inp.shape, device='cuda').half()
for _ in range(5):
    eng.forward(x)
torch.cuda.synchronize()
dbg = torch.zeros(512 * 8, dtype=torch.int64, device="cuda")
lib.fm_conv_tma_set_debug(C.c_void_p(dbg.data_ptr()))
eng.forward(x)
torch.cuda.synchronize()
lib.fm_conv_tma_set_debug(None)
d = dbg.cpu().numpy().reshape(512, 8)
d = d[d[:, 0] > 0]
d = d[d[:, 0].argsort()]
t0 = d[0, 0]
print(f"{name}: {len(d)} conv_tma launches, first entry -> last exit {(d[:, 3].max() - t0) / 1e3:.1f} us")
print(f"{'#':>3s} {'grid':>12s} {'nk':>4s} {'bn/ns':>6s} {'entry':>8s} {'waited':>7s} {'main':>6s} {'epi':>6s} {'gap':>6s}"
      f" | epi = {'stage':>6s} {'rows':>6s} {'tail':>6s}")
prev_exit = None
tot = {"waited": 0.0, "main": 0.0, "epi": 0.0, "gap": 0.0}
for i, r in enumerate(d):
    gx, gy, gz, nk = r[4] >> 40, (r[4] >> 24) & 0xffff, (r[4] >> 16) & 0xff, r[4] & 0xffff
    bn, ns = r[5] >> 32, r[5] & 0xffffffff
    waited, main, epi = (r[1] - r[0]) / 1e3, (r[2] - r[1]) / 1e3, (r[3] - r[2]) / 1e3
    gap = (r[1] - prev_exit) / 1e3 if prev_exit is not None else 0.0
    prev_exit = r[3]
    tot["waited"] += waited; tot["main"] += main; tot["epi"] += epi; tot["gap"] += gap
    print(f"{i:3d} {f'{gx}x{gy}x{gz}':>12s} {nk:4d} {f'{bn}/{ns}':>6s} {(r[0] - t0) / 1e3:8.1f} {waited:7.1f} {main:6.1f} {epi:6.1f} {gap:6.1f}"
          f" | {(r[6] - r[2]) / 1e3:12.1f} {(r[7] - r[6]) / 1e3:6.1f} {(r[3] - r[7]) / 1e3:6.1f}")
print("sums (us):", {k: round(v, 1) for k, v in tot.items()})
