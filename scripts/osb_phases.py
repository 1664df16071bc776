"""Phase stamps of one fm_osb_streams CTA (clock64 at the level boundaries): where a level's time goes.
usage: python scripts/osb_phases.py [w mid h cin n]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastmot_b200 import _lib  # noqa: E402
from test_gpu_osnet_fused import run_osb_streams, _random_block  # noqa: E402

w, mid, h, cin, n = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (32, 64, 64, 64, 200)
lib = _lib.require_device()
dbg = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.fm_osb_set_debug.argtypes = [C.c_void_p]
w1, b1, pws, dws = _random_block(cin, mid, 1)
x = (torch.randn(n, h, w, cin).abs() * 0.7).half().cuda()
run_osb_streams(x, w1, b1, pws, dws)            # warm-up
lib.fm_osb_set_debug(C.c_void_p(dbg.data_ptr()))
run_osb_streams(x, w1, b1, pws, dws)
lib.fm_osb_set_debug(None)
d = dbg.cpu().numpy()
t0 = d[0]
print(f"geometry w={w} mid={mid} h={h} cin={cin} n={n};  conv1 done (control) at {d[1] - t0} cycles")
names = ["ctl:pw_full", "ctl:act_ready", "ctl:mma_issued", "ctl:pw_empty", "cmp:start", "cmp:acc_full0", "cmp:epi_done",
         "cmp:bar", "cmp:dw_done"]
print("lvl " + " ".join(f"{k:>14s}" for k in names))
for lvl in range(10):
    row = d[16 + lvl * 16: 16 + lvl * 16 + 9] - t0
    print(f"{lvl:3d} " + " ".join(f"{int(v):14d}" for v in row))
