"""Micro-benchmark of the conv kernels on OSNet / YOLO layer shapes (CUDA events, L2 flushed between launches)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmot_b200 import _lib
from fastmot_b200.devmem import ptr, stream_ptr
from fastmot_b200.engine import _conv_desc
lib = _lib.require_device()
WS = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
DBG = "--dbg" in sys.argv
dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
lib.fm_conv_set_debug.argtypes = [C.c_void_p]
if DBG:
    lib.fm_conv_set_debug(C.c_void_p(dbg.data_ptr()))
SHAPES = [  # n, h, w, cin, cout, k, stride
    (224, 64, 32, 64, 64, 1, 1), (224, 64, 32, 64, 256, 1, 1), (224, 64, 32, 256, 64, 1, 1), (224, 64, 32, 256, 256, 1, 1),
    (224, 32, 16, 96, 96, 1, 1), (224, 32, 16, 96, 384, 1, 1), (224, 16, 8, 128, 128, 1, 1), (224, 256, 128, 8, 64, 7, 2),
    (1, 80, 80, 128, 256, 3, 1), (1, 40, 40, 256, 512, 3, 1), (1, 20, 20, 512, 1024, 3, 1), (1, 160, 160, 64, 64, 3, 1),
    (1, 320, 320, 32, 64, 3, 2), (1, 640, 640, 8, 32, 3, 1),
]
ONLY = [int(a.split('=')[1]) for a in sys.argv if a.startswith('--only=')]
if ONLY:
    SHAPES = [SHAPES[i] for i in ONLY]
for (n, h, w, cin, cout, k, st) in SHAPES:
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    x = torch.randn(n, h, w, cin, device="cuda").half()
    wt = torch.randn(cout, k, k, cin, device="cuda").half()
    b = torch.zeros(cout, device="cuda")
    y = torch.zeros(n, ho, wo, cout, device="cuda").half()
    d = _conv_desc(n, h, w, cin, cin, 0, ho, wo, cout, cout, 0, k, st, pad, 5, ws=WS)
    ts = []
    for it in range(6):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.fm_conv2d_tc(C.byref(d), ptr(x), ptr(wt), ptr(b), None, ptr(y), stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[1:])[len(ts[1:]) // 2]
    flops = 2.0 * n * ho * wo * cout * cin * k * k
    byts = (x.numel() + y.numel() + wt.numel()) * 2
    if DBG:
        t = dbg.cpu().numpy().reshape(4096, 8)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        ph = (t[:, 1:7] - t[:, 0:6]).mean(0)
        print("   CTAs", len(t), "span(us)", (t[:, 6].max() - t0) / 1e3, "start spread(us)", (t[:, 0].max() - t0) / 1e3,
              "phases ns [alloc+sync, plan+prologue, mainloop, commit-wait, epilogue, dealloc]:", [int(x) for x in ph],
              "first tmem_ld32 ns:", int((t[:, 7] - t[:, 4]).mean()))
        dbg.zero_()
    print(f"{(n,h,w,cin,cout,k,st)}: {us:8.1f} us  {flops/us/1e6:8.1f} TFLOP/s  {byts/us/1e3:7.1f} GB/s(min traffic)")

# ---- depthwise 3x3 (OSNet Lite 3x3) ----
if not ONLY:
    for (n, h, w, c) in [(224, 64, 32, 64), (224, 32, 16, 96), (224, 16, 8, 128)]:
        x = torch.randn(n, h, w, c, device="cuda").half()
        wt = torch.randn(9, c, device="cuda").half()
        b = torch.zeros(c, device="cuda")
        y = torch.empty_like(x)
        ts = []
        for it in range(6):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.fm_dwconv3(ptr(x), ptr(wt), ptr(b), ptr(y), n, h, w, c, 5, stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts[1:])[len(ts[1:]) // 2]
        print(f"dw3x3 {(n,h,w,c)}: {us:8.1f} us   {2 * x.numel() * 2 / us / 1e3:7.1f} GB/s(min traffic)")
