import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastmot_b200 import _lib
from fastmot_b200.devmem import ptr, stream_ptr
from fastmot_b200.packing import pack_b_sw64
lib = _lib.require_device()
n = 2
rng = np.random.default_rng(7)
mode = sys.argv[1] if len(sys.argv) > 1 else "rand"
w7 = rng.normal(0, np.sqrt(2.0 / 147), (64, 7, 7, 3)).astype(np.float32)
if mode == "center":        # only the centre tap, channel 0 -> out = relu(x[2oy][2ox][0] * w)
    w7[:] = 0
    w7[:, 3, 3, 0] = 1.0
b7 = np.zeros(64, np.float32)
g = torch.Generator().manual_seed(1)
x = torch.randn(n, 256, 128, 3, generator=g).half()
xin = torch.zeros(n, 264, 136, 4, dtype=torch.float16)
xin[:, 4:-4, 4:-4, :3] = x
wk = np.zeros((64, 7, 8, 4), np.float32)
wk[:, :, 1:8, :3] = w7
img = torch.as_tensor(pack_b_sw64(wk.reshape(64, 224))).cuda()
out = torch.full((n, 64, 32, 64), float('nan'), dtype=torch.float16, device="cuda")
xin_d, b_d = xin.cuda(), torch.as_tensor(b7).cuda()
import ctypes as C
dbg = torch.zeros(128 * 64 + 7 * 2048, dtype=torch.float32, device="cuda")
lib.fm_osnet_stem_set_debug.argtypes = [C.c_void_p]
lib.fm_osnet_stem_set_debug(C.c_void_p(dbg.data_ptr()))
_lib.check(lib.fm_osnet_stem(ptr(xin_d), n, ptr(img), ptr(b_d), ptr(out), stream_ptr()), "stem")
torch.cuda.synchronize()
lib.fm_osnet_stem_set_debug(None)
d_ = dbg.cpu().numpy()
acc = d_[:128 * 64].reshape(128, 64)
# expected A (tile 0 of crop 0): A[m=(oyl,ox)][k=r*32+j*4+c] = xin[0][2*oy+1+r][2*ox+j][c]
xp = xin[0].float().numpy()
A = np.zeros((128, 224), np.float32)
for oyl in range(2):
    for ox_ in range(64):
        for r in range(7):
            A[oyl * 64 + ox_, r * 32:(r + 1) * 32] = xp[2 * oyl + 1 + r, 2 * ox_:2 * ox_ + 8, :].reshape(-1)
Wm = wk.reshape(64, 224).astype(np.float16).astype(np.float32)
want_acc = A @ Wm.T
print("acc max err", float(np.abs(acc - want_acc).max()), "acc[5,:4]", acc[5, :4], "want", want_acc[5, :4])
# un-swizzle the dumped A slices: slice ks = [128 rows][128 B]
raw = d_[128 * 64:].view(np.uint32).view(np.float16).reshape(7, 128, 4, 8)
Ag = np.zeros((128, 224), np.float32)
for ks in range(7):
    for m in range(128):
        for c in range(4):
            Ag[m, ks * 32 + c * 8: ks * 32 + c * 8 + 8] = raw[ks, m, c ^ ((m >> 1) & 3)]
print("A max err", float(np.abs(Ag - A).max()))
bad = np.argwhere(np.abs(Ag - A) > 1e-3)
print("bad A entries", len(bad), bad[:6].tolist())
if len(bad):
    m, k = bad[0]
    print("A row", m, "k", k, "got", Ag[m, k - k % 8:k - k % 8 + 8], "want", A[m, k - k % 8:k - k % 8 + 8])
wt = torch.as_tensor(w7).half().float().permute(0, 3, 1, 2).contiguous()
y = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt, torch.as_tensor(b7), stride=2, padding=3))
want = F.max_pool2d(y.half().float(), 3, 2, 1).permute(0, 2, 3, 1)
got = out.float().cpu()
d = (got - want).abs()
print("finite", bool(torch.isfinite(got).all()), "max err", float(d.max()), "want max", float(want.abs().max()))
print("err by pooled row (crop 0):", [round(float(d[0, r].max()), 3) for r in range(0, 64, 4)])
print("err by col:", [round(float(d[0, :, c].max()), 3) for r, c in enumerate(range(0, 32, 2))])
print("err by channel:", [round(float(d[0, :, :, c].max()), 3) for c in range(0, 64, 8)])
print("got[0,5,5,:4]", got[0, 5, 5, :4].tolist(), "want", want[0, 5, 5, :4].tolist())
print("got[0,0,0,:4]", got[0, 0, 0, :4].tolist(), "want", want[0, 0, 0, :4].tolist())
