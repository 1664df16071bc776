"""Conv-stack executors that take the place of the reference's TensorRT engines
(fastmot/utils/inference.py:39-125): `YoloEngine` runs a Darknet layer list, `OSNetEngine` the OSNet op list,
both on NHWC fp16 device tensors through the C-ABI layer kernels (csrc/nn.cu, csrc/conv_tc.cu).

The launch sequence of a network is recorded once (static shapes) and replayed per call; with `use_graph=True`
the replay is captured into a CUDA graph so a 170-layer detector costs one launch.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .devmem import ptr, stream_ptr
from .models import darknet, osnet

_ACT = darknet.ACTS
ACT_AFTER_RESIDUAL = 0x100
IN_C_PAD = 8      # network inputs are NHWC with 3 real + 5 zero channels: one 16-byte chunk per pixel


class _Launch:
    """A recorded C-ABI call: function + fully bound arguments."""
    __slots__ = ("fn", "args", "what")

    def __init__(self, fn, args, what):
        self.fn, self.args, self.what = fn, args, what

    def __call__(self, s):
        rc = self.fn(*self.args, s)
        if rc:
            _lib.check(rc, self.what)


def _conv_desc(n, hi, wi, cin, cin_stride, cin_off, ho, wo, cout, cout_stride, cout_off, k, stride, pad, act, ws=None):
    d = _lib.FmConvDesc()
    if ws is not None:       # fp32 split-K scratch (a torch uint8 tensor owned by the caller)
        d.ws, d.ws_bytes = ws.data_ptr(), ws.numel()
    d.n, d.hi, d.wi, d.cin, d.cin_stride, d.cin_offset = n, hi, wi, cin, cin_stride, cin_off
    d.ho, d.wo, d.cout, d.cout_stride, d.cout_offset = ho, wo, cout, cout_stride, cout_off
    d.kh = d.kw = k
    d.stride, d.pad, d.act = stride, pad, act
    d.res_stride = d.res_offset = 0
    return d


class _Net:
    """Shared plumbing: recorded launches, optional CUDA-graph replay, conv dispatch (tcgen05 when supported)."""
    WS_BYTES = 32 << 20

    def __init__(self, use_tc=True, use_graph=False):
        self._lib = _lib.require_device()
        self.use_tc = use_tc
        self.use_graph = use_graph
        self.launches = []
        self._graph = None
        self._keep = []
        self.n_tc = self.n_simt = self.n_tma = 0
        self.use_tma = os.environ.get("FM_CONV_TMA", "1") != "0"     # A/B switch: 0 = every conv through conv_tc.cu
        self.layer_bytes = 0        # algorithmic HBM bytes of the conv / depthwise layers (each tensor moved once)
        self.dev = torch.device("cuda")
        # split-K scratch of the tcgen05 conv: one per engine, so engines on different streams never share partials
        self.ws = torch.empty(self.WS_BYTES, dtype=torch.uint8, device=self.dev)

    def _conv(self, desc, x, w, b, out, residual=None):
        lib = self._lib
        desc.ws, desc.ws_bytes = self.ws.data_ptr(), self.ws.numel()
        self._keep.append(desc)
        self.layer_bytes += 2 * (desc.n * desc.hi * desc.wi * desc.cin + desc.n * desc.ho * desc.wo * desc.cout
                                 * (2 if residual is not None else 1) + desc.kh * desc.kw * desc.cin * desc.cout)
        if self.use_tc and self.use_tma and lib.fm_conv2d_tma_supported(C.byref(desc)):
            fn = lib.fm_conv2d_tma          # TMA-fed, cluster split-K (csrc/conv_tma.cu)
            self.n_tc += 1
            self.n_tma += 1
        elif self.use_tc and lib.fm_conv2d_tc_supported(C.byref(desc)):
            fn = lib.fm_conv2d_tc
            self.n_tc += 1
        else:
            fn = lib.fm_conv2d_simt
            self.n_simt += 1
        self.launches.append(_Launch(fn, (C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(residual), ptr(out)), "conv"))

    def kernels_per_replay(self):
        extra = {'fm_channel_gate': 2, 'fm_channel_gate4': 2, 'fm_channel_gate4_pooled': 1, 'fm_osb_merge': 1}
        return sum(1 + extra.get(l.what, 0) for l in self.launches)

    def _add(self, fn_name, *args):
        self.launches.append(_Launch(getattr(self._lib, fn_name), args, fn_name))

    def warm(self, n=3):
        """Replays the recorded network n times (graph capture included) so the first timed call is steady state."""
        for _ in range(n):
            self.replay()
        torch.cuda.synchronize()

    def replay(self):
        _lib.count_graph_kernels(self.kernels_per_replay() if self.use_graph and self._graph is not None else 0)
        if self.use_graph:
            if self._graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    sp = stream_ptr()
                    for l in self.launches:      # warm-up outside capture
                        l(sp)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sp = stream_ptr()
                    for l in self.launches:
                        l(sp)
                self._graph = g
            self._graph.replay()
        else:
            sp = stream_ptr()
            for l in self.launches:
                l(sp)


class YoloEngine(_Net):
    """Darknet graph executor.  forward(inp NHWC8 fp16) -> list of head tensors [H, W, (5+C)*A] fp16."""
    heads_nhwc = True

    def __init__(self, layers, input_hw, weights, use_tc=True, use_graph=False):
        super().__init__(use_tc, use_graph)
        H, W = input_hw
        self.layers, self.shapes = darknet.infer_shapes(layers, 3, H, W)
        self.inp = torch.zeros(H, W, IN_C_PAD, dtype=torch.float16, device=self.dev)
        self.flops = darknet.count_flops(layers, 3, H, W)
        L = self.layers
        n = len(L)
        # ---- plan: which layers are written straight into a concat buffer ----
        home = {}       # layer -> (route index, channel offset)
        strided_ok = ('convolutional', 'maxpool', 'upsample')
        for i, l in enumerate(L):
            if l['type'] == 'route' and len(l['layers_abs']) > 1 and l.get('groups', 1) == 1:
                off = 0
                for s in l['layers_abs']:
                    if s not in home and L[s]['type'] in strided_ok:
                        home[s] = (i, off)
                    off += self.shapes[s][0]
        bufs = {}       # route index -> tensor

        def route_buf(i):
            if i not in bufs:
                c, h, w = self.shapes[i]
                bufs[i] = torch.zeros(h, w, c, dtype=torch.float16, device=self.dev)
            return bufs[i]

        self.views = []     # per layer: (tensor, c, c_stride, c_off, h, w)
        self.params = {}
        self.heads = []
        # shortcut layers (yolo2onnx.py:707-731) whose first operand is the convolution right before them and is read
        # by nothing else: the add moves into that conv's epilogue (residual after the activation), one launch less
        refs = {}
        for i, l in enumerate(L):
            if l['type'] == 'route':
                for s_ in l['layers_abs']:
                    refs.setdefault(s_, set()).add(i)
            elif l['type'] == 'shortcut':
                refs.setdefault(l['from_abs'], set()).add(i)
        fuse_sc = os.environ.get("FM_FUSE_SHORTCUT", "1") != "0"
        self.fused_shortcuts = set()
        for i, l in enumerate(L[:-1]):
            nx = L[i + 1]
            if fuse_sc and l['type'] == 'convolutional' and nx['type'] == 'shortcut' and i not in home and \
                    nx.get('activation', 'linear') == 'linear' and not refs.get(i) and nx['from_abs'] < i and \
                    self.shapes[nx['from_abs']] == self.shapes[i]:
                self.fused_shortcuts.add(i + 1)
        spp_prev = {}
        for i, l in enumerate(L):
            t = l['type']
            c, h, w = self.shapes[i]
            if i in home:
                ri, off = home[i]
                out = (route_buf(ri), c, self.shapes[ri][0], off, h, w)
            elif t in ('convolutional', 'maxpool', 'upsample', 'shortcut') and i not in self.fused_shortcuts:
                out = (torch.zeros(h, w, c, dtype=torch.float16, device=self.dev), c, c, 0, h, w)
            else:
                out = None
            src = self.views[i - 1] if i else (self.inp, IN_C_PAD, IN_C_PAD, 0, H, W)
            if t == 'convolutional':
                wt, bs = weights[i]
                k = l['size']
                cin = src[1]
                if i == 0:   # physical input has IN_C_PAD channels (the extra ones are zero)
                    wt = np.concatenate([wt, np.zeros(wt.shape[:3] + (IN_C_PAD - wt.shape[3],), np.float32)], -1)
                wd = torch.as_tensor(np.ascontiguousarray(wt)).to(self.dev).half().contiguous()
                bd = torch.as_tensor(bs).to(self.dev).float().contiguous()
                self.params[i] = (wd, bd)
                pad = k // 2 if l.get('pad', 0) else 0
                d = _conv_desc(1, src[4], src[5], cin, src[2], src[3], h, w, c, out[2], out[3], k, l.get('stride', 1),
                               pad, _ACT[l.get('activation', 'linear')])
                if i + 1 in self.fused_shortcuts:
                    b = self.views[L[i + 1]['from_abs']]
                    d.res_stride, d.res_offset = b[2], b[3]
                    self._conv(d, src[0], wd, bd, out[0], residual=b[0])
                else:
                    self._conv(d, src[0], wd, bd, out[0])
            elif t == 'maxpool':
                ksz, psrc = l['size'], src
                key = (src[0].data_ptr(), src[1], src[2], src[3])
                if l['stride'] == 1 and ksz % 2 == 1:
                    # SPP (5 / 9 / 13 on the same tensor): a k x k stride-1 max over a k0 x k0 max is the
                    # (k + k0 - 1) window, so 9 = 5 o 5 and 13 = 5 o 9 -- each pool reads 25 taps instead of 81 / 169
                    prev = spp_prev.get(key)
                    if prev is not None and ksz > prev[0]:
                        ksz, psrc = ksz - prev[0] + 1, prev[1]
                    spp_prev[key] = (l['size'], out)
                self._add('fm_maxpool', ptr(psrc[0]), ptr(out[0]), 1, psrc[4], psrc[5], psrc[1], psrc[2], psrc[3],
                          ksz, l['stride'], out[2], out[3])
            elif t == 'upsample':
                self._add('fm_upsample_copy', ptr(src[0]), ptr(out[0]), 1, src[4], src[5], src[1], src[2], src[3],
                          l['stride'], out[2], out[3])
            elif t == 'shortcut' and i in self.fused_shortcuts:
                out = self.views[i - 1]          # already holds conv + residual
            elif t == 'shortcut':
                a, b = self.views[i - 1], self.views[l['from_abs']]
                self._add('fm_add_act_strided', ptr(a[0]), a[2], a[3], ptr(b[0]), b[2], b[3], ptr(out[0]), out[2],
                          out[3], h * w, c, _ACT[l.get('activation', 'linear')])
            elif t == 'route':
                srcs = l['layers_abs']
                g = l.get('groups', 1)
                if len(srcs) == 1:
                    sv = self.views[srcs[0]]
                    cg = sv[1] // g
                    out = (sv[0], cg, sv[2], sv[3] + l.get('group_id', 0) * cg, sv[4], sv[5])
                else:
                    buf = route_buf(i)
                    off = 0
                    for s in srcs:
                        sv = self.views[s]
                        cs = sv[1] // g
                        if home.get(s, (None,))[0] != i:
                            self._add('fm_upsample_copy', ptr(sv[0]), ptr(buf), 1, sv[4], sv[5], cs, sv[2],
                                      sv[3] + l.get('group_id', 0) * cs, 1, c, off)
                        off += cs
                    out = (buf, c, c, 0, h, w)
            elif t == 'yolo':
                out = self.views[i - 1]
                assert out[2] == out[1] and out[3] == 0
                self.heads.append(out[0])
            self.views.append(out)

    def forward(self, inp):
        if inp.data_ptr() != self.inp.data_ptr():
            self.inp.copy_(inp)
        self.replay()
        return self.heads


def build_yolo_engine(model, weights=None, use_tc=True, use_graph=True, head_obj_bias=-5.0):
    """Engine for a `models.YOLO` descriptor.  Without a Darknet .weights file the weights are synthetic
    (seeded He-normal, detection-prior objectness bias) — there are no trained weights offline."""
    if isinstance(model.CFG, str) and model.CFG in darknet.BUILDERS:
        layers = darknet.BUILDERS[model.CFG](num_classes=model.NUM_CLASSES,
                                             anchors_per_head=len(model.ANCHORS[0]) // 2)
    else:
        _, layers = darknet.parse_cfg(open(model.CFG).read())
    if weights is None:
        if model.WEIGHTS_PATH:
            weights = darknet.load_weights(model.WEIGHTS_PATH, layers, 3)
        else:
            # FM_SYNTH_OBJ_BIAS: objectness prior of the SYNTHETIC heads (how many random-weight candidates pass
            # conf_thresh); bench.py lowers it for the big-input models so the candidate count stays realistic
            # FM_SYNTH_HEAD_GAIN: scale of the synthetic head weights (logit variance on real frames)
            head_obj_bias = float(os.environ.get("FM_SYNTH_OBJ_BIAS", head_obj_bias))
            weights = darknet.synthetic_weights(layers, 3, head_obj_bias=head_obj_bias,
                                                num_classes=model.NUM_CLASSES,
                                                head_gain=float(os.environ.get("FM_SYNTH_HEAD_GAIN", 1.0)))
    return YoloEngine(layers, model.INPUT_SHAPE[1:], weights, use_tc=use_tc, use_graph=use_graph)


class OSNetEngine(_Net):
    """OSNet executor for up to `max_batch` crops per call: forward(x [n,256,128,8] fp16, n) -> [n, 512] f32
    L2-normalised embeddings (feature_extractor.py:62-74)."""

    def __init__(self, width, weights=None, input_hw=(256, 128), feature_dim=512, max_batch=256, use_tc=True,
                 use_graph=False, ops=None):
        """`ops` (+ `weights`): an op list in the vocabulary of models/osnet.py, e.g. from
        models.onnx_import.import_reid_onnx; default: the published OSNet of the given width."""
        super().__init__(use_tc, use_graph)
        # The ReID stack's stand-alone convs are thousands of tiles with 1-8 K slices each: the one-tile-per-CTA TMA
        # kernel pays its per-CTA set-up 20 times per SM there (OSNet x1.0 @200 crops: 2.89 ms with it, 2.66 ms with
        # the cp.async kernel, profiles/r02_summary.md); it is the batch-1 detector layers it was written for.
        self.use_tma = os.environ.get("FM_OSNET_TMA", "0") == "1"
        if ops is not None and weights is None:
            raise ValueError("a custom op list needs its weights")
        self.ops = list(ops) if ops is not None else osnet.build_osnet(width, feature_dim)
        self.weights = weights if weights is not None else osnet.synthetic_weights(self.ops)
        self.max_batch = max_batch
        self.feature_dim = feature_dim
        self.macs_per_crop = osnet.count_macs(self.ops, *input_hw)
        B, (H, W) = max_batch, input_hw
        dev = self.dev
        # FM_OSB_FUSED=0 falls back to one launch per layer (r01 path); default: fused stem (fm_osnet_stem), one
        # fm_osb_streams + fm_osb_merge pair per OSBlock (csrc/osnet_fused.cu, csrc/osnet_stem.cu)
        self.fuse_osb = os.environ.get("FM_OSB_FUSED", "1") != "0" and use_tc
        first, second = self.ops[0], self.ops[1]
        self.fuse_stem = (self.fuse_osb and os.environ.get("FM_OSB_STEM", "1") != "0" and (H, W) == (256, 128)
                          and first[0] == 'conv' and first[2:8] == (3, 64, 7, 2, 3, 'relu') and second[0] == 'maxpool3s2'
                          and second[1] == first[9])
        # network input: NHWC8 (layout 1 of fm_roi_resize_norm) or, for the fused stem, NHWC4 inside a zero border
        self.inp_layout = 2 if self.fuse_stem else 1
        if self.fuse_stem:
            self.inp = torch.zeros(B, H + 8, W + 8, 4, dtype=torch.float16, device=dev)
        else:
            self.inp = torch.zeros(B, H, W, IN_C_PAD, dtype=torch.float16, device=dev)
        self.out = torch.zeros(B, feature_dim, dtype=torch.float32, device=dev)
        self.n_dev = None
        # last use of every symbolic buffer -> simple size-keyed recycling
        last = {}
        for k, op in enumerate(self.ops):
            for name in self._reads(op):
                last[name] = k
        pool = {}
        live = {'input': (self.inp, IN_C_PAD, H, W)}
        params = {}

        self._bufs = []     # every activation buffer must outlive the recorded launches (raw pointers!)

        def alloc(numel, dtype=torch.float16):
            key = (numel, dtype)
            if pool.get(key):
                return pool[key].pop()
            t = torch.zeros(numel, dtype=dtype, device=dev)
            self._bufs.append(t)
            return t

        def release(name):
            t = live.pop(name, None)
            if t is not None and name not in ('input',):
                pool.setdefault((t[0].numel(), t[0].dtype), []).append(t[0])

        def dparam(name):
            if name not in params:
                params[name] = tuple(torch.as_tensor(a).to(dev) for a in self.weights[name])
            return params[name]

        self.pooled = torch.zeros(4 * B, 512, dtype=torch.float32, device=dev)
        self.gate_tmp = torch.zeros(4 * B, 512, dtype=torch.float32, device=dev)
        self._params = params
        fused_add = {}
        self.n_osb = 0
        skip_until = -1
        pooled_by_tail = {}
        for k, op in enumerate(self.ops):
            kind = op[0]
            if k <= skip_until:
                continue
            if k == 0 and self.fuse_stem:
                from .packing import pack_b_sw64
                w7, b7 = self.weights[op[1]]                        # [64][7][7][3]
                wk = np.zeros((64, 7, 8, 4), np.float32)
                wk[:, :, 1:8, :3] = w7
                img = torch.as_tensor(pack_b_sw64(wk.reshape(64, 224))).to(dev)
                b_d = torch.as_tensor(np.ascontiguousarray(b7, np.float32)).to(dev)
                y = alloc(B * 64 * 32 * 64)
                self._keep += [img, b_d]
                self._add('fm_osnet_stem', ptr(self.inp), B, ptr(img), ptr(b_d), ptr(y))
                self.n_tc += 1
                self.layer_bytes += 2 * (B * (H + 8) * (W + 8) * 4 + B * 64 * 32 * 64)
                live[self.ops[1][2]] = (y, 64, 64, 32)
                skip_until = 1
                continue
            if kind == 'conv' and self.fuse_osb and op[4] == 1 and op[7] == 'relu':     # OSBlock candidate (structural)
                blk = self._match_osblock(k)
                x, xc, h, w = live[op[8]]
                if blk is not None and xc == op[2] and xc % 64 == 0 and \
                        self._lib.fm_osb_streams_strips(h, w, op[3]) > 0:
                    tails_names, gate_k = blk
                    mid = op[3]
                    strips = self._lib.fm_osb_streams_strips(h, w, mid)
                    from .packing import pack_b_sw128
                    w1, b1 = self.weights[op[1]]
                    w1_img = torch.as_tensor(pack_b_sw128(w1.reshape(mid, xc))).to(dev)
                    b1_d = torch.as_tensor(np.ascontiguousarray(b1, np.float32)).to(dev)
                    pw_imgs, dw_blobs = [], []
                    for i in range(10):
                        pw_op, dw_op = self.ops[k + 1 + 2 * i], self.ops[k + 2 + 2 * i]
                        wp, bp = self.weights[pw_op[1]]
                        wd, bd = self.weights[dw_op[1]]
                        pw_imgs.append(pack_b_sw128(wp.reshape(mid, mid)))
                        dw_blobs.append(np.concatenate([np.ascontiguousarray(wd, np.float32).astype(np.float16)
                                                        .reshape(-1).view(np.uint8),
                                                        np.ascontiguousarray(bp, np.float32).view(np.uint8),
                                                        np.ascontiguousarray(bd, np.float32).view(np.uint8)]))
                    pw_d = torch.as_tensor(np.concatenate(pw_imgs)).to(dev)
                    dw_d = torch.as_tensor(np.concatenate(dw_blobs)).to(dev)
                    tails = [alloc(B * h * w * mid) for _ in range(4)]
                    gap = torch.zeros(B * strips * 4 * mid, dtype=torch.float32, device=dev)
                    d = _lib.FmOsbStreams()
                    d.x, d.n, d.h, d.w, d.cin, d.mid = x.data_ptr(), B, h, w, xc, mid
                    d.w1, d.b1, d.pw, d.dw = w1_img.data_ptr(), b1_d.data_ptr(), pw_d.data_ptr(), dw_d.data_ptr()
                    for i in range(4):
                        d.tails[i] = tails[i].data_ptr()
                    d.gap_part = gap.data_ptr()
                    self._keep += [d, w1_img, b1_d, pw_d, dw_d, gap]
                    self._add('fm_osb_streams', C.byref(d))
                    self.n_tc += 1
                    self.n_osb += 1
                    self.layer_bytes += 2 * (B * h * w * (xc + 4 * mid))
                    for i, tn in enumerate(tails_names):
                        live[tn] = (tails[i], mid, h, w)
                    pooled_by_tail[tails_names[0]] = (gap, strips)
                    # the block input may die here (no identity / downsample use): same bookkeeping as below
                    for name in self._reads(op):
                        if last.get(name) == k:
                            release(name)
                    skip_until = gate_k - 1
                    continue
            if kind == 'conv':
                _, name, cin, cout, ks, stride, pad, act, src, dst = op
                x, xc, h, w = live[src]
                ho, wo = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
                wt, bs = self.weights[name]
                if xc != cin:  # stem: physical 4 channels
                    wt = np.concatenate([wt, np.zeros(wt.shape[:3] + (xc - cin,), np.float32)], -1)
                wd = torch.as_tensor(np.ascontiguousarray(wt)).to(dev).half().contiguous()
                bd = torch.as_tensor(bs).to(dev).float().contiguous()
                params[name] = (wd, bd)
                y = alloc(B * ho * wo * cout)
                nxt = self.ops[k + 1] if k + 1 < len(self.ops) else None
                if nxt is not None and nxt[0] == 'add_relu' and nxt[1] == dst and act == 'linear' and nxt[2] in live:
                    # relu(conv3(x) + identity): residual + activation in the conv epilogue, no extra pass
                    d = _conv_desc(B, h, w, xc, xc, 0, ho, wo, cout, cout, 0, ks, stride, pad,
                                   _ACT['relu'] | ACT_AFTER_RESIDUAL)
                    d.res_stride, d.res_offset = cout, 0
                    self._conv(d, x, wd, bd, y, residual=live[nxt[2]][0])
                    fused_add[k + 1] = nxt[2]
                    new = (nxt[3], (y, cout, ho, wo))
                else:
                    d = _conv_desc(B, h, w, xc, xc, 0, ho, wo, cout, cout, 0, ks, stride, pad, _ACT[act])
                    self._conv(d, x, wd, bd, y)
                    new = (dst, (y, cout, ho, wo))
            elif kind == 'dw':
                _, name, c, act, src, dst = op
                x, xc, h, w = live[src]
                wd, bd = dparam(name)
                wd = wd.half().contiguous()
                params[name] = (wd, bd)
                y = alloc(B * h * w * c)
                self._add('fm_dwconv3', ptr(x), ptr(wd), ptr(bd), ptr(y), B, h, w, c, _ACT[act])
                self.layer_bytes += 2 * (2 * B * h * w * c + 9 * c)
                new = (dst, (y, c, h, w))
            elif kind == 'maxpool3s2':
                x, xc, h, w = live[op[1]]
                ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
                y = alloc(B * ho * wo * xc)
                self._add('fm_maxpool_pad', ptr(x), ptr(y), B, h, w, xc, 3, 2, 1)
                new = (op[2], (y, xc, ho, wo))
            elif kind == 'avgpool2':
                x, xc, h, w = live[op[1]]
                y = alloc(B * (h // 2) * (w // 2) * xc)
                self._add('fm_avgpool2', ptr(x), ptr(y), B, h, w, xc)
                new = (op[2], (y, xc, h // 2, w // 2))
            elif kind == 'gate':
                _, name, c, src, acc, accumulate = op
                x, xc, h, w = live[src]
                w1, b1, w2, b2 = dparam(name)
                if acc not in live:
                    live[acc] = (alloc(B * h * w * c), c, h, w)
                a = live[acc][0]
                self._add('fm_channel_gate', ptr(x), ptr(self.pooled), ptr(self.gate_tmp), ptr(w1), ptr(b1), ptr(w2),
                          ptr(b2), ptr(a), B, h * w, c, w1.shape[0], accumulate)
                new = None
            elif kind == 'gate4' and op[3][0] in pooled_by_tail and self._match_merge(k) is not None and \
                    self._lib.fm_osb_merge_ncta(op[2], self._match_merge(k)[1][3]) > 0:
                # gate + conv3 (+ downsample) + residual + ReLU in one launch (kernel G)
                _, name, c, srcs, acc = op
                ds_op, c3_op, add_op, last_k = self._match_merge(k)
                from .packing import pack_b_sw128
                xs = [live[s_][0] for s_ in srcs]
                _, _, h, w = live[srcs[0]]
                gap, strips = pooled_by_tail[srcs[0]]
                gw1, gb1, gw2, gb2 = dparam(name)
                cout = c3_op[3]
                ncta = self._lib.fm_osb_merge_ncta(c, cout)
                w3, b3 = self.weights[c3_op[1]]
                wcat = w3.reshape(cout, c)
                bias = np.asarray(b3, np.float32).copy()
                ident_name = add_op[2] if add_op[1] == c3_op[9] else add_op[1]
                d = _lib.FmOsbMerge()
                if ds_op is not None:
                    wdn, bdn = self.weights[ds_op[1]]
                    cin = ds_op[2]
                    wcat = np.concatenate([wdn.reshape(cout, cin), wcat], 1)
                    bias += np.asarray(bdn, np.float32)
                    xin = live[ds_op[8]]
                    d.x, d.res, d.cin = xin[0].data_ptr(), None, cin
                else:
                    d.x, d.res, d.cin = None, live[ident_name][0].data_ptr(), cout
                img = np.concatenate([pack_b_sw128(wcat[r:r + ncta]) for r in range(0, cout, ncta)])
                img_d = torch.as_tensor(img).to(dev)
                bias_d = torch.as_tensor(bias).to(dev)
                y = alloc(B * h * w * cout)
                d.n, d.hw, d.cout, d.mid, d.cr, d.strips = B, h * w, cout, c, gw1.shape[0], strips
                for i in range(4):
                    d.tails[i] = xs[i].data_ptr()
                d.gap_part = gap.data_ptr()
                d.gw1, d.gb1, d.gw2, d.gb2 = (t_.data_ptr() for t_ in (gw1, gb1, gw2, gb2))
                d.wimg, d.bias, d.out = img_d.data_ptr(), bias_d.data_ptr(), y.data_ptr()
                d.gate_scratch = self.gate_tmp.data_ptr()
                self._keep += [d, img_d, bias_d]
                self._add('fm_osb_merge', C.byref(d))
                self.n_tc += 1
                self.n_merge = getattr(self, 'n_merge', 0) + 1
                self.layer_bytes += 2 * (B * h * w * (4 * c + 2 * cout))
                # bookkeeping of the skipped ops' reads
                for kk in range(k, last_k + 1):
                    for nm in self._reads(self.ops[kk]):
                        if last.get(nm) == kk:
                            release(nm)
                if add_op[3] in live:
                    release(add_op[3])
                live[add_op[3]] = (y, cout, h, w)
                skip_until = last_k
                continue
            elif kind == 'gate4':
                _, name, c, srcs, acc = op
                xs = [live[s_][0] for s_ in srcs]
                _, xc, h, w = live[srcs[0]]
                w1, b1, w2, b2 = dparam(name)
                a = alloc(B * h * w * c)
                if srcs[0] in pooled_by_tail:      # channel sums already produced by fm_osb_streams
                    gap, strips = pooled_by_tail[srcs[0]]
                    self._add('fm_channel_gate4_pooled', ptr(xs[0]), ptr(xs[1]), ptr(xs[2]), ptr(xs[3]), ptr(gap),
                              strips, ptr(self.gate_tmp), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(a), B, h * w, c,
                              w1.shape[0])
                else:
                    self._add('fm_channel_gate4', ptr(xs[0]), ptr(xs[1]), ptr(xs[2]), ptr(xs[3]), ptr(self.pooled),
                              ptr(self.gate_tmp), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(a), B, h * w, c,
                              w1.shape[0])
                new = (acc, (a, c, h, w))
            elif kind == 'add_relu':
                if k in fused_add:        # folded into the preceding conv's epilogue
                    continue
                a, ac, h, w = live[op[1]]
                b = live[op[2]][0]
                y = alloc(B * h * w * ac)
                self._add('fm_add_act', ptr(a), ptr(b), ptr(y), B * h * w * ac, _ACT['relu'])
                new = (op[3], (y, ac, h, w))
            elif kind == 'gap':
                x, xc, h, w = live[op[1]]
                y = alloc(B * xc, torch.float32)
                self._add('fm_global_avgpool', ptr(x), ptr(y), B, h * w, xc)
                new = (op[2], (y, xc, 1, 1))
            elif kind == 'fc':
                _, name, cin, cout, src, dst = op
                x = live[src][0]
                wd, bd = dparam(name)
                self._add('fm_fc_norm', ptr(x), ptr(wd), ptr(bd), ptr(self.out), B, cin, cout, 1, 1)
                new = None
            else:
                raise NotImplementedError(kind)
            for name in self._reads(op):
                if last.get(name) == k and not (kind == 'gate' and name == op[4]):
                    release(name)
            if new is not None:
                if new[0] in live:
                    release(new[0])
                live[new[0]] = new[1]

    def _match_osblock(self, k):
        """ops[k] = '<blk>.conv1'; returns (tail buffer names of the four streams, index of the gate4 op) when the next
        twenty ops are the Lite-3x3 chains a.0, b.0, b.1, c.0 .. d.3 feeding that gate, else None."""
        ops = self.ops
        if k + 21 >= len(ops):
            return None
        x1 = ops[k][9]
        mid = ops[k][3]
        tails = []
        i = k + 1
        for s_ in range(4):
            prev = x1
            for j in range(s_ + 1):
                pw, dw = ops[i], ops[i + 1]
                if pw[0] != 'conv' or dw[0] != 'dw' or pw[2] != mid or pw[3] != mid or pw[4] != 1 or pw[7] != 'linear' \
                        or pw[8] != prev or dw[4] != pw[9] or dw[2] != mid or dw[3] != 'relu':
                    return None
                prev = dw[5]
                i += 2
            tails.append(prev)
        g = ops[i]
        if g[0] != 'gate4' or tuple(g[3]) != tuple(tails) or ops[k][7] != 'relu':
            return None
        return tails, i

    def _match_merge(self, k):
        """ops[k] = gate4 of an OSBlock; returns (downsample conv or None, conv3, add_relu, index of add_relu) when
        the ops that follow are [downsample 1x1] conv3 1x1 (linear, reads the gate output) and add_relu."""
        ops = self.ops
        acc = ops[k][4]
        i = k + 1
        ds = None
        if i < len(ops) and ops[i][0] == 'conv' and ops[i][8] != acc:      # a conv beside conv3: the downsample branch
            ds = ops[i]
            i += 1
        if i + 1 >= len(ops):
            return None
        c3, add = ops[i], ops[i + 1]
        if c3[0] != 'conv' or c3[4] != 1 or c3[7] != 'linear' or c3[8] != acc or add[0] != 'add_relu':
            return None
        if c3[9] not in (add[1], add[2]):
            return None
        other = add[2] if add[1] == c3[9] else add[1]
        if ds is not None and (ds[4] != 1 or ds[7] != 'linear' or ds[9] != other or ds[2] % 64):
            return None
        return ds, c3, add, i + 1

    @staticmethod
    def _reads(op):
        kind = op[0]
        if kind == 'conv':
            return [op[8]]
        if kind == 'dw':
            return [op[4]]
        if kind in ('maxpool3s2', 'avgpool2', 'gap'):
            return [op[1]]
        if kind == 'gate':
            return [op[3], op[4]]
        if kind == 'gate4':
            return list(op[3])
        if kind == 'add_relu':
            return [op[1], op[2]]
        if kind == 'fc':
            return [op[4]]
        return []

    def load_nhwc8(self, x):
        """Copies crops given as [n][256][128][>= 3] fp16 (RGB first) into the engine's input buffer, whatever its
        layout (tests and tools; the product path writes the buffer with fm_roi_resize_norm)."""
        n = x.shape[0]
        if self.inp_layout == 2:
            self.inp[:n, 4:-4, 4:-4, :3].copy_(x[..., :3])
        else:
            self.inp[:n].copy_(x)

    def forward(self, n=None):
        """Runs the recorded network on self.inp (all max_batch rows; rows >= n are don't-care)."""
        self.replay()
        return self.out if n is None else self.out[:n]


def build_reid_engine(model, max_batch=256, use_tc=True, use_graph=True):
    """Engine for a `models.ReID` descriptor.  `MODEL_PATH` (an ONNX file, the reference's reid.py:20-23) is lowered
    by models.onnx_import; without it the descriptor's `ARCH` selects the built-in OSNet with synthetic weights."""
    if getattr(model, 'MODEL_PATH', None) is not None:
        from .models.onnx_import import import_reid_onnx
        ops, weights, in_shape, dim = import_reid_onnx(str(model.MODEL_PATH))
        if tuple(in_shape) != tuple(model.INPUT_SHAPE):
            raise ValueError(f"{model.__name__}: ONNX input {in_shape} != INPUT_SHAPE {model.INPUT_SHAPE}")   # reid.py:66
        if dim != model.OUTPUT_LAYOUT:
            raise ValueError(f"{model.__name__}: ONNX feature dim {dim} != OUTPUT_LAYOUT {model.OUTPUT_LAYOUT}")
        return OSNetEngine(None, weights=weights, input_hw=in_shape[1:], feature_dim=dim, max_batch=max_batch,
                           use_tc=use_tc, use_graph=use_graph, ops=ops)
    arch, width = model.ARCH
    assert arch == 'osnet'
    return OSNetEngine(width, input_hw=model.INPUT_SHAPE[1:], feature_dim=model.OUTPUT_LAYOUT, max_batch=max_batch,
                       use_tc=use_tc, use_graph=use_graph)


# ------------------------------------------------------------------------------------------------ stage profiling
class _Profiler:
    """CUDA-event timers around the conv stacks (recorded on the stream each stack is launched on)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.records = []

    def add(self, kind, e0, e1, flops, path, nbytes=0.0):
        self.records.append((kind, e0, e1, flops, path, nbytes))

    def summary(self):
        torch.cuda.synchronize()
        out = {"yolo_ms": 0.0, "osnet_ms": 0.0, "yolo_flops": 0.0, "osnet_flops": 0.0, "yolo_bytes": 0.0,
               "osnet_bytes": 0.0, "yolo_calls": 0, "osnet_calls": 0, "detector_frames": 0}
        path = set()
        for kind, e0, e1, flops, p, nbytes in self.records:
            out[kind + "_ms"] += e0.elapsed_time(e1)
            out[kind + "_flops"] += flops
            out[kind + "_bytes"] += nbytes
            out[kind + "_calls"] += 1
            out["detector_frames"] += kind == "yolo"
            path.add(p)
        out["conv_path"] = "+".join(sorted(path)) if path else None
        return out


_PROF = None


def enable_profiling():
    global _PROF
    _PROF = _Profiler()
    return _PROF


def _profiled(kind):
    def deco(fn):
        def wrapper(self, *a, **k):
            if _PROF is None:
                return fn(self, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(self, *a, **k)
            e1.record()
            if kind == "yolo":
                flops = float(self.flops)
            else:
                n = a[0] if a and a[0] is not None else self.max_batch
                flops = 2.0 * self.macs_per_crop * n
            path = "tcgen05" if self.n_tc >= self.n_simt else "simt"
            _PROF.add(kind, e0, e1, flops, f"{kind}:{path}({self.n_tc}tc/{self.n_simt}simt)", float(self.layer_bytes))
            return r
        return wrapper
    return deco


YoloEngine.forward = _profiled("yolo")(YoloEngine.forward)
OSNetEngine.forward = _profiled("osnet")(OSNetEngine.forward)
