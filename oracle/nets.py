"""TEST INFRASTRUCTURE ONLY.  fp32 PyTorch-CPU executors of the same network descriptions the GPU engine runs
(fastmot_b200/models/darknet.py layer lists, fastmot_b200/models/osnet.py op lists) with the same weights.

The reference's networks are TensorRT engines built from downloaded ONNX files (fastmot/models/yolo.py:106-151,
reid.py:48-92) and cannot run here (no TensorRT, no weights, no network): these executors are the floating-point
oracle for the conv stacks — "parity unpinned" with respect to the reference itself, pinned only to the layer
semantics of scripts/yolo2onnx.py:558-870.  Tolerance for the fp16 tensor-core path: ~2e-2 relative.
"""
import numpy as np
import torch
import torch.nn.functional as F

from fastmot_b200.models import darknet


def _act(x, name):
    if name == 'leaky':
        return F.leaky_relu(x, 0.1)
    if name == 'relu':
        return F.relu(x)
    if name == 'mish':
        return x * torch.tanh(F.softplus(x))
    if name == 'swish':
        return x * torch.sigmoid(x)
    if name == 'logistic':
        return torch.sigmoid(x)
    return x


def _same_upper_pool(x, k, s):
    h, w = x.shape[-2:]
    ho, wo = -(-h // s), -(-w // s)
    ph, pw = max((ho - 1) * s + k - h, 0), max((wo - 1) * s + k - w, 0)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float('-inf'))
    return F.max_pool2d(x, k, s)


def run_darknet(layers, weights, x, quantize=None):
    """x: (1,3,H,W) float32 in [0,1].  Returns the list of raw head tensors [(5+C)*A, H, W] float32.
    quantize: optional callable applied to every layer output (e.g. fp16 round trip) to model storage precision."""
    q = quantize or (lambda t: t)
    res, _ = darknet.infer_shapes(layers, x.shape[1], x.shape[2], x.shape[3])
    outs, heads = [], []
    cur = x
    for i, l in enumerate(res):
        t = l['type']
        if t == 'convolutional':
            w, b = weights[i]
            wt = q(torch.as_tensor(w).permute(0, 3, 1, 2).contiguous())
            k = l['size']
            cur = F.conv2d(cur, wt, torch.as_tensor(b), stride=l.get('stride', 1), padding=k // 2 if l.get('pad', 0) else 0)
            cur = q(_act(cur, l.get('activation', 'linear')))
        elif t == 'maxpool':
            cur = _same_upper_pool(cur, l['size'], l['stride'])
        elif t == 'upsample':
            cur = F.interpolate(cur, scale_factor=l['stride'], mode='nearest')
        elif t == 'shortcut':
            cur = q(_act(cur + outs[l['from_abs']], l.get('activation', 'linear')))
        elif t == 'route':
            g, gid = l.get('groups', 1), l.get('group_id', 0)
            parts = []
            for s in l['layers_abs']:
                o = outs[s]
                c = o.shape[1] // g
                parts.append(o[:, gid * c:(gid + 1) * c])
            cur = torch.cat(parts, 1) if len(parts) > 1 else parts[0]
        elif t == 'yolo':
            heads.append(cur[0].clone())
        outs.append(cur)
    return heads


def run_osnet(ops, weights, x, quantize=None):
    """x: (N,3,256,128) float32 normalised crops -> (N, 512) L2-normalised float32 embeddings."""
    q = quantize or (lambda t: t)
    bufs = {'input': x}
    for op in ops:
        kind = op[0]
        if kind == 'conv':
            _, name, cin, cout, ks, stride, pad, act, src, dst = op
            w, b = weights[name]
            wt = q(torch.as_tensor(w).permute(0, 3, 1, 2).contiguous())
            bufs[dst] = q(_act(F.conv2d(bufs[src], wt, torch.as_tensor(b), stride=stride, padding=pad), act))
        elif kind == 'dw':
            _, name, c, act, src, dst = op
            w, b = weights[name]
            wt = q(torch.as_tensor(w).reshape(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous())
            bufs[dst] = q(_act(F.conv2d(bufs[src], wt, torch.as_tensor(b), padding=1, groups=c), act))
        elif kind == 'maxpool3s2':
            bufs[op[2]] = F.max_pool2d(bufs[op[1]], 3, 2, 1)
        elif kind == 'avgpool2':
            bufs[op[2]] = q(F.avg_pool2d(bufs[op[1]], 2))
        elif kind == 'gate':
            _, name, c, src, acc, accumulate = op
            w1, b1, w2, b2 = (torch.as_tensor(a) for a in weights[name])
            xx = bufs[src]
            p = xx.mean((2, 3))
            g = torch.sigmoid(F.relu(p @ w1.T + b1) @ w2.T + b2)
            y = xx * g[:, :, None, None]
            bufs[acc] = q(y + bufs[acc]) if accumulate else q(y)
        elif kind == 'gate4':
            _, name, c, srcs, acc = op
            w1, b1, w2, b2 = (torch.as_tensor(a) for a in weights[name])
            tot = 0
            for s_ in srcs:
                xx = bufs[s_]
                g = torch.sigmoid(F.relu(xx.mean((2, 3)) @ w1.T + b1) @ w2.T + b2)
                tot = tot + xx * g[:, :, None, None]
            bufs[acc] = q(tot)
        elif kind == 'add_relu':
            bufs[op[3]] = q(F.relu(bufs[op[1]] + bufs[op[2]]))
        elif kind == 'gap':
            bufs[op[2]] = bufs[op[1]].mean((2, 3))
        elif kind == 'fc':
            _, name, cin, cout, src, dst = op
            w, b = (torch.as_tensor(a) for a in weights[name])
            v = F.relu(bufs[src] @ w.T + b)
            bufs[dst] = v / v.norm(dim=1, keepdim=True)
    return bufs[ops[-1][5]]     # the 'fc' op's destination ('feat' for build_osnet; any name for imported graphs)


def fp16_roundtrip(t):
    return t.half().float()
