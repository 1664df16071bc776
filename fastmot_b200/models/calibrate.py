"""Data-dependent rescaling of SYNTHETIC weights (init time only, not on the hot path).

Trained YOLO / OSNet weights carry batch-norm statistics that keep every activation O(1); plain He-normal weights
do not — through 100+ layers of residual adds and concats the activations leave the fp16 range.  `calibrate_*`
does what BN does at initialisation: it pushes one random input through the network in fp32 (PyTorch CPU, at a
small resolution) and rescales each conv's output channels to zero mean / unit variance, folding the result into
weight + bias.  Real Darknet weights (darknet.load_weights) never go through this.
"""
import numpy as np
import torch
import torch.nn.functional as F


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


def _normalise(y, w, b, rng_bias):
    """y: conv output (N,C,H,W) with weights w [C][kh][kw][cin]; returns rescaled (y, w, b)."""
    m = y.mean((0, 2, 3))
    s = y.std((0, 2, 3)).clamp_min(1e-6)
    scale = (1.0 / s).numpy()
    w = w * scale[:, None, None, None]
    b = (b - m.numpy()) * scale + rng_bias
    y = (y - m[None, :, None, None]) / s[None, :, None, None] + torch.as_tensor(rng_bias)[None, :, None, None]
    return y, w.astype(np.float32), b.astype(np.float32)


def calibrate_darknet(res_layers, weights, in_c, size=128, seed=77):
    """res_layers: darknet.infer_shapes(...)[0]; weights: {i: (w, b)} modified in place and returned."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    x = torch.rand(2, in_c, size, size, generator=g)
    outs = []
    cur = x
    with torch.no_grad():
        for i, l in enumerate(res_layers):
            t = l['type']
            if t == 'convolutional':
                w, b = weights[i]
                k = l['size']
                y = F.conv2d(cur, torch.as_tensor(w).permute(0, 3, 1, 2).contiguous(), None, stride=l.get('stride', 1),
                             padding=k // 2 if l.get('pad', 0) else 0)
                is_head = i + 1 < len(res_layers) and res_layers[i + 1]['type'] == 'yolo'
                y, w, b2 = _normalise(y, w, np.zeros_like(b), rng.normal(0, 0.1, len(b)).astype(np.float32))
                if is_head:
                    b2 = b2 + b        # keep the detection-prior biases chosen by the caller
                    y = y + torch.as_tensor(b)[None, :, None, None]
                weights[i] = (w, b2)
                cur = _act(y, l.get('activation', 'linear'))
            elif t == 'maxpool':
                k, s = l['size'], l['stride']
                h, wd = cur.shape[-2:]
                ho, wo = -(-h // s), -(-wd // s)
                ph, pw = max((ho - 1) * s + k - h, 0), max((wo - 1) * s + k - wd, 0)
                cur = F.max_pool2d(F.pad(cur, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float('-inf')), k, s)
            elif t == 'upsample':
                cur = F.interpolate(cur, scale_factor=l['stride'], mode='nearest')
            elif t == 'shortcut':
                cur = _act(cur + outs[l['from_abs']], l.get('activation', 'linear'))
            elif t == 'route':
                gq, gid = l.get('groups', 1), l.get('group_id', 0)
                parts = []
                for s in l['layers_abs']:
                    o = outs[s]
                    c = o.shape[1] // gq
                    parts.append(o[:, gid * c:(gid + 1) * c])
                cur = torch.cat(parts, 1) if len(parts) > 1 else parts[0]
            outs.append(cur)
    return weights


def calibrate_osnet(ops, weights, hw=(256, 128), seed=78):
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    bufs = {'input': torch.randn(2, 3, hw[0], hw[1], generator=g)}
    with torch.no_grad():
        for op in ops:
            kind = op[0]
            if kind == 'conv':
                _, name, cin, cout, ks, stride, pad, act, src, dst = op
                w, b = weights[name]
                y = F.conv2d(bufs[src], torch.as_tensor(w).permute(0, 3, 1, 2).contiguous(), None, stride=stride,
                             padding=pad)
                y, w, b = _normalise(y, w, np.zeros_like(b), rng.normal(0, 0.1, len(b)).astype(np.float32))
                weights[name] = (w, b)
                bufs[dst] = _act(y, act)
            elif kind == 'dw':
                _, name, c, act, src, dst = op
                w, b = weights[name]
                wt = torch.as_tensor(w).reshape(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous()
                y = F.conv2d(bufs[src], wt, None, padding=1, groups=c)
                m, s = y.mean((0, 2, 3)), y.std((0, 2, 3)).clamp_min(1e-6)
                nb = rng.normal(0, 0.1, c).astype(np.float32)
                weights[name] = ((w / s.numpy()[None, :]).astype(np.float32),
                                 ((-m / s).numpy() + nb).astype(np.float32))
                bufs[dst] = _act((y - m[None, :, None, None]) / s[None, :, None, None] +
                                 torch.as_tensor(nb)[None, :, None, None], act)
            elif kind == 'maxpool3s2':
                bufs[op[2]] = F.max_pool2d(bufs[op[1]], 3, 2, 1)
            elif kind == 'avgpool2':
                bufs[op[2]] = F.avg_pool2d(bufs[op[1]], 2)
            elif kind == 'gate':
                _, name, c, src, acc, accumulate = op
                w1, b1, w2, b2 = (torch.as_tensor(a) for a in weights[name])
                xx = bufs[src]
                gt = torch.sigmoid(F.relu(xx.mean((2, 3)) @ w1.T + b1) @ w2.T + b2)
                y = xx * gt[:, :, None, None]
                bufs[acc] = y + bufs[acc] if accumulate else y
            elif kind == 'gate4':
                _, name, c, srcs, acc = op
                w1, b1, w2, b2 = (torch.as_tensor(a) for a in weights[name])
                tot = 0
                for s_ in srcs:
                    xx = bufs[s_]
                    gt = torch.sigmoid(F.relu(xx.mean((2, 3)) @ w1.T + b1) @ w2.T + b2)
                    tot = tot + xx * gt[:, :, None, None]
                bufs[acc] = tot
            elif kind == 'add_relu':
                bufs[op[3]] = F.relu(bufs[op[1]] + bufs[op[2]])
            elif kind == 'gap':
                bufs[op[2]] = bufs[op[1]].mean((2, 3))
            elif kind == 'fc':
                pass
    return weights
