"""ONNX -> ReID engine importer, and the matching exporter.

The reference hands `ReID.MODEL_PATH` (an ONNX file, `fastmot/models/reid.py:20-23,55-63`) to TensorRT's ONNX
parser; here `import_reid_onnx` lowers the same kind of file to the op list `fastmot_b200.engine.OSNetEngine`
executes (vocabulary: `fastmot_b200/models/osnet.py`), so a custom ReID descriptor (`class MyNet(ReID): MODEL_PATH =
...`) runs on the CUDA path without TensorRT.

Supported graph (what `torch.onnx.export` emits for torchreid's OSNet family and for plain conv / residual ReID
backbones in eval mode):
  Conv (dense, or depthwise 3x3 s1 p1) [+ BatchNormalization] [+ Relu]        -> 'conv' / 'dw' (BN folded)
  MaxPool 3x3 s2 p1, AveragePool 2x2 s2, GlobalAveragePool                     -> 'maxpool3s2', 'avgpool2', 'gap'
  GlobalAveragePool -> Conv1x1 -> Relu -> Conv1x1 -> Sigmoid -> Mul(x, .)      -> channel gate; a sum (Add tree) of
      gated tensors lowers to 'gate4' (four streams sharing one gate: the OSBlock) or a chain of 'gate' ops
  Add + Relu                                                                   -> 'add_relu'
  Flatten / Reshape / Squeeze of the pooled vector, Identity, Dropout,
  all-zero Pad (opset-9 AveragePool), Shape / Gather / Unsqueeze / Concat / Constant feeding that Reshape -> aliases / ignored
  Gemm | MatMul + Add  [+ BatchNormalization] + Relu                           -> 'fc' (the engine L2-normalises the
      rows afterwards, feature_extractor.py:62-74 `_normalize`)
Anything else raises `UnsupportedOnnx` naming the node — never a silent skip.
"""
import numpy as np

from . import onnx_io
from .onnx_io import Node, Graph, ValueInfo


class UnsupportedOnnx(ValueError):
    pass


def _fold_bn(w, b, bn, axis0=True):
    """Folds BatchNormalization(scale, bias, mean, var, eps) into the producer's per-output-channel weight / bias."""
    scale, beta, mean, var, eps = bn
    k = (scale / np.sqrt(var.astype(np.float64) + eps)).astype(np.float32)
    w = w * k.reshape((-1,) + (1,) * (w.ndim - 1))
    b = (b - mean) * k + beta
    return w.astype(np.float32), b.astype(np.float32)


class _Lowering:
    def __init__(self, g):
        self.g = g
        self.init = g.initializers
        self.consumers = {}
        for n in g.nodes:
            for t in n.inputs:
                self.consumers.setdefault(t, []).append(n)
        for vi in g.outputs:
            self.consumers.setdefault(vi.name, []).append(None)      # graph outputs count as a use
        self.done = set()           # ids of nodes already absorbed by a pattern
        self.alias = {}             # tensor -> tensor (Flatten / Identity / input)
        self.gated = {}             # tensor -> (src tensor, gate weight key)
        self.gsum = {}              # tensor -> list of gated tensors (Add tree of gated streams)
        self.channels = {}          # activation tensor -> channel count
        self.ops, self.weights = [], {}
        self.used_names = set()
        if len(g.inputs) != 1:
            raise UnsupportedOnnx(f"expected one graph input, found {[v.name for v in g.inputs]}")
        vi = g.inputs[0]
        if vi.name != 'input':
            self.alias[vi.name] = 'input'
        shp = vi.shape
        if len(shp) != 4 or not all(isinstance(d, int) for d in shp[1:]):
            raise UnsupportedOnnx(f"input '{vi.name}' must be (N, C, H, W) with static C, H, W; got {shp}")
        self.input_shape = tuple(int(d) for d in shp[1:])
        self.channels['input'] = self.input_shape[0]

    # -------------------------------------------------------------------------------------------- helpers
    def _sole_consumer(self, tensor, op_type):
        c = self.consumers.get(tensor, [])
        if len(c) == 1 and c[0] is not None and c[0].op_type == op_type and id(c[0]) not in self.done:
            return c[0]
        return None

    def _w(self, name, node):
        if name not in self.init:
            raise UnsupportedOnnx(f"{node.op_type} '{node.name}': operand '{name}' is not a constant initialiser")
        return np.asarray(self.init[name])

    def _name(self, node, fallback):
        base = node.name or fallback
        name, k = base, 1
        while name in self.used_names:
            k += 1
            name = f"{base}#{k}"
        self.used_names.add(name)
        return name

    def _bn_params(self, bn):
        scale, beta, mean, var = (self._w(t, bn).astype(np.float32) for t in bn.inputs[1:5])
        return scale, beta, mean, var, float(bn.attrs.get('epsilon', 1e-5))

    def _src(self, tensor, node):
        """Resolves an activation operand to an engine buffer name, materialising a pending gate sum first."""
        while tensor in self.alias:
            tensor = self.alias[tensor]
        if tensor in self.gated:
            self.gsum[tensor] = [tensor]
        if tensor in self.gsum:
            self._emit_gates(tensor)
            return tensor
        if tensor not in self.channels:
            raise UnsupportedOnnx(f"{node.op_type} '{node.name}' reads '{tensor}', which no supported node produced")
        return tensor

    def _emit_gates(self, out):
        parts = self.gsum.pop(out)
        srcs = [self.gated[p][0] for p in parts]
        keys = [self.gated[p][1] for p in parts]
        for p in parts:
            self.gated.pop(p, None)
        c = self.channels[srcs[0]]
        if len(parts) == 4 and len(set(keys)) == 1:
            self.ops.append(('gate4', keys[0], c, tuple(srcs), out))
        else:
            for i, (s, k) in enumerate(zip(srcs, keys)):
                self.ops.append(('gate', k, c, s, out, i > 0))
        self.channels[out] = c

    # -------------------------------------------------------------------------------------------- node handlers
    def _conv(self, n):
        w = self._w(n.inputs[1], n).astype(np.float32)
        cout, cin_g, kh, kw = w.shape
        b = self._w(n.inputs[2], n).astype(np.float32) if len(n.inputs) > 2 and n.inputs[2] else np.zeros(cout, np.float32)
        group = int(n.attrs.get('group', 1))
        strides = list(n.attrs.get('strides', [1, 1]))
        pads = list(n.attrs.get('pads', [0, 0, 0, 0]))
        dil = list(n.attrs.get('dilations', [1, 1]))
        if n.attrs.get('auto_pad', 'NOTSET') not in ('NOTSET', b'NOTSET'):
            raise UnsupportedOnnx(f"Conv '{n.name}': auto_pad {n.attrs['auto_pad']} (export with explicit pads)")
        if kh != kw or strides[0] != strides[1] or len(set(pads)) != 1 or dil != [1, 1]:
            raise UnsupportedOnnx(f"Conv '{n.name}': kernel {kh}x{kw} strides {strides} pads {pads} dilations {dil}")
        src = self._src(n.inputs[0], n)
        cin = self.channels[src]
        out = n.outputs[0]
        bn = self._sole_consumer(out, 'BatchNormalization')
        if bn is not None:
            w, b = _fold_bn(w, b, self._bn_params(bn))
            self.done.add(id(bn))
            out = bn.outputs[0]
        act = 'linear'
        relu = self._sole_consumer(out, 'Relu')
        if relu is not None:
            act = 'relu'
            self.done.add(id(relu))
            out = relu.outputs[0]
        name = self._name(n, f"conv_{len(self.ops)}")
        if group == 1:
            if cin_g != cin:
                raise UnsupportedOnnx(f"Conv '{n.name}': weight expects {cin_g} input channels, tensor has {cin}")
            self.weights[name] = (np.ascontiguousarray(w.transpose(0, 2, 3, 1)), b)     # [out][kh][kw][in]
            self.ops.append(('conv', name, cin, cout, kh, strides[0], pads[0], act, src, out))
        elif group == cin == cout and cin_g == 1 and kh == 3 and strides[0] == 1 and pads[0] == 1:
            self.weights[name] = (np.ascontiguousarray(w.reshape(cout, 9).T), b)        # [tap][c]
            self.ops.append(('dw', name, cout, act, src, out))
        else:
            raise UnsupportedOnnx(f"Conv '{n.name}': group {group} with {cin}->{cout} channels, k {kh}, stride "
                                  f"{strides[0]} (only dense convs and depthwise 3x3 s1 p1 are supported)")
        self.channels[out] = cout

    def _pool(self, n):
        k = list(n.attrs.get('kernel_shape', []))
        s = list(n.attrs.get('strides', [1, 1]))
        p = list(n.attrs.get('pads', [0, 0, 0, 0]))
        src = self._src(n.inputs[0], n)
        out = n.outputs[0]
        if n.op_type == 'MaxPool' and k == [3, 3] and s == [2, 2] and p == [1, 1, 1, 1] \
                and not n.attrs.get('ceil_mode', 0):
            self.ops.append(('maxpool3s2', src, out))
        elif n.op_type == 'AveragePool' and k == [2, 2] and s == [2, 2] and p == [0, 0, 0, 0] \
                and not n.attrs.get('ceil_mode', 0):
            self.ops.append(('avgpool2', src, out))
        else:
            raise UnsupportedOnnx(f"{n.op_type} '{n.name}': kernel {k} strides {s} pads {p} "
                                  "(supported: MaxPool 3x3 s2 p1, AveragePool 2x2 s2)")
        self.channels[out] = self.channels[src]

    def _gap(self, n):
        """Either the squeeze of a channel gate (GAP -> fc1 -> Relu -> fc2 -> Sigmoid -> Mul) or the final pooling."""
        x = n.inputs[0]
        while x in self.alias:
            x = self.alias[x]
        fc1 = self._sole_consumer(n.outputs[0], 'Conv')
        if fc1 is not None:
            relu = self._sole_consumer(fc1.outputs[0], 'Relu')
            fc2 = self._sole_consumer(relu.outputs[0], 'Conv') if relu is not None else None
            sig = self._sole_consumer(fc2.outputs[0], 'Sigmoid') if fc2 is not None else None
            mul = self._sole_consumer(sig.outputs[0], 'Mul') if sig is not None else None
            if mul is None or x not in [self.alias.get(t, t) for t in mul.inputs]:
                raise UnsupportedOnnx(f"GlobalAveragePool '{n.name}' feeds a Conv that is not a channel gate "
                                      "(GAP -> Conv1x1 -> Relu -> Conv1x1 -> Sigmoid -> Mul with the pooled tensor)")
            src = self._src(x, n)
            c = self.channels[src]
            key = 'gate:' + fc1.inputs[1] + '|' + fc2.inputs[1]
            if key not in self.weights:
                w1 = self._w(fc1.inputs[1], fc1).astype(np.float32)
                w2 = self._w(fc2.inputs[1], fc2).astype(np.float32)
                if w1.shape[1:] != (c, 1, 1) or w2.shape != (c, w1.shape[0], 1, 1):
                    raise UnsupportedOnnx(f"channel gate at '{n.name}': fc shapes {w1.shape} / {w2.shape} for {c} channels")
                b1 = self._w(fc1.inputs[2], fc1).astype(np.float32) if len(fc1.inputs) > 2 else np.zeros(len(w1), np.float32)
                b2 = self._w(fc2.inputs[2], fc2).astype(np.float32) if len(fc2.inputs) > 2 else np.zeros(c, np.float32)
                self.weights[key] = (w1.reshape(len(w1), c).copy(), b1, w2.reshape(c, len(w1)).copy(), b2)
            for m in (fc1, relu, fc2, sig, mul):
                self.done.add(id(m))
            self.gated[mul.outputs[0]] = (src, key)
            return
        src = self._src(x, n)
        self.ops.append(('gap', src, n.outputs[0]))
        self.channels[n.outputs[0]] = self.channels[src]

    def _add(self, n):
        a, b = (self._resolve_alias(t) for t in n.inputs)
        parts = []
        for t in (a, b):
            if t in self.gsum:
                parts.append(self.gsum[t])
            elif t in self.gated:
                parts.append([t])
            else:
                parts = None
                break
        if parts is not None:
            for t in (a, b):
                self.gsum.pop(t, None)
            self.gsum[n.outputs[0]] = parts[0] + parts[1]
            return
        sa, sb = self._src(a, n), self._src(b, n)
        relu = self._sole_consumer(n.outputs[0], 'Relu')
        if relu is None:
            raise UnsupportedOnnx(f"Add '{n.name}' without a following Relu (residual blocks end in Add + Relu)")
        self.done.add(id(relu))
        if self.channels[sa] != self.channels[sb]:
            raise UnsupportedOnnx(f"Add '{n.name}': {self.channels[sa]} vs {self.channels[sb]} channels")
        self.ops.append(('add_relu', sa, sb, relu.outputs[0]))
        self.channels[relu.outputs[0]] = self.channels[sa]

    def _resolve_alias(self, t):
        while t in self.alias:
            t = self.alias[t]
        return t

    def _fc(self, n):
        src = self._src(n.inputs[0], n)
        w = self._w(n.inputs[1], n).astype(np.float32)
        if n.op_type == 'Gemm':
            if n.attrs.get('transA', 0) or float(n.attrs.get('alpha', 1.0)) != 1.0 or float(n.attrs.get('beta', 1.0)) != 1.0:
                raise UnsupportedOnnx(f"Gemm '{n.name}': transA / alpha / beta other than the Linear defaults")
            if not n.attrs.get('transB', 0):
                w = w.T
            b = self._w(n.inputs[2], n).astype(np.float32) if len(n.inputs) > 2 else np.zeros(len(w), np.float32)
            out = n.outputs[0]
        else:                                   # MatMul [+ Add bias]
            w = w.T
            out = n.outputs[0]
            add = self._sole_consumer(out, 'Add')
            b = np.zeros(len(w), np.float32)
            if add is not None:
                other = [t for t in add.inputs if t != out]
                if len(other) == 1 and other[0] in self.init:
                    b = self._w(other[0], add).astype(np.float32).reshape(-1)
                    self.done.add(id(add))
                    out = add.outputs[0]
        w = np.ascontiguousarray(w)
        cout, cin = w.shape
        if cin != self.channels[src]:
            raise UnsupportedOnnx(f"{n.op_type} '{n.name}': weight is {cout}x{cin}, pooled vector has {self.channels[src]}")
        bn = self._sole_consumer(out, 'BatchNormalization')
        if bn is not None:
            w, b = _fold_bn(w, b, self._bn_params(bn))
            self.done.add(id(bn))
            out = bn.outputs[0]
        relu = self._sole_consumer(out, 'Relu')
        if relu is None:
            raise UnsupportedOnnx(f"{n.op_type} '{n.name}': the embedding head must end in Relu (Linear + BN + ReLU)")
        self.done.add(id(relu))
        out = relu.outputs[0]
        name = self._name(n, 'fc')
        self.weights[name] = (w, b)
        self.ops.append(('fc', name, cin, cout, src, out))
        self.channels[out] = cout

    # -------------------------------------------------------------------------------------------- driver
    _SHAPE_JUNK = ('Shape', 'Gather', 'Unsqueeze', 'Concat', 'Cast')

    def run(self):
        for n in self.g.nodes:
            if id(n) in self.done:
                continue
            t = n.op_type
            if t == 'Conv':
                self._conv(n)
            elif t in ('MaxPool', 'AveragePool'):
                self._pool(n)
            elif t == 'GlobalAveragePool':
                self._gap(n)
            elif t == 'Add':
                self._add(n)
            elif t in ('Gemm', 'MatMul'):
                self._fc(n)
            elif t in ('Flatten', 'Reshape', 'Squeeze', 'Identity', 'Dropout'):
                self.alias[n.outputs[0]] = n.inputs[0]
            elif t == 'Constant':
                if isinstance(n.attrs.get('value'), np.ndarray):
                    self.init[n.outputs[0]] = n.attrs['value']      # e.g. the pads operand of a Pad node
            elif t == 'Pad':
                # opset-9 exporters put an explicit all-zero Pad in front of AveragePool: a no-op
                pads = n.attrs.get('pads')
                if pads is None and len(n.inputs) > 1 and n.inputs[1] in self.init:
                    pads = np.asarray(self.init[n.inputs[1]]).reshape(-1).tolist()
                if pads is None or any(int(p) != 0 for p in pads):
                    raise UnsupportedOnnx(f"Pad '{n.name}' with pads {pads} (only the all-zero Pad of old exporters)")
                self.alias[n.outputs[0]] = n.inputs[0]
            elif t in self._SHAPE_JUNK:
                continue            # shape arithmetic feeding a Reshape of the pooled vector (old exporters)
            else:
                raise UnsupportedOnnx(f"unsupported ONNX node {t} '{n.name}'")
        if len(self.g.outputs) != 1:
            raise UnsupportedOnnx(f"expected one graph output, found {[v.name for v in self.g.outputs]}")
        out = self._resolve_alias(self.g.outputs[0].name)
        if not self.ops or self.ops[-1][0] != 'fc' or self.ops[-1][5] != out:
            raise UnsupportedOnnx("the graph output is not the embedding head (GAP -> Linear [+BN] + ReLU)")
        return _canonical_order(self.ops), self.weights


def _canonical_order(ops):
    """torchreid's OSBlock.forward evaluates conv3 before the downsample branch; the engine's OSBlock matcher (and
    `build_osnet`) list the downsample first.  Swap `conv3, downsample, add_relu` triples into that order."""
    ops = list(ops)
    for i in range(len(ops) - 2):
        a, b, c = ops[i], ops[i + 1], ops[i + 2]
        if a[0] == 'conv' and b[0] == 'conv' and c[0] == 'add_relu' and b[8] != a[9] \
                and {a[9], b[9]} == {c[1], c[2]} and i > 0 and ops[i - 1][0] in ('gate4', 'gate') \
                and a[8] == ops[i - 1][4]:
            ops[i], ops[i + 1] = b, a
    return ops


def import_reid_onnx(src):
    """src: path, bytes or onnx_io.Graph.  Returns (ops, weights, input_shape (c, h, w), feature_dim)."""
    g = src if isinstance(src, Graph) else onnx_io.parse_model(src) if isinstance(src, (bytes, bytearray, memoryview)) \
        else onnx_io.load(src)
    low = _Lowering(g)
    ops, weights = low.run()
    return ops, weights, low.input_shape, ops[-1][3]


# ------------------------------------------------------------------------------------------------ exporter
def export_reid_onnx(ops, weights, input_shape=(3, 256, 128), unfold_bn=False, seed=0, torchreid_order=True):
    """Op list + weights -> onnx_io.Graph in the form torch.onnx.export gives torchreid models: NCHW Conv nodes with
    [out][in][kh][kw] weights, gates as GAP/Conv/Relu/Conv/Sigmoid/Mul, stream sums as Add chains, conv3 before the
    downsample branch.  `unfold_bn=True` writes every conv / dw / fc as <op without bias> + BatchNormalization with
    seeded statistics whose fold reproduces the given weights up to fp32 rounding (exercises the importer's BN fold).
    """
    rng = np.random.default_rng(seed)
    nodes, init = [], {}
    c0, h0, w0 = input_shape
    tname = {'input': 'input'}

    def T(buf):
        return tname.setdefault(buf, f"t_{buf}")

    def bn_split(name, w, b):
        """w, b -> (w', scale, beta, mean, var) with fold(w') == w (up to rounding)."""
        cout = w.shape[0]
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        var = rng.uniform(0.5, 2.0, cout).astype(np.float32)
        mean = rng.normal(0, 0.1, cout).astype(np.float32)
        k = scale / np.sqrt(var.astype(np.float64) + 1e-5)
        w2 = (w / k.reshape((-1,) + (1,) * (w.ndim - 1))).astype(np.float32)
        beta = (b + mean * k).astype(np.float32)
        init[name + '.bn.weight'], init[name + '.bn.bias'] = scale, beta
        init[name + '.bn.running_mean'], init[name + '.bn.running_var'] = mean, var
        return w2

    def emit_affine(op_type, name, x, w, b, attrs, act, out):
        """Conv / Gemm (+BN) (+Relu) writing tensor `out`."""
        cur = f"{name}_raw" if (unfold_bn or act == 'relu') else out
        if unfold_bn:
            w = bn_split(name, w, b)
            init[name + '.weight'] = w
            nodes.append(Node(op_type, [x, name + '.weight'], [cur], name, attrs))
            nxt = f"{name}_bn" if act == 'relu' else out
            nodes.append(Node('BatchNormalization', [cur, name + '.bn.weight', name + '.bn.bias',
                                                     name + '.bn.running_mean', name + '.bn.running_var'], [nxt],
                              name + '.bn', {'epsilon': 1e-5, 'momentum': 0.9}))
            cur = nxt
        else:
            init[name + '.weight'], init[name + '.bias'] = w, b
            nodes.append(Node(op_type, [x, name + '.weight', name + '.bias'], [cur], name, attrs))
        if act == 'relu':
            nodes.append(Node('Relu', [cur], [out], name + '.relu'))

    ops = list(ops)
    if torchreid_order:         # undo the canonical order: conv3 first, then the downsample branch
        for i in range(len(ops) - 2):
            a, b_, c = ops[i], ops[i + 1], ops[i + 2]
            if a[0] == 'conv' and b_[0] == 'conv' and c[0] == 'add_relu' and i > 0 and ops[i - 1][0] == 'gate4' \
                    and b_[8] == ops[i - 1][4] and a[8] != ops[i - 1][4]:
                ops[i], ops[i + 1] = b_, a
    for op in ops:
        kind = op[0]
        if kind == 'conv':
            _, name, cin, cout, k, stride, pad, act, src, dst = op
            w, b = weights[name]
            emit_affine('Conv', name, T(src), np.ascontiguousarray(np.asarray(w).transpose(0, 3, 1, 2)), np.asarray(b),
                        {'dilations': [1, 1], 'group': 1, 'kernel_shape': [k, k], 'pads': [pad] * 4,
                         'strides': [stride, stride]}, act, T(dst))
        elif kind == 'dw':
            _, name, c, act, src, dst = op
            w, b = weights[name]
            emit_affine('Conv', name, T(src), np.ascontiguousarray(np.asarray(w).T.reshape(c, 1, 3, 3)), np.asarray(b),
                        {'dilations': [1, 1], 'group': c, 'kernel_shape': [3, 3], 'pads': [1] * 4, 'strides': [1, 1]},
                        act, T(dst))
        elif kind == 'maxpool3s2':
            nodes.append(Node('MaxPool', [T(op[1])], [T(op[2] + "'") if op[1] == op[2] else T(op[2])], f"maxpool_{len(nodes)}",
                              {'kernel_shape': [3, 3], 'pads': [1] * 4, 'strides': [2, 2]}))
            if op[1] == op[2]:
                tname[op[2]] = tname.pop(op[2] + "'")
        elif kind == 'avgpool2':
            nodes.append(Node('AveragePool', [T(op[1])], [T(op[2])], f"avgpool_{len(nodes)}",
                              {'kernel_shape': [2, 2], 'pads': [0] * 4, 'strides': [2, 2]}))
        elif kind in ('gate', 'gate4'):
            name, c = op[1], op[2]
            srcs = list(op[3]) if kind == 'gate4' else [op[3]]
            acc = op[4]
            w1, b1, w2, b2 = (np.asarray(a) for a in weights[name])
            if name + '.fc1.weight' not in init:
                init[name + '.fc1.weight'] = w1.reshape(w1.shape[0], c, 1, 1).copy()
                init[name + '.fc1.bias'] = b1
                init[name + '.fc2.weight'] = w2.reshape(c, w1.shape[0], 1, 1).copy()
                init[name + '.fc2.bias'] = b2
            gated = []
            for s in srcs:
                p = f"{name}.{s}"
                x = T(s)
                nodes.append(Node('GlobalAveragePool', [x], [p + '_gap'], p + '.gap'))
                nodes.append(Node('Conv', [p + '_gap', name + '.fc1.weight', name + '.fc1.bias'], [p + '_fc1'], p + '.fc1',
                                  {'dilations': [1, 1], 'group': 1, 'kernel_shape': [1, 1], 'pads': [0] * 4,
                                   'strides': [1, 1]}))
                nodes.append(Node('Relu', [p + '_fc1'], [p + '_r'], p + '.relu'))
                nodes.append(Node('Conv', [p + '_r', name + '.fc2.weight', name + '.fc2.bias'], [p + '_fc2'], p + '.fc2',
                                  {'dilations': [1, 1], 'group': 1, 'kernel_shape': [1, 1], 'pads': [0] * 4,
                                   'strides': [1, 1]}))
                nodes.append(Node('Sigmoid', [p + '_fc2'], [p + '_s'], p + '.sigmoid'))
                nodes.append(Node('Mul', [x, p + '_s'], [p + '_g'], p + '.mul'))
                gated.append(p + '_g')
            if kind == 'gate' and op[5]:
                gated.insert(0, T(acc))
                tname.pop(acc)
            cur = gated[0]
            for i, gname in enumerate(gated[1:]):
                nxt = f"{name}.sum{i}.{acc}"
                nodes.append(Node('Add', [cur, gname], [nxt], nxt))
                cur = nxt
            tname[acc] = cur
        elif kind == 'add_relu':
            _, a, b_, dst = op
            s = f"add_{len(nodes)}"
            nodes.append(Node('Add', [T(a), T(b_)], [s], s))
            nodes.append(Node('Relu', [s], [T(dst)], s + '.relu'))
        elif kind == 'gap':
            nodes.append(Node('GlobalAveragePool', [T(op[1])], [T(op[2]) + '_4d'], 'global_avgpool'))
            nodes.append(Node('Flatten', [T(op[2]) + '_4d'], [T(op[2])], 'flatten', {'axis': 1}))
        elif kind == 'fc':
            _, name, cin, cout, src, dst = op
            w, b = weights[name]
            emit_affine('Gemm', name, T(src), np.asarray(w), np.asarray(b), {'alpha': 1.0, 'beta': 1.0, 'transB': 1},
                        'relu', T(dst))
            if unfold_bn:       # Gemm needs its C operand even when the BN carries the bias
                pass
        else:
            raise ValueError(f"cannot export op {op}")
    feat = ops[-1][5]
    return Graph(nodes, init, [ValueInfo('input', onnx_io.FLOAT, ('batch', c0, h0, w0))],
                 [ValueInfo(T(feat), onnx_io.FLOAT, ('batch', ops[-1][3]))], name='reid', opset=11,
                 producer='fastmot_b200.export_reid_onnx')
