"""§8 f4: ONNX front end for ReID models (role of trt.OnnxParser in fastmot/models/reid.py:47-63).

CPU: wire-format codec round trips; exporter -> importer reproduces the op list and weights; an independent ONNX
interpreter (torch functional ops applied node by node with ONNX's own NCHW / [out][in][kh][kw] conventions) agrees
with the oracle executor run on the imported op list; unsupported nodes raise by name.
GPU: a `ReID` descriptor with MODEL_PATH runs through FeatureExtractor on the fused CUDA path and reproduces the
built-in engine fed the same weights.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fastmot_b200.models import osnet, onnx_io
from fastmot_b200.models.onnx_import import export_reid_onnx, import_reid_onnx, UnsupportedOnnx


def _interpret(g, x):
    """Reference semantics of the ONNX nodes the exporter / torchreid emit (ONNX operator spec, opset 11)."""
    env = {g.inputs[0].name: x}
    env.update({k: torch.as_tensor(np.array(v)) for k, v in g.initializers.items()})
    for n in g.nodes:
        i = [env[t] for t in n.inputs]
        a = n.attrs
        t = n.op_type
        if t == 'Conv':
            y = F.conv2d(i[0], i[1], i[2] if len(i) > 2 else None, stride=a.get('strides', [1, 1]),
                         padding=a.get('pads', [0, 0, 0, 0])[:2],
                         groups=a.get('group', 1))
        elif t == 'BatchNormalization':
            shape = (1, -1) + (1,) * (i[0].dim() - 2)
            y = (i[0] - i[3].reshape(shape)) / torch.sqrt(i[4].reshape(shape) + a['epsilon']) * i[1].reshape(shape) \
                + i[2].reshape(shape)
        elif t == 'Relu':
            y = F.relu(i[0])
        elif t == 'Sigmoid':
            y = torch.sigmoid(i[0])
        elif t == 'Mul':
            y = i[0] * i[1]
        elif t == 'Add':
            y = i[0] + i[1]
        elif t == 'MaxPool':
            y = F.max_pool2d(i[0], a['kernel_shape'], a['strides'], a['pads'][:2])
        elif t == 'AveragePool':
            y = F.avg_pool2d(i[0], a['kernel_shape'], a['strides'], a['pads'][:2])
        elif t == 'GlobalAveragePool':
            y = i[0].mean((2, 3), keepdim=True)
        elif t == 'Flatten':
            y = i[0].flatten(1)
        elif t == 'Gemm':
            w = i[1].T if a.get('transB', 0) else i[1]
            y = i[0] @ w + (i[2] if len(i) > 2 else 0)
        else:
            raise AssertionError(t)
        env[n.outputs[0]] = y
    return env[g.outputs[0].name]


def test_wire_codec_roundtrip():
    rng = np.random.default_rng(0)
    init = {'w': rng.normal(size=(4, 3, 1, 1)).astype(np.float32), 'h': rng.normal(size=(5,)).astype(np.float16),
            'idx': np.array([-1, 2, 1 << 40], np.int64), 'scalar': np.float32(2.5).reshape(())}
    nodes = [onnx_io.Node('Conv', ['x', 'w'], ['y'], 'c0', {'strides': [2, 2], 'group': 1, 'alpha': 0.25,
                                                            'auto_pad': 'NOTSET', 'pads': [1, 1, 1, 1],
                                                            'value': np.arange(3, dtype=np.int64), 'neg': -7}),
             onnx_io.Node('Relu', ['y'], ['z'], '')]
    g = onnx_io.Graph(nodes, init, [onnx_io.ValueInfo('x', onnx_io.FLOAT, ('N', 3, 8, 8))],
                      [onnx_io.ValueInfo('z', onnx_io.FLOAT, ('N', 4, 4, 4))], opset=13)
    for typed in (False, True):
        g2 = onnx_io.parse_model(onnx_io.serialize(g, typed_float_data=typed))
        assert g2.opset == 13 and [n.op_type for n in g2.nodes] == ['Conv', 'Relu']
        assert g2.nodes[0].inputs == ['x', 'w'] and g2.nodes[0].outputs == ['y'] and g2.nodes[0].name == 'c0'
        at = g2.nodes[0].attrs
        assert at['strides'] == [2, 2] and at['group'] == 1 and at['alpha'] == 0.25 and at['auto_pad'] == 'NOTSET'
        assert at['neg'] == -7 and at['value'].tolist() == [0, 1, 2]
        for k, v in init.items():
            assert g2.initializers[k].dtype == v.dtype and g2.initializers[k].shape == v.shape
            np.testing.assert_array_equal(g2.initializers[k], v)
        assert g2.inputs[0].shape == ('N', 3, 8, 8) and g2.outputs[0].name == 'z'
    # a weight listed among the graph inputs (IR version 3 files) is not a network input
    g.inputs.append(onnx_io.ValueInfo('w', onnx_io.FLOAT, (4, 3, 1, 1)))
    assert [v.name for v in onnx_io.parse_model(onnx_io.serialize(g)).inputs] == ['x']
    with pytest.raises(ValueError):
        onnx_io.parse_model(onnx_io.serialize(g)[:-7] + b'\xff')


def _same_structure(ops, ops2, w, w2, tol):
    ren, wn = {}, {}

    def same(a, b):
        return ren.setdefault(a, b) == b
    assert len(ops) == len(ops2)
    for a, b in zip(ops, ops2):
        assert a[0] == b[0], (a, b)
        k = a[0]
        if k == 'conv':
            assert a[2:8] == b[2:8] and same(a[8], b[8]) and same(a[9], b[9]), (a, b)
            wn[a[1]] = b[1]
        elif k == 'dw':
            assert a[2:4] == b[2:4] and same(a[4], b[4]) and same(a[5], b[5]), (a, b)
            wn[a[1]] = b[1]
        elif k in ('maxpool3s2', 'avgpool2', 'gap'):
            assert ren[a[1]] == b[1]
            ren[a[2]] = b[2]
        elif k == 'gate4':
            assert a[2] == b[2] and all(same(x, y) for x, y in zip(a[3], b[3])) and same(a[4], b[4]), (a, b)
            wn[a[1]] = b[1]
        elif k == 'add_relu':
            assert {ren[a[1]], ren[a[2]]} == {b[1], b[2]} and same(a[3], b[3]), (a, b)
        elif k == 'fc':
            assert a[2:4] == b[2:4] and same(a[4], b[4]), (a, b)
            wn[a[1]] = b[1]
    for n1, n2 in wn.items():
        for x, y in zip(w[n1], w2[n2]):
            assert x.shape == y.shape and x.dtype == y.dtype == np.float32
            assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-6), n1


@pytest.mark.parametrize("unfold_bn", [False, True])
def test_export_import_reproduces_osnet(unfold_bn):
    ops = osnet.build_osnet(0.25)
    w = osnet.synthetic_weights(ops, calibrate=False)
    data = onnx_io.serialize(export_reid_onnx(ops, w, unfold_bn=unfold_bn), typed_float_data=unfold_bn)
    ops2, w2, in_shape, dim = import_reid_onnx(data)
    assert in_shape == (3, 256, 128) and dim == 512
    _same_structure(ops, ops2, w, w2, 0.0 if not unfold_bn else 2e-6)
    # the six OSBlocks lower to the fused form the engine matches: shared gate -> gate4, downsample before conv3
    assert sum(o[0] == 'gate4' for o in ops2) == 6 and not any(o[0] == 'gate' for o in ops2)


@pytest.mark.parametrize("unfold_bn", [False, True])
def test_imported_graph_matches_onnx_semantics(unfold_bn):
    """exporter and importer are checked against a third party: node-by-node ONNX semantics in torch."""
    from oracle import nets
    ops = osnet.build_osnet(0.25)
    w = osnet.synthetic_weights(ops)
    g = export_reid_onnx(ops, w, unfold_bn=unfold_bn, seed=3)
    x = torch.randn(2, 3, 256, 128, generator=torch.Generator().manual_seed(5))
    y = _interpret(onnx_io.parse_model(onnx_io.serialize(g)), x)
    y = (y / y.norm(dim=1, keepdim=True)).numpy()               # feature_extractor.py:88-98 _normalize
    ops2, w2, _, _ = import_reid_onnx(onnx_io.serialize(g))
    got = nets.run_osnet(ops2, w2, x).numpy()
    want = nets.run_osnet(ops, w, x).numpy()
    assert np.abs(y - want).max() < 2e-5, np.abs(y - want).max()         # exporter == ONNX semantics of the op list
    assert np.abs(got - y).max() < 2e-5, np.abs(got - y).max()           # importer == ONNX semantics of the file


def _custom_graph(ch=16):
    """A non-OSNet custom model: stem conv + ReLU, maxpool, a residual block with a depthwise conv, a two-stream gate
    sum (lowered to 'gate' + 'gate' accumulate), MatMul + Add head with BatchNorm."""
    rng = np.random.default_rng(7)
    N = onnx_io.Node
    init, nodes = {}, []

    def conv(name, x, cin, cout, k, stride, pad, relu=True, group=1):
        init[name + '.w'] = rng.normal(0, np.sqrt(2 / (k * k * cin / group)), (cout, cin // group, k, k)).astype(np.float32)
        init[name + '.b'] = rng.normal(0, 0.05, cout).astype(np.float32)
        nodes.append(N('Conv', [x, name + '.w', name + '.b'], [name + '_o'], name,
                       {'kernel_shape': [k, k], 'strides': [stride, stride], 'pads': [pad] * 4, 'group': group}))
        if relu:
            nodes.append(N('Relu', [name + '_o'], [name + '_r'], name + '.relu'))
        return name + ('_r' if relu else '_o')
    t = conv('stem', 'images', 3, ch, 3, 2, 1)
    nodes.append(N('MaxPool', [t], ['p'], 'pool', {'kernel_shape': [3, 3], 'strides': [2, 2], 'pads': [1] * 4}))
    a = conv('b1.pw', 'p', ch, ch, 1, 1, 0, relu=False)
    a = conv('b1.dw', a, ch, ch, 3, 1, 1, relu=True, group=ch)
    nodes.append(N('Add', [a, 'p'], ['b1_sum'], 'b1.add'))
    nodes.append(N('Relu', ['b1_sum'], ['b1'], 'b1.relu'))
    s0 = conv('s0', 'b1', ch, ch, 1, 1, 0)
    s1 = conv('s1', 'b1', ch, ch, 3, 1, 1)
    gated = []
    for i, s in enumerate((s0, s1)):
        for nm, shp in ((f'g{i}.fc1', (ch // 4, ch, 1, 1)), (f'g{i}.fc2', (ch, ch // 4, 1, 1))):
            init[nm + '.w'] = rng.normal(0, 0.3, shp).astype(np.float32)
            init[nm + '.b'] = rng.normal(0, 0.05, shp[0]).astype(np.float32)
        nodes += [N('GlobalAveragePool', [s], [f'g{i}_p'], f'g{i}.gap'),
                  N('Conv', [f'g{i}_p', f'g{i}.fc1.w', f'g{i}.fc1.b'], [f'g{i}_1'], f'g{i}.fc1', {'kernel_shape': [1, 1]}),
                  N('Relu', [f'g{i}_1'], [f'g{i}_2'], f'g{i}.r'),
                  N('Conv', [f'g{i}_2', f'g{i}.fc2.w', f'g{i}.fc2.b'], [f'g{i}_3'], f'g{i}.fc2', {'kernel_shape': [1, 1]}),
                  N('Sigmoid', [f'g{i}_3'], [f'g{i}_4'], f'g{i}.s'),
                  N('Mul', [f'g{i}_4', s], [f'g{i}_y'], f'g{i}.mul')]
        gated.append(f'g{i}_y')
    nodes.append(N('Add', gated, ['gs'], 'gsum'))
    t = conv('tail', 'gs', ch, 2 * ch, 1, 1, 0)
    nodes += [N('AveragePool', [t], ['ap'], 'ap', {'kernel_shape': [2, 2], 'strides': [2, 2], 'pads': [0] * 4}),
              N('GlobalAveragePool', ['ap'], ['gp'], 'gap'), N('Flatten', ['gp'], ['v'], 'flat', {'axis': 1})]
    init['head.w'] = rng.normal(0, 0.2, (2 * ch, 24)).astype(np.float32)        # MatMul: [in][out]
    init['head.b'] = rng.normal(0, 0.05, 24).astype(np.float32)
    for k, v in (('bn.s', rng.uniform(0.5, 1.5, 24)), ('bn.b', rng.normal(0, 0.1, 24)), ('bn.m', rng.normal(0, 0.1, 24)),
                 ('bn.v', rng.uniform(0.5, 2, 24))):
        init[k] = v.astype(np.float32)
    nodes += [N('MatMul', ['v', 'head.w'], ['mm'], 'head'), N('Add', ['mm', 'head.b'], ['mmb'], 'head.bias'),
              N('BatchNormalization', ['mmb', 'bn.s', 'bn.b', 'bn.m', 'bn.v'], ['bnout'], 'head.bn', {'epsilon': 1e-5}),
              N('Relu', ['bnout'], ['emb'], 'head.relu')]
    g = onnx_io.Graph(nodes, init, [onnx_io.ValueInfo('images', onnx_io.FLOAT, ('N', 3, 64, 32))],
                      [onnx_io.ValueInfo('emb', onnx_io.FLOAT, ('N', 24))])
    return g


def test_plain_residual_backbone_and_gate_chain():
    from oracle import nets
    g = _custom_graph()
    ops, w, in_shape, dim = import_reid_onnx(onnx_io.serialize(g))
    assert in_shape == (3, 64, 32) and dim == 24
    kinds = [o[0] for o in ops]
    assert kinds == ['conv', 'maxpool3s2', 'conv', 'dw', 'add_relu', 'conv', 'conv', 'gate', 'gate', 'conv', 'avgpool2',
                     'gap', 'fc'], kinds
    assert ops[7][5] is False and ops[8][5] is True and ops[7][4] == ops[8][4]
    x = torch.randn(3, 3, 64, 32, generator=torch.Generator().manual_seed(1))
    # the checker has no MatMul: present it as Gemm(transB=0), the same product
    g2 = onnx_io.parse_model(onnx_io.serialize(g))
    for n in g2.nodes:
        if n.op_type == 'MatMul':
            n.op_type, n.attrs = 'Gemm', {'transB': 0}
    y = _interpret(g2, x)
    y = (y / y.norm(dim=1, keepdim=True)).numpy()
    got = nets.run_osnet(ops, w, x).numpy()
    assert np.abs(got - y).max() < 2e-5, np.abs(got - y).max()


# ---- a torchreid-style OSNet written with torch.nn (Zhou et al. ICCV'19 / torchreid osnet.py module structure), used
# only to obtain a GENUINE torch.onnx.export file: the importer is then checked against the torch forward itself.
import torch.nn as nn


class _ConvLayer(nn.Module):
    def __init__(self, i, o, k, stride=1, pad=0):
        super().__init__()
        self.conv, self.bn = nn.Conv2d(i, o, k, stride, pad, bias=False), nn.BatchNorm2d(o)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class _Conv1x1Linear(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.conv, self.bn = nn.Conv2d(i, o, 1, bias=False), nn.BatchNorm2d(o)

    def forward(self, x):
        return self.bn(self.conv(x))


class _LightConv3x3(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.conv1 = nn.Conv2d(i, o, 1, bias=False)
        self.conv2 = nn.Conv2d(o, o, 3, 1, 1, bias=False, groups=o)
        self.bn = nn.BatchNorm2d(o)

    def forward(self, x):
        return F.relu(self.bn(self.conv2(self.conv1(x))))


class _ChannelGate(nn.Module):
    def __init__(self, c, r=16):
        super().__init__()
        self.gap, self.fc1, self.fc2 = nn.AdaptiveAvgPool2d(1), nn.Conv2d(c, c // r, 1), nn.Conv2d(c // r, c, 1)

    def forward(self, x):
        return x * torch.sigmoid(self.fc2(F.relu(self.fc1(self.gap(x)))))


class _OSBlock(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        m = o // 4
        self.conv1 = _ConvLayer(i, m, 1)
        self.conv2a = _LightConv3x3(m, m)
        self.conv2b = nn.Sequential(*[_LightConv3x3(m, m) for _ in range(2)])
        self.conv2c = nn.Sequential(*[_LightConv3x3(m, m) for _ in range(3)])
        self.conv2d = nn.Sequential(*[_LightConv3x3(m, m) for _ in range(4)])
        self.gate, self.conv3 = _ChannelGate(m), _Conv1x1Linear(m, o)
        self.downsample = _Conv1x1Linear(i, o) if i != o else None

    def forward(self, x):
        identity, x1 = x, self.conv1(x)
        x2 = self.gate(self.conv2a(x1)) + self.gate(self.conv2b(x1)) + self.gate(self.conv2c(x1)) + \
            self.gate(self.conv2d(x1))
        x3 = self.conv3(x2)
        if self.downsample is not None:
            identity = self.downsample(identity)
        return F.relu(x3 + identity)


class _TorchOSNet(nn.Module):
    def __init__(self, ch=(16, 64, 96, 128), dim=512):
        super().__init__()
        self.conv1, self.maxpool = _ConvLayer(3, ch[0], 7, 2, 3), nn.MaxPool2d(3, 2, 1)

        def stage(i, o, trans):
            mods = [_OSBlock(i, o), _OSBlock(o, o)]
            if trans:
                mods.append(nn.Sequential(_ConvLayer(o, o, 1), nn.AvgPool2d(2, 2)))
            return nn.Sequential(*mods)
        self.conv2, self.conv3, self.conv4 = stage(ch[0], ch[1], True), stage(ch[1], ch[2], True), stage(ch[2], ch[3], False)
        self.conv5, self.gap = _ConvLayer(ch[3], ch[3], 1), nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(ch[3], dim), nn.BatchNorm1d(dim), nn.ReLU())

    def forward(self, x):
        x = self.conv5(self.conv4(self.conv3(self.conv2(self.maxpool(self.conv1(x))))))
        v = self.gap(x)
        return self.fc(v.view(v.size(0), -1))


def _torch_export(model, x, opset):
    import io
    import warnings
    try:
        from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    except Exception:
        pytest.skip("this torch has no TorchScript ONNX exporter")
    # the exporter only needs the `onnx` package for custom onnxscript functions, which this graph has none of
    keep = onnx_proto_utils._add_onnxscript_fn
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto
    try:
        buf = io.BytesIO()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.onnx.export(model, x, buf, opset_version=opset, input_names=['images'], output_names=['features'],
                              dynamic_axes={'images': {0: 'batch'}, 'features': {0: 'batch'}}, dynamo=False)
        return buf.getvalue()
    finally:
        onnx_proto_utils._add_onnxscript_fn = keep


@pytest.mark.parametrize("opset", [9, 11, 17])
def test_genuine_torch_onnx_export_imports_and_matches_the_torch_forward(opset):
    """Third-party pin of the importer: a file written by torch.onnx.export (eval-mode BN folded into the convs,
    Shape / Gather / Concat / Reshape for the flatten, Gemm + BatchNormalization + Relu head, opset-9's zero Pad) must
    lower to the OSNet op list and reproduce the torch module's own forward."""
    from oracle import nets
    torch.manual_seed(0)
    model = _TorchOSNet().eval()
    for mod in model.modules():
        if isinstance(mod, (nn.BatchNorm2d, nn.BatchNorm1d)):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 2)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    x = torch.randn(2, 3, 256, 128)
    data = _torch_export(model, x, opset)
    g = onnx_io.parse_model(data)
    assert g.producer == 'pytorch' and g.opset == opset
    ops, w, in_shape, dim = import_reid_onnx(data)
    assert in_shape == (3, 256, 128) and dim == 512
    # the same structure as the built-in x0.25 network: the engine's fused OSBlock / stem matchers apply
    ref_ops = osnet.build_osnet(0.25)
    assert [o[0] for o in ops] == [o[0] for o in ref_ops]
    assert [o[2:8] for o in ops if o[0] == 'conv'] == [o[2:8] for o in ref_ops if o[0] == 'conv']
    with torch.no_grad():
        want = model(x)
    want = want / want.norm(dim=1, keepdim=True)
    got = nets.run_osnet(ops, w, x)
    assert float((got - want).abs().max()) < 1e-6, float((got - want).abs().max())


def test_unsupported_nodes_raise_by_name():
    ops = osnet.build_osnet(0.25)
    w = osnet.synthetic_weights(ops, calibrate=False)
    g = export_reid_onnx(ops, w)
    g.nodes[1] = onnx_io.Node('LeakyRelu', g.nodes[1].inputs, g.nodes[1].outputs, 'stem.act', {'alpha': 0.1})
    with pytest.raises(UnsupportedOnnx, match="LeakyRelu 'stem.act'"):
        import_reid_onnx(g)
    g = export_reid_onnx(ops, w)
    g.nodes[0].attrs['strides'] = [2, 1]
    with pytest.raises(UnsupportedOnnx, match="conv1"):
        import_reid_onnx(g)
    g = export_reid_onnx(ops, w)
    g.nodes = g.nodes[:-1]                  # embedding head without its ReLU
    g.nodes[-1].outputs = [g.outputs[0].name]
    with pytest.raises(UnsupportedOnnx, match="must end in Relu"):
        import_reid_onnx(g)


@pytest.mark.gpu
def test_reid_descriptor_with_onnx_runs_on_fused_cuda_path(tmp_path):
    from fastmot_b200 import FeatureExtractor, models
    from fastmot_b200.engine import OSNetEngine
    from fastmot_b200.synth import SyntheticScene
    ops = osnet.build_osnet(1.0)
    w = osnet.synthetic_weights(ops)
    path = tmp_path / "custom_osnet.onnx"
    onnx_io.save(export_reid_onnx(ops, w), str(path))

    class CustomOnnxReID(models.ReID):       # the reference's plugin API: subclass + class attributes (reid.py:10-45)
        MODEL_PATH = path
        INPUT_SHAPE = (3, 256, 128)
        OUTPUT_LAYOUT = 512
        METRIC = 'cosine'

    sc = SyntheticScene(24, seed=4)
    frame, tl = sc.frame(0), sc.detections(0)[0]
    fe = FeatureExtractor('CustomOnnxReID', use_graph=True)
    emb = np.asarray(fe(frame, tl))
    assert emb.shape == (24, 512) and fe.metric == 'cosine'
    eng = fe._engine(24)
    assert eng.n_osb == 6 and eng.fuse_stem, "imported graph must take the fused OSBlock / stem kernels"
    ref = OSNetEngine(1.0, weights=w, max_batch=24, use_graph=False)
    ref.inp.copy_(eng.inp)
    want = ref.forward(24).cpu().numpy()
    np.testing.assert_array_equal(emb, want)        # same kernels, same weights: bit-identical
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-4)


@pytest.mark.gpu
def test_custom_backbone_onnx_vs_oracle_on_gpu():
    """The non-OSNet graph of the CPU test (64 channels wide) through the generic layer kernels vs the fp32 oracle."""
    from fastmot_b200.engine import OSNetEngine
    from oracle import nets
    ops, w, in_shape, dim = import_reid_onnx(onnx_io.serialize(_custom_graph(64)))
    eng = OSNetEngine(None, weights=w, input_hw=in_shape[1:], feature_dim=dim, max_batch=8, use_graph=False, ops=ops)
    x = torch.randn(8, 3, 64, 32, generator=torch.Generator().manual_seed(11))
    inp = torch.zeros(8, 64, 32, 8, dtype=torch.float16)
    inp[..., :3] = x.permute(0, 2, 3, 1).half()
    eng.load_nhwc8(inp.cuda())
    got = eng.forward().cpu()
    want = nets.run_osnet(ops, w, inp[..., :3].float().permute(0, 3, 1, 2), nets.fp16_roundtrip)
    assert got.shape == want.shape == (8, 24)
    assert float((got - want).abs().max()) < 5e-3, float((got - want).abs().max())


def test_wire_codec_property_roundtrip():
    """Randomised graphs (hypothesis): every tensor dtype / shape, attribute kind and name survives serialize -> parse."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    dtypes = st.sampled_from([np.float32, np.float16, np.float64, np.int64, np.int32, np.uint8, np.int8])
    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz0123456789_./:", min_size=1, max_size=24)

    @st.composite
    def tensors(draw):
        dt = np.dtype(draw(dtypes))
        shape = draw(hnp.array_shapes(min_dims=0, max_dims=4, max_side=5))
        if dt.kind == 'f':
            elems = st.floats(-1e3, 1e3, width=16 if dt.itemsize == 2 else 32)
        else:
            info = np.iinfo(dt)
            elems = st.integers(int(info.min), int(info.max))
        return draw(hnp.arrays(dt, shape, elements=elems))

    attr_vals = st.one_of(st.integers(-2**62, 2**62), st.floats(-1e6, 1e6, width=32), names,
                          st.lists(st.integers(-2**40, 2**40), min_size=1, max_size=6),
                          st.lists(st.floats(-1e3, 1e3, width=32), min_size=1, max_size=6), tensors())

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(st.dictionaries(names, tensors(), max_size=5),
           st.lists(st.tuples(names, st.lists(names, max_size=3), st.lists(names, min_size=1, max_size=2),
                              st.dictionaries(names, attr_vals, max_size=4)), max_size=5),
           st.integers(1, 21), st.booleans())
    def check(init, nodes, opset, typed):
        g = onnx_io.Graph([onnx_io.Node(op, list(i), list(o), nm, dict(a)) for (nm, i, o, a), op in
                           zip(nodes, ["Conv", "Relu", "Add", "Gemm", "Custom"] * 2)],
                          dict(init), [onnx_io.ValueInfo("x", onnx_io.FLOAT, ("N", 3, 8, 8))],
                          [onnx_io.ValueInfo("y", onnx_io.FLOAT16, (2, "M"))], opset=opset)
        g.inputs = [v for v in g.inputs if v.name not in g.initializers]
        g2 = onnx_io.parse_model(onnx_io.serialize(g, typed_float_data=typed))
        assert g2.opset == opset and len(g2.nodes) == len(g.nodes)
        for a, b in zip(g.nodes, g2.nodes):
            assert (a.op_type, a.inputs, a.outputs, a.name) == (b.op_type, b.inputs, b.outputs, b.name)
            assert set(a.attrs) == set(b.attrs)
            for k, v in a.attrs.items():
                w = b.attrs[k]
                if isinstance(v, np.ndarray):
                    assert w.dtype == v.dtype and w.shape == v.shape and np.array_equal(w, v, equal_nan=True)
                elif isinstance(v, float):
                    assert w == np.float32(v)
                elif isinstance(v, list) and v and isinstance(v[0], float):
                    assert w == [float(np.float32(x)) for x in v]
                else:
                    assert w == v, (k, v, w)
        assert set(g2.initializers) == set(g.initializers)
        for k, v in g.initializers.items():
            w = g2.initializers[k]
            assert w.dtype == v.dtype and w.shape == v.shape and np.array_equal(w, v, equal_nan=True)
        assert [(v.name, v.elem_type, v.shape) for v in g2.outputs] == [("y", onnx_io.FLOAT16, (2, "M"))]

    check()
