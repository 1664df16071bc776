"""CPU: host-side pieces of the drop-in boundary — Darknet cfg / .weights reader (scripts/yolo2onnx.py:86-205, 283-400),
model registries (fastmot/models/yolo.py:52-58, reid.py:39-45), label map, config decoder, benchmark scene."""
import json
import os
import struct
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

CFG = """
[net]
width=64
height=64
channels=3

[convolutional]
batch_normalize=1
filters=8
size=3
stride=2
pad=1
activation=leaky

[convolutional]
batch_normalize=1
filters=16
size=1
stride=1
pad=1
activation=mish

[route]
layers=-1
groups=2
group_id=1

[maxpool]
size=2
stride=2

[convolutional]
filters=18   # head: conv bias instead of BN
size=1
stride=1
pad=1
activation=linear

[yolo]
mask = 0,1,2
anchors = 10,14, 23,27, 37,58
classes=1
"""


def test_parse_cfg_fields_and_shapes():
    from fastmot_b200.models import darknet
    net, layers = darknet.parse_cfg(CFG)
    assert net['width'] == 64 and net['channels'] == 3
    assert [l['type'] for l in layers] == ['convolutional', 'convolutional', 'route', 'maxpool', 'convolutional', 'yolo']
    assert layers[0]['batch_normalize'] == 1 and layers[0]['activation'] == 'leaky' and layers[4]['filters'] == 18
    assert layers[2]['layers'] == [-1] and layers[2]['groups'] == 2 and layers[2]['group_id'] == 1
    assert layers[5]['anchors'] == [10, 14, 23, 27, 37, 58] and layers[5]['mask'] == [0, 1, 2]
    res, shapes = darknet.infer_shapes(layers, 3, 64, 64)
    assert shapes == [(8, 32, 32), (16, 32, 32), (8, 32, 32), (8, 16, 16), (18, 16, 16), (18, 16, 16)]
    assert res[2]['layers_abs'] == [1]
    # Darknet BFLOPs convention
    assert darknet.count_flops(layers, 3, 64, 64) == 2 * (3 * 9 * 8 * 32 * 32 + 8 * 16 * 32 * 32 + 8 * 18 * 16 * 16)


def test_load_weights_folds_batchnorm_like_the_converter(tmp_path):
    """Write a Darknet .weights file by hand (header, then per conv: BN beta, gamma, mean, var | conv bias, then
    weights [out][in][kh][kw]) and check the loaded, BN-folded layer against torch conv + batch_norm."""
    from fastmot_b200.models import darknet
    _, layers = darknet.parse_cfg(CFG)
    rng = np.random.default_rng(3)
    blobs, truth = [], {}
    specs = [(0, 3, 8, 3, True), (1, 8, 16, 1, True), (4, 8, 18, 1, False)]
    for idx, cin, cout, k, bn in specs:
        w = rng.normal(size=(cout, cin, k, k)).astype(np.float32)
        if bn:
            beta, gamma = rng.normal(size=cout).astype(np.float32), rng.uniform(0.5, 1.5, cout).astype(np.float32)
            mean, var = rng.normal(size=cout).astype(np.float32), rng.uniform(0.5, 2.0, cout).astype(np.float32)
            blobs += [beta, gamma, mean, var, w.ravel()]
            truth[idx] = (w, None, (beta, gamma, mean, var))
        else:
            b = rng.normal(size=cout).astype(np.float32)
            blobs += [b, w.ravel()]
            truth[idx] = (w, b, None)
    path = tmp_path / "net.weights"
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", 0, 2, 5))          # major, minor, revision
        f.write(struct.pack("<q", 12345))              # `seen` is 64-bit from version 0.2 on
        f.write(np.concatenate(blobs).astype(np.float32).tobytes())
    got = darknet.load_weights(str(path), layers, 3)
    assert sorted(got) == [0, 1, 4]
    x_by_layer = {0: torch.randn(1, 3, 9, 9), 1: torch.randn(1, 8, 5, 5), 4: torch.randn(1, 8, 5, 5)}
    for idx, (w, b, bn) in truth.items():
        x = x_by_layer[idx]
        k = w.shape[-1]
        ref = F.conv2d(x, torch.as_tensor(w), None if b is None else torch.as_tensor(b), padding=k // 2)
        if bn is not None:
            beta, gamma, mean, var = (torch.as_tensor(a) for a in bn)
            ref = F.batch_norm(ref, mean, var, gamma, beta, training=False, eps=1e-5)
        gw, gb = got[idx]
        assert gw.shape == (w.shape[0], k, k, w.shape[1])            # engine layout [out][kh][kw][in]
        out = F.conv2d(x, torch.as_tensor(gw).permute(0, 3, 1, 2), torch.as_tensor(gb), padding=k // 2)
        assert float((out - ref).abs().max()) < 1e-4
    # a file that is too short for the cfg is an error, not a silent zero fill
    short = tmp_path / "short.weights"
    short.write_bytes(open(path, "rb").read()[:200])
    with pytest.raises(ValueError):
        darknet.load_weights(str(short), layers, 3)


def test_published_flops_of_the_builders():
    """The cfg files are not in the reference tree; the builders restate the published nets.  Darknet prints 6.9
    BFLOPs for yolov4-tiny@416 and ~119 for yolov4-csp@640."""
    from fastmot_b200.models import darknet
    tiny = darknet.count_flops(darknet.BUILDERS['yolov4-tiny'](num_classes=80), 3, 416, 416) / 1e9
    csp = darknet.count_flops(darknet.BUILDERS['yolov4-csp'](num_classes=80), 3, 640, 640) / 1e9
    assert 6.7 < tiny < 7.1, tiny
    assert 110 < csp < 130, csp


def test_model_registries_and_label_map():
    from fastmot_b200 import models
    csp = models.YOLO.get_model('YOLOv4CSP')
    assert csp.INPUT_SHAPE == (3, 640, 640) and csp.LETTERBOX and csp.NEW_COORDS       # yolo.py:171-182
    tiny = models.YOLO.get_model('YOLOv4Tiny')
    assert tiny.INPUT_SHAPE == (3, 416, 416) and not tiny.LETTERBOX                    # yolo.py:256-264

    class MyYOLO(models.YOLO):                       # the plugin API: subclassing registers the model by name
        CFG = 'yolov4-tiny'
        NUM_CLASSES = 3
        INPUT_SHAPE = (3, 320, 320)
        LAYER_FACTORS = [32, 16]
        SCALES = [1.05, 1.05]
        ANCHORS = [[81, 82, 135, 169, 344, 319], [23, 27, 37, 58, 81, 82]]
    assert models.YOLO.get_model('MyYOLO') is MyYOLO
    osnet = models.ReID.get_model('OSNet10')
    assert osnet.OUTPUT_LAYOUT == 512 and osnet.INPUT_SHAPE == (3, 256, 128)           # reid.py:95-109
    with pytest.raises(KeyError):
        models.YOLO.get_model('NoSuchModel')
    old = list(models.label.LABEL_MAP)
    try:
        models.set_label_map(['pedestrian', 'cyclist'])
        assert models.label.get_label_name(1) == 'cyclist' and models.label.get_label_name(7) == 'class7'
    finally:
        models.set_label_map(old)


def test_config_decoder_gives_tuples_for_namespace_splat():
    from fastmot_b200.utils import ConfigDecoder
    txt = '{"resize_to": [1280, 720], "mot_cfg": {"class_ids": [1], "tracker_cfg": {"max_age": 6}}}'
    cfg = json.loads(txt, cls=ConfigDecoder, object_hook=lambda d: SimpleNamespace(**d))    # app.py:57-58
    assert cfg.resize_to == (1280, 720) and cfg.mot_cfg.class_ids == (1,)
    assert vars(cfg.mot_cfg.tracker_cfg) == {"max_age": 6}


def test_benchmark_scene_bounce_keeps_objects_in_their_cells():
    from fastmot_b200.synth import SyntheticScene
    lin = SyntheticScene(40, seed=1, label=0, dropout_frames=())
    bnc = SyntheticScene(40, seed=1, label=0, dropout_frames=(), bounce_radius=16)
    for t in (0, 5, 17):                                  # identical while every displacement is inside +-16 px
        if np.abs(lin.vel * t).max() <= 16:
            np.testing.assert_array_equal(np.stack(lin.positions(t)), np.stack(bnc.positions(t)))
    x0, y0 = bnc.positions(0)
    for t in (100, 441, 5000):
        x, y = bnc.positions(t)
        assert np.abs(x - x0).max() <= 16 and np.abs(y - y0).max() <= 16
        assert len(bnc.detections(t)[0]) == 40             # nothing leaves the frame
    # consecutive frames move by at most |v| (+ rounding): no jumps at the turning points
    xa, ya = bnc.positions(63)
    xb, yb = bnc.positions(64)
    assert np.abs(xb - xa).max() <= np.ceil(np.abs(bnc.vel[:, 0]).max()) + 1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the metric, the bounded
    sample description and the zero-copy e2e block.  Config 1 (association only) keeps it to a few seconds."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1",
                        "--steps", "6", "--warmup", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("frames/sec") and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 6 and d["warmup"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[0]" in d["config"]["workload"]
