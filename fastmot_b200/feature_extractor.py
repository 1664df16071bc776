"""ReID feature extractor with the reference's API (fastmot/feature_extractor.py:11-98):
`FeatureExtractor(model, batch_size)`, `extract_async(frame, tlbrs)`, `postprocess() -> (N, 512)` unit-norm rows,
`__call__`, `metric`, `null_embeddings`.

All crops of a frame are pre-processed by one kernel launch straight from the device-resident frame (the reference
crops + cv2.resize's on CPU threads and uploads 16 crops at a time) and run through the OSNet engine in one batch;
the embeddings stay on the GPU (`DeviceEmbeddings`) and feed the association kernels directly.  The reference's
aliasing defect for > batch_size crops is not reproduced.
"""
import numpy as np
import torch

from . import _lib, models
from .devmem import ptr, stream_ptr, FrameUploader
from .tracker import DeviceEmbeddings


class FeatureExtractor:
    def __init__(self, model='OSNet025', batch_size=16, max_crops=512, size=(1920, 1080), use_tc=True,
                 use_graph=True):
        self._lib = _lib.require_device()
        self.model = models.ReID.get_model(model)
        assert batch_size >= 1
        self.batch_size = batch_size
        self.feature_dim = self.model.OUTPUT_LAYOUT
        self.max_crops = max_crops
        self.size = size
        self._use_tc, self._use_graph = use_tc, use_graph
        self._engines = {}
        dev = torch.device("cuda")
        self._tlbr_dev = torch.zeros(max_crops, 4, dtype=torch.float64, device=dev)
        self._tlbr_host = torch.zeros(max_crops, 4, dtype=torch.float64).pin_memory()
        self._uploader = None
        self.last_num_features = 0
        self._out = None

    def _engine(self, n):
        """Engines are planned per batch bucket (multiples of 8 crops): a frame pays for its own crops, not for
        max_crops."""
        from .engine import build_reid_engine
        b = max(8, -(-n // 8) * 8)
        if b not in self._engines:
            self._engines[b] = build_reid_engine(self.model, max_batch=b, use_tc=self._use_tc,
                                                 use_graph=self._use_graph)
        return self._engines[b]

    def __call__(self, frame, tlbrs):
        self.extract_async(frame, tlbrs)
        return self.postprocess()

    @property
    def metric(self):
        return self.model.METRIC

    def extract_async(self, frame, tlbrs):
        """frame: HxWx3 u8 host array or cuda tensor; tlbrs: (N,4)."""
        tlbrs = np.ascontiguousarray(tlbrs, np.float64).reshape(-1, 4)
        n = len(tlbrs)
        self.last_num_features = n
        self._out = None
        if n == 0:
            return
        if n > self.max_crops:
            raise MemoryError(f"{n} crops > max_crops {self.max_crops}")
        if torch.is_tensor(frame):
            frame_dev = frame
            h, w = frame.shape[:2]
        else:
            h, w = frame.shape[:2]
            if self._uploader is None or self._uploader.shape != (h, w, 3):
                self._uploader = FrameUploader((w, h))
            frame_dev = self._uploader.upload(frame)
        eng = self._engine(n)
        self._tlbr_host[:n] = torch.as_tensor(tlbrs)
        self._tlbr_dev[:n].copy_(self._tlbr_host[:n], non_blocking=True)
        c, ih, iw = self.model.INPUT_SHAPE
        rc = self._lib.fm_roi_resize_norm(ptr(frame_dev), w, h, ptr(self._tlbr_dev), None, n, iw, ih,
                                          eng.inp_layout, ptr(eng.inp), stream_ptr())
        _lib.check(rc, "fm_roi_resize_norm")
        self._out = eng.forward(n)

    def postprocess(self):
        """Returns DeviceEmbeddings (N, feature_dim) — numpy-convertible, rows L2-normalised."""
        if self.last_num_features == 0:
            return np.empty((0, self.feature_dim))
        return DeviceEmbeddings(self._out)

    def null_embeddings(self, detections):
        embeddings = np.ones((len(detections), self.feature_dim))
        embeddings /= np.linalg.norm(embeddings, axis=1, keepdims=True)
        return embeddings
