"""KLT flow front-end (API of fastmot/flow.py:16-264).  Placeholder until csrc/klt_*.cu lands."""
import numpy as np


class Flow:
    def __init__(self, size, bg_feat_scale_factor=(0.1, 0.1), opt_flow_scale_factor=(0.5, 0.5), feat_density=0.005,
                 feat_dist_factor=0.06, ransac_max_iter=500, ransac_conf=0.99, max_error=100, inlier_thresh=4,
                 bg_feat_thresh=10, obj_feat_params=None, opt_flow_params=None):
        self.size = size
        self.bg_keypoints = np.empty((0, 2), np.float32)
        self.prev_bg_keypoints = np.empty((0, 2), np.float32)
        self.pool = None

    def bind_pool(self, pool):
        self.pool = pool

    def init(self, frame):
        pass

    def predict_device(self, frame, tracks, h_dev, h_ok_dev):
        raise NotImplementedError("KLT kernels not built yet")

    def fetch_klt_bboxes(self, order):
        return {}
