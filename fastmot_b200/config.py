"""Default configuration of the tracker stages: the values of the reference's cfg/mot.json:44-96, as plain data
(benchmarks, tests and examples build their `MOT(...)` kwargs from here)."""
from types import SimpleNamespace as NS


def default_tracker_cfg():
    return dict(max_age=6, age_penalty=2, motion_weight=0.2, max_assoc_cost=0.8, max_reid_cost=0.6, iou_thresh=0.4,
                duplicate_thresh=0.8, occlusion_thresh=0.7, conf_thresh=0.5, confirm_hits=1, history_size=50,
                kalman_filter_cfg=NS(std_factor_acc=2.25, std_offset_acc=78.5, std_factor_det=(0.08, 0.08),
                                     std_factor_klt=(0.14, 0.14), min_std_det=(4.0, 4.0), min_std_klt=(5.0, 5.0),
                                     init_pos_weight=5, init_vel_weight=12, vel_coupling=0.6, vel_half_life=2),
                flow_cfg=NS(bg_feat_scale_factor=(0.1, 0.1), opt_flow_scale_factor=(0.5, 0.5), feat_density=0.005,
                            feat_dist_factor=0.06, ransac_max_iter=500, ransac_conf=0.99, max_error=100,
                            inlier_thresh=4, bg_feat_thresh=10,
                            obj_feat_params=NS(maxCorners=1000, qualityLevel=0.06, blockSize=3),
                            opt_flow_params=NS(winSize=(5, 5), maxLevel=5, criteria=(3, 10, 0.03))))
