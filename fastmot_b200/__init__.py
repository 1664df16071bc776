"""fastmot_b200 — B200-native implementation of the FastMOT per-frame hot path.

Public names mirror fastmot/__init__.py:1-7 of the reference.
"""
from .track import Track
from .kalman_filter import KalmanFilter, MeasType
from .flow import Flow
from .tracker import MultiTracker, DeviceEmbeddings
