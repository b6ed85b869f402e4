"""Camera calibration of the CARLA front camera (drop-in for the two helpers of the reference's
``simlingo_training/utils/projection.py:24-61`` that ``dl_collate_fn`` needs)."""
import math

import torch


def get_camera_intrinsics(w, h, fov):
    """float32 [3, 3] pinhole matrix of a ``w`` x ``h`` image with horizontal field of view ``fov`` (degrees)."""
    focal = w / (2.0 * math.tan(fov * math.pi / 360.0))
    return torch.tensor([[focal, 0.0, w / 2.0], [0.0, focal, h / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64).to(torch.float32)


def get_camera_extrinsics():
    """float32 [4, 4] homogeneous ``[R t; 0 1]``: identity rotation, camera mounted at x = -1.5, y = 0, z = 2.0."""
    m = torch.eye(4, dtype=torch.float32)
    m[:3, 3] = torch.tensor([-1.5, 0.0, 2.0])
    return m
