"""Mirror of mlx_nerf/ops/pose.py: synthetic camera poses on a sphere.  Host code (a 4x4 matrix)."""
import numpy as np
import torch


def pose_spherical(theta, phi, radius):
    """pose_spherical (ops/pose.py:7-58): camera-to-world matrix looking at the origin from (theta, phi, radius),
    angles in degrees.  fp32 factors multiplied in the reference's order; returns a [4, 4] fp32 CPU tensor."""
    def _trans_radius(r):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, r], [0, 0, 0, 1]], dtype=np.float32)

    def _rotate_phi(a):
        return np.array([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0], [0, 0, 0, 1]],
                        dtype=np.float32)

    def _rotate_theta(a):
        return np.array([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]],
                        dtype=np.float32)

    pose = _trans_radius(radius)
    pose = _rotate_phi(phi / 180.0 * np.pi) @ pose
    pose = _rotate_theta(theta / 180.0 * np.pi) @ pose
    swap = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return torch.from_numpy((swap @ pose).astype(np.float32))
