"""Shared helpers of the parity tests."""
import numpy as np


def ulp_diff(a, b):
    """Distance in units-in-the-last-place between two float32 arrays (same shape)."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def block_rel(a, b):
    """max|a-b| / max|b| over a block (relative error is per block, not element-wise)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.max(np.abs(b)) if b.size else 0.0
    num = np.max(np.abs(a - b)) if b.size else 0.0
    return num / den if den > 0 else num


def random_scene_cloud(rng, n, extent=20.0, planes=True):
    """Points on a few planes plus clutter: exercises multi-point voxels."""
    pts = np.zeros((n, 4), np.float32)
    k = n // 2 if planes else 0
    pts[:k, 0] = rng.uniform(-extent, extent, k)
    pts[:k, 1] = rng.uniform(-extent, extent, k)
    pts[:k, 2] = rng.normal(0, 0.02, k)
    pts[k:, :3] = rng.uniform(-extent, extent, (n - k, 3)) * np.array([1, 1, 0.3])
    pts[:, 3] = rng.uniform(0, 255, n)
    return pts


def rand_pose(rng, trans=10.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R, rng.uniform(-trans, trans, 3)
