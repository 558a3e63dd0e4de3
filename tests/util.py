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


def build_window(ctx, oracle, seq, n_keys, frames_per_key, first_id=1000):
    """Drive GPU and oracle side by side over the same frames; the GPU side is fed the ORACLE's
    undistorted clouds (staged parity)."""
    from fflins import api
    leaf = 0.5
    F = seq.filter_num
    keys = []
    fid = first_id
    fi = 0
    for k in range(n_keys):
        group = []
        for j in range(frames_per_key):
            fr = seq.frame(fi)
            fi += 1
            und, R, t, _ = oracle.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl,
                                            seq.t_bl, fr["stamp"], fr["td"], threads=8)
            dec = oracle.decimate(und, F)
            down = oracle.voxel_filter(dec, leaf)
            world = oracle.project(R, t, down)
            ctx.upload(fid, api.CLOUD_MAP_FULL, und)
            ctx.upload(fid, api.CLOUD_LIDAR_FULL, dec)
            ctx.upload(fid, api.CLOUD_LIDAR, down)
            ctx.upload(fid, api.CLOUD_WORLD, world)
            ctx.set_pose(fid, R, t)
            group.append(dict(id=fid, und=und, down=down, world=world, R=R, t=t, fr=fr))
            fid += 1
        key = group[-1]
        # oracle merge (lidar_map.cc:190-211)
        clouds = [key["und"]]
        for g in group[:-1]:
            Rr, tr = oracle.relative_pose(key["R"], key["t"], g["R"], g["t"])
            clouds.append(oracle.project(Rr, tr, g["und"]))
        full = np.concatenate(clouds)
        key["map_full"] = full
        key["map"] = oracle.voxel_filter(full, leaf)
        info = ctx.keyframe_merge(key["id"], [g["id"] for g in group[:-1]])
        assert info.n_map == len(key["map"])
        for g in group[:-1]:
            ctx.release(g["id"])
        ctx.keyframe_add(key["id"])
        ctx.set_motion(key["id"], key["fr"]["v"], key["fr"]["omega"], key["fr"]["td"])
        keys.append(key)
    return keys


def window_params(seq, keys, rng):
    cur = keys[-1]
    pose_refs = np.stack([seq.imu_pose7(k["fr"]["stamp"], rng, 0.01, 0.002) for k in keys[:-1]])
    pose_cur = seq.imu_pose7(cur["fr"]["stamp"], rng, 0.01, 0.002)
    return pose_refs, pose_cur, seq.ext7(), cur["fr"]["td"] + 0.003
