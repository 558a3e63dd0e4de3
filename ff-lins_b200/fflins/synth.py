"""Seeded synthetic sequences for the FF-LINS LiDAR hot path.

A sensor rig drives a closed circular trajectory inside a 60 x 60 x 10 m room of axis-aligned
planes and boxes. Each raw point is ray-cast from the LiDAR pose AT ITS OWN TIMESTAMP, so the
raw cloud carries real motion distortion; the INS window is the analytic trajectory sampled at
the IMU rate plus small noise. Shapes follow the four reference configs
(config/ff_lins_{robot,lili,mcd_ouster64,i2nav_at128}.yaml): point count, decimation factor,
IMU rate, extrinsic and time delay.

The trajectory is periodic with `period` frames, so a benchmark can cycle over the frames of
one lap for ever: every frame call is self-contained (its own window, stamp and td).
"""
import math

import numpy as np

GPS_T0 = 345600.0  # absolute second-of-week style time base (fp32 time would not be enough)
FRAME_DT = 0.1

CONFIGS = {
    # name: n, filter_num, imu_hz, speed, td, q_b_l(xyzw), t_b_l, pattern
    "robot": dict(n=20000, filter_num=10, imu_hz=200, radius=6.0, period=250, td=0.0,
                  q_b_l=(0.0, 0.0, 0.0, 1.0), t_b_l=(0.0, 0.0, 0.0), pattern="rosette", index=0),
    "lili": dict(n=24000, filter_num=10, imu_hz=200, radius=6.0, period=250, td=0.0,
                 q_b_l=(0.0, 0.0, 0.0, 1.0), t_b_l=(0.05, 0.02, -0.06), pattern="horizon", index=1),
    "ouster64": dict(n=65536, filter_num=20, imu_hz=400, radius=10.0, period=125, td=-0.12,
                     q_b_l=(0.0, 0.0, 0.0, 1.0), t_b_l=(-0.05, 0.0, 0.055), pattern="spin64", index=2),
    "at128": dict(n=153600, filter_num=80, imu_hz=200, radius=10.0, period=125, td=0.0,
                  q_b_l=(1.0, 0.0, 0.0, 0.0), t_b_l=(0.0, 0.0, 0.0), pattern="at128", index=3),
    # BASELINE.json config 4 read as the "dense-map association stress case": the same AT128 scan in a street-sized
    # scene (400 x 400 x 40 m, building-sized boxes, ~20 m/s): surfaces 50-200 m away put only a few points into each
    # 0.5 m voxel, so a 5-frame keyframe map holds ~1.2 x 10^5 centroids instead of ~1.5 x 10^4 (Q ~ 1.4 k)
    "dense": dict(n=153600, filter_num=80, imu_hz=200, radius=50.0, period=160, td=0.0,
                  q_b_l=(1.0, 0.0, 0.0, 0.0), t_b_l=(0.0, 0.0, 0.0), pattern="at128", index=4,
                  scene=dict(half=200.0, height=40.0, boxes=90, size_lo=(8.0, 8.0, 6.0), size_hi=(40.0, 40.0, 30.0), clear=6.0)),
}


def quat_xyzw_to_R(q):
    x, y, z, w = q
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat_xyzw(R):
    """Batch rotation matrices (...,3,3) -> quaternions (...,4) xyzw, w >= 0."""
    R = np.asarray(R)
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    w = np.sqrt(np.maximum(0.0, 1 + m00 + m11 + m22)) / 2
    x = np.sqrt(np.maximum(0.0, 1 + m00 - m11 - m22)) / 2
    y = np.sqrt(np.maximum(0.0, 1 - m00 + m11 - m22)) / 2
    z = np.sqrt(np.maximum(0.0, 1 - m00 - m11 + m22)) / 2
    x = np.copysign(x, R[..., 2, 1] - R[..., 1, 2])
    y = np.copysign(y, R[..., 0, 2] - R[..., 2, 0])
    z = np.copysign(z, R[..., 1, 0] - R[..., 0, 1])
    q = np.stack([x, y, z, w], -1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


class Scene:
    """Room (inside-out box) plus solid boxes; all axis-aligned."""

    def __init__(self, rng, ring_radius, half=30.0, height=10.0, boxes=14, size_lo=(1.5, 1.5, 1.0), size_hi=(6.0, 6.0, 6.0),
                 clear=2.5):
        n_boxes = boxes
        self.room_lo = np.array([-half, -half, 0.0])
        self.room_hi = np.array([half, half, height])
        boxes = []
        tries = 0
        while len(boxes) < n_boxes and tries < 1000:
            tries += 1
            c = rng.uniform([-(half - 4), -(half - 4), 0], [half - 4, half - 4, 0])
            s = rng.uniform(list(size_lo), list(size_hi))
            lo = np.array([c[0] - s[0] / 2, c[1] - s[1] / 2, 0.0])
            hi = np.array([c[0] + s[0] / 2, c[1] + s[1] / 2, s[2]])
            # keep the driving ring clear
            corners = np.array([[lo[0], lo[1]], [lo[0], hi[1]], [hi[0], lo[1]], [hi[0], hi[1]], c[:2]])
            r = np.linalg.norm(corners, axis=1)
            near = np.abs(r - ring_radius).min() < clear or (r.min() < ring_radius < r.max())
            if not near:
                boxes.append((lo, hi))
        self.boxes = boxes

    def cast(self, o, d):
        """o, d: (n,3). Returns range (n,) to the first surface."""
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            t1 = (self.room_lo - o) * inv
            t2 = (self.room_hi - o) * inv
            rng_room = np.nanmin(np.maximum(t1, t2), axis=1)
            best = rng_room
            for lo, hi in self.boxes:
                a = (lo - o) * inv
                b = (hi - o) * inv
                tn = np.nanmax(np.minimum(a, b), axis=1)
                tf = np.nanmin(np.maximum(a, b), axis=1)
                hit = (tn <= tf) & (tn > 0)
                best = np.where(hit & (tn < best), tn, best)
        return best


class Sequence:
    def __init__(self, name="robot", seed=None, n_points=None, ins_entries=1000):
        cfg = dict(CONFIGS[name])
        self.name = name
        self.cfg = cfg
        self.n = int(n_points or cfg["n"])
        self.seed = (0xFF11D5 + cfg["index"]) if seed is None else seed
        self.filter_num = cfg["filter_num"]
        self.imu_hz = cfg["imu_hz"]
        self.radius = cfg["radius"]
        self.period = cfg["period"]          # frames per lap
        self.td = cfg["td"]
        self.W = ins_entries
        self.omega_z = 2 * math.pi / (self.period * FRAME_DT)
        self.speed = self.omega_z * self.radius
        self.R_bl = quat_xyzw_to_R(cfg["q_b_l"])
        self.t_bl = np.array(cfg["t_b_l"], np.float64)
        self.scene = Scene(np.random.default_rng(self.seed), self.radius, **cfg.get("scene", {}))
        self._dirs_cache = {}

    # ---- analytic trajectory of the IMU body frame ------------------------------------
    def body_pose(self, t):
        """t: array of times relative to GPS_T0 -> (p (n,3), R (n,3,3))."""
        t = np.atleast_1d(np.asarray(t, np.float64))
        a = self.omega_z * t
        p = np.stack([self.radius * np.cos(a), self.radius * np.sin(a), 1.5 + 0.05 * np.sin(3 * a)], -1)
        yaw = a + math.pi / 2
        pitch = 0.03 * np.sin(5 * a)
        roll = 0.04 * np.sin(4 * a + 0.3)
        cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
        R = np.empty((len(t), 3, 3))
        R[:, 0, 0] = cy * cp
        R[:, 0, 1] = cy * sp * sr - sy * cr
        R[:, 0, 2] = cy * sp * cr + sy * sr
        R[:, 1, 0] = sy * cp
        R[:, 1, 1] = sy * sp * sr + cy * cr
        R[:, 1, 2] = sy * sp * cr - cy * sr
        R[:, 2, 0] = -sp
        R[:, 2, 1] = cp * sr
        R[:, 2, 2] = cp * cr
        return p, R

    def body_velocity(self, t):
        h = 1e-4
        p1, R1 = self.body_pose(np.array([t + h]))
        p0, R0 = self.body_pose(np.array([t - h]))
        v = (p1[0] - p0[0]) / (2 * h)
        dR = R0[0].T @ R1[0]
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (4 * h)
        return v, w

    # ---- scan patterns: unit directions in the LiDAR frame + per-point time fraction ----
    def scan_dirs(self, frame):
        pat = self.cfg["pattern"]
        n = self.n
        key = frame if pat in ("rosette", "horizon") else 0
        if key in self._dirs_cache:
            return self._dirs_cache[key]
        j = np.arange(n, dtype=np.float64)
        if pat == "rosette":      # Livox Mid-70 class: circular 70 deg FoV, non-repetitive
            phi = 2 * math.pi * (j / n) * 40 + 0.7 * frame
            alpha = np.radians(35.0) * np.abs(np.cos(3.1 * phi + 0.13 * frame))
            beta = 1.7 * phi
            d = np.stack([np.cos(alpha), np.sin(alpha) * np.cos(beta), np.sin(alpha) * np.sin(beta)], -1)
            frac = j / n
        elif pat == "horizon":    # Livox Horizon: 81.7 x 25.1 deg, 6 lines
            line = (j % 6)
            s = (j // 6) / (n // 6)
            az = np.radians(81.7 / 2) * np.sin(2 * math.pi * (s * 25 + 0.37 * frame))
            el = np.radians(25.1 / 2) * np.sin(2 * math.pi * (s * 61 + 0.11 * frame)) + np.radians(0.28) * (line - 2.5)
            d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
            frac = j / n
        elif pat == "spin64":     # Ouster OS1-64: 64 rings x 1024 columns, column-ordered times
            rings, cols = 64, n // 64
            col = j // rings
            ring = j % rings
            az = 2 * math.pi * col / cols
            el = np.radians(-22.5 + 45.0 * ring / (rings - 1))
            d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
            frac = col / cols
        elif pat == "at128":      # Hesai AT128: 128 rings x 1200 columns over 120 deg
            rings = 128
            cols = n // rings
            col = j // rings
            ring = j % rings
            az = np.radians(-60.0 + 120.0 * col / max(cols - 1, 1))
            el = np.radians(-12.5 + 25.4 * ring / (rings - 1))
            d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
            frac = col / max(cols, 1)
        else:
            raise ValueError(pat)
        self._dirs_cache[key] = (d, frac)
        return d, frac

    # ---- one frame ----------------------------------------------------------------------
    def frame(self, i):
        """Returns dict(xyzi f32 (n,4), time f64 (n,), start, end, stamp, td, ins_t, ins_p, ins_q, v, omega)."""
        rng = np.random.default_rng(self.seed * 1000003 + i)
        d, frac = self.scan_dirs(i % self.period)
        n = self.n
        t_start = i * FRAME_DT
        t_l = t_start + frac * FRAME_DT * (1 - 1.0 / 4096)     # lidar-clock time of each point, relative
        p_b, R_b = self.body_pose(t_l + self.td)               # IMU-clock time t_b = t_l + td
        R_l = R_b @ self.R_bl
        o = p_b + R_b @ self.t_bl
        dw = np.einsum("nij,nj->ni", R_l, d)
        rng_true = self.scene.cast(o, dw)
        rng_meas = rng_true + rng.normal(0, 0.01, n)
        clutter = rng.random(n) < 0.02
        rng_meas = np.where(clutter, rng.uniform(1.0, np.maximum(rng_true, 1.5)), rng_meas)
        rng_meas = np.clip(rng_meas, 1.0, 250.0)
        pts = d * rng_meas[:, None]
        xyzi = np.empty((n, 4), np.float32)
        xyzi[:, :3] = pts
        xyzi[:, 3] = rng.uniform(0, 255, n)
        time = GPS_T0 + t_l
        end = float(time[-1])
        stamp = end + self.td
        # INS window: W entries at the IMU rate ending a few epochs after the frame end
        dt = 1.0 / self.imu_hz
        k_last = math.ceil((stamp - GPS_T0) / dt) + 5
        ks = np.arange(k_last - self.W + 1, k_last + 1)
        t_ins = ks * dt
        p_ins, R_ins = self.body_pose(t_ins)
        irng = np.random.default_rng(self.seed * 7919 + 17)       # noise tied to the epoch, not the frame
        noise_p = irng.normal(0, 1e-4, (4096, 3))
        noise_r = irng.normal(0, 1e-4, (4096, 3))
        sel = np.mod(ks, 4096)
        p_ins = p_ins + noise_p[sel]
        q_ins = R_to_quat_xyzw(R_ins)
        q_ins[:, :3] += 0.5 * noise_r[sel]
        q_ins /= np.linalg.norm(q_ins, axis=1, keepdims=True)
        v, w = self.body_velocity(stamp - GPS_T0)
        return dict(xyzi=xyzi, time=time, start=float(time[0]), end=end, stamp=stamp, td=self.td,
                    ins_t=GPS_T0 + t_ins, ins_p=np.ascontiguousarray(p_ins), ins_q=np.ascontiguousarray(q_ins),
                    v=v, omega=w)

    # ---- raw IMU increments of the analytic trajectory (what the INS mechanization consumes) -----------------------
    def body_accel(self, t):
        """Second derivative of body_pose's position (analytic)."""
        t = np.atleast_1d(np.asarray(t, np.float64))
        a, w = self.omega_z * t, self.omega_z
        return np.stack([-self.radius * w * w * np.cos(a), -self.radius * w * w * np.sin(a),
                         -0.05 * 9 * w * w * np.sin(3 * a)], -1)

    def body_rate(self, t, h=1e-5):
        """Body-frame angular rate: vee(log(R(t-h)^T R(t+h))) / 2h."""
        t = np.atleast_1d(np.asarray(t, np.float64))
        _, R0 = self.body_pose(t - h)
        _, R1 = self.body_pose(t + h)
        dR = np.einsum("nji,njk->nik", R0, R1)
        v = np.stack([dR[:, 2, 1] - dR[:, 1, 2], dR[:, 0, 2] - dR[:, 2, 0], dR[:, 1, 0] - dR[:, 0, 1]], -1) / 2
        s = np.linalg.norm(v, axis=1, keepdims=True)
        ang = np.arcsin(np.clip(s, 0, 1))
        return v * np.where(s > 1e-12, ang / np.maximum(s, 1e-300), 1.0) / (2 * h)

    def imu_samples(self, k0, k1, gravity=-9.80, bg=(0, 0, 0), ba=(0, 0, 0), rng=None, arw=0.0, vrw=0.0):
        """IMU epochs k0..k1 (inclusive) at the IMU rate, absolute times GPS_T0 + k dt: (time, dt, dtheta[3], dvel[3]) with
        dtheta = integral of the body rate and dvel = integral of the specific force over (t - dt, t] (4-point Gauss-
        Legendre), plus optional constant biases and white noise (angle / velocity random walk, per sqrt(s))."""
        dt = 1.0 / self.imu_hz
        ks = np.arange(k0, k1 + 1)
        t1 = ks * dt
        gx = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
        gw = np.array([0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538])
        g = np.array([0.0, 0.0, gravity])
        dth = np.zeros((len(ks), 3))
        dv = np.zeros((len(ks), 3))
        for x, w in zip(gx, gw):
            tau = t1 - dt / 2 + x * dt / 2
            _, R = self.body_pose(tau)
            f = np.einsum("nji,nj->ni", R, self.body_accel(tau) - g)
            dv += w * dt / 2 * f
            dth += w * dt / 2 * self.body_rate(tau)
        dth += np.asarray(bg) * dt
        dv += np.asarray(ba) * dt
        if rng is not None:
            dth += rng.normal(0, arw * np.sqrt(dt), dth.shape)
            dv += rng.normal(0, vrw * np.sqrt(dt), dv.shape)
        return GPS_T0 + t1, np.full(len(ks), dt), dth, dv

    def body_state(self, t):
        """(p, q xyzw, v) of the body at relative time t (truth)."""
        p, R = self.body_pose(np.array([t]))
        h = 1e-5
        p1, _ = self.body_pose(np.array([t + h]))
        p0, _ = self.body_pose(np.array([t - h]))
        return p[0], R_to_quat_xyzw(R[0]), (p1[0] - p0[0]) / (2 * h)

    def imu_pose7(self, stamp, rng=None, sigma_p=0.0, sigma_r=0.0):
        """IMU pose block [t, q xyzw] at absolute time `stamp`, optionally perturbed."""
        p, R = self.body_pose(np.array([stamp - GPS_T0]))
        q = R_to_quat_xyzw(R[0])
        p = p[0].copy()
        if rng is not None:
            p += rng.normal(0, sigma_p, 3)
            q[:3] += 0.5 * rng.normal(0, sigma_r, 3)
            q /= np.linalg.norm(q)
        return np.concatenate([p, q])

    def ext7(self):
        return np.concatenate([self.t_bl, np.array(self.cfg["q_b_l"], np.float64)])


def pack_aos(xyzi, time):
    """Reference raw layout: 32-byte PointXYZIT records (lidar/lidar.h:30-38)."""
    n = len(xyzi)
    rec = np.zeros(n, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad", "<f4"),
                                       ("intensity", "<f4"), ("pad2", "<f4"), ("time", "<f8")]))
    rec["x"], rec["y"], rec["z"], rec["intensity"] = xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]
    rec["pad"] = 1.0
    rec["time"] = time
    return rec
