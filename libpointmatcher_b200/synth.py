"""Deterministic synthetic "Velodyne-like" clouds for the benchmark and the parity tests.

SURVEY.md §8(d): a spinning multi-ring LiDAR ray-cast into a box-shaped scene with a ground
plane, walls, axis-aligned boxes and vertical cylinders; range noise N(0, 2 cm).  The
*reading* is a re-scan of the same scene from a displaced pose (not a transformed copy of
the reference), so there are no exact-zero distances.  Clouds are (N, 4) float32 with w = 1:
the memory layout of the reference's 4 x N column-major `DataPoints::features`
(PointMatcher.h:169,331; IO.cpp:999-1008).  Points are emitted in firing order
(azimuth-major, ring-minor) like a real sensor.
"""
import hashlib
import os

import numpy as np

SEED = 20261018
GROUND_Z = -1.8
HALF_X, HALF_Y = 40.0, 30.0
MAX_RANGE = 120.0
ELEV_MIN_DEG, ELEV_MAX_DEG = -24.8, 2.0
RANGE_SIGMA = 0.02
_CACHE_DIR = os.environ.get("PMB200_SYNTH_CACHE", "/tmp/pmb200_synth_cache")


def pose_matrix(t=(0.0, 0.0, 0.0), yaw_deg=0.0, pitch_deg=0.0, roll_deg=0.0):
    """4x4 float64 pose: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    y, p, r = np.deg2rad([yaw_deg, pitch_deg, roll_deg])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


# pose of the reading scan relative to the reference scan
READING_POSE = pose_matrix((0.60, -0.35, 0.05), 4.0, 0.5, -0.3)


class Scene:
    """Ground plane + 4 walls + 40 boxes + 30 cylinders, fixed by the seed."""

    def __init__(self, seed=SEED):
        rng = np.random.default_rng(seed)
        boxes = []
        while len(boxes) < 40:
            size = rng.uniform(1.0, 8.0, 3)
            size[2] = rng.uniform(1.0, 5.0)
            c = np.array([rng.uniform(-HALF_X + 5, HALF_X - 5), rng.uniform(-HALF_Y + 5, HALF_Y - 5)])
            # keep the sensor track (around the origin, along +x) free
            if abs(c[1]) < 4.0 + size[1] / 2 and -8.0 - size[0] / 2 < c[0] < 30.0 + size[0] / 2:
                continue
            lo = np.array([c[0] - size[0] / 2, c[1] - size[1] / 2, GROUND_Z])
            hi = np.array([c[0] + size[0] / 2, c[1] + size[1] / 2, GROUND_Z + size[2]])
            boxes.append((lo, hi))
        cyls = []
        while len(cyls) < 30:
            r = rng.uniform(0.15, 0.6)
            c = np.array([rng.uniform(-HALF_X + 2, HALF_X - 2), rng.uniform(-HALF_Y + 2, HALF_Y - 2)])
            if abs(c[1]) < 3.0 and -8.0 < c[0] < 30.0:
                continue
            cyls.append((c, r, GROUND_Z + rng.uniform(2.0, 8.0)))
        self.boxes, self.cyls = boxes, cyls

    def raycast(self, origin, dirs):
        """Nearest hit distance along unit directions `dirs` (M, 3) from `origin` (3,)."""
        o = origin.astype(np.float64)
        d = dirs
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            t = np.full(d.shape[0], MAX_RANGE)
            # ground
            tg = (GROUND_Z - o[2]) * inv[:, 2]
            t = np.where((tg > 0) & (tg < t), tg, t)
            # walls (infinite height, inner faces of the room)
            for axis, val in ((0, HALF_X), (0, -HALF_X), (1, HALF_Y), (1, -HALF_Y)):
                tw = (val - o[axis]) * inv[:, axis]
                t = np.where((tw > 0) & (tw < t), tw, t)
            # boxes: slab test
            for lo, hi in self.boxes:
                t0 = (lo - o) * inv
                t1 = (hi - o) * inv
                tn = np.minimum(t0, t1).max(axis=1)
                tf = np.maximum(t0, t1).min(axis=1)
                hit = (tn <= tf) & (tn > 0) & (tn < t)
                t = np.where(hit, tn, t)
            # vertical cylinders (side surface only, finite height)
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            for c, r, ztop in self.cyls:
                ox, oy = o[0] - c[0], o[1] - c[1]
                b = ox * d[:, 0] + oy * d[:, 1]
                cc = ox * ox + oy * oy - r * r
                disc = b * b - a * cc
                sq = np.sqrt(np.maximum(disc, 0.0))
                tc = (-b - sq) / a
                z = o[2] + tc * d[:, 2]
                hit = (disc > 0) & (tc > 0) & (tc < t) & (z <= ztop) & (z >= GROUND_Z)
                t = np.where(hit, tc, t)
        return t


def _threads():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    # several ranks of one node generate their clouds at the same time
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, min(16, n // max(1, world)))


def _scan_uncached(n_points, pose, rings, seed, scene_seed):
    scene = Scene(scene_seed)
    az_steps = -(-n_points // rings)
    el = np.deg2rad(np.linspace(ELEV_MIN_DEG, ELEV_MAX_DEG, rings))
    rng = np.random.default_rng(seed)
    out = np.empty((az_steps * rings, 4), np.float32)
    R, origin = pose[:3, :3], pose[:3, 3]
    chunk = max(1, (1 << 18) // rings)
    starts = list(range(0, az_steps, chunk))
    # the range noise is one sequential stream (drawn chunk after chunk, as always); the ray casting of the chunks is
    # independent and runs on a thread pool (numpy releases the GIL)
    noise = [rng.normal(0.0, RANGE_SIGMA, (min(az_steps, a0 + chunk) - a0) * rings) for a0 in starts]

    def cast(i):
        a0 = starts[i]
        a1 = min(az_steps, a0 + chunk)
        az = 2 * np.pi * (np.arange(a0, a1) + 0.5) / az_steps
        azg, elg = np.meshgrid(az, el, indexing="ij")  # firing order: azimuth-major
        d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=-1).reshape(-1, 3)
        t = scene.raycast(origin, d @ R.T)
        t = t + noise[i]
        sl = slice(a0 * rings, a1 * rings)
        out[sl, :3] = (d * t[:, None]).astype(np.float32)
        out[sl, 3] = 1.0

    nthreads = _threads()
    if nthreads > 1 and len(starts) > 1:
        import concurrent.futures
        with concurrent.futures.ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(cast, range(len(starts))))
    else:
        for i in range(len(starts)):
            cast(i)
    return out[:n_points]


def scan(n_points, pose=None, rings=None, seed=SEED, scene_seed=SEED, cache=True):
    """One LiDAR scan of `n_points` points taken from `pose`, in the sensor frame."""
    pose = np.eye(4) if pose is None else np.asarray(pose, np.float64)
    if rings is None:
        # isotropic angular sampling: ring spacing ~ azimuth spacing over the 26.8 deg x 360 deg
        # field of view (64 rings x 15 625 azimuth steps would put 1 M points on 64 thin circles,
        # where 2 cm range noise >> 2.5 mm point spacing makes local surface normals meaningless)
        rings = 64
        while rings * 2 <= np.sqrt(n_points * (ELEV_MAX_DEG - ELEV_MIN_DEG) / 360.0) * 1.5:
            rings *= 2
    key = hashlib.sha1(repr((n_points, pose.round(9).tolist(), rings, seed, scene_seed, 2)).encode()).hexdigest()[:20]
    path = os.path.join(_CACHE_DIR, key + ".npy")
    if cache and n_points >= 100_000 and os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    out = _scan_uncached(n_points, pose, rings, seed, scene_seed)
    if cache and n_points >= 100_000:
        try:
            os.makedirs(_CACHE_DIR, exist_ok=True)
            tmp = path + ".%d.tmp.npy" % os.getpid()
            np.save(tmp, out)
            os.replace(tmp, path)
        except OSError:
            pass
    return out


def scan_pair(n_reading, n_reference=None, pair_seed=0, rings=None):
    """(reading, reference, T_gt): `reading` is a re-scan from READING_POSE (pair_seed = 0) or
    from a random pose with |t| <= 1 m and |angle| <= 6 deg (pair_seed > 0), `T_gt` (float64 4x4)
    maps reading coordinates into the reference frame.  rings: None = isotropic angular sampling (see scan), 64 = the
    64 x (N / 64) layout SURVEY 8d names."""
    n_reference = n_reading if n_reference is None else n_reference
    if pair_seed == 0:
        pose = READING_POSE
    else:
        rng = np.random.default_rng(SEED + pair_seed)
        t = rng.normal(size=3)
        t *= rng.uniform(0.2, 1.0) / np.linalg.norm(t)
        t[2] *= 0.1
        ang = rng.uniform(-6.0, 6.0, 3) * np.array([1.0, 0.15, 0.15])
        pose = pose_matrix(t, *ang)
    reference = scan(n_reference, np.eye(4), rings=rings, seed=SEED + 2 * pair_seed)
    reading = scan(n_reading, pose, rings=rings, seed=SEED + 2 * pair_seed + 1)
    return reading, reference, pose


def world_map(n_points, n_scans=10, spacing=2.0):
    """Concatenation of `n_scans` scans taken `spacing` metres apart along +x, in the world frame."""
    per = n_points // n_scans
    parts = []
    for s in range(n_scans):
        pose = pose_matrix((s * spacing, 0.0, 0.0))
        pts = scan(per, pose, seed=SEED + 100 + s).astype(np.float64)
        w = pts @ pose.T
        w[:, 3] = 1.0
        parts.append(w.astype(np.float32))
    return np.ascontiguousarray(np.concatenate(parts, axis=0))
