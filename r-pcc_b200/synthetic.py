"""Synthetic KITTI-like LiDAR frames (SURVEY.md section 8(d) input spec) for tests and bench.py.

There is no dataset in the image, so the workload is generated: a tilted ground plane at about
z = -1.73 m, 30-60 random vertical walls between 3 and 70 m, per-point angular jitter (so that
about a quarter of the pixels collide, as on KITTI), a few beams above the configured maximum
elevation (row clamping), range noise, 2 % dropouts, 1.5 m <= r <= 80 m.  frame(seed=i) is
deterministic: rng = default_rng(1234 + i).  Output rows are (x, y, z, intensity) float32 in
beam-major order, like a KITTI .bin; `ground` is the true plane [a,b,c,d] (f64, unit normal),
which tests and the bench inject identically into the oracle and the device path.
"""
import numpy as np

from .lidar import LidarConfig

_AZ_STEPS = {"Velodyne64E": 2083, "Velodyne32E": 2170, "VelodyneVLP16": 1800, "Velodyne64E_unofficial": 2083}


def frame(seed, lidar="Velodyne64E"):
    """-> (points (N,4) f32, ground (4,) f64)"""
    cfg = lidar if isinstance(lidar, LidarConfig) else LidarConfig(lidar)
    rng = np.random.default_rng(1234 + int(seed))
    H = cfg.H
    nz = _AZ_STEPS.get(cfg.name, cfg.W)
    vmin, vmax = cfg.vertical_min, cfg.vertical_max
    elev = vmin + (vmax - vmin) * np.arange(H) / (H - 1)
    # a few top beams reach above the configured maximum (real 64E beams reach +4.96 deg)
    n_hi = max(1, H // 16)
    elev[-n_hi:] = vmax + np.radians(np.linspace(0.8, 3.0, n_hi))

    az = (np.arange(nz) + 0.0) * (2 * np.pi / nz)
    ca, sa = np.cos(az), np.sin(az)

    # ground: z = -h0 + tx*x + ty*y
    h0 = 1.73 + rng.normal(0, 0.03)
    tx, ty = rng.normal(0, 0.01, 2)
    # walls: segments P->Q with a top height
    M = int(rng.integers(30, 61))
    dist = rng.uniform(3.0, 70.0, M)
    ang = rng.uniform(0, 2 * np.pi, M)
    cx, cy = dist * np.cos(ang), dist * np.sin(ang)
    yaw = rng.uniform(0, np.pi, M)
    half = rng.uniform(0.8, 8.0, M)
    n_long = int(rng.integers(8, 15))               # building facades / hedges
    half[:n_long] = rng.uniform(10.0, 40.0, n_long)
    dist[:n_long] = rng.uniform(8.0, 60.0, n_long)
    cx, cy = dist * np.cos(ang), dist * np.sin(ang)
    px, py = cx - half * np.cos(yaw), cy - half * np.sin(yaw)
    qx, qy = cx + half * np.cos(yaw), cy + half * np.sin(yaw)
    top = -h0 + rng.uniform(1.2, 9.0, M)
    top[:n_long] = -h0 + rng.uniform(2.5, 15.0, n_long)
    # ray (ca,sa)*t = P + u*(Q-P)
    ex, ey = (qx - px)[None, :], (qy - py)[None, :]
    den = ca[:, None] * ey - sa[:, None] * ex
    den = np.where(np.abs(den) < 1e-9, 1e-9, den)
    t = (px[None, :] * ey - py[None, :] * ex) / den
    u = (px[None, :] * sa[:, None] - py[None, :] * ca[:, None]) / den
    hit = (t > 0.5) & (u >= 0) & (u <= 1)
    t = np.where(hit, t, np.inf)
    first = np.argmin(t, axis=1)
    d_wall = t[np.arange(nz), first]                # planar distance of nearest wall per azimuth
    z_top = np.where(np.isfinite(d_wall), top[first], -np.inf)

    te = np.tan(elev)[:, None]                      # (H,1)
    ce = np.cos(elev)[:, None]
    # ground hit: t*tan(e) = -h0 + t*(tx*ca + ty*sa)
    slope = (tx * ca + ty * sa)[None, :]
    dd = slope - te
    with np.errstate(divide="ignore", invalid="ignore"):
        d_g = np.where(dd > 1e-6, h0 / dd, np.inf)
    z_at_wall = d_wall[None, :] * te
    wall_hit = (d_wall[None, :] < d_g) & (z_at_wall <= z_top[None, :])
    d_plan = np.where(wall_hit, d_wall[None, :], d_g)
    r = d_plan / ce
    r = r + rng.normal(0, 0.015, r.shape)
    keep = np.isfinite(r) & (r >= 1.5) & (r <= 80.0) & (rng.random(r.shape) > 0.02)

    # per-point angular jitter (exercises rounding and pixel collisions)
    az_j = az[None, :] + rng.uniform(-0.45, 0.45, r.shape) * (2 * np.pi / nz)
    el_j = elev[:, None] + rng.normal(0, np.radians(0.03), r.shape)
    x = r * np.cos(el_j) * np.cos(az_j)
    y = r * np.cos(el_j) * np.sin(az_j)
    z = r * np.sin(el_j)
    inten = rng.random(r.shape)
    pts = np.stack([x[keep], y[keep], z[keep], inten[keep]], -1).astype(np.float32)

    n = np.array([-tx, -ty, 1.0])
    nn = np.linalg.norm(n)
    ground = np.array([n[0] / nn, n[1] / nn, n[2] / nn, h0 / nn])
    return pts, ground


def batch(seeds, lidar="Velodyne64E"):
    """-> (points (sum N,4) f32, offsets (B+1,) int64, grounds (B,4) f64)"""
    pts, gs, off = [], [], [0]
    for s in seeds:
        p, g = frame(s, lidar)
        pts.append(p)
        gs.append(g)
        off.append(off[-1] + p.shape[0])
    return np.concatenate(pts, 0), np.asarray(off, np.int64), np.stack(gs, 0)
