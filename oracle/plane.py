"""TEST INFRASTRUCTURE -- numpy restatement of the reference's plane modelling branch.

PointCloudSegment.cluster_modeling(model_method='plane') (utils/segment_utils.py:188-216) and
plane_angle_validation (:84-93), with a STAND-IN for open3d's PointCloud.segment_plane, which is a
third-party dependency absent from /root/reference and from this image (SURVEY Appendix G restates its
published algorithm from memory; version unpinned by the reference => PARITY UNPINNED).  The stand-in is
seeded, so the "reference" numbers of config 3 (bitstream size, error bound) are reproducible."""
import numpy as np


def get_plane_from_points(pts):
    """open3d GetPlaneFromPoints: centroid + second moments, normal along the axis with the largest 2x2
    determinant, normalised; None if degenerate."""
    pts = np.asarray(pts, np.float64)
    if pts.shape[0] < 3:
        return None
    c = pts.mean(0)
    r = pts - c
    xx, xy, xz = (r[:, 0] * r[:, 0]).sum(), (r[:, 0] * r[:, 1]).sum(), (r[:, 0] * r[:, 2]).sum()
    yy, yz, zz = (r[:, 1] * r[:, 1]).sum(), (r[:, 1] * r[:, 2]).sum(), (r[:, 2] * r[:, 2]).sum()
    dx, dy, dz = yy * zz - yz * yz, xx * zz - xz * xz, xx * yy - xy * xy
    dmax = max(dx, dy, dz)
    if not dmax > 0:
        return None
    if dmax == dx:
        n = np.array([dx, xz * yz - xy * zz, xy * yz - xz * yy])
    elif dmax == dy:
        n = np.array([xz * yz - xy * zz, dy, xy * xz - yz * xx])
    else:
        n = np.array([xy * yz - xz * yy, xy * xz - yz * xx, dz])
    nn = np.linalg.norm(n)
    if not nn > 0:
        return None
    n = n / nn
    return np.array([n[0], n[1], n[2], -float(n @ c)])


def segment_plane(points, distance_threshold, ransac_n, num_iterations, rng):
    """Stand-in for open3d.geometry.PointCloud.segment_plane -> (plane_model f64[4], inlier indices)."""
    pts = np.asarray(points, np.float64)
    n = pts.shape[0]
    best, best_fit, best_rmse = np.zeros(4), 0.0, np.inf
    for _ in range(num_iterations):
        idx = rng.choice(n, ransac_n, replace=False)
        pl = get_plane_from_points(pts[idx])
        if pl is None:
            continue
        d = np.abs(pts @ pl[:3] + pl[3])
        inl = d < distance_threshold
        cnt = int(inl.sum())
        if cnt == 0:
            continue
        fit, rmse = cnt / n, float(np.sqrt((d[inl] ** 2).sum() / cnt))
        if fit > best_fit or (fit == best_fit and rmse < best_rmse):
            best, best_fit, best_rmse = pl, fit, rmse
    inl = np.where(np.abs(pts @ best[:3] + best[3]) < distance_threshold)[0]
    refit = get_plane_from_points(pts[inl]) if inl.size >= 3 else None
    return (refit if refit is not None else best), inl


def plane_angle_validation(lut, plane_model, idx, angle_threshold):
    """utils/segment_utils.py:84-93, expression for expression (including the precedence of / and *)."""
    scan_vector = lut[idx]
    with np.errstate(invalid="ignore", divide="ignore"):
        alpha = np.arccos(np.abs(np.sum(np.expand_dims(plane_model[:3], 0) * scan_vector, -1)) /
                          np.linalg.norm(plane_model[:3]) * np.linalg.norm(scan_vector, ord=2, axis=-1))
    return not (alpha.max() > np.pi * (angle_threshold / 180))


def cluster_modeling_plane(lut, range_image, seg_idx, angle_threshold=75, seed=0):
    """utils/segment_utils.py:188-216 -> cluster_models (K-1, 4) f64 (rows of labels 1 .. K-1)."""
    rng = np.random.default_rng(seed)
    H, W = lut.shape[:2]
    ri = np.asarray(range_image, np.float32).reshape(H, W, 1)
    seg = np.asarray(seg_idx).reshape(H, W)
    pc = ri * lut                                   # dataset/transformer.py:94-98
    models = []
    for i in range(int(seg.max()) + 1):
        if i == 0:
            continue
        if i == 1:
            models.append([0, 0, 0, 0.0])
            continue
        idx = np.where(seg == i)
        cur_range = ri[idx]
        if idx[0].shape[0] < 30:
            models.append([0, 0, 0, cur_range.mean()])
            continue
        plane_model, _ = segment_plane(pc[idx], 0.1, 4, 10, rng)
        if plane_angle_validation(lut, plane_model, idx, angle_threshold):
            models.append(list(plane_model))
        else:
            models.append([0, 0, 0, cur_range.mean()])
    return np.asarray(models)
