"""Synthetic two-view verification workload (SURVEY 8d, C3 flavour): many image pairs, each
with inlier matches that follow a random relative pose of a SIMPLE_RADIAL camera pair plus
uniformly random outlier matches; half the images carry a prior focal length so both the
calibrated (E+F+H) and the uncalibrated (F+H) paths run.  Vectorised numpy."""
import numpy as np


def make_pairs(n_pairs=1000, n_in=(120, 400), n_out=(40, 250), seed=0, noise=0.5, width=1000, planar_every=7):
    """Returns dict(cams [2*n_pairs] tuples, keypoints list of [n,2] arrays (2 images per pair),
    pairs [n_pairs,2], match_offsets [n_pairs+1], matches [total,2], prior [2*n_pairs] bool)."""
    rng = np.random.default_rng(seed)
    f, c = 1.2 * width, width / 2.0
    ni = rng.integers(n_in[0], n_in[1] + 1, n_pairs)
    no = rng.integers(n_out[0], n_out[1] + 1, n_pairs)
    m = ni + no
    off = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
    total = int(off[-1])
    pid = np.repeat(np.arange(n_pairs), m)
    local = np.arange(total) - off[pid]
    is_in = local < ni[pid]
    # random relative poses
    ang = rng.uniform(-0.3, 0.3, (n_pairs, 3))
    t = rng.normal(0, 1, (n_pairs, 3)); t /= np.linalg.norm(t, axis=1, keepdims=True); t *= rng.uniform(0.5, 1.5, (n_pairs, 1))
    cx, sx = np.cos(ang[:, 0]), np.sin(ang[:, 0]); cy, sy = np.cos(ang[:, 1]), np.sin(ang[:, 1]); cz, sz = np.cos(ang[:, 2]), np.sin(ang[:, 2])
    Rx = np.stack([np.ones_like(cx), 0 * cx, 0 * cx, 0 * cx, cx, -sx, 0 * cx, sx, cx], 1).reshape(-1, 3, 3)
    Ry = np.stack([cy, 0 * cy, sy, 0 * cy, np.ones_like(cy), 0 * cy, -sy, 0 * cy, cy], 1).reshape(-1, 3, 3)
    Rz = np.stack([cz, -sz, 0 * cz, sz, cz, 0 * cz, 0 * cz, 0 * cz, np.ones_like(cz)], 1).reshape(-1, 3, 3)
    R = Rz @ Ry @ Rx
    X = rng.uniform(-1, 1, (total, 3)) * [2.5, 2.5, 1.5] + [0, 0, 8]
    planar = (pid % planar_every) == (planar_every - 1)
    X[planar, 2] = 8 + 0.05 * X[planar, 0]
    x1 = f * X[:, :2] / X[:, 2:] + c + rng.normal(0, noise, (total, 2))
    Xc = np.einsum("nij,nj->ni", R[pid], X) + t[pid]
    x2 = f * Xc[:, :2] / Xc[:, 2:] + c + rng.normal(0, noise, (total, 2))
    o1 = rng.uniform(0, width, (total, 2)); o2 = rng.uniform(0, width, (total, 2))
    p1 = np.where(is_in[:, None], x1, o1)
    p2 = np.where(is_in[:, None], x2, o2)
    # image 2k holds p1 of pair k in match order; image 2k+1 holds p2 permuted
    keypoints, matches = [], np.zeros((total, 2), np.uint32)
    for k in range(n_pairs):
        a, b = off[k], off[k + 1]
        perm = rng.permutation(b - a)
        kp2 = np.empty((b - a, 2)); kp2[perm] = p2[a:b]
        keypoints += [np.ascontiguousarray(p1[a:b]), kp2]
        matches[a:b, 0] = np.arange(b - a)
        matches[a:b, 1] = perm
    prior = np.repeat(rng.random(n_pairs) < 0.5, 2)
    pairs = np.stack([2 * np.arange(n_pairs), 2 * np.arange(n_pairs) + 1], 1).astype(np.uint32)
    return {"keypoints": keypoints, "pairs": pairs, "match_offsets": off, "matches": matches, "prior": prior,
            "cam_params": (f, c, c, 0.0), "width": width, "n_inliers_true": ni}


def scene(rng, n_in, n_out, planar=False, noise=0.3, f=1200.0, ang=0.15, t=(-1.0, 0.1, 0.2)):
    c = 500.0
    X = rng.uniform(-1, 1, (n_in, 3)) * [2, 2, 1] + [0, 0, 8]
    if planar:
        X[:, 2] = 8 + 0.1 * X[:, 0]
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    x1 = f * X[:, :2] / X[:, 2:] + c + rng.normal(0, noise, (n_in, 2))
    Xc = X @ R.T + np.array(t)
    x2 = f * Xc[:, :2] / Xc[:, 2:] + c + rng.normal(0, noise, (n_in, 2))
    o1 = rng.uniform(0, 1000, (n_out, 2))
    o2 = rng.uniform(0, 1000, (n_out, 2))
    return np.r_[x1, o1], np.r_[x2, o2]
