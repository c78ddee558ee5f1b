"""Synthetic bundle-adjustment problems shared by bench.py, __graft_entry__.smoke() and the tests (tests/ba_scene.py
re-exports this module and supplies the projector for the general camera models).

Follows the reference's BA fixture conventions (src/optim/bundle_adjustment_test.cc:123-184:
SIMPLE_RADIAL f = 1.2 * width, cx = cy = width / 2, k = 0, observation noise U(-2, 2) px) and
SURVEY 8d's C4 shape: cameras on a ring looking at the scene, every point seen by
`track_len` cameras chosen by locality, perturbed initial poses / points; gauge fixed as the
reference's callers do (pose of image 0 constant, tvec.x of image 1 constant:
distributed_mapper_controller.cpp:883-884)."""
import numpy as np


def quat_from_R(R):
    """Robust rotation matrix -> (w, x, y, z): branch on the largest diagonal term."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = 2 * np.sqrt(1 + tr)
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = 2 * np.sqrt(1 + R[0, 0] - R[1, 1] - R[2, 2])
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = 2 * np.sqrt(1 + R[1, 1] - R[0, 0] - R[2, 2])
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = 2 * np.sqrt(1 + R[2, 2] - R[0, 0] - R[1, 1])
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return np.array(q)


def R_from_quat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def make_ba_problem(n_img=20, n_pts=400, track_len=6, seed=0, noise_px=2.0, pose_noise=(0.005, 0.02),
                    pt_noise=0.02, shared_camera=False, n_const_pts=0, width=1000, camera=None, world_to_image=None):
    """camera: None = the fixture's SIMPLE_RADIAL (cam_params [n_cam, 4]); or (model_id, params) of any model of
    camera_models.h -- observations are then projected with that model (the oracle's WorldToImage) and cam_params is
    [n_cam, 12] (the C ABI's camera_params_stride = 12), extra parameters starting 10 % off.  `world_to_image(model,
    params, uv) -> xy` projects normalised points for that case (the tests pass the checker's WorldToImage)."""
    rng = np.random.default_rng(seed)
    f = 1.2 * width
    ang = np.linspace(0, 2 * np.pi, n_img, endpoint=False)
    radius = 10.0
    q_true, t_true = np.zeros((n_img, 4)), np.zeros((n_img, 3))
    for i, a in enumerate(ang):
        C = np.array([radius * np.sin(a), 0.3 * np.sin(3 * a), -radius * np.cos(a)])
        zc = -C / np.linalg.norm(C)
        xc = np.cross([0, 1, 0], zc); xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc])          # world -> camera
        q_true[i] = quat_from_R(R)
        t_true[i] = -R @ C
    # points: each belongs to a ring position and is seen by the track_len nearest cameras
    # (a contiguous window of the ring, chosen by locality) -- vectorised
    pa = rng.uniform(0, 2 * np.pi, n_pts)
    X = rng.uniform(-1, 1, (n_pts, 3)) * [2.0, 1.5, 2.0]
    L = min(track_len, n_img)
    first = np.floor(pa / (2 * np.pi) * n_img - (L - 1) / 2.0 + 0.5).astype(np.int64)
    cams = np.sort((first[:, None] + np.arange(L)[None, :]) % n_img, axis=1)       # [n_pts, L]
    obs_img = cams.reshape(-1).astype(np.int32)
    obs_pt = np.repeat(np.arange(n_pts, dtype=np.int32), L)
    R_all = np.stack([R_from_quat(q) for q in q_true])
    pc = np.einsum("nij,nj->ni", R_all[obs_img], X[obs_pt]) + t_true[obs_img]
    if camera is None:
        obs_xy = f * pc[:, :2] / pc[:, 2:] + width / 2
    else:
        assert world_to_image is not None, "a projector is needed for camera models other than SIMPLE_RADIAL"
        obs_xy = world_to_image(camera[0], camera[1], pc[:, :2] / pc[:, 2:])
    obs_xy = obs_xy + rng.uniform(-noise_px, noise_px, (len(obs_img), 2))
    n_cam = 1 if shared_camera else n_img
    prob = {
        "qvec": q_true.copy(), "tvec": t_true.copy(),
        "img_cam": (np.zeros(n_img) if shared_camera else np.arange(n_img)).astype(np.int32),
        "pose_const": np.zeros(n_img, np.uint8), "tvec_const": np.zeros(n_img, np.uint8),
        "cam_model": np.full(n_cam, 2, np.int32),
        "cam_params": np.tile(np.array([f, width / 2, width / 2, 0.0]), (n_cam, 1)),
        "cam_const": np.zeros(n_cam, np.uint8),
        "xyz": X + rng.normal(0, pt_noise, X.shape),
        "pt_const": np.zeros(n_pts, np.uint8),
        "obs_img": np.ascontiguousarray(obs_img), "obs_pt": np.ascontiguousarray(obs_pt),
        "obs_xy": np.ascontiguousarray(obs_xy),
        "refine": (1, 0, 1),
    }
    if camera is not None:
        kp = np.zeros(12)
        kp[:len(camera[1])] = camera[1]
        n_lin = 4 if camera[0] in (1, 4, 5, 6, 7, 10) else 3
        kp[n_lin:] *= 0.9
        prob["cam_model"][:] = camera[0]
        prob["cam_params"] = np.tile(kp, (n_cam, 1))
    # perturb poses (not the gauge image)
    for i in range(1, n_img):
        dq = np.r_[1.0, rng.normal(0, pose_noise[0], 3)]
        q = prob["qvec"][i]
        w0, x0, y0, z0 = dq / np.linalg.norm(dq)
        w1, x1, y1, z1 = q
        prob["qvec"][i] = [w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                           w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1]
        prob["tvec"][i] += rng.normal(0, pose_noise[1], 3)
    prob["pose_const"][0] = 1          # SetConstantPose(image 0)
    prob["tvec_const"][1] = 1          # SetConstantTvec(image 1, {0})
    if n_const_pts:
        prob["pt_const"][rng.choice(n_pts, n_const_pts, replace=False)] = 1
    return prob


def copy_problem(prob):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in prob.items()}


def _project(prob, world_to_image=None):
    """WorldToImage of every observation with the problem's current parameters (numpy; general camera models through
    the `world_to_image(model, params, uv)` callable, camera by camera)."""
    q = prob["qvec"] / np.linalg.norm(prob["qvec"], axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], 1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], 1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    i, p = prob["obs_img"], prob["obs_pt"]
    pc = np.einsum("nij,nj->ni", R[i], prob["xyz"][p]) + prob["tvec"][i]
    uv = pc[:, :2] / pc[:, 2:]
    cam = prob["img_cam"][i]
    model = prob["cam_model"]
    if (model == 2).all():
        k = prob["cam_params"][cam]
        r2 = (uv ** 2).sum(1, keepdims=True)
        return k[:, :1] * uv * (1 + k[:, 3:4] * r2) + k[:, 1:3]
    assert world_to_image is not None, "a projector is needed for camera models other than SIMPLE_RADIAL"
    xy = np.zeros_like(uv)
    for c in range(len(model)):
        sel = cam == c
        if sel.any():
            xy[sel] = world_to_image(int(model[c]), list(prob["cam_params"][c]), uv[sel])
    return xy


def reprojection_rms(prob, world_to_image=None):
    """sqrt(sum ||r||^2 / N_obs) in px, evaluated independently of the solvers."""
    xy = _project(prob, world_to_image)
    return float(np.sqrt(((xy - prob["obs_xy"]) ** 2).sum() / len(xy)))


def mean_reprojection_error(prob, world_to_image=None):
    """Mean ||r|| in px: Reconstruction::ComputeMeanReprojectionError (base/reconstruction.cc:814-858)."""
    xy = _project(prob, world_to_image)
    return float(np.sqrt(((xy - prob["obs_xy"]) ** 2).sum(1)).mean())
