"""Host-side mirror of the reference's bundle-adjustment set-up:

  BundleAdjustmentConfig               src/optim/bundle_adjustment.h:105-161, .cc:88-246
  BundleAdjuster::SetUp and helpers    src/optim/bundle_adjustment.cc:316-526

`pack_problem` turns (reconstruction, config, options) into the flat arrays of `b2_ba_problem`
exactly the way SetUp builds the Ceres problem: which observations become residuals, which images
enter with a constant pose (images outside the config that observe an explicitly added point),
which cameras and points are held constant.  `unpack_problem` writes the solution back.  The
`Reconstruction` here is the minimal slice of colmap's class that SetUp touches.
"""
from __future__ import annotations

import numpy as np

MODEL_ID = {"SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4, "OPENCV_FISHEYE": 5,
            "FULL_OPENCV": 6, "FOV": 7, "SIMPLE_RADIAL_FISHEYE": 8, "RADIAL_FISHEYE": 9, "THIN_PRISM_FISHEYE": 10}   # camera_models.h


class Reconstruction:
    """cameras[id] = {model, params}; images[id] = {camera_id, qvec, tvec, points2D = [[x, y, point3D_id | -1]]};
    points3D[id] = {xyz, track = [(image_id, point2D_idx)]}."""

    def __init__(self):
        self.cameras, self.images, self.points3D = {}, {}, {}

    def add_camera(self, camera_id, model, params):
        self.cameras[camera_id] = {"model": MODEL_ID.get(model, model), "params": np.array(params, dtype=np.float64)}

    def add_image(self, image_id, camera_id, qvec, tvec, points2D_xy):
        self.images[image_id] = {"camera_id": camera_id, "qvec": np.array(qvec, dtype=np.float64),
                                 "tvec": np.array(tvec, dtype=np.float64),
                                 "points2D": [[float(x), float(y), -1] for x, y in points2D_xy]}

    def add_point3D(self, point3D_id, xyz):
        self.points3D[point3D_id] = {"xyz": np.array(xyz, dtype=np.float64), "track": []}

    def add_observation(self, point3D_id, image_id, point2D_idx):      # Reconstruction::AddObservation
        self.images[image_id]["points2D"][point2D_idx][2] = point3D_id
        self.points3D[point3D_id]["track"].append((image_id, point2D_idx))

    def delete_observation(self, image_id, point2D_idx):               # Reconstruction::DeleteObservation
        pid = self.images[image_id]["points2D"][point2D_idx][2]
        self.images[image_id]["points2D"][point2D_idx][2] = -1
        self.points3D[pid]["track"].remove((image_id, point2D_idx))

    def copy(self) -> "Reconstruction":
        r = Reconstruction()
        r.cameras = {k: {"model": v["model"], "params": v["params"].copy()} for k, v in self.cameras.items()}
        r.images = {k: {"camera_id": v["camera_id"], "qvec": v["qvec"].copy(), "tvec": v["tvec"].copy(),
                        "points2D": [list(p) for p in v["points2D"]]} for k, v in self.images.items()}
        r.points3D = {k: {"xyz": v["xyz"].copy(), "track": list(v["track"])} for k, v in self.points3D.items()}
        return r


class BundleAdjustmentConfig:
    """bundle_adjustment.h:105-161.  Sets keep insertion order irrelevant; iteration is by sorted id."""

    def __init__(self):
        self._images, self._const_cameras, self._const_poses = set(), set(), set()
        self._const_tvecs, self._variable_points, self._constant_points = {}, set(), set()

    # images
    def AddImage(self, image_id): self._images.add(image_id)
    def HasImage(self, image_id): return image_id in self._images
    def RemoveImage(self, image_id): self._images.discard(image_id)
    def NumImages(self): return len(self._images)
    def Images(self): return sorted(self._images)
    # cameras
    def SetConstantCamera(self, camera_id): self._const_cameras.add(camera_id)
    def SetVariableCamera(self, camera_id): self._const_cameras.discard(camera_id)
    def IsConstantCamera(self, camera_id): return camera_id in self._const_cameras
    def NumConstantCameras(self): return len(self._const_cameras)
    # poses
    def SetConstantPose(self, image_id):
        assert self.HasImage(image_id) and not self.HasConstantTvec(image_id)   # .cc:150-154
        self._const_poses.add(image_id)
    def SetVariablePose(self, image_id): self._const_poses.discard(image_id)
    def HasConstantPose(self, image_id): return image_id in self._const_poses
    def NumConstantPoses(self): return len(self._const_poses)
    def SetConstantTvec(self, image_id, idxs):
        idxs = list(idxs)
        assert 0 < len(idxs) <= 3 and self.HasImage(image_id) and not self.HasConstantPose(image_id)  # .cc:164-172
        assert len(set(idxs)) == len(idxs), "Tvec indices must not contain duplicates"
        self._const_tvecs[image_id] = idxs
    def RemoveConstantTvec(self, image_id): self._const_tvecs.pop(image_id, None)
    def HasConstantTvec(self, image_id): return image_id in self._const_tvecs
    def ConstantTvec(self, image_id): return self._const_tvecs[image_id]
    def NumConstantTvecs(self): return len(self._const_tvecs)
    # points
    def AddVariablePoint(self, point3D_id):
        assert point3D_id not in self._constant_points          # .cc:197-200
        self._variable_points.add(point3D_id)
    def AddConstantPoint(self, point3D_id):
        assert point3D_id not in self._variable_points          # .cc:202-205
        self._constant_points.add(point3D_id)
    def HasPoint(self, p): return p in self._variable_points or p in self._constant_points
    def HasVariablePoint(self, p): return p in self._variable_points
    def HasConstantPoint(self, p): return p in self._constant_points
    def RemoveVariablePoint(self, p): self._variable_points.discard(p)
    def RemoveConstantPoint(self, p): self._constant_points.discard(p)
    def NumPoints(self): return len(self._variable_points) + len(self._constant_points)
    def NumVariablePoints(self): return len(self._variable_points)
    def NumConstantPoints(self): return len(self._constant_points)
    def VariablePoints(self): return sorted(self._variable_points)
    def ConstantPoints(self): return sorted(self._constant_points)

    def NumResiduals(self, recon: Reconstruction) -> int:      # .cc:207-246
        n = 0
        for image_id in self._images:
            n += sum(1 for p in recon.images[image_id]["points2D"] if p[2] >= 0)
        def outside(pid):
            return sum(1 for (i, _) in recon.points3D[pid]["track"] if i not in self._images)
        n += sum(outside(p) for p in self._variable_points) + sum(outside(p) for p in self._constant_points)
        return 2 * n


def pack_problem(recon: Reconstruction, config: BundleAdjustmentConfig, refine_extrinsics: bool = True):
    """BundleAdjuster::SetUp.  Returns (problem dict for BundleAdjuster.Solve / the oracle, maps)."""
    config_const_cameras = set(config._const_cameras)      # SetUp mutates the config (.cc:447-450); work on a copy
    obs = []                                               # (point3D_id, image_id, x, y)
    num_obs = {}                                           # point3D_num_observations_
    camera_ids, prob_images, const_pose = [], [], {}
    for image_id in config.Images():                       # AddImageToProblem, .cc:338-420
        im = recon.images[image_id]
        im["qvec"] = im["qvec"] / np.linalg.norm(im["qvec"])    # NormalizeQvec
        cp = (not refine_extrinsics) or config.HasConstantPose(image_id)
        n = 0
        for x, y, pid in im["points2D"]:
            if pid < 0:
                continue
            n += 1
            num_obs[pid] = num_obs.get(pid, 0) + 1
            obs.append((pid, image_id, x, y))
        prob_images.append(image_id)
        const_pose[image_id] = cp
        if n > 0 and im["camera_id"] not in camera_ids:
            camera_ids.append(im["camera_id"])
    for pid in config.VariablePoints() + config.ConstantPoints():     # AddPointToProblem, .cc:422-472
        track = recon.points3D[pid]["track"]
        if num_obs.get(pid, 0) == len(track):
            continue
        for image_id, idx in track:
            if config.HasImage(image_id):
                continue
            num_obs[pid] = num_obs.get(pid, 0) + 1
            im = recon.images[image_id]
            if im["camera_id"] not in camera_ids:
                camera_ids.append(im["camera_id"])
                config_const_cameras.add(im["camera_id"])
            if image_id not in const_pose:
                prob_images.append(image_id)
                const_pose[image_id] = True            # BundleAdjustmentConstantPoseCostFunction
            x, y, _ = im["points2D"][idx]
            obs.append((pid, image_id, x, y))
    point_ids = sorted(num_obs)
    img_idx = {i: k for k, i in enumerate(prob_images)}
    used_cams = sorted({recon.images[i]["camera_id"] for i in prob_images})
    cam_idx = {c: k for k, c in enumerate(used_cams)}
    pt_idx = {p: k for k, p in enumerate(point_ids)}
    obs.sort(key=lambda o: pt_idx[o[0]])               # CSR by point (stable: keeps the order within a point)
    n_img, n_cam, n_pts = len(prob_images), len(used_cams), len(point_ids)
    stride = max([4] + [len(recon.cameras[c]["params"]) for c in used_cams])    # b2_ba_problem::camera_params_stride
    prob = {
        "qvec": np.stack([recon.images[i]["qvec"] for i in prob_images]) if n_img else np.zeros((0, 4)),
        "tvec": np.stack([recon.images[i]["tvec"] for i in prob_images]) if n_img else np.zeros((0, 3)),
        "img_cam": np.array([cam_idx[recon.images[i]["camera_id"]] for i in prob_images], dtype=np.int32),
        "pose_const": np.array([1 if const_pose[i] else 0 for i in prob_images], dtype=np.uint8),
        "tvec_const": np.array([sum(1 << k for k in config.ConstantTvec(i))
                                if (config.HasConstantTvec(i) and not const_pose[i]) else 0 for i in prob_images],
                               dtype=np.uint8),
        "cam_model": np.array([recon.cameras[c]["model"] for c in used_cams], dtype=np.int32),
        "cam_params": np.stack([np.r_[recon.cameras[c]["params"], np.zeros(stride)][:stride] for c in used_cams]) if n_cam else np.zeros((0, 4)),
        # ParameterizeCameras (.cc:474-512): constant if flagged -- cameras that never entered camera_ids_
        # have no residuals and stay untouched either way
        "cam_const": np.array([1 if (c in config_const_cameras or c not in camera_ids) else 0 for c in used_cams], dtype=np.uint8),
        "xyz": np.stack([recon.points3D[p]["xyz"] for p in point_ids]) if n_pts else np.zeros((0, 3)),
        # ParameterizePoints (.cc:514-526): not all of the track is in the problem, or explicitly constant
        "pt_const": np.array([1 if (len(recon.points3D[p]["track"]) > num_obs[p] or config.HasConstantPoint(p)) else 0
                              for p in point_ids], dtype=np.uint8),
        "obs_img": np.array([img_idx[o[1]] for o in obs], dtype=np.int32),
        "obs_pt": np.array([pt_idx[o[0]] for o in obs], dtype=np.int32),
        "obs_xy": np.array([[o[2], o[3]] for o in obs], dtype=np.float64).reshape(-1, 2),
    }
    for k in ("qvec", "tvec", "cam_params", "xyz"):
        prob[k] = np.ascontiguousarray(prob[k], dtype=np.float64)
    return prob, {"images": prob_images, "cameras": used_cams, "points": point_ids}


def unpack_problem(prob, maps, recon: Reconstruction) -> None:
    for k, i in enumerate(maps["images"]):
        recon.images[i]["qvec"] = prob["qvec"][k].copy()
        recon.images[i]["tvec"] = prob["tvec"][k].copy()
    for k, c in enumerate(maps["cameras"]):
        n = len(recon.cameras[c]["params"])
        recon.cameras[c]["params"] = prob["cam_params"][k][:n].copy()
    for k, p in enumerate(maps["points"]):
        recon.points3D[p]["xyz"] = prob["xyz"][k].copy()
