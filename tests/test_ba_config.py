"""The reference's own bundle-adjustment tests (src/optim/bundle_adjustment_test.cc:186-645) replayed through
the host-side mirror of BundleAdjustmentConfig / BundleAdjuster::SetUp (dagsfm_b200/ba_config.py) with the
ORACLE as solver: residual / effective-parameter counts and the CheckConstant* / CheckVariable* assertions."""
import numpy as np
import pytest

from dagsfm_b200.ba_config import BundleAdjustmentConfig, Reconstruction, pack_problem, unpack_problem
from oracle import pyoracle as orc


def generate_reconstruction(num_images, num_points, seed=0):
    """GenerateReconstruction, bundle_adjustment_test.cc:122-184."""
    rng = np.random.default_rng(seed)
    r = Reconstruction()
    for p in range(num_points):
        r.add_point3D(p, rng.uniform(-1, 1, 3))
    for i in range(num_images):
        r.add_camera(i, "SIMPLE_RADIAL", [1.2 * 1000, 500, 500, 0])
        tvec = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), 10.0])
        xy = []
        for p in range(num_points):
            X = r.points3D[p]["xyz"] + tvec
            xy.append(1200.0 * X[:2] / X[2] + 500.0 + rng.uniform(-2, 2, 2))
        r.add_image(i, i, [1, 0, 0, 0], tvec, xy)
    for i in range(num_images):
        for p in range(num_points):
            r.add_observation(p, i, p)
    return r


def solve(recon, config, refine=(1, 0, 1), refine_extrinsics=True):
    """BundleAdjuster(options, config).Solve(&reconstruction) with the oracle; BundleAdjustmentOptions defaults
    (bundle_adjustment.h:75-88: all tolerances 0, 100 iterations)."""
    prob, maps = pack_problem(recon, config, refine_extrinsics)
    prob["refine"] = refine
    s = orc.ba_solve(prob, max_num_iterations=100, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    unpack_problem(prob, maps, recon)
    return s


def variable_camera(r, o, c): return r.cameras[c]["params"][0] != o.cameras[c]["params"][0] and r.cameras[c]["params"][3] != o.cameras[c]["params"][3]
def constant_camera(r, o, c): return r.cameras[c]["params"][0] == o.cameras[c]["params"][0] and r.cameras[c]["params"][3] == o.cameras[c]["params"][3]
def variable_image(r, o, i): return (r.images[i]["qvec"] != o.images[i]["qvec"]).any() and (r.images[i]["tvec"] != o.images[i]["tvec"]).any()
def constant_image(r, o, i): return (r.images[i]["qvec"] == o.images[i]["qvec"]).all() and (r.images[i]["tvec"] == o.images[i]["tvec"]).all()
def constant_x_image(r, o, i): return variable_image(r, o, i) and r.images[i]["tvec"][0] == o.images[i]["tvec"][0]
def variable_point(r, o, p): return (r.points3D[p]["xyz"] != o.points3D[p]["xyz"]).any()
def constant_point(r, o, p): return (r.points3D[p]["xyz"] == o.points3D[p]["xyz"]).all()


def test_config_num_observations():                         # :186-208
    r = generate_reconstruction(4, 100)
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1)
    assert c.NumResiduals(r) == 400
    c.AddVariablePoint(1)
    assert c.NumResiduals(r) == 404
    c.AddConstantPoint(2)
    assert c.NumResiduals(r) == 408
    c.AddImage(2)
    assert c.NumResiduals(r) == 604
    c.AddImage(3)
    assert c.NumResiduals(r) == 800


def test_two_view():                                        # :210-245
    r = generate_reconstruction(2, 100); o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantTvec(1, [0])
    s = solve(r, c)
    assert (s.num_residuals, s.num_effective_parameters) == (400, 309)
    assert variable_camera(r, o, 0) and constant_image(r, o, 0)
    assert variable_camera(r, o, 1) and constant_x_image(r, o, 1)
    assert all(variable_point(r, o, p) for p in r.points3D)


def test_two_view_constant_camera():                        # :247-282
    r = generate_reconstruction(2, 100); o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantPose(1); c.SetConstantCamera(0)
    s = solve(r, c)
    assert (s.num_residuals, s.num_effective_parameters) == (400, 302)
    assert constant_camera(r, o, 0) and constant_image(r, o, 0)
    assert variable_camera(r, o, 1) and constant_image(r, o, 1)
    assert all(variable_point(r, o, p) for p in r.points3D)


def test_partially_contained_tracks():                      # :284-330
    r = generate_reconstruction(3, 100)
    variable_pid = r.images[2]["points2D"][0][2]
    r.delete_observation(2, 0)
    o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantPose(1)
    s = solve(r, c)
    assert (s.num_residuals, s.num_effective_parameters) == (400, 7)
    assert variable_camera(r, o, 0) and constant_image(r, o, 0)
    assert variable_camera(r, o, 1) and constant_image(r, o, 1)
    assert constant_camera(r, o, 2) and constant_image(r, o, 2)
    for p in r.points3D:
        assert variable_point(r, o, p) if p == variable_pid else constant_point(r, o, p)


def test_partially_contained_tracks_force_to_optimize_point():   # :332-388
    r = generate_reconstruction(3, 100)
    variable_pid = r.images[2]["points2D"][0][2]
    add_variable_pid = r.images[2]["points2D"][1][2]
    add_constant_pid = r.images[2]["points2D"][2][2]
    r.delete_observation(2, 0)
    o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantPose(1)
    c.AddVariablePoint(add_variable_pid); c.AddConstantPoint(add_constant_pid)
    s = solve(r, c)
    # + 2 residuals in the 3rd image for the added variable point; the added CONSTANT point's block in the
    # (constant) 3rd image has no free parameter and is dropped from the reduced program
    assert (s.num_residuals, s.num_effective_parameters) == (402, 10)
    assert variable_camera(r, o, 0) and constant_image(r, o, 0)
    assert variable_camera(r, o, 1) and constant_image(r, o, 1)
    assert constant_camera(r, o, 2) and constant_image(r, o, 2)
    for p in r.points3D:
        if p in (variable_pid, add_variable_pid):
            assert variable_point(r, o, p)
        else:
            assert constant_point(r, o, p)


def test_constant_points():                                 # :390-435
    r = generate_reconstruction(2, 100); o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantPose(1)
    c.AddConstantPoint(1); c.AddConstantPoint(2)
    s = solve(r, c)
    assert (s.num_residuals, s.num_effective_parameters) == (400, 298)
    for p in r.points3D:
        assert constant_point(r, o, p) if p in (1, 2) else variable_point(r, o, p)


def test_variable_image():                                  # :437-477
    r = generate_reconstruction(3, 100); o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.AddImage(2); c.SetConstantPose(0); c.SetConstantTvec(1, [0])
    s = solve(r, c)
    assert (s.num_residuals, s.num_effective_parameters) == (600, 317)
    assert constant_image(r, o, 0) and constant_x_image(r, o, 1) and variable_image(r, o, 2)
    assert all(variable_camera(r, o, k) for k in range(3)) and all(variable_point(r, o, p) for p in r.points3D)


@pytest.mark.parametrize("refine,n_eff,f_changes,pp_changes,k_changes", [
    ((0, 0, 1), 307, False, False, True),      # TestConstantFocalLength  :479-529
    ((1, 1, 1), 313, True, True, True),        # TestVariablePrincipalPoint :531-593
    ((1, 0, 0), 307, True, False, False),      # TestConstantExtraParam :595-645
])
def test_refine_flags(refine, n_eff, f_changes, pp_changes, k_changes):
    r = generate_reconstruction(2, 100); o = r.copy()
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantTvec(1, [0])
    s = solve(r, c, refine=refine)
    assert (s.num_residuals, s.num_effective_parameters) == (400, n_eff)
    for cam in (0, 1):
        a, b = r.cameras[cam]["params"], o.cameras[cam]["params"]
        assert (a[0] != b[0]) == f_changes and (a[3] != b[3]) == k_changes
        assert ((a[1] != b[1]) and (a[2] != b[2])) == pp_changes
    assert constant_image(r, o, 0) and constant_x_image(r, o, 1)


def test_config_invariants():                               # bundle_adjustment.cc:150-205 CHECKs
    c = BundleAdjustmentConfig()
    c.AddImage(0)
    c.SetConstantPose(0)
    with pytest.raises(AssertionError):
        c.SetConstantTvec(0, [0])
    c.AddVariablePoint(5)
    with pytest.raises(AssertionError):
        c.AddConstantPoint(5)
