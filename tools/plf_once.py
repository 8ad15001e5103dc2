"""One PoseInertialOptimizationLastFrame call on cuda:0 (for `ncu -k regex:pose_inertial_lf` and ORBX_PIO_PROFILE=1)."""
import sys; sys.path.insert(0, "tests"); sys.path.insert(0, "awesome-orb-slam3-3dvisioncraft-version_b200")
import orbx, scenarios as sc
ctx = orbx.Context(0); cam = orbx.make_camera(); opt = orbx.Optimizer(ctx)
j = sc.inertial_lf_scenario(300, 300, 0.6)
opt.PoseInertialOptimizationLastFrame(j["xw"], j["obs"], j["isg"], j["close"], cam, j["Tcw"], j["Tcb"], j["Tbc"], j["state"], j["prev"], j["preint"], j["preint_jac"], j["preint_bias"], j["infoI"], j["infoG"], j["infoA"], j["prior_state"], j["prior_H"])
