"""One PoseInertialOptimizationLastKeyFrame call on cuda:0 (for `ncu -k regex:pose_inertial_kernel`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
import orbx  # noqa: E402
import scenarios as sc  # noqa: E402

ctx = orbx.Context(0)
cam = orbx.make_camera()
opt = orbx.Optimizer(ctx)
i = sc.inertial_scenario(300, 300, 0.6)
opt.PoseInertialOptimizationLastKeyFrame(i["xw"], i["obs"], i["isg"], i["close"], cam, i["Tcw"], i["Tcb"], i["Tbc"], i["state"], i["kf"],
                                         i["preint"], i["infoI"], i["infoG"], i["infoA"])
