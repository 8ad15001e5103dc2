"""CPU: the file:line citations the tracker harness and the docs lean on still point at what they claim, in the mounted
reference tree (skipped where /root/reference is absent).  Guards against citing the wrong lines (one such slip was found
and fixed in round 2: TrackLocalMap's inertial branch is src/Tracking.cc:2466-2490)."""
import os
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference sources not mounted")


def _lines(rel, a, b):
    with open(os.path.join(REF, rel), errors="replace") as f:
        L = f.readlines()
    return "".join(L[a - 1:b])


@pytest.mark.parametrize("rel,a,b,needles", [
    # TrackWithMotionModel: th = 7 (stereo) / 15 (mono), SearchByProjection(Cur, Last), retry with 2*th, PoseOptimization
    ("src/Tracking.cc", 2360, 2378, ["th=7", "th=15", "matcher.SearchByProjection(mCurrentFrame,mLastFrame,th", "2*th"]),
    ("src/Tracking.cc", 2395, 2395, ["Optimizer::PoseOptimization(&mCurrentFrame)"]),
    # TrackLocalMap: visual / LastFrame / LastKeyFrame branch
    ("src/Tracking.cc", 2466, 2490, ["isImuInitialized", "!mbMapUpdated", "PoseInertialOptimizationLastFrame(&mCurrentFrame)",
                                      "PoseInertialOptimizationLastKeyFrame(&mCurrentFrame)"]),
    # SearchLocalPoints: nnratio 0.8, SearchByProjection(F, local points, th)
    ("src/Tracking.cc", 2915, 2964, ["ORBmatcher matcher(0.8)", "matcher.SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th"]),
    # stereo Frame constructor: two extraction threads; ComputeStereoMatches BEFORE mb is assigned
    ("src/Frame.cc", 111, 114, ["thread threadLeft(&Frame::ExtractORB,this,0,imLeft,0,0)", "threadRight"]),
    ("src/Frame.cc", 132, 132, ["ComputeStereoMatches()"]),
    ("src/Frame.cc", 166, 166, ["mb = mbf/fx"]),
    ("src/Frame.cc", 985, 987, ["minZ = mb", "maxD = mbf/minZ"]),
    ("src/Frame.cc", 349, 349, ["ExtractORB(0,imGray,0,1000)"]),
    # pose helpers of the stereo-inertial glue
    ("src/Frame.cc", 520, 530, ["SetImuPoseVelocity", "tbw = -Rbw*twb", "mTcw = mImuCalib.Tcb*Tbw"]),
    ("src/Frame.cc", 534, 554, ["mOw = -mRcw.t()*mtcw", "GetImuPosition", "mRwc*mImuCalib.Tcb.rowRange(0,3).col(3)+mOw", "GetImuRotation"]),
    ("src/KeyFrame.cc", 136, 143, ["cv::Mat Rwc = Rcw.t();", "Ow = -Rwc*tcw;"]),
    # transposed products in the matcher
    ("src/ORBmatcher.cc", 1179, 1180, ["R12 = R1w*R2w.t();", "t12 = -R1w*R2w.t()*t2w+t1w;"]),
    ("src/ORBmatcher.cc", 2263, 2263, ["twc = -Rcw.t()*tcw"]),
    ("src/ORBmatcher.cc", 36, 38, ["TH_HIGH = 100", "TH_LOW = 50", "HISTO_LENGTH = 30"]),
    # chi2 schedules of the three pose optimisers
    ("src/Optimizer.cc", 1139, 1141, ["chi2Mono[4]={5.991,5.991,5.991,5.991}", "chi2Stereo[4]={7.815,7.815,7.815, 7.815}", "its[4]={10,10,10,10}"]),
    ("src/Optimizer.cc", 7882, 7885, ["chi2Mono[4]={12,7.5,5.991,5.991}", "chi2Stereo[4]={15.6,9.8,7.815,7.815}"]),
    ("src/Optimizer.cc", 8369, 8370, ["chi2Mono[4]={5.991,5.991,5.991,5.991}", "chi2Stereo[4]={15.6f,9.8f,7.815f,7.815f}"]),
    # the quadtree's pointer tie-break
    ("src/ORBextractor.cc", 674, 690, ["sort(vPrevSizeAndPointerToNode.begin(),vPrevSizeAndPointerToNode.end())"]),
])
def test_citation(rel, a, b, needles):
    text = _lines(rel, a, b)
    for n in needles:
        assert n in text, (rel, a, b, n)


@pytest.mark.parametrize("rel,a,sig", [
    # every hot-path row cites the line its reference function starts at
    ("src/ORBextractor.cc", 408, "ORBextractor::ORBextractor("),
    ("src/ORBextractor.cc", 1158, "void ORBextractor::ComputePyramid("),
    ("src/ORBextractor.cc", 763, "void ORBextractor::ComputeKeyPointsOctTree("),
    ("src/ORBextractor.cc", 537, "ORBextractor::DistributeOctTree("),
    ("src/ORBextractor.cc", 479, "void ExtractorNode::DivideNode("),
    ("src/ORBextractor.cc", 75, "static float IC_Angle("),
    ("src/ORBextractor.cc", 106, "static void computeOrbDescriptor("),
    ("src/ORBextractor.cc", 1074, "int ORBextractor::operator()("),
    ("src/Frame.cc", 444, "void Frame::AssignFeaturesToGrid("),
    ("src/Frame.cc", 755, "Frame::GetFeaturesInArea("),
    ("src/Frame.cc", 955, "void Frame::ComputeStereoMatches("),
    ("src/Frame.cc", 571, "bool Frame::isInFrustum("),
    ("src/Frame.cc", 874, "void Frame::UndistortKeyPoints("),
    ("src/ORBmatcher.cc", 2700, "int ORBmatcher::DescriptorDistance("),
    ("src/ORBmatcher.cc", 59, "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints"),
    ("src/ORBmatcher.cc", 2244, "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame"),
    ("src/ORBmatcher.cc", 1138, "int ORBmatcher::SearchForTriangulation("),
    ("src/ORBmatcher.cc", 323, "int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F"),
    ("src/ORBmatcher.cc", 1630, "int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints"),
    ("src/Optimizer.cc", 907, "int Optimizer::PoseOptimization(Frame *pFrame)"),
    ("src/Optimizer.cc", 1811, "void Optimizer::LocalBundleAdjustment(KeyFrame *pKF, bool* pbStopFlag"),
    ("src/Optimizer.cc", 7665, "int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame *pFrame"),
    ("src/Optimizer.cc", 8068, "int Optimizer::PoseInertialOptimizationLastFrame(Frame *pFrame"),
])
def test_function_starts_where_cited(rel, a, sig):
    text = _lines(rel, max(a - 12, 1), a + 12)
    norm = lambda t: "".join(t.split())   # noqa: E731
    assert norm(sig) in norm(text), (rel, a, sig)


def test_every_cited_range_exists():
    """Every `src/File.cc:a-b` / `include/File.h:a-b` citation in the C header and the docs lies inside that file."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"((?:src|include|Thirdparty)/[A-Za-z0-9_/\.]+\.(?:cc|cpp|h|hpp)):(\d+)(?:-(\d+))?")
    nlines = {}
    bad, total = [], 0
    for doc in ("include/orbx.h", "DESIGN.md", "INTEGRATION.md", "README.md"):
        text = open(os.path.join(root, doc), errors="replace").read()
        for m in pat.finditer(text):
            rel, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            path = os.path.join(REF, rel)
            if not os.path.exists(path):
                bad.append((doc, m.group(0), "no such file"))
                continue
            if rel not in nlines:
                with open(path, errors="replace") as f:
                    nlines[rel] = sum(1 for _ in f)
            total += 1
            if not (1 <= a <= b <= nlines[rel]):
                bad.append((doc, m.group(0), nlines[rel]))
    assert total > 100 and not bad, bad[:10]
