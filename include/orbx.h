/*
 * orbx.h — C ABI of the B200-native ORB-SLAM3 tracking hot path.
 *
 * This is the drop-in boundary: flat POD arguments, caller-owned
 * buffers, `int` status returns, no C++ / torch / OpenCV types.  The C++ shim classes
 * under awesome-orb-slam3-3dvisioncraft-version_b200/shim/ keep the reference's
 * class signatures and forward to these entry points.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * reference tree).
 *
 * Status codes: ORBX_OK (0) success; ORBX_EMPTY (-1) empty image (the reference's
 * ORBextractor::operator() returns -1, src/ORBextractor.cc:1078); ORBX_EINVAL (-2) bad
 * argument; ORBX_ECUDA (-3) CUDA runtime failure (orbx_last_error() has the text);
 * ORBX_ECAP (-4) a caller-provided or internal capacity was exceeded.
 */
#ifndef ORBX_H_
#define ORBX_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_EMPTY (-1)
#define ORBX_EINVAL (-2)
#define ORBX_ECUDA (-3)
#define ORBX_ECAP (-4)

#define ORBX_MAX_LEVELS 16
#define ORBX_DESC_BYTES 32
#define ORBX_GRID_COLS 64 /* FRAME_GRID_COLS, include/Frame.h:39 */
#define ORBX_GRID_ROWS 48 /* FRAME_GRID_ROWS, include/Frame.h:38 */

/* The six cv::KeyPoint fields the reference fills (class_id is always -1). */
typedef struct orbx_keypoint {
  float x, y;     /* pt, in level-0 pixel coordinates */
  float size;     /* (int)(31*scale[octave]) */
  float angle;    /* degrees, cv::fastAtan2 */
  float response; /* FAST score */
  int32_t octave;
} orbx_keypoint;

typedef struct orbx_ctx orbx_ctx; /* one per (process, device): stream, scratch */
typedef struct orbx_ext orbx_ext; /* one per ORBextractor instance */

int orbx_abi_version(void);
/* Thread-local text of the last failure on the calling thread. */
const char *orbx_last_error(void);

/* Context.  Fails (returns NULL) when no CUDA device is usable: there is no CPU
 * fallback anywhere behind this ABI. */
orbx_ctx *orbx_create(int device);
void orbx_destroy(orbx_ctx *ctx);
/* cudaStream_t the context launches on (as void*), for callers that time with events. */
void *orbx_stream(orbx_ctx *ctx);
int orbx_synchronize(orbx_ctx *ctx);
/* Number of kernel launches issued through this context since creation. */
uint64_t orbx_launch_count(const orbx_ctx *ctx);

/* Page-locked host memory.  Host-pointer entry points accept any host memory; when the images they are given
 * live in page-locked memory (from here, cudaHostAlloc or cudaHostRegister) they are DMA'd to the device
 * directly instead of being staged through an internal pinned buffer. */
void *orbx_host_alloc(size_t bytes);
void orbx_host_free(void *p);

/* ------------------------------------------------------------------------------------
 * ORBextractor — replaces ORBextractor::ORBextractor (src/ORBextractor.cc:408-468,
 * include/ORBextractor.h:49-50).  max_w/max_h/max_batch size the device-resident
 * pyramid and scratch (max_batch images of at most max_w x max_h per call).
 * ---------------------------------------------------------------------------------- */
orbx_ext *orbx_extractor_create(orbx_ctx *ctx, int nfeatures, float scaleFactor, int nlevels,
                                int iniThFAST, int minThFAST, int max_w, int max_h,
                                int max_batch);
void orbx_extractor_destroy(orbx_ext *ext);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:59-81).  Arrays of nlevels floats;
 * any pointer may be NULL. */
int orbx_extractor_levels(const orbx_ext *ext);
int orbx_extractor_scale_tables(const orbx_ext *ext, float *scale, float *inv_scale,
                                float *sigma2, float *inv_sigma2);
/* mnFeaturesPerLevel (src/ORBextractor.cc:433-444). */
int orbx_extractor_features_per_level(const orbx_ext *ext, int *n_per_level);
/* Upper bound on the keypoints one image can yield (sum over levels of N_l + slack). */
int orbx_extractor_max_keypoints(const orbx_ext *ext);
/* cudaStream_t (as void*) this extractor enqueues on.  Every extractor owns its stream so
 * the Left/Right instances the reference drives from two threads (src/Frame.cc:111-114)
 * overlap on the device. */
void *orbx_extractor_stream(orbx_ext *ext);

/* ORBextractor::operator() (src/ORBextractor.cc:1074-1156) for one 8-bit gray image in
 * HOST memory.  kps/desc have room for `cap` keypoints; the quadtree may return a few more
 * than nfeatures (at most 2 extra per level), so cap >= orbx_extractor_max_keypoints() is
 * always enough.  *n_out = number of keypoints; *mono_index_out = the reference's return value
 * (monoIndex).  lap0/lap1 = vLappingArea[0..1] (mono ctor passes {0,1000}, stereo {0,0}:
 * src/Frame.cc:349,111-112).  Returns ORBX_EMPTY for w==0||h==0||img==NULL. */
int orbx_extract(orbx_ext *ext, const uint8_t *img, int w, int h, int stride, int lap0,
                 int lap1, orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out,
                 int *mono_index_out);

/* Many-stream mode: B same-sized images per call (B <= max_batch), HOST pointers.
 * Image b is imgs[b] with row pitch `stride`.  Outputs are [B][cap] / [B][cap][32];
 * n_out/mono_index_out are [B]. */
int orbx_extract_batch(orbx_ext *ext, int B, const uint8_t *const *imgs, int w, int h,
                       int stride, int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc,
                       int cap, int *n_out, int *mono_index_out);

/* Same, but every pointer is a DEVICE pointer (images packed [B][h][stride]); nothing is
 * copied and the call only enqueues work on orbx_stream(ctx).  Used by the resident-data
 * throughput measurement and by device-resident pipelines. */
int orbx_extract_batch_device(orbx_ext *ext, int B, const uint8_t *d_imgs, int w, int h,
                              int stride, int lap0, int lap1, orbx_keypoint *d_kps,
                              uint8_t *d_desc, int cap, int *d_n_out, int *d_mono_index_out);

/* Backs the public member `mvImagePyramid` (include/ORBextractor.h:83): copies level
 * `level` of image `b` of the LAST extract call (un-bordered; the 19-px frame the
 * reference allocates is never read by the hot path) to host memory. */
int orbx_pyramid_level(orbx_ext *ext, int b, int level, uint8_t *dst, int dst_stride,
                       int *w_out, int *h_out);

/* Per-stage device time of the LAST extract call, measured with CUDA events recorded on the
 * extractor's stream between its kernels (enable first; costs ~6 event records per call).
 * Stages: 0 pyramid (nlevels-1 resize launches), 1 FAST cells, 2 Gaussian blur, 3 quadtree,
 * 4 orientation+descriptor.  ms[] and launches[] hold ORBX_EXT_STAGES entries. */
#define ORBX_EXT_STAGES 5
int orbx_extractor_set_profiling(orbx_ext *ext, int enable);
int orbx_extractor_stage_ms(orbx_ext *ext, float *ms, int *launches);

/* Test hook: the Gaussian-blurred copy of level `level` of image `b` of the LAST extract call (the
 * plane computeOrbDescriptor samples; src/ORBextractor.cc:1120-1121), same conventions as
 * orbx_pyramid_level. */
int orbx_debug_blur_level(orbx_ext *ext, int b, int level, uint8_t *dst, int dst_stride,
                          int *w_out, int *h_out);

/* Test/diagnostic view: FAST candidates of (image b, level) of the last call, i.e. the
 * contents of `vToDistributeKeys` (src/ORBextractor.cc:775,845-851) in an unspecified
 * order.  xy = [n][2] int16 (coordinates relative to minBorder), score = [n] uint8. */
int orbx_debug_candidates(orbx_ext *ext, int b, int level, int16_t *xy, uint8_t *score,
                          int cap, int *n_out);

/* Tuning hook: per-phase SM-cycle totals of the PoseOptimization kernel (0 build pass, 1 reduction, 2 solve + exp,
 * 3 trial residual pass, 4 trial reduction, 5 accept/reject logic, 6 chi2 classification, 7 whole kernel; counts:
 * 8 builds, 9 solved trials, 10 replayed trials, 11 CTAs).  enable = 1 clears and starts, 0 stops; `out` (16 values,
 * may be NULL) receives the totals accumulated so far.  Synchronises the device. */
int orbx_debug_pose_opt_profile(orbx_ctx *ctx, int enable, unsigned long long *out);

/* ====================================================================================
 * Matchers.  A Frame / KeyFrame crosses the boundary as flat arrays (the shim flattens the
 * reference's pointer graph once per call, under the same mutexes the reference takes).
 * Only the pinhole, Nleft == -1 configuration (mono / rectified stereo / RGB-D) is covered;
 * KannalaBrandt8 stereo-fisheye (mpCamera2) is out of scope.
 * ================================================================================== */
typedef struct orbx_frame_desc {
  int32_t n;                 /* Frame::N */
  const orbx_keypoint *kps;  /* mvKeysUn: x, y, octave, angle are read */
  const uint8_t *desc;       /* mDescriptors, [n][32] */
  const float *uright;       /* mvuRight [n]; NULL = monocular (all -1) */
  float min_x, min_y, max_x, max_y; /* mnMinX, mnMinY, mnMaxX, mnMaxY (src/Frame.cc:147-152) */
} orbx_frame_desc;

typedef struct orbx_camera {
  float fx, fy, cx, cy; /* Pinhole::mvParameters (src/CameraModels/Pinhole.cpp:31-50) */
  float bf, b;          /* Frame::mbf, Frame::mb = mbf/fx (src/Frame.cc:166) */
} orbx_camera;

/* Device-resident Frame / KeyFrame (Frame::Frame + AssignFeaturesToGrid, src/Frame.cc:90-170,444-478).
 * orbx_frame_upload copies the descriptor's arrays to the device ONCE and builds the 64x48 grid there; afterwards every
 * entry point below that is handed a descriptor over the SAME host arrays (same kps / desc / uright pointers, n and
 * bounds) uses the resident copy: no per-call upload of the frame, no per-call grid build.  The host arrays must stay
 * unchanged while the handle lives (the reference's Frame does not change these members after its constructor and
 * ComputeStereoMatches); release the handle when the Frame dies.  Signatures of the matcher calls are unchanged. */
typedef struct orbx_frame orbx_frame;
orbx_frame *orbx_frame_upload(orbx_ctx *ctx, const orbx_frame_desc *frame);
void orbx_frame_release(orbx_frame *frame);
int orbx_frame_count(const orbx_ctx *ctx);

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2700-2716): Hamming distance of two
 * 256-bit descriptors.  A static CPU helper in the reference (also called from Frame.cc,
 * MapPoint.cc, LoopClosing.cc); provided here for the shim, computed on the host. */
int orbx_descriptor_distance(const uint8_t *a, const uint8_t *b);

/* Frame::AssignFeaturesToGrid + GetFeaturesInArea (src/Frame.cc:444-478,755-850) as a device
 * primitive, exposed for tests: returns, for each of nq queries (x, y, r, minLevel, maxLevel),
 * the indices GetFeaturesInArea would return, in the reference's order (cell column outer,
 * cell row inner, insertion order inside a cell).  out_idx is [nq][cap], out_n is [nq]
 * (the true count, which may exceed cap). */
int orbx_features_in_area(orbx_ctx *ctx, const orbx_frame_desc *frame, int nq, const float *x,
                          const float *y, const float *r, const int32_t *min_level,
                          const int32_t *max_level, int32_t *out_idx, int cap, int32_t *out_n);

/* Frame::ComputeStereoMatches (src/Frame.cc:955-1133).  extL/extR hold the pyramids
 * (mvImagePyramid) of the extraction that produced the keypoints: image bL of extL's last call
 * and image bR of extR's last call (extL may equal extR in batch mode).  Host arrays in,
 * uright/depth [nL] out (-1 where unmatched). */
int orbx_stereo_match(orbx_ctx *ctx, orbx_ext *extL, int bL, orbx_ext *extR, int bR,
                      const orbx_keypoint *kpL, const uint8_t *descL, int nL,
                      const orbx_keypoint *kpR, const uint8_t *descR, int nR, float bf, float b,
                      float *uright, float *depth);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:59-255), track-local-map search.
 *   kp_blocked[n]   : 1 where F.mvpMapPoints[i] != NULL && Observations() > 0 at entry
 *   per MapPoint q  : proj_x/proj_y/proj_xr = mTrackProjX/Y/XR, level = mnTrackScaleLevel,
 *                     view_cos = mTrackViewCos, mp_desc = GetDescriptor(),
 *                     flags bit0 = mbTrackInView && !isBad() && !(bFarPoints && depth > thFar),
 *                           bit1 = Observations() > 0 (an assigned keypoint then blocks later queries)
 *   th, nnratio     : the call's th and the matcher's mfNNratio
 *   scale_factors   : F.mvScaleFactors [nlevels]
 * Out: best_idx[nq] = keypoint the MapPoint was assigned to, or -1; *nmatches = return value.
 * The caller replays  F.mvpMapPoints[best_idx[q]] = pMP_q  in increasing q. */
int orbx_search_by_projection_map(orbx_ctx *ctx, const orbx_frame_desc *frame,
                                  const uint8_t *kp_blocked, int nq, const float *proj_x,
                                  const float *proj_y, const float *proj_xr, const int32_t *level,
                                  const float *view_cos, const uint8_t *mp_desc,
                                  const uint8_t *flags, float th, float nnratio,
                                  const float *scale_factors, int nlevels, int32_t *best_idx,
                                  int32_t *nmatches);

/* ORBmatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)
 * (src/ORBmatcher.cc:2244-2509), motion-model search.
 *   Tcw_cur / Tcw_last : 4x4 row-major float32 poses (mTcw)
 *   per last-frame keypoint q (nq = Last.N): flags bit0 = mvpMapPoints[q] && !mvbOutlier[q],
 *                     bit1 = Observations() > 0; xw[q][3] = GetWorldPos(); octave, angle of the
 *                     last frame's keypoint; mp_desc = GetDescriptor()
 * Out: match_idx[nq] = current keypoint chosen for q (before the rotation-histogram filter) or -1;
 *      kept[nq] = 0 where the rotation filter removed the match; cur_match[n] = final
 *      CurrentFrame.mvpMapPoints as a last-frame index or -1; *nmatches = return value. */
int orbx_search_by_projection_frame(orbx_ctx *ctx, const orbx_frame_desc *cur,
                                    const uint8_t *cur_blocked, const orbx_camera *cam,
                                    const float *Tcw_cur, const float *Tcw_last, int nq,
                                    const uint8_t *flags, const float *xw, const int32_t *octave,
                                    const float *angle, const uint8_t *mp_desc, float th, int bMono,
                                    int check_orientation, const float *scale_factors, int nlevels,
                                    int32_t *match_idx, uint8_t *kept, int32_t *cur_match,
                                    int32_t *nmatches);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, bCoarse)
 * (src/ORBmatcher.cc:1138-1428); the F12 argument is unused by the reference overload.
 *   kf1/kf2        : keypoints (mvKeysUn), descriptors, mvuRight of the two keyframes
 *   has_mp1/2      : GetMapPoint(i) != NULL
 *   fv*_node/off/idx : DBoW2::FeatureVector as CSR — node ids ascending [nn], offsets [nn+1],
 *                     feature indices in the map's vector order
 *   R1w,t1w,R2w,t2w : GetRotation()/GetTranslation() (row-major 3x3 / 3), float32
 *   level_sigma2   : mvLevelSigma2 [nlevels]; scale_factors = mvScaleFactors [nlevels]
 * Out: match12[n1] = vMatches12 after the rotation filter; *nmatches = return value. */
int orbx_search_for_triangulation(orbx_ctx *ctx, const orbx_frame_desc *kf1,
                                  const orbx_frame_desc *kf2, const uint8_t *has_mp1,
                                  const uint8_t *has_mp2, int nn1, const int32_t *fv1_node,
                                  const int32_t *fv1_off, const int32_t *fv1_idx, int nn2,
                                  const int32_t *fv2_node, const int32_t *fv2_off,
                                  const int32_t *fv2_idx, const orbx_camera *cam1,
                                  const orbx_camera *cam2, const float *R1w, const float *t1w,
                                  const float *R2w, const float *t2w, const float *level_sigma2,
                                  const float *scale_factors, int nlevels, int only_stereo,
                                  int coarse, int check_orientation, int32_t *match12,
                                  int32_t *nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
 * (src/ORBmatcher.cc:323-591; Tracking::TrackReferenceKeyFrame src/Tracking.cc:2186-2210 and relocalisation), SURVEY.md §8 f2.
 *   kf / frame   : keypoints (angle is read: kf->kps = mvKeysUn, frame->kps = mvKeys) and descriptors
 *   kf_has_mp[i] : vpMapPointsKF[i] != NULL && !isBad()
 *   fv_*         : pKF->mFeatVec and F.mFeatVec as CSR (orbx_vocabulary_transform's layout)
 *   nnratio, check_orientation : the matcher's mfNNratio and mbCheckOrientation
 * Out: match_f[frame->n] = keyframe feature i whose MapPoint vpMapPointMatches[j] receives, or -1;
 *      *nmatches = return value.  Pinhole / Nleft == -1 only. */
int orbx_search_by_bow(orbx_ctx *ctx, const orbx_frame_desc *kf, const orbx_frame_desc *frame,
                       const uint8_t *kf_has_mp, int nn_kf, const int32_t *fv_kf_node,
                       const int32_t *fv_kf_off, const int32_t *fv_kf_idx, int nn_f,
                       const int32_t *fv_f_node, const int32_t *fv_f_off, const int32_t *fv_f_idx,
                       float nnratio, int check_orientation, int32_t *match_f, int32_t *nmatches);

/* ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th, bRight = false)
 * (src/ORBmatcher.cc:1630-1883; LocalMapping::SearchInNeighbors src/LocalMapping.cc:1006-1042), SURVEY.md §8 f2 —
 * the search half: projection, distance / viewing-angle / predicted-scale gates, GetFeaturesInArea, chi2 gates,
 * best descriptor.  The map surgery after a hit (Replace / AddObservation / AddMapPoint, :1844-1867) stays with the
 * caller, which replays it over best_idx in increasing i.
 *   kf           : mvKeysUn, mDescriptors, mvuRight, image bounds of pKF
 *   Rcw[9], tcw[3], Ow[3] : GetRotation / GetTranslation / GetCameraCenter (row-major float32)
 *   flags[i] bit0 : pMP && !isBad() && !IsInKeyFrame(pKF)
 *   xw[i][3] = GetWorldPos(); mp_max_dist / mp_min_dist = the members mfMaxDistance / mfMinDistance (the 1.2 / 0.8
 *   invariance factors and PredictScale are applied inside, src/MapPoint.cc:566-593); mp_normal[i][3] = GetNormal();
 *   mp_desc = GetDescriptor(); scale_factors = mvScaleFactors; inv_level_sigma2 = mvInvLevelSigma2;
 *   log_scale_factor = mfLogScaleFactor.
 * Out: best_idx[i] = keyframe keypoint the MapPoint fuses into (bestDist <= TH_LOW) or -1; *nfused = number of hits. */
int orbx_fuse(orbx_ctx *ctx, const orbx_frame_desc *kf, const orbx_camera *cam, const float *Rcw,
              const float *tcw, const float *Ow, int nmp, const uint8_t *flags, const float *xw,
              const float *mp_max_dist, const float *mp_min_dist, const float *mp_normal,
              const uint8_t *mp_desc, float th, const float *scale_factors,
              const float *inv_level_sigma2, int nlevels, float log_scale_factor, int32_t *best_idx,
              int32_t *nfused);

/* ====================================================================================
 * DBoW2 vocabulary: the step that feeds SearchForTriangulation / SearchByBoW.
 * ================================================================================== */
typedef struct orbx_voc orbx_voc;

/* TemplatedVocabulary::loadFromBinaryFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1442-1478): the
 * reference's Vocabulary/ORBvoc.bin format — header {uint32 n_nodes (root included), uint32 node_size = 41,
 * int32 k, int32 L, int32 scoring, int32 weighting}, then per non-root node {int32 parent, uint8 desc[32],
 * float32 weight, uint8 is_leaf}.  The tree is uploaded once and stays resident in HBM.  NULL on failure. */
orbx_voc *orbx_vocabulary_load(orbx_ctx *ctx, const char *path);
orbx_voc *orbx_vocabulary_from_memory(orbx_ctx *ctx, const void *data, size_t bytes);
void orbx_vocabulary_destroy(orbx_voc *voc);
/* m_k, m_L, node count (root included), word count, ScoringType, WeightingType; any pointer may be NULL. */
int orbx_vocabulary_info(const orbx_voc *voc, int *k, int *L, int *n_nodes, int *n_words, int *scoring,
                         int *weighting);

/* TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
 * (TemplatedVocabulary.h:1140-1219; called as transform(vCurrentDesc, mBowVec, mFeatVec, 4) from
 * Frame::ComputeBoW src/Frame.cc:865-872 and KeyFrame::ComputeBoW src/KeyFrame.cc:125-134).
 *   desc[n][32]  : the frame's descriptors (HOST)
 * Out (caller-owned, room for n entries; fv_off n+1):
 *   bow_word/bow_value[*n_bow] : mBowVec in std::map order (ascending WordId), values as the reference's doubles
 *   fv_node[*n_fv], fv_off[*n_fv+1], fv_idx[...] : mFeatVec as CSR (ascending NodeId, feature indices ascending)
 *                 — the layout orbx_search_for_triangulation takes. */
int orbx_vocabulary_transform(orbx_voc *voc, const uint8_t *desc, int n, int levelsup, int32_t *bow_word,
                              double *bow_value, int32_t *n_bow, int32_t *fv_node, int32_t *fv_off,
                              int32_t *fv_idx, int32_t *n_fv);
/* Many-frame mode, DEVICE pointers, enqueues on orbx_stream(ctx): descriptors [F][cap][32] with d_n[F] valid rows
 * each; d_leaf/d_node [F][cap] receive the leaf node reached per feature (-1 = stopped word) and the node at level
 * L - levelsup; outputs [F][cap] (d_fv_off [F][cap+1]), counts [F]. */
int orbx_vocabulary_transform_batch_device(orbx_voc *voc, int F, const uint8_t *d_desc, const int32_t *d_n,
                                           int cap, int levelsup, int32_t *d_leaf, int32_t *d_node,
                                           int32_t *d_bow_word, double *d_bow_value, int32_t *d_n_bow,
                                           int32_t *d_fv_node, int32_t *d_fv_off, int32_t *d_fv_idx,
                                           int32_t *d_n_fv);

/* ====================================================================================
 * Per-frame glue between extractor and matcher.
 * ================================================================================== */

/* Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit) for nmp MapPoints at once (src/Frame.cc:571-650,
 * Nleft == -1; Tracking::SearchLocalPoints src/Tracking.cc:2884-2900 calls it per local MapPoint with 0.5).
 *   Rcw[9], tcw[3], Ow[3] : mRcw, mtcw, mOw;  min/max : mnMinX, mnMaxX, mnMinY, mnMaxY;  nlevels = mnScaleLevels;
 *   log_scale_factor = mfLogScaleFactor;  xw = GetWorldPos();  mp_max_dist / mp_min_dist = mfMaxDistance /
 *   mfMinDistance (raw members, see orbx_fuse);  mp_normal = GetNormal()
 * Out, per MapPoint: in_view = mbTrackInView; proj_x/proj_y = mTrackProjX/Y (-1 when behind the camera or outside the
 * image, otherwise the projection even if a later test fails, as in the reference).  proj_xr, depth, level, view_cos
 * (mTrackProjXR, mTrackDepth, mnTrackScaleLevel, mTrackViewCos) are IN/OUT: overwritten only where in_view = 1, the
 * caller's (stale) values survive elsewhere exactly as the MapPoint fields do. */
int orbx_is_in_frustum(orbx_ctx *ctx, const orbx_camera *cam, const float *Rcw, const float *tcw,
                       const float *Ow, float min_x, float max_x, float min_y, float max_y,
                       float viewing_cos_limit, int nlevels, float log_scale_factor, int nmp,
                       const float *xw, const float *mp_max_dist, const float *mp_min_dist,
                       const float *mp_normal, uint8_t *in_view, float *proj_x, float *proj_y,
                       float *proj_xr, float *depth, int32_t *level, float *view_cos,
                       int32_t *n_in_view);

/* Frame::UndistortKeyPoints (src/Frame.cc:874-924): cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) on the
 * n keypoint positions xy[n][2]; dist_coef = (k1, k2, p1, p2[, k3]), n_dist = 4 or 5.  When k1 == 0 the points are
 * copied (the reference's early return).  out_xy may alias xy. */
int orbx_undistort_keypoints(orbx_ctx *ctx, const float *xy, int n, const orbx_camera *cam,
                             const float *dist_coef, int n_dist, float *out_xy);

/* ====================================================================================
 * Optimisers (g2o linearisation + Levenberg-Marquardt as modified by ORB-SLAM3, all fp64 on
 * the device; one thread block runs a whole optimisation).
 * ================================================================================== */

/* Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:907-1272).
 *   n_edges     : keypoints i with pFrame->mvpMapPoints[i] != NULL, in increasing i
 *   xw[e][3]    : pMP->GetWorldPos() (float32, as the reference copies it)
 *   obs[e][3]   : kpUn.pt.x, kpUn.pt.y, mvuRight[i]   (mvuRight < 0 => monocular 2-D edge, else 3-D stereo edge)
 *   inv_sigma2  : pFrame->mvInvLevelSigma2[kpUn.octave]
 *   Tcw[16]     : pFrame->mTcw, row-major float32; overwritten with the optimised pose
 * Out: outlier[e] = pFrame->mvbOutlier; *n_inliers = the reference's return value
 *      (nInitialCorrespondences - nBad, 0 and pose untouched when n_edges < 3);
 *      iters[4] = LM iterations g2o ran in each of the four rounds (parity evidence). */
int orbx_pose_optimization(orbx_ctx *ctx, int n_edges, const float *xw, const float *obs,
                           const float *inv_sigma2, const orbx_camera *cam, float *Tcw,
                           uint8_t *outlier, int32_t *n_inliers, int32_t *iters);

/* Many-stream mode: P independent PoseOptimization problems in one launch.  edge_ofs[P+1]
 * delimits each problem's slice of xw/obs/inv_sigma2/outlier; Tcw is [P][16]; n_inliers [P];
 * iters [P][4].  All pointers are DEVICE pointers; the call only enqueues on orbx_stream(ctx). */
int orbx_pose_optimization_batch_device(orbx_ctx *ctx, int P, const int32_t *d_edge_ofs,
                                        const float *d_xw, const float *d_obs,
                                        const float *d_inv_sigma2, const orbx_camera *cam,
                                        float *d_Tcw, uint8_t *d_outlier, int32_t *d_n_inliers,
                                        int32_t *d_iters, double *d_scratch /* [3 * total_edges] */);

/* Optimizer::LocalBundleAdjustment(KeyFrame*, bool *pbStopFlag, Map*, int &num_fixedKF)
 * (src/Optimizer.cc:1811-2523) — the numeric core: graph = n_kf SE3 vertices (kf_fixed[k] != 0:
 * fixed), n_mp marginalised points, one edge per observation (mono 2-D when obs[e][2] < 0, else
 * stereo 3-D), Huber on every edge, optimize(5) then optimize(10), final chi2/depth test.
 *   kf_Tcw[n_kf][16], mp_xyz[n_mp][3] : in/out, float32 (written back unless aborted)
 *   e_kf/e_mp/e_obs/e_inv_sigma2      : per observation
 *   lambda_init : 0 => g2o's tau*max(diag) rule; the reference passes 100 for inertial maps (:1968)
 *   stop_flag   : HOST pointer polled like g2o's forceStopFlag (*pbStopFlag), may be NULL
 * Out: edge_bad[e] = 1 for observations the reference would erase (vToErase, :2295-2344);
 *      iters[2] = LM iterations of the two optimize() calls;
 *      *status = 0 done, 1 stopped before optimising, 2 rejected (>= 50 % bad: nothing written). */
int orbx_local_ba(orbx_ctx *ctx, int n_kf, float *kf_Tcw, const uint8_t *kf_fixed, int n_mp,
                  float *mp_xyz, int n_edges, const int32_t *e_kf, const int32_t *e_mp,
                  const float *e_obs, const float *e_inv_sigma2, const orbx_camera *cam,
                  double lambda_init, const volatile uint8_t *stop_flag, uint8_t *edge_bad,
                  int32_t *iters, int32_t *status);

/* ---- keyframe-rate half of the batched many-stream mode (BASELINE.json config 5: "full track+LocalBA") ----
 * LocalMapping::Run (src/LocalMapping.cc:128) processes every new keyframe with CreateNewMapPoints (:501-628: one
 * ORBmatcher::SearchForTriangulation per covisible neighbour, nn = 10 for stereo) and Optimizer::LocalBundleAdjustment
 * (:201).  One camera stream yields one such step per keyframe; S independent streams yield S of them, which these two
 * PREPARED PLANS run in one launch sequence each: prepare() uploads the problems once into one device pool, run() only
 * enqueues on a CUDA stream (cuda_stream = NULL: the context's) and may be repeated (every run restarts from the
 * inputs given to prepare), fetch() waits for the last run and copies the results into the problem structs.
 * Per-problem results are bit-identical to the single-call entry points (tests/test_keyframe_gpu.py). */
typedef struct orbx_tri_problem { /* one orbx_search_for_triangulation call; host pointers, same meaning */
  const orbx_frame_desc *kf1, *kf2;
  const uint8_t *has_mp1, *has_mp2;
  int32_t nn1;
  const int32_t *fv1_node, *fv1_off, *fv1_idx;
  int32_t nn2;
  const int32_t *fv2_node, *fv2_off, *fv2_idx;
  const orbx_camera *cam1, *cam2;
  const float *R1w, *t1w, *R2w, *t2w;
  int32_t only_stereo, coarse;
  int32_t *match12; /* out [kf1->n], may be NULL */
  int32_t nmatches; /* out */
} orbx_tri_problem;
typedef struct orbx_tri_batch orbx_tri_batch;
orbx_tri_batch *orbx_tri_batch_prepare(orbx_ctx *ctx, int n_problems, const orbx_tri_problem *problems,
                                       const float *level_sigma2, const float *scale_factors, int nlevels,
                                       int check_orientation);
int orbx_tri_batch_run(orbx_tri_batch *batch, void *cuda_stream);
int orbx_tri_batch_fetch(orbx_tri_batch *batch, orbx_tri_problem *problems);
void orbx_tri_batch_destroy(orbx_tri_batch *batch);

typedef struct orbx_lba_problem { /* one orbx_local_ba call; host pointers, same meaning */
  int32_t n_kf, n_mp, n_edges;
  float *kf_Tcw; /* in (prepare) / out (fetch) [n_kf][16] */
  const uint8_t *kf_fixed;
  float *mp_xyz; /* in / out [n_mp][3] */
  const int32_t *e_kf, *e_mp;
  const float *e_obs, *e_inv_sigma2;
  double lambda_init;
  uint8_t *edge_bad; /* out [n_edges], may be NULL */
  int32_t iters[2];  /* out */
  int32_t status;    /* out: 0 done, 2 rejected (>= 50 % bad: poses / points left as given) */
} orbx_lba_problem;
typedef struct orbx_lba_batch orbx_lba_batch;
orbx_lba_batch *orbx_lba_batch_prepare(orbx_ctx *ctx, int n_problems, const orbx_lba_problem *problems,
                                       const orbx_camera *cam);
int orbx_lba_batch_run(orbx_lba_batch *batch, void *cuda_stream);
int orbx_lba_batch_fetch(orbx_lba_batch *batch, orbx_lba_problem *problems);
void orbx_lba_batch_destroy(orbx_lba_batch *batch);
size_t orbx_lba_batch_device_bytes(const orbx_lba_batch *batch);

/* Optimizer::PoseInertialOptimizationLastKeyFrame(Frame*, bool bRecInit) (src/Optimizer.cc:7665-8066;
 * vertices/edges: include/G2oTypes.h:387-491, src/G2oTypes.cc:170-220,385-407,496-520,730-812) — the
 * visual-inertial replacement of PoseOptimization that Tracking::TrackLocalMap calls when the map was
 * updated (src/Tracking.cc:2466-2490).  15 free unknowns (pose, velocity, gyro bias, acc bias of the
 * frame), the last keyframe's four vertices fixed, Gauss-Newton + dense LDL^T, 4 x 10 iterations.
 *   xw/obs/inv_sigma2 : as orbx_pose_optimization (inv_sigma2 already divided by uncertainty2)
 *   close_pt[e]       : pMP->mTrackDepth < 10.f
 *   Tcw/Tcb/Tbc[16]   : pFrame->mTcw, mImuCalib.Tcb, mImuCalib.Tbc (row-major float32)
 *   state[21]         : in/out, double: Rwb[9], twb[3], velocity[3], gyro bias[3], acc bias[3] of the frame
 *   kf_state[21]      : the same of pFrame->mpLastKeyFrame (fixed)
 *   preint[16]        : mpImuPreintegrated->GetDeltaRotation/Velocity/Position(keyframe bias), dT
 *   info_inertial[81] : EdgeInertial's information (ctor, src/G2oTypes.cc:700-727);
 *   info_gyro/acc[9]  : C.block<3,3>(9,9).inverse(), C.block<3,3>(12,12).inverse()
 *   rec_init          : bRecInit
 * Out: outlier[e] (pFrame->mvbOutlier), H15[225] row-major (the Hessian handed to ConstraintPoseImu,
 *      :8030-8063), *n_ret = nInitialCorrespondences - nBad, iters[4] = Gauss-Newton iterations per round.
 * Conventions where the reference cannot be reproduced bit for bit (re-orthonormalisation): DESIGN.md §7. */
int orbx_pose_inertial_optimization_last_keyframe(
    orbx_ctx *ctx, int n_edges, const float *xw, const float *obs, const float *inv_sigma2,
    const uint8_t *close_pt, const orbx_camera *cam, const float *Tcw, const float *Tcb,
    const float *Tbc, double *state, const double *kf_state, const double *preint,
    const double *info_inertial, const double *info_gyro, const double *info_acc, int rec_init,
    uint8_t *outlier, double *H15, int32_t *n_ret, int32_t *iters);

/* Many-stream form of the call above: P independent frames in one launch (one CTA per problem);
 * argument layout as orbx_pose_inertial_optimization_last_frame_batch below. */
int orbx_pose_inertial_optimization_last_keyframe_batch(
    orbx_ctx *ctx, int P, const int32_t *edge_ofs, const float *xw, const float *obs,
    const float *inv_sigma2, const uint8_t *close_pt, const orbx_camera *cam, const float *Tcw,
    const float *Tcb, const float *Tbc, double *state, const double *kf_state, const double *preint,
    const double *info_inertial, const double *info_gyro, const double *info_acc, int rec_init,
    uint8_t *outlier, double *H15, int32_t *n_ret, int32_t *iters);

/* Optimizer::PoseInertialOptimizationLastFrame(Frame*, bool bRecInit) (src/Optimizer.cc:8068-8603) — what
 * Tracking::TrackLocalMap calls on every other visual-inertial frame (src/Tracking.cc:2466-2490).  The previous
 * frame's four vertices are free too (30 unknowns), tied down by EdgePriorPoseImu (src/G2oTypes.cc:941-981,
 * Huber delta 5) built from pFp->mpcpi; EdgeInertial is linearised with respect to all six vertices
 * (:752-812) with bias-corrected deltas (src/ImuTypes.cc:367-394); at the end the previous frame is
 * marginalised out of the 30x30 Hessian (Optimizer::Marginalize, :5366-5450).
 *   state[21]       : in/out, the frame (layout as above);  prev_state[21]: pFrame->mpPrevFrame (in)
 *   preint[16]      : RAW dR[9], dV[3], dP[3], dT of pFrame->mpImuPreintegratedFrame
 *   preint_jac[45]  : its JRg, JVg, JVa, JPg, JPa (3x3 row-major each);  preint_bias[6]: its bias, gyro xyz then acc xyz
 *   info_*          : as above, from mpImuPreintegratedFrame->C
 *   prior_state[21] : pFp->mpcpi->Rwb, twb, vwb, bg, ba;  prior_H[225]: pFp->mpcpi->H
 * Out: as above; H15 = H.block<15,15>(15,15) after Marginalize(H, 0, 14), i.e. the argument of the new
 *      ConstraintPoseImu. */
int orbx_pose_inertial_optimization_last_frame(
    orbx_ctx *ctx, int n_edges, const float *xw, const float *obs, const float *inv_sigma2,
    const uint8_t *close_pt, const orbx_camera *cam, const float *Tcw, const float *Tcb,
    const float *Tbc, double *state, const double *prev_state, const double *preint,
    const double *preint_jac, const double *preint_bias, const double *info_inertial,
    const double *info_gyro, const double *info_acc, const double *prior_state,
    const double *prior_H, int rec_init, uint8_t *outlier, double *H15, int32_t *n_ret,
    int32_t *iters);

/* Many-stream form: P independent frames in one launch (one CTA per problem).  edge_ofs[P+1] delimits
 * each problem's slice of xw/obs/inv_sigma2/close_pt/outlier; every other per-problem argument is the
 * single-call argument with a leading [P] dimension (Tcw [P][16], state [P][21], ... H15 [P][225],
 * n_ret [P], iters [P][4]); cam, Tcb and Tbc are shared by the rig.  Results are identical to P single calls. */
int orbx_pose_inertial_optimization_last_frame_batch(
    orbx_ctx *ctx, int P, const int32_t *edge_ofs, const float *xw, const float *obs,
    const float *inv_sigma2, const uint8_t *close_pt, const orbx_camera *cam, const float *Tcw,
    const float *Tcb, const float *Tbc, double *state, const double *prev_state,
    const double *preint, const double *preint_jac, const double *preint_bias,
    const double *info_inertial, const double *info_gyro, const double *info_acc,
    const double *prior_state, const double *prior_H, int rec_init, uint8_t *outlier, double *H15,
    int32_t *n_ret, int32_t *iters);

/* ====================================================================================
 * Many-stream tracking replay: S independent stereo streams
 * advance one frame per call, everything device-resident between stages:
 *   extract L+R (2*S images) -> ComputeStereoMatches -> SearchByProjection(Cur, Last) ->
 *   PoseOptimization -> SearchByProjection(F, local map) -> PoseOptimization.
 * No dataset is available offline, so the map a stream tracks against is synthesised from the frame
 * itself: its stereo points back-projected at Tcw_true play the role of the last frame's / local
 * map's MapPoints (real descriptors, real candidate densities), and tracking starts from
 * Tcw_prior (the motion-model guess).  The per-stage kernels are exactly the ones behind the
 * single-frame entry points above.
 * ================================================================================== */
typedef struct orbx_tracker orbx_tracker;
#define ORBX_TRACK_STATS 8 /* nL, nR, nStereo, matches(last frame), inliers, matches(local map), inliers, LM its */

/* Monocular tracker (BASELINE config 1).  Replaces Frame::Frame(mono) src/Frame.cc:308-349 ->
 * Tracking::TrackWithMotionModel with bMono (th = 15, src/Tracking.cc:2364-2378) -> PoseOptimization's monocular edges
 * (src/Optimizer.cc:961-1010) -> TrackLocalMap.  ONE image per stream ([I0, I1, ...] wherever the stereo tracker takes
 * [L0, R0, L1, R1, ...]); ext max_batch >= S; mvuRight = mvDepth = -1 for every keypoint, so there is no back-projection
 * harness: the map must be given with orbx_tracker_set_map / _upload_map before the first step (ORBX_EINVAL otherwise).
 * cam->b / cam->bf may be 0.  stats[1] (nR) and stats[2] (nStereo) are 0.  Everything else as orbx_tracker_create. */
orbx_tracker *orbx_tracker_create_mono(orbx_ctx *ctx, orbx_ext *ext /* max_batch >= S */, int S, const orbx_camera *cam,
                                       float th_frame, float th_map, float nnratio_map);
/* 2 (stereo tracker) or 1 (monocular tracker) */
int orbx_tracker_images_per_stream(const orbx_tracker *trk);

/* Visual-inertial TrackLocalMap (BASELINE config 3; src/Tracking.cc:2466-2490): with an inertial mode bound, the SECOND
 * pose optimisation of a step is Optimizer::PoseInertialOptimizationLastKeyFrame (mode 1, src/Optimizer.cc:7665-8066)
 * or PoseInertialOptimizationLastFrame (mode 2, :8068-8603) instead of PoseOptimization, on the same edges (close_pt =
 * bit 2 of orbx_track_map::map_flags, i.e. pMP->mTrackDepth < 10).  The frame's ImuCamPose is built on the device from
 * the pose the first PoseOptimization left, exactly as Frame::GetImuRotation / GetImuPosition do (src/Frame.cc:534-554,
 * float cv::Mat arithmetic), with the velocity and bias given here (pFrame->mVw, mImuBias after PredictStateIMU); the
 * optimised state goes back into the step's Tcw_out through Frame::SetImuPoseVelocity's arithmetic (src/Frame.cc:520-530).
 * Array layouts are those of orbx_pose_inertial_optimization_last_{keyframe,frame}_batch with P = S:
 *   Tcb, Tbc [16];  velocity [S][3];  bias [S][6] (gyro xyz, acc xyz);
 *   ref_state [S][21]: the last keyframe (mode 1) / the previous frame (mode 2);  preint [S][16];
 *   preint_jac [S][45], preint_bias [S][6], prior_state [S][21], prior_H [S][225]: mode 2 only;
 *   info_inertial [S][81], info_gyro [S][9], info_acc [S][9].
 * stats[6] (inliers_2) = nInitialCorrespondences - nBad, stats[7] counts the Gauss-Newton iterations. */
typedef struct orbx_track_imu {
  int mode;    /* 0 off, 1 LastKeyFrame, 2 LastFrame */
  int rec_init;
  const float *Tcb, *Tbc;
  const float *velocity, *bias;
  const double *ref_state, *preint, *preint_jac, *preint_bias;
  const double *info_inertial, *info_gyro, *info_acc;
  const double *prior_state, *prior_H;
} orbx_track_imu;
/* Bind DEVICE-resident inputs for the following steps (NULL / mode 0: back to the visual PoseOptimization); they are read
 * when the step runs and must stay valid until then. */
int orbx_tracker_set_inertial(orbx_tracker *trk, const orbx_track_imu *imu);
/* HOST-side inputs -> the next step's slot (asynchronous), then bound.  In mode 2, ref_state and (prior_state, prior_H)
 * may be NULL: the state and marginalised Hessian the PREVIOUS inertial step left on the device are used (the
 * reference's pFp->mpcpi chain, src/Optimizer.cc:8594-8599). */
int orbx_tracker_upload_inertial(orbx_tracker *trk, const orbx_track_imu *imu);
/* Results of the last inertial step: body states [S][21] (Rwb, twb, v, bg, ba) and the 15x15 Hessians [S][225] for the
 * next ConstraintPoseImu; either may be NULL.  Synchronises the tracker. */
int orbx_tracker_inertial_result(orbx_tracker *trk, double *state, double *H15);
/* Device addresses of those two arrays, for chaining them as ref_state / prior_state / prior_H with
 * orbx_tracker_set_inertial. */
const double *orbx_tracker_inertial_state_dev(orbx_tracker *trk);
const double *orbx_tracker_inertial_hessian_dev(orbx_tracker *trk);

orbx_tracker *orbx_tracker_create(orbx_ctx *ctx, orbx_ext *ext /* max_batch >= 2*S */, int S,
                                  const orbx_camera *cam, float th_frame, float th_map,
                                  float nnratio_map);
void orbx_tracker_destroy(orbx_tracker *trk);
/* images: [2*S][h][stride] device bytes, image 2s = left, 2s+1 = right of stream s; poses [S][16]
 * row-major float32 device arrays; stats [S][ORBX_TRACK_STATS] device int32.  Enqueues on the
 * extractor's stream and returns. */
int orbx_tracker_step_device(orbx_tracker *trk, const uint8_t *d_imgs, int w, int h, int stride,
                             const float *d_Tcw_true, const float *d_Tcw_prior, float *d_Tcw_out,
                             int32_t *d_stats);
/* Same through HOST buffers: copies the 2*S images in, the S poses and stats out, synchronises. */
int orbx_tracker_step(orbx_tracker *trk, const uint8_t *const *imgs, int w, int h, int stride,
                      const float *Tcw_true, const float *Tcw_prior, float *Tcw_out, int32_t *stats);
/* Asynchronous form of orbx_tracker_step for a host that keeps the device busy: submit() only ENQUEUES the step —
 * page-locked H2D of the 2*S images on a copy stream (it runs under the kernels of the previous step), the step
 * itself in overlap mode, the D2H of poses and statistics behind it — and returns; collect() blocks until the
 * OLDEST outstanding submit is complete and hands out its results.  At most two submits may be outstanding
 * (ORBX_ECAP otherwise).  The image memory must stay valid and unchanged until the matching collect() returns. */
int orbx_tracker_submit(orbx_tracker *trk, const uint8_t *const *imgs, int w, int h, int stride,
                        const float *Tcw_true, const float *Tcw_prior);
int orbx_tracker_collect(orbx_tracker *trk, float *Tcw_out, int32_t *stats);
/* CUDA graph of one step (single-frame latency path): with enable = 1, a step whose arguments (device pointers, image
 * geometry, map binding) repeat those of the previous call is captured once by stream capture of the same code path and
 * replayed with ONE graph launch from then on; any change of the arguments, overlap mode, profiling, keyframe work or an
 * inertial mode run eagerly.  Results are identical.  orbx_tracker_graph_launches counts the replays. */
int orbx_tracker_set_graph(orbx_tracker *trk, int enable);
long long orbx_tracker_graph_launches(const orbx_tracker *trk);

/* Overlap mode: step t's extraction + stereo matching run on the extractor's stream and its matching +
 * pose stages on a second stream, double-buffered, so the latency-bound fp64 optimisation of step t
 * overlaps the throughput-bound extraction of step t+1 (frames of different steps are independent
 * until Track() consumes them, exactly as Frame construction precedes Tracking::Track in the reference).
 * Results of a step are complete on orbx_tracker_result_stream(); orbx_tracker_synchronize() waits for both. */
int orbx_tracker_set_overlap(orbx_tracker *trk, int enable);
void *orbx_tracker_result_stream(orbx_tracker *trk);
int orbx_tracker_synchronize(orbx_tracker *trk);
/* The local map of every stream for the NEXT step(s), as flat DEVICE arrays with stride m_cap = orbx_tracker_map_capacity()
 * (>= 2048): what the shim would flatten from Tracking::mvpLocalMapPoints and mLastFrame (SURVEY.md §8(d) workload).
 * With a map set, SearchByProjection(Cur, Last) runs over the entries flagged in last_flags, TrackLocalMap's
 * SearchLocalPoints runs the full Frame::isInFrustum(pMP, 0.5) test (distance-invariance range, viewing angle,
 * MapPoint::PredictScale) over the entries flagged in map_flags that are not matched yet, and PoseOptimization's edges
 * take xw from here.  The arrays must stay valid until the steps using them have finished; every valid entry must have
 * bit1 (Observations() > 0) set.  NULL returns to the self-map harness of round 1. */
#define ORBX_TRACK_MAP_CAP 2048
typedef struct orbx_track_map {
  int32_t m_cap;
  const int32_t *n_map;       /* [S] */
  const float *xw;            /* [S][m_cap][3] GetWorldPos() */
  const uint8_t *desc;        /* [S][m_cap][32] GetDescriptor() */
  const uint8_t *last_flags;  /* [S][m_cap] bit0: mLastFrame.mvpMapPoints holds it && !mvbOutlier, bit1: Observations() > 0 */
  const int32_t *last_octave; /* [S][m_cap] octave / angle of the last frame's keypoint that observed it */
  const float *last_angle;
  const uint8_t *map_flags;   /* [S][m_cap] bit0: in mvpLocalMapPoints && !isBad(), bit1: Observations() > 0,
                                 bit2: mTrackDepth < 10 (read by the inertial optimisers only) */
  const float *max_dist, *min_dist; /* [S][m_cap] mfMaxDistance / mfMinDistance */
  const float *normal;        /* [S][m_cap][3] GetNormal() */
  float log_scale_factor;     /* Frame::mfLogScaleFactor; <= 0: log of the extractor's scale factor */
} orbx_track_map;
int orbx_tracker_map_capacity(const orbx_tracker *trk);
int orbx_tracker_set_map(orbx_tracker *trk, const orbx_track_map *map);
/* The same from HOST arrays, for the host-buffer entry points (orbx_tracker_step / _submit): call it right before the
 * step it belongs to; the arrays are copied into the tracker's own device buffer of that step's slot (page-locked
 * caller memory is DMA'd directly; with the asynchronous pipeline the copy runs on the copy stream under the previous
 * step's kernels).  orbx_tracker_map_bytes() = bytes copied per call. */
int orbx_tracker_upload_map(orbx_tracker *trk, const orbx_track_map *host_map);
size_t orbx_tracker_map_bytes(const orbx_tracker *trk);
/* Motion-model chaining (Tracking::TrackWithMotionModel, src/Tracking.cc:2354: SetPose(mVelocity * mLastFrame.mTcw)):
 * when enabled, the Tcw_prior argument of the step functions is the RELATIVE motion dT (last -> current, [S][16]) and
 * the prior used is dT * (pose the previous step produced), composed on the device; d_Tcw_init [S][16] (a device OR
 * host pointer) is the pose before the first chained step. */
int orbx_tracker_set_chain(orbx_tracker *trk, int enable, const float *d_Tcw_init);
/* Keyframe-rate work of the S streams (LocalMapping's share, which the reference runs in its own thread beside
 * Tracking, src/LocalMapping.cc:68-200): every `period`-th step the two prepared plans — the CreateNewMapPoints
 * searches (10 SearchForTriangulation calls per stream) and one LocalBundleAdjustment per stream — are enqueued on a
 * third CUDA stream of the lowest priority.  orbx_tracker_synchronize() waits for it as well.  NULL plans detach. */
int orbx_tracker_set_keyframe_work(orbx_tracker *trk, orbx_tri_batch *tri, orbx_lba_batch *lba, int period);
void *orbx_tracker_keyframe_stream(orbx_tracker *trk);
long long orbx_tracker_keyframe_runs(const orbx_tracker *trk);
/* Per-stage device time of the last step (CUDA events on the stream), ORBX_TRACK_STAGES entries:
 * 0 extract, 1 stereo match, 2 search-by-projection (last frame), 3 pose optimisation #1,
 * 4 search-by-projection (local map), 5 pose optimisation #2. */
#define ORBX_TRACK_STAGES 6
int orbx_tracker_set_profiling(orbx_tracker *trk, int enable);
int orbx_tracker_stage_ms(orbx_tracker *trk, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H_ */
