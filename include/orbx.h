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

/* Test/diagnostic view: FAST candidates of (image b, level) of the last call, i.e. the
 * contents of `vToDistributeKeys` (src/ORBextractor.cc:775,845-851) in an unspecified
 * order.  xy = [n][2] int16 (coordinates relative to minBorder), score = [n] uint8. */
int orbx_debug_candidates(orbx_ext *ext, int b, int level, int16_t *xy, uint8_t *score,
                          int cap, int *n_out);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H_ */
