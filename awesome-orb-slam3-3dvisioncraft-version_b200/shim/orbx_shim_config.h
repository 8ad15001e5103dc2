// orbx_shim_config.h — which declarations the shim compiles against.
//
// In a real integration the shim TUs are added to the reference's CMakeLists.txt and see the
// reference's own, UNMODIFIED headers (include/ORBextractor.h, ORBmatcher.h, Optimizer.h, Frame.h, KeyFrame.h,
// MapPoint.h, Map.h, CameraModels/GeometricCamera.h) plus OpenCV 3.x.  OpenCV/Eigen/Boost are not installed in
// the build container, so `make -C shim check` compiles the same sources with -DORBX_SHIM_SYNTAX_CHECK against
// interface-equivalent stand-ins (stubs/ref_iface.h) — a syntax and type check, not a link test.
#pragma once
#ifdef ORBX_SHIM_SYNTAX_CHECK
#include "stubs/ref_iface.h"
#else
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Optimizer.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "Map.h"
#include "CameraModels/GeometricCamera.h"
#endif
#include "../../include/orbx.h"

namespace orbx_shim {
// One orbx_ctx per process (device 0 unless ORBX_DEVICE is set); created on first use.
orbx_ctx* context();
// Aborts with the library's error text: the reference's methods have no error channel and there is no CPU
// fallback to fall back to.
[[noreturn]] void die(const char* where, int status);
inline void check(const char* where, int status) { if (status != ORBX_OK) die(where, status); }
}  // namespace orbx_shim
