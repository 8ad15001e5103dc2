"""orbx — Python host-side view of liborbx.so (hand-written sm_100a CUDA behind include/orbx.h).

The classes mirror the reference's operator interface for the tracking hot path
(ORBextractor / ORBmatcher / Optimizer) so parity tests read like the reference's call sites.
There is no CPU fallback: importing works anywhere, but creating a Context without the built
library or without a CUDA device raises.
"""
from .api import (Context, host_array, ORBextractor, KP_DTYPE, OrbxError, lib_path, load_library,
                  build_library, declared_symbols, ORBmatcher, Optimizer, Tracker, features_in_area, stereo_match, Frame,
                  Camera, make_camera, ORBVocabulary, search_by_bow, fuse, prepare_images, is_in_frustum, undistort_keypoints,
                  TriangulationBatch, LocalBABatch, ResidentFrame)  # noqa: F401
