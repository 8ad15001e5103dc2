// orbx_extract.cuh — device-visible parameter blocks of the batched ORB extractor.
#pragma once
#include "orbx_common.cuh"
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#define ORBX_EDGE 19          // EDGE_THRESHOLD, src/ORBextractor.cc:72
#define ORBX_MINB 16          // minBorderX = EDGE_THRESHOLD-3, src/ORBextractor.cc:771
#define ORBX_FAST_CELLS 8     // cells per FAST tile (one cell row x up to 8 cells)
#define ORBX_FAST_TP 240      // row pitch in bytes of the FAST shared-memory planes = TMA box width (compile-time: immediate offsets; a multiple of 16 for TMA; 60 words = -4 banks per row, so the 8 rows of a strip column hit 8 different banks: 256 B measured 5x the bank conflicts and 2.4 ms instead of 1.75)
#define ORBX_BLUR_TW 120   // output columns per warp tile of the blur kernel (30 words; 2 lanes carry the halo)
#define ORBX_BLUR_STRIP 32    // output rows per warp tile of the blur kernel

// One pyramid level of a batch of B equally sized images.
struct LevelParams {
  int w, h, pitch;                 // level size in pixels, row pitch in bytes
  size_t imgStride;                // bytes between consecutive images of the batch
  uint8_t* pyr;                    // un-blurred level (level 0 may alias the caller's input)
  uint8_t* blur;                   // 7x7 sigma-2 blurred level (own pitch: level 0 of `pyr` may be caller memory)
  int blurPitch;
  size_t blurStride;
  // FAST cell tiling (src/ORBextractor.cc:771-804)
  int nCols, nRows, wCell, hCell, maxBX, maxBY;
  int tileStart, tilesPerRow;      // flattened FAST tile ids of this level
  int fastCells;                   // cells per FAST tile (<= ORBX_FAST_CELLS, chosen so the TMA box is <= 256 B wide)
  int fastTP, fastTH;              // FAST shared-memory tile: row pitch in bytes (= TMA box width = ORBX_FAST_TP) and rows
  int useTma;                      // 1: the tile is fetched by one cp.async.bulk.tensor, 0: by 32-bit loads
  int blurTileStart, blurTilesX, blurTilesY;
  // resize tables (level l from l-1): offsets into ExtractParams::tab (int16 units)
  int tabX, tabY;
  int pyrSpan;                     // max source bytes (x0[3] + 1 - x0[0] + 1) any aligned group of 4 output columns reads
  // quadtree
  int nFeat, nIni;
  float hX;
  int candOfs, candCap;            // slice of the per-image candidate buffer
  int selOfs, selCap;              // slice of the per-image selected-keypoint buffer
  float scale, kpSize;
};

struct ExtractParams {
  int nlevels, B;
  int iniTh, minTh;
  int lap0, lap1;
  int candPerImage, selPerImage;   // per-image strides of cand / sel buffers
  int totalFastTiles, totalBlurTiles;
  int nodeCap;
  int fastTileBytes;               // bytes of one FAST shared-memory plane: ORBX_FAST_TP x (max hCell + 14) rows (8-row strips over-read)
  int fastScoreBytes;              // bytes of the FAST score plane: ORBX_FAST_TP x (max hCell + 2) rows
  int fastCandCap;                 // >= interior (evaluated) pixels of any FAST tile, multiple of 64
  const int16_t* tab;              // resize coefficient tables
  const int4* fastTiles;           // [totalFastTiles] host-built FAST tile records (see configure_geometry)
  uint32_t* cand;                  // [B][candPerImage] packed x | y<<12 | score<<24 (coords rel. to minBorder)
  int* candN;                      // [B][nlevels]
  uint16_t* keyNode;               // [B][candPerImage] quadtree scratch
  uint2* sel;                      // [B][selPerImage] packed selected keypoints
  int* selN;                       // [B][nlevels]
  int* selLap;                     // [B][nlevels] number of "lapping" keypoints per level
  int* err;                        // device error flag (capacity overflow)
  LevelParams lv[ORBX_MAX_LEVELS];
};

// One 3-D tensor map (x = bytes of a row, y = rows, z = image of the batch) per pyramid level, for the FAST tiles.
// Only the first ORBX_TMA_LEVELS levels get one (deeper levels use the 32-bit-load path): the maps are passed as the
// FIRST __grid_constant__ kernel parameter and, together with ExtractParams, must stay inside the classic 4 KB
// parameter window -- a descriptor that sits beyond it makes UTMALDG raise "illegal instruction" on sm_100a.
#define ORBX_TMA_LEVELS 8
struct alignas(128) FastTmaMaps {
  CUtensorMap m[ORBX_TMA_LEVELS];
};
static_assert(sizeof(FastTmaMaps) + sizeof(ExtractParams) <= 4096, "kernel parameters must fit the 4 KB window");
