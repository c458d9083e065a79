// Internal declarations shared by the liborbgpu.so translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/orbgpu.h"

namespace orbgpu {

constexpr int kEdge = 19;        // EDGE_THRESHOLD, ORBextractor.cc:74
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE, ORBextractor.cc:73
constexpr int kPatchSize = 31;   // PATCH_SIZE, ORBextractor.cc:72
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD - 3, ORBextractor.cc:740
constexpr int kMaxLevels = 16;
constexpr int kMaxRoots = 16;

// Device-side status word bits (OR-ed by kernels, checked on download).
enum : unsigned {
  kStatusCandOverflow = 1u,   // more FAST candidates than the per-level buffer holds
  kStatusNodeOverflow = 2u,   // octree node table too small (cannot happen by construction)
  kStatusKpOverflow = 4u,     // more keypoints than the per-image output capacity
};

// Geometry of one pyramid level; identical on host and device (passed by value in LevelTable).
struct LevelGeom {
  int w, h;                 // interior size, ORBextractor.cc:1044-1045
  int pitch;                // bytes per row of the padded plane (multiple of 16)
  int rows;                 // h + 2*kEdge
  unsigned long long planeBytes;   // pitch * rows rounded to 256
  unsigned long long pyrOffset;    // byte offset of this level's [batch] array in the pyramid buffer
  // FAST cell grid, ORBextractor.cc:740-754
  int maxBX, maxBY;         // w - 16, h - 16
  int nCols, nRows, wCell, hCell;
  // octree, ORBextractor.cc:535-554
  int quota;                // mnFeaturesPerLevel
  int nIni;
  float hX;
  int rootX0[kMaxRoots], rootX1[kMaxRoots];
  int maxNodes;             // node table size for this level
  // buffers
  int candCap;              // candidate rows per image
  unsigned long long candOffset;   // in uint2 units: level base in candidate buffer ([batch][candCap])
  int selCap;               // selected rows per image (quota + 3, padded)
  unsigned long long selOffset;    // in int4 units
  float scale, invScale;
  int patchSize;            // (int)(31 * scale), ORBextractor.cc:800
  // resize tables (level >= 1): offsets into the int2 table buffer
  int xTab, yTab;
  // TMA box (bytes x rows) of the kernel that stages this level in shared memory
  int rzBoxW, rzBoxH;       // source window of the tiled resize that PRODUCES this level (0 = generic kernel)
};

struct LevelTable {
  int nlevels;
  int iniTh, minTh;
  int kpCap;                // output keypoint rows per image
  LevelGeom lv[kMaxLevels];
};

struct Cand {   // one FAST candidate / selected keypoint: octree coordinates + response
  unsigned xy;  // x | y << 16
  unsigned resp;
};

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define ORBGPU_CUDA_CHECK(expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::orbgpu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                          __LINE__);                                                         \
      return ORBGPU_E_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

struct TmaMaps;   // tma.cuh

// ---- kernel launchers (each returns a cudaError_t from the launch) ----
struct ExtractBuffers {
  const LevelTable *dT;     // device copy of the level table
  uint8_t *pyr;             // all levels, [level][batch] padded planes
  uint8_t *blur;            // same layout
  const int2 *tabs;         // resize tables
  Cand *cand;               // [level][batch][candCap]
  int *candCount;           // [batch][kMaxLevels]
  unsigned short *nodeOf;   // octree scratch, same indexing as cand
  int4 *sel;                // [level][batch][selCap]: x_level, y_level, response, 0
  int *selCount;            // [batch][kMaxLevels]
  orbgpu_keypoint *kps;     // [batch][kpCap]
  uint8_t *desc;            // [batch][kpCap][32]
  int *kpCount;             // [batch]
  int *levelStart;          // [batch][kMaxLevels+1] prefix of selCount
  unsigned *status;         // one word
};

cudaError_t launch_import_level0(const LevelTable &T, const ExtractBuffers &B, const uint8_t *src,
                                 size_t rowStep, size_t imgStep, int imgBase, int batch, cudaStream_t s);
cudaError_t launch_resize_level(const LevelTable &T, const ExtractBuffers &B, int level, int imgBase, int batch,
                                cudaStream_t s);
void resize_box(const LevelTable &T, int l, int *boxW, int *boxH);
int resize_smem_bytes(const LevelTable &T);
cudaError_t resize_configure(int slotBytes);
void resize_make_tiles(const LevelTable &T, int level, std::vector<int4> &tiles);
cudaError_t launch_resize_tiles(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &maps, const int4 *tiles,
                                int ntiles, int slotBytes, int imgBase, int batch, cudaStream_t s);
cudaError_t launch_fill_borders(const LevelTable &T, const ExtractBuffers &B, int firstLevel, int imgBase, int batch, cudaStream_t s);
cudaError_t launch_blur(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &maps, const int4 *tiles, int ntiles,
                        int imgBase, int batch, cudaStream_t s);
int blur_box_w();
int blur_box_h();
int blur_smem_bytes();
cudaError_t blur_configure();
void blur_make_tiles(const LevelTable &T, std::vector<int4> &tiles);
#ifndef ORBGPU_FAST_WARPS
#define ORBGPU_FAST_WARPS 4          // independent warps per CTA (one strip each)
#endif
#ifndef ORBGPU_FAST_CTAS_PER_SM
#define ORBGPU_FAST_CTAS_PER_SM 3
#endif
#ifndef ORBGPU_FAST_PREFETCH
#define ORBGPU_FAST_PREFETCH 4       // rows of L1 prefetch ahead of the register loads
#endif
#ifndef ORBGPU_FAST_SHIFT
#define ORBGPU_FAST_SHIFT 1          // strips of four 32-pixel cells (2 bytes off the 4-byte grid) run shifted instead of as three cells
#endif
#ifndef ORBGPU_FAST_DEFER
#define ORBGPU_FAST_DEFER 1          // NMS survivors leave the row loop as one score word per lane and row; collected after the march
#endif
#ifndef ORBGPU_FAST_AHEAD
#define ORBGPU_FAST_AHEAD 1
#endif
struct FastConfig { int stageCap, rowsCap, smemBytes; };
int fast_cells_per_strip(int wCell);
FastConfig fast_make_config(const LevelTable &T);
bool fast_strip_shifted(const LevelGeom &G, int firstCell, int cells);
cudaError_t launch_fast(const LevelTable &T, const ExtractBuffers &B, const int4 *strips, int nPlain, int nShift, const FastConfig &c,
                        int imgBase, int batch, cudaStream_t s, cudaStream_t sShift);
cudaError_t launch_octree(const LevelTable &T, const ExtractBuffers &B, int imgBase, int batch, int smemBytes,
                          cudaStream_t s);
int hamming_dense_q(int nq);
cudaError_t launch_describe(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &pyrMaps, const TmaMaps &blurMaps, int imgBase, int batch,
                            cudaStream_t s);
cudaError_t describe_configure();
cudaError_t describe_configure_lat();
void describe_boxes(int *discW, int *discH, int *patchW, int *patchH);   // TMA boxes of the two staged neighbourhoods
cudaError_t launch_sincos_checksum(unsigned firstBits, unsigned long long count, int chunkLog2, unsigned long long *dSums, cudaStream_t s);
cudaError_t launch_sincos_values(unsigned firstBits, int n, float *dS, float *dC, cudaStream_t s);
cudaError_t describe_prepare();       // pattern table of the variant, once per device, outside stream capture
cudaError_t describe_prepare_lat();
// second thread shape of octree / describe / stereo for small batches (octree_lat.cu, describe_lat.cu, stereo_lat.cu)
constexpr int kLatencyBatch = 16;
cudaError_t launch_octree_lat(const LevelTable &T, const ExtractBuffers &B, int imgBase, int batch, int smemBytes, cudaStream_t s);
cudaError_t octree_configure_lat(int smemBytes);
cudaError_t launch_describe_lat(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &pyrMaps, const TmaMaps &blurMaps, int imgBase, int batch,
                                cudaStream_t s);
size_t octree_smem_bytes(int maxNodes);
cudaError_t octree_configure(int smemBytes);
cudaError_t fast_configure(int smemBytes);

struct StereoArgs {
  const uint8_t *pyrL, *pyrR;
  const orbgpu_keypoint *kpsL, *kpsR;
  const uint8_t *descL, *descR;
  const int *cntL, *cntR;
  int kpCapL, kpCapR;
  int firstL, firstR, pairs;
  float mbf, mb;
  float *uRight, *depth;    // [pairs][kpCapL] (indexed from 0, i.e. relative to firstL)
  int *bestSad;             // [pairs][kpCapL]
  int *rowStart;            // [pairs][rows + 1]: right keypoints bucketed by image row
  int4 *recs;               // [pairs][kpCapR]: the row-sorted right keypoints as (x, minr | maxr << 16, index, octave)
};
cudaError_t stereo_configure(int rows);
cudaError_t launch_stereo(const LevelTable &T, const LevelTable *dT, const StereoArgs &A, int maxKp, cudaStream_t s);
cudaError_t stereo_configure_lat(int rows);
cudaError_t launch_stereo_lat(const LevelTable &T, const LevelTable *dT, const StereoArgs &A, int maxKp, cudaStream_t s);

struct Top2Rec { int d1, d2; long long idx; };   // 16 bytes, the shard exchange unit
static_assert(sizeof(Top2Rec) == 16, "Top2Rec must be 16 bytes");

cudaError_t launch_hamming_dense(const uint8_t *bank, int rows, long long indexBase, const uint8_t *q,
                                 int nq, Top2Rec *partial, int chunks, Top2Rec *result, cudaStream_t s);
int hamming_dense_chunks(int rows, int nq, int smCount);
cudaError_t launch_top2_merge(const Top2Rec *parts, int nparts, int nq, Top2Rec *result, float nnratio,
                              int thLow, long long *match, cudaStream_t s);
// sharded bank: push the shard's merged records into every peer's window / wait for all shards and fold them
struct ExchangeView {
  int rank, world, maxQueries;
  char *const *peers;        // device array [world]: base of every rank's window as mapped in THIS process
  char *local;               // base of this rank's own window
  unsigned *done;            // last-CTA counter (local)
  unsigned *error;           // arrival timeout flag (local)
};
size_t exchange_window_bytes(int world, int maxQueries);
cudaError_t launch_top2_push(const Top2Rec *partial, int nparts, int nq, const ExchangeView &x, unsigned epoch, cudaStream_t s);
cudaError_t launch_top2_wait_merge(int nq, const ExchangeView &x, unsigned epoch, Top2Rec *result, float nnratio, int thLow,
                                   long long *match, cudaStream_t s);
cudaError_t launch_hamming_scan_only(const uint8_t *bank, int rows, long long indexBase, const uint8_t *q, int nq, Top2Rec *partial,
                                     int chunks, cudaStream_t s);
cudaError_t launch_hamming_csr(const uint8_t *q, int nq, const uint8_t *train, const int *offsets,
                               const int *cand, int *idx1, int *d1, int *d2, cudaStream_t s);
cudaError_t launch_pair_distance(const uint8_t *a, const uint8_t *b, int n, int *dist, cudaStream_t s);
cudaError_t run_popc_microbench(int smCount, double *popcPerSec, double *lop3PerSec);
cudaError_t run_pipe_probe(int smCount, int mode, double *opsPerSec);

}  // namespace orbgpu
