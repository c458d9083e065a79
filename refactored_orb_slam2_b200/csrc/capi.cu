// Host side of the C ABI declared in include/orbgpu.h: handle management, level geometry
// (ORBextractor::ORBextractor, ORBextractor.cc:407-464), buffer layout in HBM and kernel sequencing.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

static thread_local char g_error[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cv::resize's per-axis tables for INTER_LINEAR on 8-bit data (see pyramid.cu).
static void resize_axis_table(int srcN, int dstN, int2 *out) {
  const double inv_scale = (double)dstN / srcN;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dstN; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { f = 0; s = 0; }
    if (s >= srcN - 1) { f = 0; s = srcN - 1; }
    const int a0 = (int)lrintf((1.f - f) * 2048.f), a1 = (int)lrintf(f * 2048.f);
    out[d] = make_int2(s, a0 | (a1 << 16));
  }
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace orbgpu

using namespace orbgpu;

struct orbgpu_extractor {
  orbgpu_extractor_params p{};
  int device = 0, smCount = 0;
  double scaleFactor = 0;   // ORBextractor.h:91: the member is a double
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> quota;
  cudaStream_t stream = nullptr;

  // configured shape
  int cfgW = 0, cfgH = 0;
  LevelTable T{};
  LevelTable *dT = nullptr;
  int2 *dTabs = nullptr; size_t tabsCap = 0;
  int4 *dTiles = nullptr; size_t tilesCap = 0; int ntiles = 0, ntilesShift = 0;   // FAST strips: ntiles plain ones, then ntilesShift shifted ones
  FastConfig fastCfg{};
  TmaMaps blurMaps{}, rzMaps{}, descDiscMaps{}, descPatchMaps{};
  int4 *dRzTiles = nullptr; size_t rzTilesCap = 0; int rzTileOff[kMaxLevels] = {}, rzTileCnt[kMaxLevels] = {};
  int rzSlotBytes = 0;
  int4 *dBlurTiles = nullptr; size_t blurTilesCap = 0; int nBlurTiles = 0, nBlurTiles0 = 0;   // tiles are level-major: the first nBlurTiles0 are level 0
  int octSmem = 0;

  // buffers (sized for max_width x max_height x max_batch)
  uint8_t *dInput = nullptr; size_t inputBytes = 0;
  uint8_t *dPyr = nullptr, *dBlur = nullptr; size_t pyrBytes = 0;
  Cand *dCand = nullptr; unsigned short *dNodeOf = nullptr; size_t candRows = 0;
  int4 *dSel = nullptr; size_t selRows = 0;
  int *dCandCount = nullptr, *dSelCount = nullptr, *dKpCount = nullptr;
  orbgpu_keypoint *dKps = nullptr; uint8_t *dDesc = nullptr;
  unsigned *dStatus = nullptr;
  int kpCap = 0;
  // stereo results (owned by the LEFT handle)
  float *dURight = nullptr, *dDepth = nullptr; int *dBestSad = nullptr, *dRowStart = nullptr; int4 *dSorted = nullptr;
  cudaEvent_t evDone = nullptr;
  // the blur of a batch only needs its pyramid: it runs on a side stream next to FAST + octree and joins before describe
  cudaStream_t side = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evLevel0 = nullptr;
  int lastBatch = 0;
  // CUDA graph of the extraction kernel sequence for small batches (the frame-at-a-time path is launch bound: 13 kernels
  // and 2 memsets for one stereo pair); re-captured whenever its key changes
  struct ExtractGraph {
    cudaGraphExec_t exec = nullptr;
    const uint8_t *dImages = nullptr; int imgBase = -1, batch = 0, w = 0, h = 0; size_t rowStep = 0, imgStep = 0;
    int launches = 0;
  };
  static constexpr int kGraphSlots = 4;      // a 16-image host batch is two graphed chunks: one slot per (input, first image, count) key
  ExtractGraph graphs[kGraphSlots];
  int graphVictim = 0;
  // outputs of an enqueued, not yet awaited host-buffer batch (orbgpu_extract_batch_async .. orbgpu_wait)
  const int32_t *pendingCounts = nullptr; int pendingCount = 0, pendingCapacity = 0;
  // host-buffer pipeline: copy streams + per-chunk events
  static constexpr int kMaxChunks = 16;
  cudaStream_t copyIn = nullptr, copyOut = nullptr;
  cudaEvent_t evChunkIn[kMaxChunks] = {}, evChunkDone[kMaxChunks] = {};
  // optional per-stage timing
  static constexpr int kProfRing = 32;   // batches remembered between two orbgpu_stage_times calls
  bool profiling = false;
  cudaEvent_t evStage[kProfRing][ORBGPU_NUM_STAGES][2] = {};
  bool stageRan[kProfRing][ORBGPU_NUM_STAGES] = {};
  long profCount = 0;
  int profSlot = 0;

  ExtractBuffers buffers() const {
    ExtractBuffers B{};
    B.dT = dT; B.pyr = dPyr; B.blur = dBlur; B.tabs = dTabs; B.cand = dCand; B.candCount = dCandCount;
    B.nodeOf = dNodeOf; B.sel = dSel; B.selCount = dSelCount; B.kps = dKps; B.desc = dDesc; B.kpCount = dKpCount;
    B.levelStart = nullptr; B.status = dStatus;
    return B;
  }
};

struct orbgpu_bank {
  const uint8_t *d = nullptr;
  uint8_t *owned = nullptr;
  int rows = 0;
  long long indexBase = 0;
  int device = 0, smCount = 0;
  Top2Rec *partial = nullptr; size_t partialCap = 0;
  uint8_t *dQueries = nullptr; Top2Rec *dResult = nullptr; int qCap = 0;
  cudaStream_t stream = nullptr;
};

namespace {

// Level geometry for a w x h image.  `batchCap` sizes the per-level [batch] arrays.
int build_table(const orbgpu_extractor &e, int w, int h, int batchCap, LevelTable &T, std::vector<int2> &tabs,
                std::vector<int4> &tiles, size_t &pyrBytes, size_t &candRows, size_t &selRows) {
  const int nl = e.p.nlevels;
  std::memset(&T, 0, sizeof(T));
  T.nlevels = nl; T.iniTh = e.p.ini_th_fast; T.minTh = e.p.min_th_fast; T.kpCap = e.kpCap;
  tabs.clear(); tiles.clear();
  size_t pyrOff = 0, candOff = 0, selOff = 0;
  for (int l = 0; l < nl; l++) {
    LevelGeom &G = T.lv[l];
    G.w = (int)lrintf((float)w * e.invScale[l]);   // ORBextractor.cc:1044-1045
    G.h = (int)lrintf((float)h * e.invScale[l]);
    G.pitch = (int)round_up(G.w + 2 * kEdge, 16);
    G.rows = G.h + 2 * kEdge;
    G.planeBytes = round_up((size_t)G.pitch * G.rows, 256);
    G.pyrOffset = pyrOff;
    pyrOff += G.planeBytes * batchCap;
    G.maxBX = G.w - kMinBorder; G.maxBY = G.h - kMinBorder;
    const float width = (float)(G.maxBX - kMinBorder), height = (float)(G.maxBY - kMinBorder);
    if (width < 30.f || height < 30.f) {
      set_error("level %d of a %dx%d image is %dx%d: too small for the 30 px FAST cell grid (the reference "
                "divides by zero here, ORBextractor.cc:751-754)", l, w, h, G.w, G.h);
      return ORBGPU_E_UNSUPPORTED;
    }
    G.nCols = (int)(width / 30.f); G.nRows = (int)(height / 30.f);     // ORBextractor.cc:751-754
    G.wCell = (int)std::ceil(width / G.nCols); G.hCell = (int)std::ceil(height / G.nRows);
    G.quota = e.quota[l];
    G.nIni = (int)std::round(width / (G.maxBY - kMinBorder));          // ORBextractor.cc:535
    if (G.nIni < 1 || G.nIni > kMaxRoots) {
      set_error("level %d aspect ratio gives %d octree roots; the reference supports 1..%d (nIni == 0 is a "
                "division by zero there, ORBextractor.cc:535-537)", l, G.nIni, kMaxRoots);
      return ORBGPU_E_UNSUPPORTED;
    }
    G.hX = width / G.nIni;
    for (int i = 0; i < G.nIni; i++) {
      G.rootX0[i] = (int)(G.hX * (float)i);
      G.rootX1[i] = (int)(G.hX * (float)(i + 1));
    }
    G.maxNodes = std::max(G.quota, 4 * G.nIni) + 8;
    // NMS survivors are strict 8-neighbour maxima inside one cell's detection rectangle: at most ceil(cw / 2) * ceil(ch / 2)
    // per cell, so (W / 2 + nCols) * (H / 2 + nRows) bounds a level whatever the image (W x H = the cells' union)
    G.candCap = (int)round_up((size_t)((int)width / 2 + G.nCols + 1) * ((int)height / 2 + G.nRows + 1), 64);
    G.candOffset = candOff;
    candOff += (size_t)G.candCap * batchCap;
    G.selCap = (int)round_up(std::max(G.quota, 4 * G.nIni) + 4, 4);
    G.selOffset = selOff;
    selOff += (size_t)G.selCap * batchCap;
    G.scale = e.scale[l]; G.invScale = e.invScale[l];
    G.patchSize = (int)(kPatchSize * e.scale[l]);                     // ORBextractor.cc:800
    if (l > 0) {
      resize_box(T, l, &G.rzBoxW, &G.rzBoxH);
      if (G.rzBoxW > 256 || G.rzBoxH > 256) G.rzBoxW = G.rzBoxH = 0;   // large scale factors: generic kernel
    }
    if (l > 0) {
      const LevelGeom &P = T.lv[l - 1];
      G.xTab = (int)tabs.size(); tabs.resize(tabs.size() + G.w);
      resize_axis_table(P.w, G.w, tabs.data() + G.xTab);
      G.yTab = (int)tabs.size(); tabs.resize(tabs.size() + G.h);
      resize_axis_table(P.h, G.h, tabs.data() + G.yTab);
    }
    for (int ci = 0; ci < G.nRows; ci++) {
      if (kEdge + ci * G.hCell >= G.h - kEdge) continue;               // empty detection rows
      const int per = fast_cells_per_strip(G.wCell);                   // FAST strips: <= 128 pixels of whole cells, one warp each
      for (int cj = 0; cj < G.nCols; cj += per) {
        if (kEdge + cj * G.wCell >= G.w - kEdge) continue;
        tiles.push_back(make_int4(l, ci, cj, std::min(per, G.nCols - cj)));
      }
    }
  }
  pyrBytes = pyrOff; candRows = candOff; selRows = selOff;
  return ORBGPU_OK;
}

int configure(orbgpu_extractor *e, int w, int h) {
  if (w == e->cfgW && h == e->cfgH) return ORBGPU_OK;
  if (w > e->p.max_width || h > e->p.max_height) {
    set_error("image %dx%d exceeds the handle's max %dx%d", w, h, e->p.max_width, e->p.max_height);
    return ORBGPU_E_CAPACITY;
  }
  LevelTable T;
  std::vector<int2> tabs;
  std::vector<int4> tiles;
  size_t pyrBytes, candRows, selRows;
  int rc = build_table(*e, w, h, e->p.max_batch, T, tabs, tiles, pyrBytes, candRows, selRows);
  if (rc) return rc;
  if (pyrBytes > e->pyrBytes || candRows > e->candRows || selRows > e->selRows) {
    set_error("internal: buffers sized for %dx%d do not hold %dx%d", e->p.max_width, e->p.max_height, w, h);
    return ORBGPU_E_CAPACITY;
  }
  int maxNodes = 0;
  for (int l = 0; l < T.nlevels; l++) maxNodes = std::max(maxNodes, T.lv[l].maxNodes);
  const int octSmem = (int)octree_smem_bytes(maxNodes);
  const FastConfig fc = fast_make_config(T);
  if (octSmem > 200 * 1024 || fc.smemBytes > 200 * 1024 || maxNodes > 65000) {
    set_error("nfeatures too large for the shared-memory octree (%d nodes, %d bytes)", maxNodes, octSmem);
    return ORBGPU_E_UNSUPPORTED;
  }
  for (int l = 0; l < T.nlevels; l++)
    if (T.lv[l].wCell > 124 || T.lv[l].hCell > 63) {      // cannot happen: a cell is 30 .. 59 pixels by construction (ORBextractor.cc:751-754)
      set_error("internal: FAST cell %dx%d does not fit a strip", T.lv[l].wCell, T.lv[l].hCell);
      return ORBGPU_E_UNSUPPORTED;
    }
  // FAST strips that run in the shifted form go to the end of the list: they are a launch of their own
  const int nPlainStrips = (int)(std::stable_partition(tiles.begin(), tiles.end(), [&](const int4 &t) { return !fast_strip_shifted(T.lv[t.x], t.z, t.w); }) -
                                 tiles.begin());
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  if (tabs.size() > e->tabsCap) {
    if (e->dTabs) cudaFree(e->dTabs);
    e->tabsCap = tabs.size() + 1024;
    ORBGPU_CUDA_CHECK(cudaMalloc(&e->dTabs, e->tabsCap * sizeof(int2)));
  }
  std::vector<int4> rzTiles, one;
  int rzOff[kMaxLevels] = {}, rzCnt[kMaxLevels] = {};
  TmaMaps rzMaps;
  for (int l = 1; l < T.nlevels; l++) {
    const LevelGeom &G = T.lv[l], &P = T.lv[l - 1];
    if (!G.rzBoxW) continue;
    resize_make_tiles(T, l, one);
    rzOff[l] = (int)rzTiles.size(); rzCnt[l] = (int)one.size();
    rzTiles.insert(rzTiles.end(), one.begin(), one.end());
    ORBGPU_CUDA_CHECK(encode_plane_map(&rzMaps.m[l - 1], e->dPyr + P.pyrOffset, P.pitch, P.rows, P.planeBytes, e->p.max_batch,
                                       G.rzBoxW, G.rzBoxH));
  }
  const int rzSlot = resize_smem_bytes(T);
  if (!rzTiles.empty()) {
    if (rzTiles.size() > e->rzTilesCap) {
      if (e->dRzTiles) cudaFree(e->dRzTiles);
      e->rzTilesCap = rzTiles.size() + 64;
      ORBGPU_CUDA_CHECK(cudaMalloc(&e->dRzTiles, e->rzTilesCap * sizeof(int4)));
    }
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(e->dRzTiles, rzTiles.data(), rzTiles.size() * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
    ORBGPU_CUDA_CHECK(resize_configure(rzSlot));
  }
  std::vector<int4> blurTiles;
  blur_make_tiles(T, blurTiles);
  TmaMaps blurMaps;
  for (int l = 0; l < T.nlevels; l++) {
    const LevelGeom &G = T.lv[l];
    ORBGPU_CUDA_CHECK(encode_plane_map(&blurMaps.m[l], e->dPyr + G.pyrOffset, G.pitch, G.rows, G.planeBytes, e->p.max_batch,
                                       blur_box_w(), blur_box_h()));
  }
  TmaMaps discMaps, patchMaps;         // describe: 48 x 31 boxes on the pyramid levels, 64 x 37 boxes on the blurred levels
  {
    int dw, dh, pw, ph;
    describe_boxes(&dw, &dh, &pw, &ph);
    for (int l = 0; l < T.nlevels; l++) {
      const LevelGeom &G = T.lv[l];
      ORBGPU_CUDA_CHECK(encode_plane_map(&discMaps.m[l], e->dPyr + G.pyrOffset, G.pitch, G.rows, G.planeBytes, e->p.max_batch, dw, dh));
      ORBGPU_CUDA_CHECK(encode_plane_map(&patchMaps.m[l], e->dBlur + G.pyrOffset, G.pitch, G.rows, G.planeBytes, e->p.max_batch, pw, ph));
    }
    ORBGPU_CUDA_CHECK(describe_configure());
    ORBGPU_CUDA_CHECK(describe_configure_lat());
  }
  if (blurTiles.size() > e->blurTilesCap) {
    if (e->dBlurTiles) cudaFree(e->dBlurTiles);
    e->blurTilesCap = blurTiles.size() + 64;
    ORBGPU_CUDA_CHECK(cudaMalloc(&e->dBlurTiles, e->blurTilesCap * sizeof(int4)));
  }
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(e->dBlurTiles, blurTiles.data(), blurTiles.size() * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
  ORBGPU_CUDA_CHECK(blur_configure());
  if (tiles.size() > e->tilesCap) {
    if (e->dTiles) cudaFree(e->dTiles);
    e->tilesCap = tiles.size() + 256;
    ORBGPU_CUDA_CHECK(cudaMalloc(&e->dTiles, e->tilesCap * sizeof(int4)));
  }
  if (!tabs.empty())
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(e->dTabs, tabs.data(), tabs.size() * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(e->dTiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(e->dT, &T, sizeof(T), cudaMemcpyHostToDevice, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  {
    // the opt-in shared-memory limit is a per-function, per-device attribute shared by all handles: only raise it
    static std::mutex mu;
    static int octMax[64] = {0}, fastMax[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    const int d = e->device & 63;
    octMax[d] = std::max(octMax[d], octSmem);
    fastMax[d] = std::max(fastMax[d], fc.smemBytes);
    ORBGPU_CUDA_CHECK(octree_configure(octMax[d]));
    ORBGPU_CUDA_CHECK(octree_configure_lat(octMax[d]));
    ORBGPU_CUDA_CHECK(fast_configure(fastMax[d]));
  }
  e->T = T; e->ntiles = nPlainStrips; e->ntilesShift = (int)tiles.size() - nPlainStrips; e->fastCfg = fc; e->octSmem = octSmem;
  e->blurMaps = blurMaps; e->nBlurTiles = (int)blurTiles.size();
  e->nBlurTiles0 = (int)std::count_if(blurTiles.begin(), blurTiles.end(), [](const int4 &t) { return t.x == 0; });
  e->descDiscMaps = discMaps; e->descPatchMaps = patchMaps;
  e->rzMaps = rzMaps; e->rzSlotBytes = rzSlot;
  for (int l = 0; l < kMaxLevels; l++) { e->rzTileOff[l] = rzOff[l]; e->rzTileCnt[l] = rzCnt[l]; }
  e->cfgW = w; e->cfgH = h;
  return ORBGPU_OK;
}

struct StageTimer {
  orbgpu_extractor *e; int stage;
  StageTimer(orbgpu_extractor *e_, int stage_) : e(e_), stage(stage_) {
    if (e->profiling) { cudaEventRecord(e->evStage[e->profSlot][stage][0], e->stream); e->stageRan[e->profSlot][stage] = true; }
  }
  ~StageTimer() { if (e->profiling) cudaEventRecord(e->evStage[e->profSlot][stage][1], e->stream); }
};

// The whole ORBextractor::operator() for images [imgBase, imgBase + batch) of a device-resident batch whose image 0
// starts at dImages; asynchronous on e->stream.
int run_extract_range(orbgpu_extractor *e, const uint8_t *dImages, int imgBase, int batch, size_t rowStep, size_t imgStep) {
  const ExtractBuffers B = e->buffers();
  cudaStream_t s = e->stream;
  ORBGPU_CUDA_CHECK(cudaMemsetAsync(e->dCandCount + (size_t)imgBase * kMaxLevels, 0, sizeof(int) * kMaxLevels * batch, s));
  ORBGPU_CUDA_CHECK(cudaMemsetAsync(e->dSelCount + (size_t)imgBase * kMaxLevels, 0, sizeof(int) * kMaxLevels * batch, s));
  // Two streams per batch: the blur only needs the pyramid, the descriptor kernel needs the blur and the octree.  FAST fills
  // every SM on its own (ALU pipe, 128+ registers per thread), so a blur launched next to it only queues behind or in front
  // of it; the octree, however, is a latency-bound kernel of batch x levels CTAs that leaves most of the GPU idle.  The blur
  // therefore starts on the side stream when FAST has finished, runs next to the octree, and joins before the descriptors.
  // With per-stage timing on, everything stays on one stream so that the stage timers remain meaningful.
  static const bool noOverlap = getenv("ORBGPU_NO_OVERLAP") != nullptr;
  const bool overlap = !noOverlap && !e->profiling;
  const bool lat = batch <= kLatencyBatch;       // frame-at-a-time calls: few CTAs, give each more threads
  auto pyramid_rest = [&](cudaStream_t q) -> int {
    for (int l = 1; l < e->T.nlevels; l++) {
      if (e->T.lv[l].rzBoxW)
        ORBGPU_CUDA_CHECK(launch_resize_tiles(e->T, B, e->rzMaps, e->dRzTiles + e->rzTileOff[l], e->rzTileCnt[l], e->rzSlotBytes, imgBase, batch, q));
      else
        ORBGPU_CUDA_CHECK(launch_resize_level(e->T, B, l, imgBase, batch, q));
    }
    ORBGPU_CUDA_CHECK(launch_fill_borders(e->T, B, 1, imgBase, batch, q));
    return ORBGPU_OK;
  };
  if (overlap) {
    // Level 0 is final after the import: its blur (46 % of the blur's pixels) starts at once on the side stream and runs next to the
    // seven dependent resize launches, whose upper levels are too small to fill the GPU; the blur of levels 1 .. 7 follows on the
    // side stream when FAST has finished (next to the octree) and joins before the descriptors.
    static const bool noSplit = getenv("ORBGPU_NO_BLUR_SPLIT") != nullptr;
    const int n0 = noSplit || lat || e->T.nlevels < 2 ? 0 : e->nBlurTiles0;      // frame-at-a-time calls are launch bound: one blur launch
    ORBGPU_CUDA_CHECK(launch_import_level0(e->T, B, dImages, rowStep, imgStep, imgBase, batch, s));
    if (n0 > 0) {
      ORBGPU_CUDA_CHECK(cudaEventRecord(e->evLevel0, s));
      ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->side, e->evLevel0, 0));
      ORBGPU_CUDA_CHECK(launch_blur(e->T, B, e->blurMaps, e->dBlurTiles, n0, imgBase, batch, e->side));
    }
    int rc = pyramid_rest(s);
    if (rc) return rc;
    if (lat && e->ntiles > 0 && e->ntilesShift > 0) {
      // frame-at-a-time: the plain and the shifted FAST launch run side by side (each is one ~8 us march deep whatever its size)
      ORBGPU_CUDA_CHECK(cudaEventRecord(e->evLevel0, s));
      ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->side, e->evLevel0, 0));
      ORBGPU_CUDA_CHECK(launch_fast(e->T, B, e->dTiles, e->ntiles, e->ntilesShift, e->fastCfg, imgBase, batch, s, e->side));
      ORBGPU_CUDA_CHECK(cudaEventRecord(e->evJoin, e->side));
      ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(s, e->evJoin, 0));
    } else {
      ORBGPU_CUDA_CHECK(launch_fast(e->T, B, e->dTiles, e->ntiles, e->ntilesShift, e->fastCfg, imgBase, batch, s, s));
    }
    ORBGPU_CUDA_CHECK(cudaEventRecord(e->evFork, s));
    ORBGPU_CUDA_CHECK(lat ? launch_octree_lat(e->T, B, imgBase, batch, e->octSmem, s) : launch_octree(e->T, B, imgBase, batch, e->octSmem, s));
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->side, e->evFork, 0));
    if (e->nBlurTiles > n0)
      ORBGPU_CUDA_CHECK(launch_blur(e->T, B, e->blurMaps, e->dBlurTiles + n0, e->nBlurTiles - n0, imgBase, batch, e->side));
    ORBGPU_CUDA_CHECK(cudaEventRecord(e->evJoin, e->side));
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(s, e->evJoin, 0));
    ORBGPU_CUDA_CHECK(lat ? launch_describe_lat(e->T, B, e->descDiscMaps, e->descPatchMaps, imgBase, batch, s) : launch_describe(e->T, B, e->descDiscMaps, e->descPatchMaps, imgBase, batch, s));
    return ORBGPU_OK;
  }
  {
    StageTimer t(e, 0);
    ORBGPU_CUDA_CHECK(launch_import_level0(e->T, B, dImages, rowStep, imgStep, imgBase, batch, s));
    int rc = pyramid_rest(s);
    if (rc) return rc;
  }
  { StageTimer t(e, 1); ORBGPU_CUDA_CHECK(launch_fast(e->T, B, e->dTiles, e->ntiles, e->ntilesShift, e->fastCfg, imgBase, batch, s, s)); }
  { StageTimer t(e, 2); ORBGPU_CUDA_CHECK(lat ? launch_octree_lat(e->T, B, imgBase, batch, e->octSmem, s) : launch_octree(e->T, B, imgBase, batch, e->octSmem, s)); }
  { StageTimer t(e, 3); ORBGPU_CUDA_CHECK(launch_blur(e->T, B, e->blurMaps, e->dBlurTiles, e->nBlurTiles, imgBase, batch, s)); }
  { StageTimer t(e, 4); ORBGPU_CUDA_CHECK(lat ? launch_describe_lat(e->T, B, e->descDiscMaps, e->descPatchMaps, imgBase, batch, s) : launch_describe(e->T, B, e->descDiscMaps, e->descPatchMaps, imgBase, batch, s)); }
  return ORBGPU_OK;
}

// Small batches replay a captured CUDA graph of run_extract_range instead of re-launching its kernels one by one.
constexpr int kGraphMaxBatch = 8;
int run_extract_range_graphed(orbgpu_extractor *e, const uint8_t *dImages, int imgBase, int batch, size_t rowStep, size_t imgStep) {
  static const bool disabled = getenv("ORBGPU_NO_GRAPH") != nullptr;
  if (disabled || e->profiling || batch > kGraphMaxBatch) return run_extract_range(e, dImages, imgBase, batch, rowStep, imgStep);
  orbgpu_extractor::ExtractGraph *slot = nullptr;
  for (auto &g : e->graphs)
    if (g.exec && g.dImages == dImages && g.imgBase == imgBase && g.batch == batch && g.w == e->cfgW && g.h == e->cfgH && g.rowStep == rowStep &&
        g.imgStep == imgStep) { slot = &g; break; }
  const bool hit = slot != nullptr;
  if (!hit) {
    for (auto &g : e->graphs) if (!g.exec) { slot = &g; break; }
    if (!slot) { slot = &e->graphs[e->graphVictim]; e->graphVictim = (e->graphVictim + 1) % orbgpu_extractor::kGraphSlots; }
  }
  orbgpu_extractor::ExtractGraph &G = *slot;
  if (!hit) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    const long long before = g_launches;
    if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return run_extract_range(e, dImages, imgBase, batch, rowStep, imgStep);
    }
    const int rc = run_extract_range(e, dImages, imgBase, batch, rowStep, imgStep);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    const int captured = (int)(g_launches - before);
    g_launches = before;                                   // nothing has run yet
    if (rc != ORBGPU_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&G.exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      G.exec = nullptr;
      cudaGetLastError();
      return run_extract_range(e, dImages, imgBase, batch, rowStep, imgStep);
    }
    cudaGraphDestroy(graph);
    G.dImages = dImages; G.imgBase = imgBase; G.batch = batch; G.w = e->cfgW; G.h = e->cfgH; G.rowStep = rowStep; G.imgStep = imgStep;
    G.launches = captured;
  }
  ORBGPU_CUDA_CHECK(cudaGraphLaunch(G.exec, e->stream));
  count_launch(G.launches);
  return ORBGPU_OK;
}

void begin_batch(orbgpu_extractor *e) {
  if (e->profiling) {
    e->profSlot = (int)(e->profCount++ % orbgpu_extractor::kProfRing);
    for (bool &r : e->stageRan[e->profSlot]) r = false;
  }
}

int run_extract(orbgpu_extractor *e, const uint8_t *dImages, int batch, int w, int h, size_t rowStep, size_t imgStep) {
  int rc = configure(e, w, h);
  if (rc) return rc;
  begin_batch(e);
  ORBGPU_CUDA_CHECK(cudaMemsetAsync(e->dStatus, 0, sizeof(unsigned), e->stream));
  rc = run_extract_range_graphed(e, dImages, 0, batch, rowStep, imgStep);
  if (rc) return rc;
  e->lastBatch = batch;
  return ORBGPU_OK;
}

int check_status(orbgpu_extractor *e) {
  unsigned st = 0;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(&st, e->dStatus, sizeof(st), cudaMemcpyDeviceToHost, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  if (st) {
    set_error("device buffers overflowed (status 0x%x:%s%s%s)", st, st & kStatusCandOverflow ? " FAST candidates" : "",
              st & kStatusNodeOverflow ? " octree nodes" : "", st & kStatusKpOverflow ? " keypoints" : "");
    return ORBGPU_E_CAPACITY;
  }
  return ORBGPU_OK;
}

int check_image(orbgpu_extractor *e, int image, int level) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  if (image < 0 || image >= e->lastBatch || level < 0 || level >= e->T.nlevels) {
    set_error("image %d / level %d out of range (last batch %d, levels %d)", image, level, e->lastBatch, e->T.nlevels);
    return ORBGPU_E_INVALID;
  }
  return ORBGPU_OK;
}

}  // namespace

extern "C" {

int orbgpu_version(void) { return ORBGPU_VERSION; }
const char *orbgpu_last_error(void) { return g_error; }
int64_t orbgpu_launch_count(int reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}
int orbgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int orbgpu_extractor_create(const orbgpu_extractor_params *params, orbgpu_extractor **out) {
  if (!params || !out) { set_error("null argument"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  const orbgpu_extractor_params &p = *params;
  if (p.nfeatures < 1 || p.nlevels < 1 || p.nlevels > kMaxLevels || !(p.scale_factor > 1.0f) || p.max_width < 1 ||
      p.max_height < 1 || p.max_batch < 1 || p.max_batch > 65535 || p.min_th_fast < 0 || p.ini_th_fast < p.min_th_fast ||
      p.ini_th_fast > 254) {
    set_error("invalid extractor parameters");
    return ORBGPU_E_INVALID;
  }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  orbgpu_extractor *e = new orbgpu_extractor();
  e->p = p;
  if (p.device >= 0) e->device = p.device; else ORBGPU_CUDA_CHECK(cudaGetDevice(&e->device));
  DeviceGuard guard(e->device);
  if (!guard.ok) { set_error("cannot select CUDA device %d", e->device); delete e; return ORBGPU_E_CUDA; }
  cudaDeviceProp prop;
  ORBGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, e->device));
  e->smCount = prop.multiProcessorCount;

  // ORBextractor.cc:411-441 (float arithmetic in the reference's order; scaleFactor is a double member)
  const int nl = p.nlevels;
  e->scaleFactor = (double)p.scale_factor;
  e->scale.assign(nl, 1.f); e->sigma2.assign(nl, 1.f); e->invScale.assign(nl, 1.f); e->invSigma2.assign(nl, 1.f);
  for (int i = 1; i < nl; i++) {
    e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
    e->sigma2[i] = e->scale[i] * e->scale[i];
  }
  for (int i = 0; i < nl; i++) { e->invScale[i] = 1.0f / e->scale[i]; e->invSigma2[i] = 1.0f / e->sigma2[i]; }
  e->quota.assign(nl, 0);
  const float factor = (float)(1.0f / e->scaleFactor);
  float want = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; l++) {
    e->quota[l] = (int)lrintf(want);
    sum += e->quota[l];
    want *= factor;
  }
  e->quota[nl - 1] = std::max(p.nfeatures - sum, 0);
  int cap = 0;
  // a level returns at most max(quota + 3, 4 * roots) keypoints: the octree stops within one split of its quota, but never
  // before the first round has split every root (ORBextractor.cc:566-745), and a level has up to kMaxRoots roots
  for (int l = 0; l < nl; l++) cap += std::max(e->quota[l] + 3, 4 * kMaxRoots);
  e->kpCap = (int)round_up(cap, 8);

  ORBGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&e->evDone, cudaEventDisableTiming));
  ORBGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  ORBGPU_CUDA_CHECK(describe_prepare());
  ORBGPU_CUDA_CHECK(describe_prepare_lat());
  ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming));
  ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&e->evJoin, cudaEventDisableTiming));
  ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&e->evLevel0, cudaEventDisableTiming));
  LevelTable T;
  std::vector<int2> tabs;
  std::vector<int4> tiles;
  int rc = build_table(*e, p.max_width, p.max_height, p.max_batch, T, tabs, tiles, e->pyrBytes, e->candRows, e->selRows);
  if (rc) { orbgpu_extractor_destroy(e); return rc; }
  e->inputBytes = (size_t)p.max_width * p.max_height * p.max_batch;
  const size_t B = p.max_batch;
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dT, sizeof(LevelTable)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dInput, e->inputBytes));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dPyr, e->pyrBytes));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dBlur, e->pyrBytes));
  ORBGPU_CUDA_CHECK(cudaMemset(e->dBlur, 0, e->pyrBytes));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dCand, e->candRows * sizeof(Cand)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dNodeOf, e->candRows * sizeof(unsigned short)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dSel, e->selRows * sizeof(int4)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dCandCount, B * kMaxLevels * sizeof(int)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dSelCount, B * kMaxLevels * sizeof(int)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dKpCount, B * sizeof(int)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dKps, B * e->kpCap * sizeof(orbgpu_keypoint)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dDesc, B * e->kpCap * 32));
  ORBGPU_CUDA_CHECK(cudaMalloc(&e->dStatus, sizeof(unsigned)));
  ORBGPU_CUDA_CHECK(cudaMemset(e->dStatus, 0, sizeof(unsigned)));
  ORBGPU_CUDA_CHECK(cudaMemset(e->dKpCount, 0, B * sizeof(int)));
  rc = configure(e, p.max_width, p.max_height);
  if (rc) { orbgpu_extractor_destroy(e); return rc; }
  *out = e;
  return ORBGPU_OK;
}

void orbgpu_extractor_destroy(orbgpu_extractor *e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  void *ptrs[] = {e->dT, e->dTabs, e->dTiles, e->dBlurTiles, e->dRzTiles, e->dInput, e->dPyr, e->dBlur, e->dCand, e->dNodeOf, e->dSel, e->dCandCount,
                  e->dSelCount, e->dKpCount, e->dKps, e->dDesc, e->dStatus, e->dURight, e->dDepth, e->dBestSad, e->dRowStart, e->dSorted};
  for (void *q : ptrs) if (q) cudaFree(q);
  for (auto &g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->evDone) cudaEventDestroy(e->evDone);
  if (e->evFork) cudaEventDestroy(e->evFork);
  if (e->evJoin) cudaEventDestroy(e->evJoin);
  if (e->evLevel0) cudaEventDestroy(e->evLevel0);
  if (e->side) cudaStreamDestroy(e->side);
  for (cudaEvent_t ev : e->evChunkIn) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->evChunkDone) if (ev) cudaEventDestroy(ev);
  if (e->copyIn) cudaStreamDestroy(e->copyIn);
  if (e->copyOut) cudaStreamDestroy(e->copyOut);
  for (auto &slot : e->evStage) for (auto &p : slot) for (cudaEvent_t ev : p) if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int orbgpu_extractor_tables(const orbgpu_extractor *e, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                            int32_t *features_per_level) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  for (int i = 0; i < e->p.nlevels; i++) {
    if (scale) scale[i] = e->scale[i];
    if (inv_scale) inv_scale[i] = e->invScale[i];
    if (sigma2) sigma2[i] = e->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = e->invSigma2[i];
    if (features_per_level) features_per_level[i] = e->quota[i];
  }
  return ORBGPU_OK;
}

int orbgpu_set_profiling(orbgpu_extractor *e, int enable) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  if (enable && !e->evStage[0][0][0])
    for (auto &slot : e->evStage) for (auto &p : slot) for (cudaEvent_t &ev : p) ORBGPU_CUDA_CHECK(cudaEventCreate(&ev));
  e->profiling = enable != 0;
  e->profCount = 0;
  return ORBGPU_OK;
}

int orbgpu_stage_times(orbgpu_extractor *e, float *milliseconds, int n) {
  if (!e || !milliseconds) { set_error("null argument"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  const int slots = (int)std::min<long>(e->profCount, orbgpu_extractor::kProfRing);
  for (int i = 0; i < n; i++) {
    milliseconds[i] = 0.f;
    if (i >= ORBGPU_NUM_STAGES || !e->profiling) continue;
    double sum = 0;
    int cnt = 0;
    for (int k = 0; k < slots; k++) {
      if (!e->stageRan[k][i]) continue;
      float ms = 0.f;
      ORBGPU_CUDA_CHECK(cudaEventElapsedTime(&ms, e->evStage[k][i][0], e->evStage[k][i][1]));
      sum += ms; cnt++;
    }
    if (cnt) milliseconds[i] = (float)(sum / cnt);
  }
  e->profCount = 0;   // the next batch starts a new averaging window
  for (auto &slot : e->stageRan) for (bool &r : slot) r = false;
  return ORBGPU_OK;
}

const char *orbgpu_stage_name(int stage) {
  static const char *names[ORBGPU_NUM_STAGES] = {"pyramid", "fast", "octree", "blur", "describe", "stereo"};
  return stage >= 0 && stage < ORBGPU_NUM_STAGES ? names[stage] : "";
}

int orbgpu_extractor_capacity(const orbgpu_extractor *e) { return e ? e->kpCap : ORBGPU_E_INVALID; }
void *orbgpu_extractor_stream(orbgpu_extractor *e) { return e ? (void *)e->stream : nullptr; }

int orbgpu_extract_batch_device(orbgpu_extractor *e, const uint8_t *d_images, int count, int width, int height,
                                size_t row_step, size_t image_step) {
  if (!e || !d_images) { set_error("null argument"); return ORBGPU_E_INVALID; }
  if (count < 1 || count > e->p.max_batch) { set_error("batch %d outside 1..%d", count, e->p.max_batch); return ORBGPU_E_CAPACITY; }
  if (width < 1 || height < 1 || row_step < (size_t)width) { set_error("bad image shape"); return ORBGPU_E_INVALID; }
  if (e->pendingCounts) { set_error("a batch enqueued with orbgpu_extract_batch_async has not been awaited (orbgpu_wait)"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  return run_extract(e, d_images, count, width, height, row_step, image_step);
}

int orbgpu_results_device(orbgpu_extractor *e, const orbgpu_keypoint **d_keypoints, const uint8_t **d_descriptors,
                          const int32_t **d_counts, int *capacity) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  if (d_keypoints) *d_keypoints = e->dKps;
  if (d_descriptors) *d_descriptors = e->dDesc;
  if (d_counts) *d_counts = e->dKpCount;
  if (capacity) *capacity = e->kpCap;
  return ORBGPU_OK;
}

int orbgpu_synchronize(orbgpu_extractor *e) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  return check_status(e);
}

int orbgpu_download_results(orbgpu_extractor *e, orbgpu_keypoint *keypoints, uint8_t *descriptors, int capacity,
                            int32_t *counts) {
  if (!e || !counts) { set_error("null argument"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  const int n = e->lastBatch;
  if (n < 1) { set_error("no batch has been extracted"); return ORBGPU_E_INVALID; }
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(counts, e->dKpCount, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
  int rc = check_status(e);   // synchronises
  if (rc) return rc;
  int maxCount = 0;
  for (int i = 0; i < n; i++) maxCount = std::max(maxCount, counts[i]);
  if (maxCount > capacity) { set_error("capacity %d < %d keypoints", capacity, maxCount); return ORBGPU_E_CAPACITY; }
  if (maxCount > 0 && keypoints)
    ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(keypoints, (size_t)capacity * sizeof(orbgpu_keypoint), e->dKps,
                                        (size_t)e->kpCap * sizeof(orbgpu_keypoint), (size_t)maxCount * sizeof(orbgpu_keypoint),
                                        n, cudaMemcpyDeviceToHost, e->stream));
  if (maxCount > 0 && descriptors)
    ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(descriptors, (size_t)capacity * 32, e->dDesc, (size_t)e->kpCap * 32,
                                        (size_t)maxCount * 32, n, cudaMemcpyDeviceToHost, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return ORBGPU_OK;
}

// Host-buffer form.  Large batches are cut into chunks that flow through three streams — H2D copy, the kernel
// sequence, D2H copy — so that the PCIe transfers of chunks k+1 / k-1 overlap the kernels of chunk k.
// chunkScale: the synchronous call hides its first upload / last download better with more, smaller chunks (x2); callers that
// keep several batches in flight already overlap neighbouring batches and prefer fewer, larger ones
static int enqueue_extract_batch(orbgpu_extractor *e, const uint8_t *images, int count, int width, int height, size_t row_step,
                                 size_t image_step, orbgpu_keypoint *keypoints, uint8_t *descriptors, int capacity, int32_t *counts,
                                 int chunkScale) {
  if (!e || !images || !counts) { set_error("null argument"); return ORBGPU_E_INVALID; }
  if (count < 1 || count > e->p.max_batch) { set_error("batch %d outside 1..%d", count, e->p.max_batch); return ORBGPU_E_CAPACITY; }
  if (width < 1 || height < 1 || row_step < (size_t)width) { set_error("bad image shape"); return ORBGPU_E_INVALID; }
  if (width > e->p.max_width || height > e->p.max_height) {
    set_error("image %dx%d exceeds the handle's max %dx%d", width, height, e->p.max_width, e->p.max_height);
    return ORBGPU_E_CAPACITY;
  }
  if (capacity < 1 || ((keypoints == nullptr) != (descriptors == nullptr))) { set_error("bad output buffers"); return ORBGPU_E_INVALID; }
  // one host-buffer batch per handle: the next batch's kernels would overwrite dKps / dDesc / dKpCount under the previous
  // batch's D2H copies, which run on their own stream
  if (e->pendingCounts) { set_error("a batch enqueued with orbgpu_extract_batch_async has not been awaited (orbgpu_wait)"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  int rc = configure(e, width, height);
  if (rc) return rc;
  if (!e->copyIn) {
    ORBGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copyIn, cudaStreamNonBlocking));
    ORBGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copyOut, cudaStreamNonBlocking));
    for (auto &ev : e->evChunkIn) ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->evChunkDone) ORBGPU_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  const size_t imgBytes = (size_t)width * height;
  static const int wantChunks = [] { const char *v = getenv("ORBGPU_CHUNKS"); const int n = v ? atoi(v) : 2; return std::min(std::max(n, 1), (int)orbgpu_extractor::kMaxChunks); }();   // measured with two batches in flight: 1 -> 68.7k, 2 -> 71.5k, 3 -> 70.2k, 4 -> 69.7k, 8 -> 63.0k frames/s
  const int want = std::min(wantChunks * chunkScale, (int)orbgpu_extractor::kMaxChunks);
  const int chunks = count >= 8 * want ? want : count >= 16 ? 2 : 1;
  const int per = (count + chunks - 1) / chunks;
  const int cols = std::min(capacity, e->kpCap);
  begin_batch(e);
  ORBGPU_CUDA_CHECK(cudaMemsetAsync(e->dStatus, 0, sizeof(unsigned), e->stream));
  // the copy streams must not run ahead of work already queued on the handle's stream (a previous batch still reading dInput)
  ORBGPU_CUDA_CHECK(cudaEventRecord(e->evDone, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->copyIn, e->evDone, 0));
  for (int c = 0; c < chunks; c++) {
    const int first = c * per, n = std::min(per, count - first);
    if (n <= 0) break;
    // ---- H2D: tight device copy, rows of `width` bytes.  Contiguous inputs go as ONE linear copy (2-D copies of
    // odd-width rows fall off the DMA fast path).
    uint8_t *dst = e->dInput + (size_t)first * imgBytes;
    const uint8_t *src = images + (size_t)first * image_step;
    if (row_step == (size_t)width && (n == 1 || image_step == imgBytes)) {
      ORBGPU_CUDA_CHECK(cudaMemcpyAsync(dst, src, imgBytes * n, cudaMemcpyHostToDevice, e->copyIn));
    } else if (n == 1 || image_step == (size_t)height * row_step) {
      ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(dst, width, src, row_step, width, (size_t)height * n, cudaMemcpyHostToDevice, e->copyIn));
    } else {
      for (int i = 0; i < n; i++)
        ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(dst + (size_t)i * imgBytes, width, src + (size_t)i * image_step, row_step, width, height,
                                            cudaMemcpyHostToDevice, e->copyIn));
    }
    ORBGPU_CUDA_CHECK(cudaEventRecord(e->evChunkIn[c], e->copyIn));
    // ---- kernels
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->evChunkIn[c], 0));
    rc = run_extract_range_graphed(e, e->dInput, first, n, width, imgBytes);
    if (rc) return rc;
    ORBGPU_CUDA_CHECK(cudaEventRecord(e->evChunkDone[c], e->stream));
    // ---- D2H: whole rows of the chunk (counts are only known on the device; unused rows are unspecified)
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(e->copyOut, e->evChunkDone[c], 0));
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(counts + first, e->dKpCount + first, sizeof(int) * n, cudaMemcpyDeviceToHost, e->copyOut));
    if (keypoints) {
      ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(keypoints + (size_t)first * capacity, (size_t)capacity * sizeof(orbgpu_keypoint),
                                          e->dKps + (size_t)first * e->kpCap, (size_t)e->kpCap * sizeof(orbgpu_keypoint),
                                          (size_t)cols * sizeof(orbgpu_keypoint), n, cudaMemcpyDeviceToHost, e->copyOut));
      ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(descriptors + (size_t)first * capacity * 32, (size_t)capacity * 32,
                                          e->dDesc + (size_t)first * e->kpCap * 32, (size_t)e->kpCap * 32, (size_t)cols * 32, n,
                                          cudaMemcpyDeviceToHost, e->copyOut));
    }
  }
  e->lastBatch = count;
  e->pendingCounts = counts; e->pendingCount = count; e->pendingCapacity = capacity;
  return ORBGPU_OK;
}

int orbgpu_extract_batch_async(orbgpu_extractor *e, const uint8_t *images, int count, int width, int height, size_t row_step,
                               size_t image_step, orbgpu_keypoint *keypoints, uint8_t *descriptors, int capacity, int32_t *counts) {
  return enqueue_extract_batch(e, images, count, width, height, row_step, image_step, keypoints, descriptors, capacity, counts, 1);
}

int orbgpu_wait(orbgpu_extractor *e) {
  if (!e) { set_error("null handle"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  if (e->copyOut) ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->copyOut));
  int rc = check_status(e);      // synchronises the handle's stream (stereo results are copied on it)
  const int32_t *counts = e->pendingCounts;
  const int n = e->pendingCount, capacity = e->pendingCapacity;
  e->pendingCounts = nullptr; e->pendingCount = 0;
  if (rc) return rc;
  for (int i = 0; counts && i < n; i++)
    if (counts[i] > capacity) { set_error("capacity %d < %d keypoints", capacity, counts[i]); return ORBGPU_E_CAPACITY; }
  return ORBGPU_OK;
}

int orbgpu_extract_batch(orbgpu_extractor *e, const uint8_t *images, int count, int width, int height, size_t row_step,
                         size_t image_step, orbgpu_keypoint *keypoints, uint8_t *descriptors, int capacity, int32_t *counts) {
  int rc = enqueue_extract_batch(e, images, count, width, height, row_step, image_step, keypoints, descriptors, capacity, counts, 2);
  if (rc) return rc;
  return orbgpu_wait(e);
}

int orbgpu_extract(orbgpu_extractor *e, const uint8_t *image, int width, int height, size_t step, orbgpu_keypoint *keypoints,
                   uint8_t *descriptors, int capacity, int *n) {
  if (!e || !n) { set_error("null argument"); return ORBGPU_E_INVALID; }
  *n = 0;
  if (!image || width == 0 || height == 0) return ORBGPU_OK;   // ORBextractor.cc:981-982
  int32_t c = 0;
  int rc = orbgpu_extract_batch(e, image, 1, width, height, step, step * (size_t)height, keypoints, descriptors, capacity, &c);
  if (rc) return rc;
  *n = c;
  return ORBGPU_OK;
}

int orbgpu_level_size(const orbgpu_extractor *e, int level, int *width, int *height) {
  if (!e || level < 0 || level >= e->T.nlevels) { set_error("bad level"); return ORBGPU_E_INVALID; }
  if (width) *width = e->T.lv[level].w;
  if (height) *height = e->T.lv[level].h;
  return ORBGPU_OK;
}

static int copy_level(orbgpu_extractor *e, const uint8_t *base, int image, int level, int with_border, uint8_t *dst,
                      size_t dst_step) {
  int rc = check_image(e, image, level);
  if (rc) return rc;
  if (!dst) { set_error("null destination"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  const LevelGeom &G = e->T.lv[level];
  const int b = with_border ? kEdge : 0;
  const uint8_t *src = base + G.pyrOffset + (size_t)image * G.planeBytes + (size_t)(kEdge - b) * G.pitch + (kEdge - b);
  ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_step, src, G.pitch, G.w + 2 * b, G.h + 2 * b, cudaMemcpyDeviceToHost, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return ORBGPU_OK;
}

int orbgpu_get_pyramid_level(orbgpu_extractor *e, int image, int level, int with_border, uint8_t *dst, size_t dst_step) {
  return copy_level(e, e ? e->dPyr : nullptr, image, level, with_border, dst, dst_step);
}
int orbgpu_get_blurred_level(orbgpu_extractor *e, int image, int level, uint8_t *dst, size_t dst_step) {
  return copy_level(e, e ? e->dBlur : nullptr, image, level, 0, dst, dst_step);
}

int orbgpu_get_stage(orbgpu_extractor *e, int image, int level, int stage, int32_t *rows, int capacity) {
  int rc = check_image(e, image, level);
  if (rc) return rc;
  DeviceGuard guard(e->device);
  const LevelGeom &G = e->T.lv[level];
  int n = 0;
  const int *cntPtr = (stage == 0 ? e->dCandCount : e->dSelCount) + image * kMaxLevels + level;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(&n, cntPtr, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  if (stage == 0) {
    if (n > G.candCap) { set_error("candidate buffer overflowed: %d > %d", n, G.candCap); return ORBGPU_E_CAPACITY; }
    std::vector<Cand> c(n);
    if (n) ORBGPU_CUDA_CHECK(cudaMemcpy(c.data(), e->dCand + G.candOffset + (size_t)image * G.candCap, n * sizeof(Cand), cudaMemcpyDeviceToHost));
    // reference order: cell row, cell column, then row-major inside the cell (ORBextractor.cc:757-790)
    auto key = [&](const Cand &k) {
      const int x = (int)(k.xy & 0xffff) - 3, y = (int)(k.xy >> 16) - 3;
      const int cj = x / G.wCell, ci = y / G.hCell;
      return (((long long)ci * G.nCols + cj) * G.hCell + (y - ci * G.hCell)) * G.wCell + (x - cj * G.wCell);
    };
    std::sort(c.begin(), c.end(), [&](const Cand &a, const Cand &b) { return key(a) < key(b); });
    for (int i = 0; i < n && i < capacity && rows; i++) {
      rows[3 * i] = c[i].xy & 0xffff; rows[3 * i + 1] = c[i].xy >> 16; rows[3 * i + 2] = (int)c[i].resp;
    }
  } else if (stage == 1) {
    std::vector<int4> s(n);
    if (n) ORBGPU_CUDA_CHECK(cudaMemcpy(s.data(), e->dSel + G.selOffset + (size_t)image * G.selCap, n * sizeof(int4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < capacity && rows; i++) {
      rows[3 * i] = s[i].x - kMinBorder; rows[3 * i + 1] = s[i].y - kMinBorder; rows[3 * i + 2] = s[i].z;
    }
  } else {
    set_error("unknown stage %d", stage);
    return ORBGPU_E_INVALID;
  }
  return n;
}

// ---------------------------------------------------------------- stereo

int orbgpu_stereo_match_device(orbgpu_extractor *L, orbgpu_extractor *R, int first_left, int first_right, int pairs,
                               float mbf, float mb) {
  if (!L || !R) { set_error("null handle"); return ORBGPU_E_INVALID; }
  if (L->device != R->device) { set_error("left and right handles live on different devices"); return ORBGPU_E_INVALID; }
  if (pairs < 1 || first_left < 0 || first_right < 0 || first_left + pairs > L->lastBatch || first_right + pairs > R->lastBatch) {
    set_error("pair range outside the last batches");
    return ORBGPU_E_INVALID;
  }
  if (L->cfgW != R->cfgW || L->cfgH != R->cfgH || L->p.nlevels != R->p.nlevels || L->p.scale_factor != R->p.scale_factor ||
      L->p.max_batch != R->p.max_batch || L->p.max_width != R->p.max_width || L->p.max_height != R->p.max_height ||
      L->p.nfeatures != R->p.nfeatures || L->kpCap != R->kpCap) {      // dSorted / row buckets are sized by the LEFT handle's kpCap
    set_error("left and right extractors must share image size, pyramid parameters, nfeatures and buffer sizing");
    return ORBGPU_E_INVALID;
  }
  if (!(mb > 0.f) || !(mbf > 0.f)) { set_error("mbf and mb must be positive"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(L->device);
  if (!L->dURight) {
    const size_t n = (size_t)L->p.max_batch * L->kpCap;
    ORBGPU_CUDA_CHECK(cudaMalloc(&L->dURight, n * sizeof(float)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&L->dDepth, n * sizeof(float)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&L->dBestSad, n * sizeof(int)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&L->dSorted, n * sizeof(int4)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&L->dRowStart, (size_t)L->p.max_batch * (L->p.max_height + 2) * sizeof(int)));
    ORBGPU_CUDA_CHECK(stereo_configure(L->p.max_height + 1));
    ORBGPU_CUDA_CHECK(stereo_configure_lat(L->p.max_height + 1));
  }
  if (R != L) {
    ORBGPU_CUDA_CHECK(cudaEventRecord(R->evDone, R->stream));
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(L->stream, R->evDone, 0));
  }
  StereoArgs A{};
  A.pyrL = L->dPyr; A.pyrR = R->dPyr; A.kpsL = L->dKps; A.kpsR = R->dKps; A.descL = L->dDesc; A.descR = R->dDesc;
  A.cntL = L->dKpCount; A.cntR = R->dKpCount; A.kpCapL = L->kpCap; A.kpCapR = R->kpCap;
  A.firstL = first_left; A.firstR = first_right; A.pairs = pairs; A.mbf = mbf; A.mb = mb;
  A.uRight = L->dURight; A.depth = L->dDepth; A.bestSad = L->dBestSad; A.rowStart = L->dRowStart; A.recs = L->dSorted;
  { StageTimer t(L, 5); ORBGPU_CUDA_CHECK(pairs <= kLatencyBatch / 2 ? launch_stereo_lat(L->T, L->dT, A, L->kpCap, L->stream) : launch_stereo(L->T, L->dT, A, L->kpCap, L->stream)); }
  if (R != L) {   // the right handle must not overwrite its buffers before the match has read them
    ORBGPU_CUDA_CHECK(cudaEventRecord(L->evDone, L->stream));
    ORBGPU_CUDA_CHECK(cudaStreamWaitEvent(R->stream, L->evDone, 0));
  }
  return ORBGPU_OK;
}

int orbgpu_stereo_results_device(orbgpu_extractor *L, const float **d_u_right, const float **d_depth,
                                 const int32_t **d_best_sad) {
  if (!L || !L->dURight) { set_error("no stereo match has run on this handle"); return ORBGPU_E_INVALID; }
  if (d_u_right) *d_u_right = L->dURight;
  if (d_depth) *d_depth = L->dDepth;
  if (d_best_sad) *d_best_sad = L->dBestSad;
  return ORBGPU_OK;
}

int orbgpu_stereo_match_async(orbgpu_extractor *L, orbgpu_extractor *R, int first_left, int first_right, int pairs, float mbf,
                              float mb, float *u_right, float *depth, int capacity) {
  if (!u_right || !depth) { set_error("null output"); return ORBGPU_E_INVALID; }
  int rc = orbgpu_stereo_match_device(L, R, first_left, first_right, pairs, mbf, mb);
  if (rc) return rc;
  DeviceGuard guard(L->device);
  const int cols = std::min(capacity, L->kpCap);
  ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(u_right, (size_t)capacity * 4, L->dURight, (size_t)L->kpCap * 4, (size_t)cols * 4, pairs,
                                      cudaMemcpyDeviceToHost, L->stream));
  ORBGPU_CUDA_CHECK(cudaMemcpy2DAsync(depth, (size_t)capacity * 4, L->dDepth, (size_t)L->kpCap * 4, (size_t)cols * 4, pairs,
                                      cudaMemcpyDeviceToHost, L->stream));
  return ORBGPU_OK;
}

int orbgpu_stereo_match(orbgpu_extractor *L, orbgpu_extractor *R, int first_left, int first_right, int pairs, float mbf,
                        float mb, float *u_right, float *depth, int capacity) {
  int rc = orbgpu_stereo_match_async(L, R, first_left, first_right, pairs, mbf, mb, u_right, depth, capacity);
  if (rc) return rc;
  DeviceGuard guard(L->device);
  return check_status(L);
}

// Used by search.cu (orbgpu_features_from_extractor): device pointers of image `image` of the last batch and, when
// pair >= 0, row `pair` of the stereo results of the last orbgpu_stereo_match on this (left) handle.
extern "C" int orbgpu_internal_extractor_view(orbgpu_extractor *e, int image, int pair, const orbgpu_keypoint **kps, const uint8_t **desc,
                                              int *count, const float **uRight, int *device, void **stream) {
  if (image < 0 || image >= e->lastBatch) { set_error("image %d outside the last batch (%d)", image, e->lastBatch); return ORBGPU_E_INVALID; }
  if (pair >= 0 && (!e->dURight || pair >= e->p.max_batch)) { set_error("no stereo result row %d on this handle", pair); return ORBGPU_E_INVALID; }
  DeviceGuard guard(e->device);
  int n = 0;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(&n, e->dKpCount + image, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  int rc = check_status(e);
  if (rc) return rc;
  *kps = e->dKps + (size_t)image * e->kpCap;
  *desc = e->dDesc + (size_t)image * e->kpCap * 32;
  *count = std::min(n, e->kpCap);
  *uRight = pair >= 0 ? e->dURight + (size_t)pair * e->kpCap : nullptr;
  *device = e->device;
  *stream = (void *)e->stream;
  return ORBGPU_OK;
}

// ---------------------------------------------------------------- matcher

int orbgpu_bank_wrap_device(const uint8_t *d_descriptors, int rows, int64_t index_base, int device, orbgpu_bank **out) {
  if (!out || rows < 0 || (rows > 0 && !d_descriptors)) { set_error("bad bank arguments"); return ORBGPU_E_INVALID; }
  if (((uintptr_t)d_descriptors & 15) != 0) { set_error("bank rows must be 16-byte aligned"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  orbgpu_bank *b = new orbgpu_bank();
  if (device >= 0) b->device = device; else ORBGPU_CUDA_CHECK(cudaGetDevice(&b->device));
  DeviceGuard guard(b->device);
  cudaDeviceProp prop;
  ORBGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, b->device));
  b->smCount = prop.multiProcessorCount;
  b->d = d_descriptors; b->rows = rows; b->indexBase = index_base;
  ORBGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  *out = b;
  return ORBGPU_OK;
}

int orbgpu_bank_create(const uint8_t *descriptors, int rows, int64_t index_base, int device, orbgpu_bank **out) {
  if (!out || rows < 0 || (rows > 0 && !descriptors)) { set_error("bad bank arguments"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  int dev = device;
  if (dev < 0) ORBGPU_CUDA_CHECK(cudaGetDevice(&dev));
  uint8_t *d = nullptr;
  {
    DeviceGuard guard(dev);
    ORBGPU_CUDA_CHECK(cudaMalloc(&d, std::max<size_t>((size_t)rows * 32, 32)));
    if (rows) ORBGPU_CUDA_CHECK(cudaMemcpy(d, descriptors, (size_t)rows * 32, cudaMemcpyHostToDevice));
  }
  int rc = orbgpu_bank_wrap_device(d, rows, index_base, dev, out);
  if (rc) { cudaFree(d); return rc; }
  (*out)->owned = d;
  return ORBGPU_OK;
}

void orbgpu_bank_destroy(orbgpu_bank *b) {
  if (!b) return;
  DeviceGuard guard(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  if (b->owned) cudaFree(b->owned);
  if (b->partial) cudaFree(b->partial);
  if (b->dQueries) cudaFree(b->dQueries);
  if (b->dResult) cudaFree(b->dResult);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int orbgpu_hamming_top2_dense_device(orbgpu_bank *b, const uint8_t *d_queries, int nq, void *d_result, void *stream) {
  if (!b || nq < 0 || (nq > 0 && (!d_queries || !d_result))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (nq == 0) return ORBGPU_OK;
  if (((uintptr_t)d_queries & 15) != 0 || ((uintptr_t)d_result & 15) != 0) { set_error("queries / results must be 16-byte aligned"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(b->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : b->stream;
  const int chunks = hamming_dense_chunks(b->rows, nq, b->smCount);
  const size_t need = (size_t)chunks * nq;
  if (need > b->partialCap) {
    ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    if (b->partial) cudaFree(b->partial);
    b->partialCap = need;
    ORBGPU_CUDA_CHECK(cudaMalloc(&b->partial, need * sizeof(Top2Rec)));
  }
  ORBGPU_CUDA_CHECK(launch_hamming_dense(b->d, b->rows, b->indexBase, d_queries, nq, b->partial, chunks, (Top2Rec *)d_result, s));
  return ORBGPU_OK;
}

int orbgpu_hamming_top2_dense(orbgpu_bank *b, const uint8_t *queries, int nq, int64_t *idx1, int32_t *d1, int32_t *d2) {
  if (!b || nq < 0 || (nq > 0 && (!queries || !idx1 || !d1 || !d2))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (nq == 0) return ORBGPU_OK;
  DeviceGuard guard(b->device);
  if (nq > b->qCap) {
    ORBGPU_CUDA_CHECK(cudaStreamSynchronize(b->stream));
    if (b->dQueries) cudaFree(b->dQueries);
    if (b->dResult) cudaFree(b->dResult);
    b->qCap = nq;
    ORBGPU_CUDA_CHECK(cudaMalloc(&b->dQueries, (size_t)nq * 32));
    ORBGPU_CUDA_CHECK(cudaMalloc(&b->dResult, (size_t)nq * sizeof(Top2Rec)));
  }
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(b->dQueries, queries, (size_t)nq * 32, cudaMemcpyHostToDevice, b->stream));
  int rc = orbgpu_hamming_top2_dense_device(b, b->dQueries, nq, b->dResult, b->stream);
  if (rc) return rc;
  std::vector<Top2Rec> r(nq);
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(r.data(), b->dResult, (size_t)nq * sizeof(Top2Rec), cudaMemcpyDeviceToHost, b->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  for (int i = 0; i < nq; i++) { idx1[i] = r[i].idx; d1[i] = r[i].d1; d2[i] = r[i].d2; }
  return ORBGPU_OK;
}

struct orbgpu_exchange {
  int rank = 0, world = 1, maxQueries = 0, device = 0;
  char *window = nullptr;                 // this rank's window (exported through CUDA IPC)
  std::vector<char *> peer;               // every rank's window as mapped here (peer[rank] == window)
  char **dPeers = nullptr;                // device copy of `peer`
  unsigned *dDone = nullptr, *dError = nullptr;
  unsigned epoch = 0;
  bool connected = false;
};

int orbgpu_exchange_create(int rank, int world, int max_queries, int device, orbgpu_exchange **out, uint8_t *handle) {
  if (!out || !handle || world < 1 || world > 32 || rank < 0 || rank >= world || max_queries < 1) { set_error("bad exchange arguments"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) == ORBGPU_IPC_HANDLE_BYTES, "IPC handle size");
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  orbgpu_exchange *x = new orbgpu_exchange;
  x->rank = rank; x->world = world; x->maxQueries = max_queries;
  cudaGetDevice(&x->device);
  const size_t bytes = exchange_window_bytes(world, max_queries);
  cudaError_t e = cudaMalloc(&x->window, bytes);
  if (e == cudaSuccess) e = cudaMemset(x->window, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&x->dPeers, sizeof(char *) * world);
  if (e == cudaSuccess) e = cudaMalloc(&x->dDone, 2 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(x->dDone, 0, 2 * sizeof(unsigned));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->window);
  if (e != cudaSuccess) { set_error("exchange_create: %s", cudaGetErrorString(e)); orbgpu_exchange_destroy(x); return ORBGPU_E_CUDA; }
  x->dError = x->dDone + 1;
  std::memcpy(handle, &h, sizeof(h));
  x->peer.assign(world, nullptr);
  x->peer[rank] = x->window;
  if (world == 1) {
    if (cudaMemcpy(x->dPeers, x->peer.data(), sizeof(char *), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("exchange_create: copy failed"); orbgpu_exchange_destroy(x); return ORBGPU_E_CUDA; }
    x->connected = true;
  }
  *out = x;
  return ORBGPU_OK;
}

int orbgpu_exchange_connect(orbgpu_exchange *x, const uint8_t *handles) {
  if (!x || !handles) { set_error("null argument"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(x->device);
  for (int r = 0; r < x->world; r++) {
    if (r == x->rank || x->peer[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * ORBGPU_IPC_HANDLE_BYTES, sizeof(h));
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cannot map the window of rank %d: %s", r, cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
    x->peer[r] = (char *)p;
  }
  ORBGPU_CUDA_CHECK(cudaMemcpy(x->dPeers, x->peer.data(), sizeof(char *) * x->world, cudaMemcpyHostToDevice));
  x->connected = true;
  return ORBGPU_OK;
}

int orbgpu_exchange_check(orbgpu_exchange *x) {
  if (!x) { set_error("null exchange"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(x->device);
  unsigned err = 0;
  ORBGPU_CUDA_CHECK(cudaMemcpy(&err, x->dError, sizeof(err), cudaMemcpyDeviceToHost));
  if (err) { set_error("rank %d: the records of rank %u did not arrive within 2 s", x->rank, err - 1); return ORBGPU_E_CUDA; }
  return ORBGPU_OK;
}

void orbgpu_exchange_destroy(orbgpu_exchange *x) {
  if (!x) return;
  DeviceGuard guard(x->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < (int)x->peer.size(); r++)
    if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  cudaFree(x->window); cudaFree(x->dPeers); cudaFree(x->dDone);
  delete x;
}

int orbgpu_hamming_top2_sharded_device(orbgpu_bank *b, orbgpu_exchange *x, const uint8_t *d_queries, int nq, void *d_result, float nnratio,
                                       int th_low, int64_t *d_match, void *stream) {
  if (!b || !x || nq < 0 || (nq > 0 && !d_queries)) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (!x->connected) { set_error("exchange not connected"); return ORBGPU_E_INVALID; }
  if (nq > x->maxQueries) { set_error("%d queries exceed the exchange window (%d)", nq, x->maxQueries); return ORBGPU_E_CAPACITY; }
  if (b->device != x->device) { set_error("bank and exchange live on different devices"); return ORBGPU_E_INVALID; }
  if (nq == 0) return ORBGPU_OK;
  if (((uintptr_t)d_queries & 15) != 0 || ((uintptr_t)d_result & 15) != 0) { set_error("queries / results must be 16-byte aligned"); return ORBGPU_E_INVALID; }
  DeviceGuard guard(b->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : b->stream;
  const int chunks = hamming_dense_chunks(b->rows, nq, b->smCount);
  const size_t need = (size_t)chunks * nq;
  if (need > b->partialCap) {
    ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    if (b->partial) cudaFree(b->partial);
    b->partialCap = need;
    ORBGPU_CUDA_CHECK(cudaMalloc(&b->partial, need * sizeof(Top2Rec)));
  }
  ExchangeView V;
  V.rank = x->rank; V.world = x->world; V.maxQueries = x->maxQueries; V.peers = x->dPeers; V.local = x->window; V.done = x->dDone; V.error = x->dError;
  const unsigned epoch = ++x->epoch;
  ORBGPU_CUDA_CHECK(launch_hamming_scan_only(b->d, b->rows, b->indexBase, d_queries, nq, b->partial, chunks, s));
  ORBGPU_CUDA_CHECK(launch_top2_push(b->partial, b->rows > 0 ? chunks : 0, nq, V, epoch, s));
  ORBGPU_CUDA_CHECK(launch_top2_wait_merge(nq, V, epoch, (Top2Rec *)d_result, nnratio, th_low, (long long *)d_match, s));
  return ORBGPU_OK;
}

int orbgpu_top2_merge_device(const void *d_parts, int parts, int nq, void *d_result, float nnratio, int th_low,
                             int64_t *d_match, int device, void *stream) {
  if (parts < 0 || nq < 0 || (nq > 0 && parts > 0 && !d_parts)) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  ORBGPU_CUDA_CHECK(launch_top2_merge((const Top2Rec *)d_parts, parts, nq, (Top2Rec *)d_result, nnratio, th_low,
                                      (long long *)d_match, (cudaStream_t)stream));
  return ORBGPU_OK;
}

int orbgpu_hamming_top2_csr(const uint8_t *queries, int nq, const uint8_t *train, int nt, const int32_t *offsets,
                            const int32_t *cand, int32_t *idx1, int32_t *d1, int32_t *d2, int device) {
  if (nq < 0 || nt < 0 || (nq > 0 && (!queries || !offsets || !idx1 || !d1 || !d2))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (nq == 0) return ORBGPU_OK;
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  const int ncand = offsets[nq];
  for (int i = 0; i < nq; i++)
    if (offsets[i] > offsets[i + 1] || offsets[i] < 0) { set_error("offsets must be non-decreasing"); return ORBGPU_E_INVALID; }
  for (int i = 0; i < ncand; i++)
    if (cand[i] < 0 || cand[i] >= nt) { set_error("candidate %d out of range", cand[i]); return ORBGPU_E_INVALID; }
  DeviceGuard guard(device);
  uint8_t *dq = nullptr, *dt = nullptr;
  int *doff = nullptr, *dcand = nullptr, *dout = nullptr;
  auto cleanup = [&] { cudaFree(dq); cudaFree(dt); cudaFree(doff); cudaFree(dcand); cudaFree(dout); };
  cudaError_t ce;
#define CSR_TRY(x) if ((ce = (x)) != cudaSuccess) { set_error("%s failed: %s", #x, cudaGetErrorString(ce)); cleanup(); return ORBGPU_E_CUDA; }
  CSR_TRY(cudaMalloc(&dq, (size_t)nq * 32));
  CSR_TRY(cudaMalloc(&dt, std::max<size_t>((size_t)nt * 32, 32)));
  CSR_TRY(cudaMalloc(&doff, (size_t)(nq + 1) * 4));
  CSR_TRY(cudaMalloc(&dcand, std::max<size_t>((size_t)ncand * 4, 4)));
  CSR_TRY(cudaMalloc(&dout, (size_t)nq * 12));
  CSR_TRY(cudaMemcpy(dq, queries, (size_t)nq * 32, cudaMemcpyHostToDevice));
  if (nt) CSR_TRY(cudaMemcpy(dt, train, (size_t)nt * 32, cudaMemcpyHostToDevice));
  CSR_TRY(cudaMemcpy(doff, offsets, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice));
  if (ncand) CSR_TRY(cudaMemcpy(dcand, cand, (size_t)ncand * 4, cudaMemcpyHostToDevice));
  CSR_TRY(launch_hamming_csr(dq, nq, dt, doff, dcand, dout, dout + nq, dout + 2 * nq, 0));
  CSR_TRY(cudaMemcpy(idx1, dout, (size_t)nq * 4, cudaMemcpyDeviceToHost));
  CSR_TRY(cudaMemcpy(d1, dout + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost));
  CSR_TRY(cudaMemcpy(d2, dout + 2 * nq, (size_t)nq * 4, cudaMemcpyDeviceToHost));
#undef CSR_TRY
  cleanup();
  return ORBGPU_OK;
}

int orbgpu_descriptor_distance(const uint8_t *a, const uint8_t *b, int n, int32_t *dist, int device) {
  if (n < 0 || (n > 0 && (!a || !b || !dist))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (n == 0) return ORBGPU_OK;
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  uint8_t *da = nullptr, *db = nullptr;
  int *dd = nullptr;
  cudaError_t ce = cudaMalloc(&da, (size_t)n * 32);
  if (ce == cudaSuccess) ce = cudaMalloc(&db, (size_t)n * 32);
  if (ce == cudaSuccess) ce = cudaMalloc(&dd, (size_t)n * 4);
  if (ce == cudaSuccess) ce = cudaMemcpy(da, a, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(db, b, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = launch_pair_distance(da, db, n, dd, 0);
  if (ce == cudaSuccess) ce = cudaMemcpy(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dd);
  if (ce != cudaSuccess) { set_error("descriptor_distance: %s", cudaGetErrorString(ce)); return ORBGPU_E_CUDA; }
  return ORBGPU_OK;
}

int orbgpu_measure_popc_peak(int device, double *popc_per_second, double *lop3_per_second) {
  if (!popc_per_second || !lop3_per_second) { set_error("null output"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  int dev = 0;
  ORBGPU_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ORBGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  ORBGPU_CUDA_CHECK(run_popc_microbench(prop.multiProcessorCount, popc_per_second, lop3_per_second));
  return ORBGPU_OK;
}

int orbgpu_selftest_sincos(uint32_t first_bits, uint64_t count, int chunk_log2, uint64_t *sums, float *sin_values, float *cos_values, int device) {
  if (chunk_log2 < 5 || chunk_log2 > 30 || count == 0 || (uint64_t)first_bits + count > 0x100000000ull) { set_error("bad range"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  if (sums) {
    const size_t chunks = (size_t)((count + (1ull << chunk_log2) - 1) >> chunk_log2);
    unsigned long long *d = nullptr;
    ORBGPU_CUDA_CHECK(cudaMalloc(&d, chunks * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d, 0, chunks * sizeof(unsigned long long));
    if (e == cudaSuccess) e = launch_sincos_checksum(first_bits, count, chunk_log2, d, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(sums, d, chunks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("sincos selftest: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  }
  if (sin_values && cos_values) {
    if (count > (1u << 26)) { set_error("value dump limited to 2^26 patterns"); return ORBGPU_E_INVALID; }
    float *ds = nullptr, *dc = nullptr;
    cudaError_t e = cudaMalloc(&ds, count * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&dc, count * sizeof(float));
    if (e == cudaSuccess) e = launch_sincos_values(first_bits, (int)count, ds, dc, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(sin_values, ds, count * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cos_values, dc, count * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(ds); cudaFree(dc);
    if (e != cudaSuccess) { set_error("sincos selftest: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  }
  return ORBGPU_OK;
}

int orbgpu_measure_pipe(int device, int mode, double *ops_per_second) {
  if (!ops_per_second) { set_error("null output"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceGuard guard(device);
  int dev = 0;
  ORBGPU_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ORBGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  ORBGPU_CUDA_CHECK(run_pipe_probe(prop.multiProcessorCount, mode, ops_per_second));
  return ORBGPU_OK;
}

}  // extern "C"
