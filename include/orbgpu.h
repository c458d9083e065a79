/* orbgpu.h — C ABI of the B200-native ORB front-end (liborbgpu.so).
 *
 * Drop-in boundary for the ONE hot path of sjulier/Refactored_ORB_SLAM2 (paths below are relative
 * to Source/Libraries/ORB_SLAM2 in the reference tree):
 *
 *   ORBextractor::operator()            src/ORBextractor.cc:978-1039   (include/ORBextractor.h:47-57)
 *   ORBextractor getters / tables       include/ORBextractor.h:59-75, src/ORBextractor.cc:407-464
 *   ORBmatcher::DescriptorDistance +    src/ORBmatcher.cc:1542-1556, best/second-best loops :201-243
 *     the best / second-best searches
 *   Frame::ComputeStereoMatches         src/Frame.cc:477-646
 *
 * Plain pointers and sizes only; no C++/torch types.  Every entry point returns ORBGPU_OK (0) or a
 * negative error code; orbgpu_last_error() returns the message for the calling thread.  Nothing in
 * this library falls back to the CPU: without a usable CUDA device every compute call fails with
 * ORBGPU_E_CUDA.
 *
 * Threading: one handle = one CUDA stream + private device buffers.  Distinct handles may be used
 * concurrently from distinct host threads (the reference runs the left and right extractor on two
 * std::threads, src/Frame.cc:87-90); one handle must not be used from two threads at once.
 */
#ifndef ORBGPU_H
#define ORBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBGPU_VERSION 160

enum {
  ORBGPU_OK = 0,
  ORBGPU_E_INVALID = -1,     /* bad argument */
  ORBGPU_E_CUDA = -2,        /* CUDA runtime error, or no device */
  ORBGPU_E_CAPACITY = -3,    /* a caller- or handle-sized buffer is too small */
  ORBGPU_E_UNSUPPORTED = -4  /* shape the reference itself cannot process (e.g. aspect < 0.5) */
};

/* cv::KeyPoint memory layout (7 x 4 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct orbgpu_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orbgpu_keypoint;

/* ORBextractor constructor arguments (include/ORBextractor.h:50-51) + device buffer sizing. */
typedef struct orbgpu_extractor_params {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels;
  int32_t ini_th_fast;
  int32_t min_th_fast;
  int32_t max_width;   /* largest image the handle will see */
  int32_t max_height;
  int32_t max_batch;   /* images processed per call (a stereo pair = 2) */
  int32_t device;      /* CUDA device ordinal, -1 = current device */
} orbgpu_extractor_params;

typedef struct orbgpu_extractor orbgpu_extractor;
typedef struct orbgpu_bank orbgpu_bank;

int orbgpu_version(void);
const char *orbgpu_last_error(void);
int orbgpu_device_count(void);

/* ---- ORBextractor ------------------------------------------------------------------------- */

/* ORBextractor::ORBextractor (src/ORBextractor.cc:407-464). */
int orbgpu_extractor_create(const orbgpu_extractor_params *params, orbgpu_extractor **out);
void orbgpu_extractor_destroy(orbgpu_extractor *e);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:63-73) and mnFeaturesPerLevel; each array holds nlevels entries, any
 * pointer may be NULL. */
int orbgpu_extractor_tables(const orbgpu_extractor *e, float *scale, float *inv_scale, float *sigma2,
                            float *inv_sigma2, int32_t *features_per_level);
/* Upper bound on keypoints per image: per level max(quota + 3, 4 x 16 roots) — a careful octree split adds at
 * most 3 nodes before the >= N break (src/ORBextractor.cc:701-702), and the first round splits every root
 * whatever the quota (src/ORBextractor.cc:566-600), so a wide level with a small quota returns 4 x roots. */
int orbgpu_extractor_capacity(const orbgpu_extractor *e);

/* ORBextractor::operator()(image, mask, keypoints, descriptors) (src/ORBextractor.cc:978-1039):
 * one CV_8UC1 host image with row stride `step`; mask is ignored by the reference and absent here.
 * Writes *n keypoints (level-major, reference order) and *n rows of 32 descriptor bytes.
 * An empty image (width or height 0, or image NULL) returns ORBGPU_OK with *n = 0, like :981. */
int orbgpu_extract(orbgpu_extractor *e, const uint8_t *image, int width, int height, size_t step,
                   orbgpu_keypoint *keypoints, uint8_t *descriptors, int capacity, int *n);

/* The same call over `count` equally sized host images (image i starts at images + i*image_step).
 * keypoints / descriptors hold `capacity` rows per image; counts[i] receives image i's count. */
int orbgpu_extract_batch(orbgpu_extractor *e, const uint8_t *images, int count, int width, int height,
                         size_t row_step, size_t image_step, orbgpu_keypoint *keypoints,
                         uint8_t *descriptors, int capacity, int32_t *counts);

/* Split form for callers that keep several batches in flight on several handles from ONE host thread
 * (refactored_orb_slam2_b200/pipeline.py): _async enqueues the chunked H2D copies, the kernels and the D2H copies and
 * returns; orbgpu_wait blocks until everything queued on the handle (including an orbgpu_stereo_match_async) has
 * landed in the host buffers, then reports device-side overflow and counts[i] > capacity.  Host buffers should be
 * pinned, and must stay valid until orbgpu_wait returns.  orbgpu_extract_batch == _async + orbgpu_wait.
 * ONE batch per handle: a second _async (or orbgpu_extract_batch_device) before orbgpu_wait returns ORBGPU_E_INVALID —
 * its kernels would overwrite the device results under the first batch's D2H copies. */
int orbgpu_extract_batch_async(orbgpu_extractor *e, const uint8_t *images, int count, int width, int height,
                               size_t row_step, size_t image_step, orbgpu_keypoint *keypoints,
                               uint8_t *descriptors, int capacity, int32_t *counts);
int orbgpu_wait(orbgpu_extractor *e);

/* Device-resident form: images already in HBM; results stay in the handle's device buffers (see
 * orbgpu_results_device) until the next extract call.  Asynchronous on the handle's stream. */
int orbgpu_extract_batch_device(orbgpu_extractor *e, const uint8_t *d_images, int count, int width,
                                int height, size_t row_step, size_t image_step);
/* Device pointers to the last batch: keypoints [count][capacity], descriptors [count][capacity][32],
 * counts [count].  *capacity receives the row capacity per image. */
int orbgpu_results_device(orbgpu_extractor *e, const orbgpu_keypoint **d_keypoints,
                          const uint8_t **d_descriptors, const int32_t **d_counts, int *capacity);
/* Copies the last batch's results to host buffers laid out as in orbgpu_extract_batch; checks the
 * device-side status word (candidate / node capacity) and synchronises the stream. */
int orbgpu_download_results(orbgpu_extractor *e, orbgpu_keypoint *keypoints, uint8_t *descriptors,
                            int capacity, int32_t *counts);
int orbgpu_synchronize(orbgpu_extractor *e);
/* The handle's cudaStream_t (as void*), so callers can order their own work against it. */
void *orbgpu_extractor_stream(orbgpu_extractor *e);

/* mvImagePyramid (include/ORBextractor.h:75): level size and bytes of image `image` of the last
 * batch.  with_border != 0 also copies the 19 px BORDER_REFLECT_101 apron (dst is then
 * (h+38) x (w+38)). */
int orbgpu_level_size(const orbgpu_extractor *e, int level, int *width, int *height);
int orbgpu_get_pyramid_level(orbgpu_extractor *e, int image, int level, int with_border, uint8_t *dst,
                             size_t dst_step);
/* Stage outputs for parity tests: stage 0 = FAST candidates after per-cell NMS (returned sorted in
 * the reference's cell-major order), 1 = octree selection in list order, 2 = blurred level bytes
 * are fetched with orbgpu_get_blurred_level.  Rows are (x, y, response) in octree coordinates
 * (level coordinates - 16).  Returns the row count (>= 0) or an error code. */
int orbgpu_get_stage(orbgpu_extractor *e, int image, int level, int stage, int32_t *rows, int capacity);
int orbgpu_get_blurred_level(orbgpu_extractor *e, int image, int level, uint8_t *dst, size_t dst_step);

/* ---- Frame::ComputeStereoMatches ------------------------------------------------------------ */

/* Frame::ComputeStereoMatches (src/Frame.cc:477-646) for `pairs` stereo pairs whose left images are
 * images [first_left, first_left+pairs) of `left`'s last batch and right images are
 * [first_right, ...) of `right`'s last batch (left and right may be the same handle).
 * `mb` is the baseline in metres (= mbf / fx): the reference reads the member before assigning it
 * (src/Frame.cc:505 vs :124), so it is explicit here.  u_right / depth hold `capacity` floats per
 * pair (host memory), -1 where there is no match (mvuRight / mvDepth). */
int orbgpu_stereo_match(orbgpu_extractor *left, orbgpu_extractor *right, int first_left,
                        int first_right, int pairs, float mbf, float mb, float *u_right, float *depth,
                        int capacity);
/* The same without the final synchronisation (pair it with orbgpu_wait on `left`). */
int orbgpu_stereo_match_async(orbgpu_extractor *left, orbgpu_extractor *right, int first_left, int first_right,
                              int pairs, float mbf, float mb, float *u_right, float *depth, int capacity);
/* Device-resident form; results at orbgpu_stereo_results_device of the LEFT handle. */
int orbgpu_stereo_match_device(orbgpu_extractor *left, orbgpu_extractor *right, int first_left,
                               int first_right, int pairs, float mbf, float mb);
int orbgpu_stereo_results_device(orbgpu_extractor *left, const float **d_u_right, const float **d_depth,
                                 const int32_t **d_best_sad);

/* ---- ORBmatcher Hamming searches ------------------------------------------------------------ */

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1542-1556) for n pairs of 32-byte rows (host). */
int orbgpu_descriptor_distance(const uint8_t *a, const uint8_t *b, int n, int32_t *dist, int device);

/* A descriptor bank (keyframe descriptors) resident in HBM.  `index_base` is added to row numbers
 * in results, so a row-sharded bank reports global indices. */
int orbgpu_bank_create(const uint8_t *descriptors, int rows, int64_t index_base, int device,
                       orbgpu_bank **out);
/* Wraps device memory owned by the caller (must stay alive, 16-byte aligned). */
int orbgpu_bank_wrap_device(const uint8_t *d_descriptors, int rows, int64_t index_base, int device,
                            orbgpu_bank **out);
void orbgpu_bank_destroy(orbgpu_bank *b);

/* Best / second-best scan of src/ORBmatcher.cc:201-221 over the whole bank for every query:
 * strict '<' (lowest index wins ties), best1 = best2 = 256 and idx = -1 when the bank is empty.
 * Host queries [nq][32]; outputs idx1 (int64, global index), d1, d2 (int32). */
int orbgpu_hamming_top2_dense(orbgpu_bank *bank, const uint8_t *queries, int nq, int64_t *idx1,
                              int32_t *d1, int32_t *d2);
/* Device form: d_queries [nq][32] (16-byte aligned); d_result receives nq records of
 * { int32 d1, int32 d2, int64 idx1 } (16 bytes) — the unit exchanged between bank shards.
 * `stream` is a cudaStream_t; NULL selects the bank's own (non-blocking) stream. */
int orbgpu_hamming_top2_dense_device(orbgpu_bank *bank, const uint8_t *d_queries, int nq,
                                     void *d_result, void *stream);
/* Merges `parts` shard results ([parts][nq] records, shard order = ascending bank index) exactly
 * as one sequential scan over the concatenated bank would; optionally applies the acceptance test
 * of src/ORBmatcher.cc:224-226 (d1 <= th_low && d1 < nnratio*d2): d_match[i] = idx1 or -1. */
int orbgpu_top2_merge_device(const void *d_parts, int parts, int nq, void *d_result, float nnratio,
                             int th_low, int64_t *d_match, int device, void *stream);

/* ---- row-sharded bank across the GPUs of one node: scan + exchange + merge without a collective library ----
 * One process per GPU.  Every rank owns a window in its HBM ([2 epochs][world][max_queries] 16-byte records + arrival
 * flags) that its peers map through CUDA IPC.  orbgpu_hamming_top2_sharded_device then runs, on `stream`:
 *   1. the bank scan of this rank's shard (as orbgpu_hamming_top2_dense_device);
 *   2. ONE kernel that folds the shard's chunk partials per query and stores the 16-byte record straight into EVERY
 *      peer's window over NVLink (peer stores), the last CTA publishing this rank's arrival flag in every window;
 *   3. ONE kernel that waits for the `world` arrival flags of the current epoch in the LOCAL window and folds the
 *      records in rank order (= ascending bank index), applying the acceptance test of src/ORBmatcher.cc:224-226.
 * The exchange therefore starts the moment a rank's scan ends and costs no host round trip and no NCCL launch; the
 * result is identical on every rank and identical to one sequential scan of the whole bank.  Windows alternate between
 * two epochs, so a fast rank can run one step ahead of a slow one, never two.  A peer that does not arrive within
 * ~2 s makes the call's NEXT invocation (or orbgpu_exchange_check) fail with ORBGPU_E_CUDA instead of hanging. */
typedef struct orbgpu_exchange orbgpu_exchange;
#define ORBGPU_IPC_HANDLE_BYTES 64
/* Allocates this rank's window and writes its IPC handle (to be all-gathered by the caller's process group). */
int orbgpu_exchange_create(int rank, int world, int max_queries, int device, orbgpu_exchange **out,
                           uint8_t *handle /* [ORBGPU_IPC_HANDLE_BYTES] */);
/* Maps the peers' windows: handles = [world][ORBGPU_IPC_HANDLE_BYTES] in rank order (own entry ignored). */
int orbgpu_exchange_connect(orbgpu_exchange *x, const uint8_t *handles);
int orbgpu_exchange_check(orbgpu_exchange *x);     /* synchronises nothing; reports a recorded arrival timeout */
void orbgpu_exchange_destroy(orbgpu_exchange *x);
/* d_result: nq merged { d1, d2, idx } records (may be NULL); d_match: nq accepted global indices or -1 (may be NULL). */
int orbgpu_hamming_top2_sharded_device(orbgpu_bank *bank, orbgpu_exchange *x, const uint8_t *d_queries, int nq,
                                       void *d_result, float nnratio, int th_low, int64_t *d_match, void *stream);

/* Candidate-list form used by the projection / BoW searches: query i scans train rows
 * cand[offsets[i] .. offsets[i+1]) in list order.  Host pointers; idx1 is a train row or -1. */
int orbgpu_hamming_top2_csr(const uint8_t *queries, int nq, const uint8_t *train, int nt,
                            const int32_t *offsets, const int32_t *cand, int32_t *idx1, int32_t *d1,
                            int32_t *d2, int device);

/* ------------------------------------------------------------------------------------------------
 * Device-resident feature sets and the ORBmatcher searches
 *
 * Every search of src/ORBmatcher.cc has the same shape: for each query (a projected MapPoint, a
 * keypoint of another frame) enumerate candidate keypoints — the window walk of
 * Frame::GetFeaturesInArea (src/Frame.cc:341-397) / KeyFrame::GetFeaturesInArea
 * (src/KeyFrame.cc:526-567) over the 64x48 grid of Frame::AssignFeaturesToGrid (src/Frame.cc:250-263,
 * 399-411), or the features of one vocabulary node — gate them, keep the best / second-best
 * Hamming distance, accept by threshold and ratio, and in six of the eleven routines let an
 * accepted match take its keypoint away from LATER queries.  One feature set = the undistorted
 * keypoints, right coordinates, descriptors and grid of one Frame / KeyFrame, resident in HBM;
 * orbgpu_search runs one whole routine's candidate loops on it.  What stays on the host (shim/
 * ORBmatcher.cc) is what the reference computes per MapPoint BEFORE its candidate loop (projection,
 * frustum and distance gates, PredictScale) and the pointer-graph updates AFTER it.
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbgpu_features orbgpu_features;

typedef struct orbgpu_features_desc {
  int32_t n;                          /* Frame::N */
  const orbgpu_keypoint *keys_un;     /* Frame::mvKeysUn (x, y, octave, angle are read) */
  const float *u_right;               /* Frame::mvuRight, or NULL for "all -1" (monocular) */
  const uint8_t *descriptors;         /* Frame::mDescriptors, [n][32] */
  float min_x, min_y;                 /* Frame::mnMinX / mnMinY used by PosInGrid (src/Frame.cc:399-411) */
  float grid_w_inv, grid_h_inv;       /* Frame::mfGridElementWidthInv / HeightInv */
  int32_t grid_cols, grid_rows;       /* FRAME_GRID_COLS / ROWS (64 x 48, include/Frame.h:35-36) */
  float query_min_x, query_min_y;     /* mnMinX / mnMinY as the window walk reads them: the same floats for a Frame,
                                         the truncated ints of include/KeyFrame.h:190-193 for a KeyFrame */
  int32_t device;                     /* CUDA device, -1 = current */
} orbgpu_features_desc;

/* Uploads one frame's features and builds its grid on the device (cell lists in ascending keypoint
 * index, exactly the push_back order of AssignFeaturesToGrid). */
int orbgpu_features_create(const orbgpu_features_desc *desc, orbgpu_features **out);
/* The same from the device-resident output of the last extract_batch call on `e` (image `image`),
 * device to device: keypoints and descriptors are taken from the extractor, u_right from the stereo
 * results of orbgpu_stereo_match when `pair` >= 0 (else all -1).  Valid only for undistorted input
 * (mDistCoef[0] == 0: Frame::UndistortKeyPoints copies mvKeys, src/Frame.cc:419-423). */
int orbgpu_features_from_extractor(orbgpu_extractor *e, int image, int pair, const orbgpu_features_desc *geometry,
                                   orbgpu_features **out);
/* Pinhole + radial-tangential camera as the reference's settings files give it (Camera.fx .. Camera.k3, e.g.
 * Examples/Monocular/TUM1.yaml:8-17): mK and mDistCoef of src/Frame.cc, float like there. */
typedef struct orbgpu_camera {
  float fx, fy, cx, cy;
  float dist[5];                      /* k1 k2 p1 p2 k3 (k3 = 0 when the file has none) */
  int32_t ndist;                      /* 4 or 5 */
} orbgpu_camera;
/* orbgpu_features_from_extractor for DISTORTED input: Frame::UndistortKeyPoints (src/Frame.cc:419-445) runs on the device
 * copy of the keypoints — cv::undistortPoints(mvKeys, mK, mDistCoef, R = I, P = mK), OpenCV 4.x's five double-precision
 * fixed-point iterations, bit-exact against cv2 — before the grid is built, so distorted monocular / RGB-D frames need no
 * D2H -> host -> H2D round trip between extraction and the first search.  dist[0] == 0 copies (Frame.cc:420-423). */
int orbgpu_features_from_extractor_undistorted(orbgpu_extractor *e, int image, int pair, const orbgpu_features_desc *geometry,
                                               const orbgpu_camera *camera, orbgpu_features **out);
/* The same operation on host arrays: `keys_un` receives the records of `keys` with pt replaced (may alias `keys`);
 * orbgpu_undistort_points takes bare (x, y) pairs — the four image corners of Frame::ComputeImageBounds
 * (src/Frame.cc:447-475) go through it. */
int orbgpu_undistort_keypoints(const orbgpu_keypoint *keys, int n, const orbgpu_camera *camera, orbgpu_keypoint *keys_un, int device);
int orbgpu_undistort_points(const float *xy, int n, const orbgpu_camera *camera, float *xy_undistorted, int device);
void orbgpu_features_destroy(orbgpu_features *f);
int orbgpu_features_count(const orbgpu_features *f);
/* Grid as CSR: cell (ix, iy) is row ix * grid_rows + iy (mGrid[ix][iy]); cell_start has cols*rows+1 entries,
 * cell_items `n` entries (keypoints outside the grid appear in no cell; *n_items receives the number stored). */
int orbgpu_features_grid(orbgpu_features *f, int32_t *cell_start, int32_t *cell_items, int32_t *n_items);
/* Download (for callers that built the set from an extractor). */
int orbgpu_features_download(orbgpu_features *f, orbgpu_keypoint *keys_un, float *u_right, uint8_t *descriptors);

enum {  /* orbgpu_search_policy.source */
  ORBGPU_SRC_WINDOW = 0,      /* GetFeaturesInArea(u, v, r, min_level, max_level) */
  ORBGPU_SRC_LIST = 1         /* cand[list_begin .. list_end) in list order (one vocabulary node) */
};
enum {  /* orbgpu_search_policy.ratio_mode */
  ORBGPU_RATIO_NONE = 0,
  ORBGPU_RATIO_SAME_LEVEL = 1, /* reject if level1 == level2 && d1 > ratio * d2            (ORBmatcher.cc:118-120) */
  ORBGPU_RATIO_LESS = 2        /* accept only if (float)d1 < ratio * (float)d2             (ORBmatcher.cc:225, 454, 565) */
};
enum {  /* orbgpu_search_policy.greedy */
  ORBGPU_GREEDY_NONE = 0,      /* queries are independent */
  ORBGPU_GREEDY_TAKEN = 1,     /* an accepted match of a query with ORBGPU_Q_BLOCKS set makes its keypoint unavailable to
                                  every later query (F.mvpMapPoints[idx] / vpMatched[idx] / vbMatched2[idx] tests) */
  ORBGPU_GREEDY_DISTANCE = 2   /* SearchForInitialization: a keypoint matched earlier at distance d is unavailable to a later
                                  query whose distance to it is >= d (vMatchedDistance, ORBmatcher.cc:447) */
};
enum {  /* orbgpu_search_policy.gate */
  ORBGPU_GATE_NONE = 0,
  ORBGPU_GATE_FUSE = 1,        /* chi-square reprojection gates of Fuse (ORBmatcher.cc:853-876): needs inv_level_sigma2 */
  ORBGPU_GATE_EPIPOLAR = 2     /* SearchForTriangulation (ORBmatcher.cc:664-699): '<=' update (last minimum wins), stereo flags, epipole distance,
                                  CheckDistEpipolarLine (:137-159): needs f12, ex, ey, level_sigma2, scale_factors */
};
#define ORBGPU_Q_BLOCKS 1      /* orbgpu_query.flags: this query's accepted match blocks later queries */
#define ORBGPU_Q_STEREO 2      /* EPIPOLAR gate: the query keypoint has a right coordinate (bStereo1) */

typedef struct orbgpu_query {
  float u, v, r;                    /* window centre / half-size; EPIPOLAR gate: (u, v) = kp1.pt */
  int32_t min_level, max_level;     /* octave < min_level skipped; max_level >= 0 && octave > max_level skipped */
  float ur, ur_tol;                 /* stereo gate (policy.stereo_gate) and FUSE gate: projected right coordinate; tolerance */
  int32_t list_begin, list_end;     /* ORBGPU_SRC_LIST */
  int32_t flags;
} orbgpu_query;                     /* 40 bytes */

typedef struct orbgpu_search_policy {
  int32_t source, ratio_mode, greedy, gate;
  int32_t accept_max;               /* accept when d1 <= accept_max (TH_HIGH, TH_LOW, TH_LOW - 1 for '<', ORBdist) */
  float ratio;                      /* ORBmatcher::mfNNratio */
  int32_t stereo_gate;              /* 1: candidates with u_right > 0 need |q.ur - u_right| <= q.ur_tol (ORBmatcher.cc:94-98, 1307-1312) */
  int32_t only_stereo;              /* EPIPOLAR gate: bOnlyStereo */
  float ex, ey;                     /* EPIPOLAR gate: epipole in the candidate image */
  float f12[9];                     /* EPIPOLAR gate: fundamental matrix, row-major */
  float level_sigma2[16], inv_level_sigma2[16], scale_factors[16];
} orbgpu_search_policy;

typedef struct orbgpu_match {
  int32_t idx;                      /* best candidate keypoint, -1 if none */
  int32_t dist, dist2;              /* best / second-best distance (256 = none) */
  int8_t level, level2;             /* their octaves (-1 = none) */
  int8_t accepted;                  /* threshold + ratio test passed */
  int8_t pad;
} orbgpu_match;                     /* 16 bytes */

/* Runs the candidate loops of one ORBmatcher routine: queries [nq] (host), their descriptors
 * [nq][32] (host), `occupied` [features.n] bytes or NULL (keypoints unavailable from the start),
 * `cand` [ncand] for ORBGPU_SRC_LIST.  Query order is the reference's loop order; the result is
 * exactly what the sequential loops produce (greedy modes are resolved on the device by iterating
 * the parallel search to its unique fixed point).  *iterations (optional) receives the passes used. */
int orbgpu_search(orbgpu_features *f, const orbgpu_query *queries, const uint8_t *query_descriptors, int nq,
                  const uint8_t *occupied, const int32_t *cand, int ncand, const orbgpu_search_policy *policy,
                  orbgpu_match *result, int *iterations);
/* Batched GetFeaturesInArea: result as CSR (offsets [nq+1], items up to `cap`); returns
 * ORBGPU_E_CAPACITY (with offsets filled) if `cap` is too small. */
int orbgpu_features_in_area(orbgpu_features *f, const orbgpu_query *queries, int nq, int32_t *offsets,
                            int32_t *items, int cap);

/* ------------------------------------------------------------------------------------------------
 * DBoW2 vocabulary tree descent
 *
 * Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:412-417, src/KeyFrame.cc:63-72) call
 * TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup = 4), whose cost is the
 * per-feature descent of Source/ThirdParty/DBoW2/DBoW2-local/include/DBoW2/TemplatedVocabulary.h:1216-1258:
 * at every level FORB::distance (…/src/FORB.cpp:77-100, a 256-bit popcount) to each of the k children,
 * strict '<' (first child wins ties), until a leaf; the word is the leaf's word id and the
 * FeatureVector node is the node passed at level L - levelsup.  The tree lives in HBM; the
 * BowVector / FeatureVector std::maps and their normalisation stay host-side DBoW2 code
 * (shim/ORBVocabularyGPU.h).
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbgpu_vocabulary orbgpu_vocabulary;
/* Flat tree as DBoW2 holds it: node 0 is the root, node i's children are
 * children[child_start[i] .. child_start[i+1]) in m_nodes[i].children order, descriptors [nodes][32],
 * word_id[i] = DBoW2 word id of a leaf (a node without children), ignored for inner nodes.
 * depth_levels = m_L. */
int orbgpu_vocabulary_create(int32_t nodes, const int32_t *child_start, const int32_t *children,
                             const uint8_t *descriptors, const int32_t *word_id, int32_t depth_levels, int device,
                             orbgpu_vocabulary **out);
void orbgpu_vocabulary_destroy(orbgpu_vocabulary *v);
/* n host descriptors [n][32] -> word_ids[n], node_ids[n].  node_ids[i] = 0 when depth_levels - levelsup <= 0
 * (TemplatedVocabulary.h:1225); when the descent reaches a leaf ABOVE that level the reference leaves
 * its output uninitialised — here the leaf's own node id is returned. */
int orbgpu_vocabulary_transform(orbgpu_vocabulary *v, const uint8_t *descriptors, int n, int levelsup,
                                int32_t *word_ids, int32_t *node_ids);
/* The same for the descriptors of a device-resident feature set (no upload). */
int orbgpu_vocabulary_transform_features(orbgpu_vocabulary *v, orbgpu_features *f, int levelsup, int32_t *word_ids,
                                         int32_t *node_ids);

/* ------------------------------------------------------------------------------------------------
 * Input staging: what the reference does to an image right before
 * the extractor sees it, on the device, so colour / unrectified frames need only one H2D.
 *   cvtColor(…, COLOR_{RGB,BGR,RGBA,BGRA}2GRAY)   src/Tracking.cc:166-182, 200-211, 228-236
 *   cv::remap(im, imRect, M1, M2, INTER_LINEAR)    Source/Examples/Stereo/stereo_euroc.cc:159-160
 *     (M1 / M2 = CV_32F maps of cv::initUndistortRectifyMap, :109-111, computed once on the host)
 * Arithmetic (OpenCV's, pinned bit-exact against cv2 4.13.0 in tests/test_staging.py):
 *   gray  = (B*3735 + G*19235 + R*9798 + 2^14) >> 15
 *   remap: sx = cvRound(mapx*32), ix = sx >> 5, fx = sx & 31 (same in y); weights (32-fx)(32-fy)*32, …
 *          out = (sum w*p + 2^14) >> 15, taps outside the source read 0 (BORDER_CONSTANT).
 * All pointers are DEVICE pointers; `stream` is a cudaStream_t (NULL = the legacy default stream).
 * Images are [batch] planes `*_img_step` bytes apart with `*_step` bytes per row.
 * ------------------------------------------------------------------------------------------------ */
enum { ORBGPU_BGR2GRAY = 0, ORBGPU_RGB2GRAY = 1 };   /* 4-channel input: the alpha byte is ignored */
int orbgpu_cvt_gray_device(const uint8_t *d_src, int channels, int code, int width, int height, size_t src_step,
                           size_t src_img_step, uint8_t *d_dst, size_t dst_step, size_t dst_img_step, int batch,
                           void *stream);
int orbgpu_remap_device(const uint8_t *d_src, int src_width, int src_height, size_t src_step, size_t src_img_step,
                        const float *d_map_x, const float *d_map_y, int width, int height, uint8_t *d_dst,
                        size_t dst_step, size_t dst_img_step, int batch, void *stream);
/* Host-buffer conveniences (upload, run, download; contiguous images). */
int orbgpu_cvt_gray(const uint8_t *src, int channels, int code, int width, int height, uint8_t *dst, int batch, int device);
int orbgpu_remap(const uint8_t *src, int src_width, int src_height, const float *map_x, const float *map_y, int width,
                 int height, uint8_t *dst, int batch, int device);

/* ------------------------------------------------------------------------------------------------
 * MapPoint::ComputeDistinctiveDescriptors, batched
 *
 * src/MapPoint.cc:229-321: among the N descriptors observing a MapPoint take the one whose MEDIAN
 * Hamming distance to all N (itself included, distance 0) is least — median = sorted[(size_t)(0.5*(N-1))],
 * first minimum wins (:301-309).  One MapPoint is ~N^2/2 popcounts, far below the cost of a device call, so the
 * entry point takes MANY MapPoints at once (LocalMapping::ProcessNewKeyFrame / CreateNewMapPoints update hundreds
 * per keyframe): group g owns rows [offsets[g], offsets[g+1]) of `descriptors` ([rows][32], host), in the order the
 * reference pushes them (iteration order of mObservations, bad keyframes skipped).
 * best_index[g] = winning row relative to the group (-1 for an empty group), best_median[g] its median.
 * ------------------------------------------------------------------------------------------------ */
int orbgpu_distinctive_descriptors(const uint8_t *descriptors, const int32_t *offsets, int groups, int32_t *best_index,
                                   int32_t *best_median, int device);

/* ------------------------------------------------------------------------------------------------
 * KeyFrameDatabase candidate scoring (second half)
 *
 * src/KeyFrameDatabase.cc:72-306 (DetectLoopCandidates, DetectRelocalizationCandidates): walk the inverted file over
 * the query's words to count, per keyframe, the words it shares with the query (mnLoopWords / mnRelocWords), keep
 * the keyframes above 0.8 x the best count, and score each against the query with DBoW2's L1 score
 * (ThirdParty/DBoW2/DBoW2-local/src/ScoringObject.cpp:23-67).  Here the database keeps every keyframe's BowVector
 * resident in HBM and ONE query call returns, for every slot at once: the shared-word count, the smallest shared word
 * id (0xffffffff if none) — with the slot number it orders the keyframes exactly as the inverted-file walk meets them
 * first — and the L1 score as the double DBoW2 computes (terms added in ascending word order), which the caller narrows
 * to float like `float si = mpVoc->score(...)`.  The covisibility accumulation that follows (:140-166, :248-275) walks
 * the KeyFrame pointer graph and stays on the host (shim/KeyFrameDatabase.cc).
 * Slots are handed out in add() order and never reused (that order IS the push_back order of each inverted-file list);
 * erase marks a slot dead; clear drops everything.  Word ids must be strictly ascending (a BowVector is a std::map).
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbgpu_bowdb orbgpu_bowdb;
int orbgpu_bowdb_create(int device, orbgpu_bowdb **out);
void orbgpu_bowdb_destroy(orbgpu_bowdb *db);
/* KeyFrameDatabase::add (:38-46): returns the new slot (>= 0) or a negative error code. */
int orbgpu_bowdb_add(orbgpu_bowdb *db, const uint32_t *word_ids, const double *weights, int n);
int orbgpu_bowdb_erase(orbgpu_bowdb *db, int slot);           /* KeyFrameDatabase::erase (:48-66) */
int orbgpu_bowdb_clear(orbgpu_bowdb *db);                     /* KeyFrameDatabase::clear (:68-71) */
int orbgpu_bowdb_slots(const orbgpu_bowdb *db);               /* slots handed out so far (dead ones included) */
/* Result arrays hold `capacity` >= orbgpu_bowdb_slots(db) entries; dead slots report 0 / 0xffffffff / 0. */
int orbgpu_bowdb_query(orbgpu_bowdb *db, const uint32_t *word_ids, const double *weights, int n, int32_t *common_words,
                       uint32_t *first_common_word, double *score, int capacity);

/* Parity hook for computeOrbDescriptor's cos / sin (src/ORBextractor.cc:102-106): evaluates the DEVICE sincos of the
 * descriptor kernel (glibc's sincosf algorithm, see csrc/describe.cu) on the `count` consecutive float32 bit patterns from
 * `first_bits`.  sums (optional, host, one uint64 per 2^chunk_log2 patterns): sum over the chunk of
 * ((sin bits << 32) | cos bits) * (0x9E3779B97F4A7C15 + 2 * pattern) mod 2^64 — the host forms the same sum with libm, so
 * every float of [0, 2 pi] is compared without moving it; sin_values / cos_values (optional, count <= 2^26): the raw values. */
int orbgpu_selftest_sincos(uint32_t first_bits, uint64_t count, int chunk_log2, uint64_t *sums, float *sin_values, float *cos_values, int device);
/* Measured integer-pipe peak (register-resident POPC loop): returns POPC lane-ops per second. */
int orbgpu_measure_popc_peak(int device, double *popc_per_second, double *lop3_per_second);
/* Issue rate of one instruction kind in lane-operations per second (register-resident loop): mode 0 POPC, 1 LOP3,
 * 2 VIMNMX3.S16x2, 3 VIMNMX.S16x2, 4 VIADD.16x2, 5 PRMT, 6 IMAD, 7 VIMNMX3 (s32), 8 IADD3. */
int orbgpu_measure_pipe(int device, int mode, double *ops_per_second);

/* Per-stage device timing (CUDA events on the handle's stream, recorded only while enabled).
 * Stages: 0 pyramid (import + 7 resizes), 1 FAST cells, 2 octree, 3 blur, 4 orientation+descriptor,
 * 5 stereo (match + median filter; recorded on the LEFT handle).  orbgpu_stage_times synchronises the
 * stream and writes each stage's mean milliseconds over the batches run since the previous call (at most
 * the last 32; 0 for a stage that did not run), then starts a new averaging window. */
#define ORBGPU_NUM_STAGES 6
int orbgpu_set_profiling(orbgpu_extractor *e, int enable);
int orbgpu_stage_times(orbgpu_extractor *e, float *milliseconds, int n);
const char *orbgpu_stage_name(int stage);

/* Kernel launches issued by this library on the calling thread since the last reset (bench's
 * gpu_launches counter). */
int64_t orbgpu_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* ORBGPU_H */
