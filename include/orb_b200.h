/* orb_b200.h — C ABI of the B200-native ORB front-end (liborb_b200.so).
 *
 * Drop-in boundary for the reference's feature front-end hot path
 * (electech6/ORB_SLAM2_detailed_comments). Each entry point names the reference
 * interface it replaces (file:line under the reference checkout). Plain pointers and
 * sizes only; no C++/torch types. All functions return an orb_status (0 = ok) and never
 * abort or throw. One CUDA stream + workspace per handle; distinct handles are
 * thread-safe against each other, a single handle is not re-entrant (like the
 * reference extractor, whose pyramid buffers are overwritten per call,
 * include/ORBextractor.h:162).
 *
 * There is NO CPU fallback: every compute entry point runs hand-written sm_100a
 * kernels and fails with ORB_ERR_CUDA when no device is usable.
 */
#ifndef ORB_B200_H_
#define ORB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orb_status {
  ORB_OK = 0,
  ORB_ERR_INVALID = 1,     /* bad argument (null pointer, non-positive size, ...) */
  ORB_ERR_CUDA = 2,        /* CUDA runtime error; see orb_last_error() */
  ORB_ERR_CAPACITY = 3,    /* an output or internal list overflowed its capacity */
  ORB_ERR_UNSUPPORTED = 4  /* geometry the reference itself cannot process (e.g. level < 2 cells) */
} orb_status;

/* Binary layout of cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct orb_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orb_keypoint;

/* ORBextractor constructor arguments (include/ORBextractor.h:93, src/ORBextractor.cc:469). */
typedef struct orb_params {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels;
  int32_t ini_th_fast;
  int32_t min_th_fast;
} orb_params;

/* One level of mvImagePyramid (include/ORBextractor.h:162): `data` points at interior pixel
 * (0,0); at least 19 readable pixels of REFLECT_101 border surround it (Frame.cc:855,967). */
typedef struct orb_level_view {
  uint8_t* data;
  int32_t width, height;
  int64_t step;
} orb_level_view;

typedef struct orb_extractor orb_extractor;

const char* orb_last_error(void);
int orb_device_count(void);

/* ---- ORBextractor ------------------------------------------------------------------- */

/* Replaces ORBextractor::ORBextractor (src/ORBextractor.cc:469-571). `max_batch` bounds the
 * frames processed per internal chunk (workspace ~5 bytes per pyramid pixel per frame). */
int orb_create(const orb_params* params, int device, int max_batch, orb_extractor** out);
int orb_destroy(orb_extractor* h);

/* Replaces GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:119-159) and exposes
 * mnFeaturesPerLevel. Any pointer may be NULL. Arrays hold nlevels entries. */
int orb_get_scale_tables(const orb_extractor* h, float* scale, float* inv_scale, float* sigma2,
                         float* inv_sigma2, int32_t* features_per_level);

/* Upper bound on keypoints per frame: size outputs with this. Per level the quadtree returns at most
 * max(mnFeaturesPerLevel + 3, 4 * nIni) keypoints, nIni = round(width' / height') roots (src/ORBextractor.cc:695).
 * orb_max_keypoints: exact for the image size the handle last processed; before the first image a bound valid for
 * aspect ratios up to 8:1. orb_max_keypoints_for_size: exact for the given image size (0 if the size is unsupported). */
int orb_max_keypoints(const orb_extractor* h);
int orb_max_keypoints_for_size(const orb_extractor* h, int width, int height);

/* Replaces ORBextractor::operator()(image, mask, keypoints, descriptors)
 * (src/ORBextractor.cc:1533-1649) for one frame in HOST memory. `image` is CV_8UC1 with row
 * stride `step`; the mask argument of the reference is ignored there and absent here.
 * Writes *n keypoints (ascending octave, level-0 coordinates) and n x 32 descriptor bytes.
 * Empty image (w or h == 0) returns ORB_OK with outputs untouched (:1537). If `pyramid` is
 * non-NULL it receives nlevels views into host memory owned by the handle (valid until the
 * next call), i.e. the mvImagePyramid contract. */
int orb_extract(orb_extractor* h, const uint8_t* image, int width, int height, size_t step,
                orb_keypoint* keypoints, int capacity, int* n, uint8_t* descriptors,
                orb_level_view* pyramid);

/* Batch of B same-sized frames in HOST memory (frame b at images + b*frame_stride, rows at
 * `step`). Outputs are strided by `capacity`: keypoints[b*capacity + i], descriptors
 * [(b*capacity + i)*32], counts[b]. Host<->device copies happen inside. */
int orb_extract_batch_host(orb_extractor* h, const uint8_t* images, int batch, int width, int height,
                           size_t step, size_t frame_stride, orb_keypoint* keypoints, int capacity,
                           int32_t* counts, uint8_t* descriptors);

/* The same, but returns as soon as the copies and kernels are enqueued: consecutive calls continue one
 * upload / compute / download pipeline (the upload of call k+1 overlaps the kernels of call k). The host
 * buffers of a call must stay valid, and its outputs must not be read, until orb_synchronize(h, NULL)
 * returns; capacity overflow is reported there. */
int orb_extract_batch_host_async(orb_extractor* h, const uint8_t* images, int batch, int width, int height,
                           size_t step, size_t frame_stride, orb_keypoint* keypoints, int capacity,
                           int32_t* counts, uint8_t* descriptors);

/* Same, with inputs and outputs RESIDENT IN DEVICE MEMORY of the handle's device. `stream`
 * is a cudaStream_t (NULL = the handle's own stream); the call is asynchronous on it. */
int orb_extract_batch_device(orb_extractor* h, const uint8_t* d_images, int batch, int width, int height,
                             size_t step, size_t frame_stride, orb_keypoint* d_keypoints, int capacity,
                             int32_t* d_counts, uint8_t* d_descriptors, void* stream);

/* Stereo front-end of Frame::Frame(imLeft, imRight, ...) (src/Frame.cc:121-158): extraction of both
 * eyes (the reference runs two extractor threads, :146-154) followed by
 * Frame::ComputeStereoMatches (src/Frame.cc:831-1082): row-band candidates, best Hamming within one
 * octave and the disparity range [0, mbf/mb), 11x11 SAD slide of +-5 px on the keypoint's pyramid
 * level with parabola sub-pixel fit, 1.5*1.4*median SAD outlier cut. uright / depth receive
 * mvuRight / mvDepth for the n_left left keypoints (-1 = no stereo match). The handle must have been
 * created with max_batch >= 2. */
int orb_extract_stereo(orb_extractor* h, const uint8_t* left, const uint8_t* right, int width, int height,
                       size_t step, float mbf, float mb, orb_keypoint* kps_left, int capacity, int* n_left,
                       uint8_t* desc_left, orb_keypoint* kps_right, int* n_right, uint8_t* desc_right,
                       float* uright, float* depth);

/* Frame::ComputeStereoMatches (src/Frame.cc:831-1082) for the reference's own call sequence: the stereo Frame constructor
 * has run one orb_extract per eye on two handles (from two threads, src/Frame.cc:146-154) and joined them; the pyramids,
 * keypoints and descriptors of both calls are still on the device. This call pairs them (the right eye's results are copied
 * device-to-device into the left handle's second frame slot) and writes mvuRight / mvDepth for the left keypoints (-1 = no
 * stereo match); *n_left (may be NULL) = number of left keypoints. Both handles: same device, parameters, image size and
 * capacity; `left` created with max_batch >= 2; no other call on either handle in between. */
int orb_stereo_match(orb_extractor* left, orb_extractor* right, float mbf, float mb, float* uright, float* depth, int* n_left);

/* Same for `pairs` stereo pairs RESIDENT IN DEVICE MEMORY, frames interleaved L0,R0,L1,R1,...;
 * d_keypoints / d_descriptors / d_counts as in orb_extract_batch_device over 2*pairs frames,
 * d_uright / d_depth are pairs x capacity floats. Asynchronous on `stream`. */
int orb_extract_stereo_batch_device(orb_extractor* h, const uint8_t* d_images, int pairs, int width, int height,
                                    size_t step, size_t frame_stride, orb_keypoint* d_keypoints, int capacity,
                                    int32_t* d_counts, uint8_t* d_descriptors, float mbf, float mb,
                                    float* d_uright, float* d_depth, void* stream);

/* Blocks until the handle's work (on `stream`, NULL = own stream) is done; reports sticky
 * capacity overflows of internal candidate lists as ORB_ERR_CAPACITY. */
int orb_synchronize(orb_extractor* h, void* stream);

/* Number of kernel launches the last batch call issued (for bench accounting). */
int orb_last_launch_count(const orb_extractor* h);

/* Host wall-clock breakdown of the last orb_extract call, microseconds: us4 = {staging copy of the image into pinned
 * memory, enqueue (H2D + CUDA-graph launch + D2H requests), waiting for the device, copying the results out}. */
int orb_last_call_breakdown(const orb_extractor* h, double* us4);

/* Per-stage device timing with CUDA events recorded on the launching stream between the
 * stages of every chunk (no host synchronisation is added). orb_get_stage_times synchronises,
 * returns the milliseconds and launch counts accumulated since the previous call for the 5
 * stages {pyramid, fast, quadtree, blur, describe} and resets them. */
int orb_set_profiling(orb_extractor* h, int enable);

/* Number of workspace lanes (1 or 2, default 2; env ORB_B200_LANES overrides the default): with 2,
 * consecutive chunks of a batch call run on two internal streams so that kernels of neighbouring
 * chunks can overlap. Results are identical; per-stage timings are only meaningful with 1 lane. */
int orb_set_lanes(orb_extractor* h, int lanes);
int orb_get_stage_times(orb_extractor* h, double* ms5, long long* launches5);

/* Stage outputs of the last call, for parity tests (frame index inside the last chunk).
 *   level geometry:           orb_stage_level_size
 *   bordered pyramid level:   (h+38) x (w+38) tightly packed
 *   blurred level:            h x w tightly packed
 *   FAST candidates entering the quadtree (unordered): x, y (level image coords), score
 *   kept keypoints of a level in list order: x, y, score */
int orb_stage_level_size(const orb_extractor* h, int level, int* width, int* height);
int orb_stage_copy_level(orb_extractor* h, int frame, int level, uint8_t* dst);
int orb_stage_copy_blur(orb_extractor* h, int frame, int level, uint8_t* dst);
int orb_stage_copy_candidates(orb_extractor* h, int frame, int level, int32_t* xs, int32_t* ys,
                              int32_t* score, int capacity, int* n);
int orb_stage_copy_kept(orb_extractor* h, int frame, int level, int32_t* xs, int32_t* ys,
                        int32_t* score, int capacity, int* n);

/* ---- Frame: undistortion and the 64x48 keypoint grid (device resident) ------------------ */

/* mK / mDistCoef of Frame (src/Frame.cc:88-118 read them from the settings file). */
typedef struct orb_camera {
  float fx, fy, cx, cy;
  float k1, k2, p1, p2, k3;
} orb_camera;

/* One Frame::GetFeaturesInArea call (src/Frame.cc:590): keypoints of `frame` within the circular
 * window of radius r around (x, y), octave in [min_level, max_level] (max_level < 0: no upper bound). */
typedef struct orb_area_query {
  int32_t frame;
  float x, y, r;
  int32_t min_level, max_level;
} orb_area_query;

/* Replaces Frame::ComputeImageBounds (src/Frame.cc:779-829): bounds4 = {mnMinX, mnMaxX, mnMinY,
 * mnMaxY}. Host side (four points); same arithmetic as the device path. */
int orb_compute_image_bounds(const orb_camera* cam, int width, int height, float* bounds4);

/* Replaces Frame::UndistortKeyPoints (src/Frame.cc:724-776), i.e. cv::undistortPoints(K, dist, P=K),
 * for `batch` frames of extraction output (capacity-strided). Only pt changes; identity if k1 == 0. */
int orb_undistort_keypoints_device(int device, const orb_keypoint* d_keypoints, const int32_t* d_counts, int batch,
                                   int capacity, const orb_camera* cam, orb_keypoint* d_keypoints_un, void* stream);

/* Replaces Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:399-423, 682-698): mGrid as CSR.
 * d_cell_start: batch x 3073 ints (cell = ix*48 + iy), d_cell_items: batch x capacity keypoint
 * indices, in insertion order inside every cell. */
int orb_assign_features_to_grid_device(int device, const orb_keypoint* d_keypoints_un, const int32_t* d_counts, int batch,
                                       int capacity, const float* bounds4, int32_t* d_cell_start,
                                       int32_t* d_cell_items, void* stream);

/* Replaces Frame::GetFeaturesInArea (src/Frame.cc:590-670) for n_queries queries: d_out
 * (n_queries x out_capacity) receives the indices in the reference's order (ix outer, iy inner,
 * cell order), d_out_counts the number found (may exceed out_capacity: then the list is truncated). */
int orb_get_features_in_area_device(int device, const orb_keypoint* d_keypoints_un, int capacity, const float* bounds4,
                                    const int32_t* d_cell_start, const int32_t* d_cell_items,
                                    const orb_area_query* d_queries, int n_queries, int32_t* d_out, int out_capacity,
                                    int32_t* d_out_counts, void* stream);

/* ---- ORBmatcher --------------------------------------------------------------------- */

/* Replaces ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103): 256-bit Hamming
 * distance of two 32-byte descriptors. Host-side (a device round trip per call would be
 * absurd); the batched kernels below compute the same quantity on the GPU. */
int orb_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* The slice of Frame that SearchForInitialization reads (Frame.h mvKeysUn, mDescriptors,
 * grid bounds). For device entry points all pointers are device pointers. */
typedef struct orb_frame_view {
  int32_t n;
  const float* xy;            /* n x 2: mvKeysUn[i].pt */
  const int32_t* octave;      /* n */
  const float* angle;         /* n, degrees */
  const uint8_t* descriptors; /* n x 32 */
} orb_frame_view;

typedef struct orb_match_params {
  float nnratio;        /* ORBmatcher::mfNNratio (ORBmatcher.h ctor, default 0.6; 0.9 at Tracking.cc:915) */
  int32_t check_orientation; /* ORBmatcher::mbCheckOrientation */
  int32_t window;       /* windowSize (Tracking.cc:926 passes 100); ignored when mode = 1 */
  int32_t mode;         /* 0 = reference-faithful (octave-0 rows, 64x48 grid cell range + circular
                           window, Frame.cc:590-670); 1 = brute force over all of frame 2 */
  float min_x, max_x, min_y, max_y; /* Frame::mnMinX.. (undistorted image bounds) */
} orb_match_params;

typedef struct orb_matcher orb_matcher;

int orb_matcher_create(int device, int max_pairs, int max_keypoints, orb_matcher** out);
int orb_matcher_destroy(orb_matcher* m);

/* Replaces ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:573-717) for one frame
 * pair in HOST memory. prev_matched (n1 x 2 floats) is vbPrevMatched, read and updated;
 * matches12 (n1 ints, -1 = none) is vnMatches12. *nmatches is the return value of the
 * reference. best/second (n1 ints each, may be NULL) receive each row's 2-NN distances
 * (INT_MAX where no candidate). */
int orb_search_for_initialization(orb_matcher* m, const orb_frame_view* f1, const orb_frame_view* f2,
                                  const orb_match_params* mp, float* prev_matched, int32_t* matches12,
                                  int* nmatches, int32_t* best, int32_t* second);

/* Batch of P independent pairs RESIDENT IN DEVICE MEMORY, brute-force mode (mode 1), every
 * frame holding exactly n keypoints: descriptors (2P x n x 32; pair p = frames 2p, 2p+1),
 * angles (2P x n). Outputs: matches12 (P x n), nmatches (P). Asynchronous on `stream`. */
int orb_match_pairs_device(orb_matcher* m, const uint8_t* d_descriptors, const float* d_angles, int pairs,
                           int n, float nnratio, int check_orientation, int32_t* d_matches12,
                           int32_t* d_nmatches, void* stream);

/* All-pairs keyframe matching: `d_all` holds n_kf x n_desc descriptors
 * (all keyframes, e.g. after an all-gather); rows [row_begin,row_end) are this rank's
 * keyframes, [col_begin,col_end) the column keyframes to match against in this call (so that
 * peers' descriptor blocks can be consumed as they arrive). d_counts is the rank's
 * (row_end-row_begin) x n_kf int matrix; entry (i-row_begin, j) receives, per ordered keyframe
 * pair, the number of row descriptors whose 2-NN passes best <= TH_LOW and best < nnratio*second. */
int orb_match_allpairs_device(orb_matcher* m, const uint8_t* d_all, int n_kf, int n_desc, int row_begin,
                              int row_end, int col_begin, int col_end, float nnratio, int32_t* d_counts,
                              void* stream);

/* ---- multi-GPU entry of the all-pairs workload: the exchange over NCCL inside the library --------
 * One process per GPU. NCCL is bound at run time (dlopen of libnccl.so.2), so single-GPU users never need it.
 * orb_shard_range: contiguous block [begin, end) of `total` units for `rank` (blocks differ by at most one).
 * orb_nccl_unique_id / orb_nccl_comm_create / orb_nccl_comm_destroy: ncclGetUniqueId (128 bytes; call on one rank and
 * hand the bytes to the others by any means) / ncclCommInitRank / ncclCommDestroy, for hosts that do not have a
 * communicator yet. A host that already owns an ncclComm_t of the same NCCL library passes it directly.
 * orb_match_allpairs_nccl: rank's keyframes are rows [begin, end) = orb_shard_range(n_kf, rank, world); d_local_desc holds
 * their descriptors ((end-begin) x n_desc x 32), d_all is n_kf x n_desc x 32 of device scratch that receives everyone's,
 * d_counts the rank's (end-begin) x n_kf block of orb_match_allpairs_device. Every rank's block is broadcast
 * (ncclBroadcast, root = owner) on a communication stream of the matcher; the all-pairs kernel runs on `stream` for the
 * own block at once and for every peer block as soon as it has landed, so transfers overlap the compute. Asynchronous
 * on `stream` (NULL = the matcher's own stream); all ranks must call it with the same n_kf / n_desc. */
void orb_shard_range(int total, int rank, int world, int* begin, int* end);
int orb_nccl_unique_id(void* id128);
int orb_nccl_comm_create(int device, int rank, int world, const void* id128, void** comm);
int orb_nccl_comm_destroy(void* comm);
int orb_match_allpairs_nccl(orb_matcher* m, void* nccl_comm /* ncclComm_t */, int rank, int world, const uint8_t* d_local_desc,
                            int n_kf, int n_desc, float nnratio, uint8_t* d_all, int32_t* d_counts, void* stream);

/* Plain Hamming distance matrix (na x nb ints) on the device: parity aid for the kernels. */
int orb_hamming_matrix_device(orb_matcher* m, const uint8_t* d_a, int na, const uint8_t* d_b, int nb,
                              int32_t* d_out, void* stream);

int orb_matcher_synchronize(orb_matcher* m, void* stream);

/* ---- ORBmatcher: ordered candidate-set searches (tracking) ----------------------------
 * SearchByProjection (local map, src/ORBmatcher.cc:72-169; last frame, :1710-1860) and SearchByBoW
 * (keyframe -> frame, :247-420) share one structure: map points are visited IN ORDER, each takes its
 * best (and second best) candidate among the keypoints that no earlier map point occupies, an accepted
 * match occupies its keypoint. The device path scores all queries in parallel (four best candidates
 * each) and then commits them in order, re-scoring only the queries whose candidates were taken. */

/* `batch` current frames as these matchers read them. Device pointers, capacity-strided. */
typedef struct orb_device_frames {
  const orb_keypoint* keypoints_un; /* (batch, capacity)      Frame::mvKeysUn */
  const uint8_t* descriptors;       /* (batch, capacity, 32)  Frame::mDescriptors */
  const float* uright;              /* (batch, capacity)      Frame::mvuRight, NULL = monocular */
  const uint8_t* occupied;          /* (batch, capacity)      1 = mvpMapPoints[i] && Observations() > 0; NULL = none */
  const int32_t* counts;            /* (batch)                Frame::N */
  const int32_t* cell_start;        /* (batch, 3073)          from orb_assign_features_to_grid_device */
  const int32_t* cell_items;        /* (batch, capacity) */
  float bounds[4];                  /* mnMinX, mnMaxX, mnMinY, mnMaxY */
  int32_t batch, capacity;
} orb_device_frames;

/* One projected map point (32 bytes). */
typedef struct orb_proj_query {
  float u, v;       /* projection in the current frame (mTrackProjX/Y, or u/v of ORBmatcher.cc:1752-1753) */
  float radius;     /* window radius r * mvScaleFactors[level]; also the tolerance on the right coordinate */
  float ur;         /* projected right-image coordinate (mTrackProjXR, or u - mbf*invz) */
  float angle;      /* angle of the source keypoint, for the rotation histogram */
  int32_t min_level, max_level;  /* GetFeaturesInArea level range (max_level < 0: open) */
  int32_t flags;    /* bit 0: query is live (mbTrackInView && !isBad / projects inside the image);
                       bit 1: its map point has Observations() > 0, i.e. a match occupies the keypoint */
} orb_proj_query;

enum {
  ORB_SEARCH_BEST = 0,        /* last frame (:1807): bestDist <= th */
  ORB_SEARCH_RATIO_LEVEL = 1, /* local map (:155-158): bestDist <= th && !(bestLevel == bestLevel2 && bestDist > ratio*bestDist2) */
  ORB_SEARCH_RATIO = 2        /* BoW (:331-336): bestDist <= th && bestDist < ratio*bestDist2 */
};

typedef struct orb_search_params {
  int32_t mode;               /* ORB_SEARCH_* */
  int32_t th;                 /* TH_HIGH = 100 / TH_LOW = 50 (src/ORBmatcher.cc:47-49) */
  float nn_ratio;             /* mfNNratio */
  int32_t check_orientation;  /* mbCheckOrientation: rotation histogram + ComputeThreeMaxima */
} orb_search_params;

/* Bytes of device scratch the two searches below need (pass the same geometry). */
size_t orb_search_scratch_bytes(int batch, int query_capacity, int capacity);

/* The projection part of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)
 * (src/ORBmatcher.cc:1734-1775): map point i of last frame b (d_world_pos (batch, query_capacity, 3);
 * d_mp_flags bit 0 = mvpMapPoints[i] && !mvbOutlier[i], bit 1 = Observations() > 0) is moved by
 * d_Tcw[b] (row-major 4x4) and projected with cam4 = {fx, fy, cx, cy}; radius = th*scale_factors[octave
 * of mvKeys[i]]; d_direction[b]: 0 = levels octave+-1, 1 = bForward, 2 = bBackward (the caller
 * evaluates :1723-1730 once per frame). Writes one orb_proj_query per map point. */
int orb_project_last_frame_device(int device, const float* d_world_pos, const uint8_t* d_mp_flags,
                                  const orb_keypoint* d_last_keypoints, const int32_t* d_last_counts, int batch,
                                  int query_capacity, const float* d_Tcw, const int32_t* d_direction, const float* cam4,
                                  const float* bounds4, float mbf, float th, const float* scale_factors, int nlevels,
                                  orb_proj_query* d_queries, void* stream);

/* Replaces the search loops of ORBmatcher::SearchByProjection (local map :72-169 with
 * ORB_SEARCH_RATIO_LEVEL, last frame :1776-1860 with ORB_SEARCH_BEST). d_queries / d_query_descriptors
 * ((batch, query_capacity[, 32]), the map points' descriptors) are visited in index order.
 * Outputs: d_match_of_keypoint (batch, capacity) = index of the query assigned to each keypoint
 * (CurrentFrame.mvpMapPoints) or -1; d_match_of_query (batch, query_capacity) = keypoint or -1;
 * d_nmatches (batch) = the function's return value. */
int orb_search_by_projection_device(int device, const orb_device_frames* frames, const orb_proj_query* d_queries,
                                    const uint8_t* d_query_descriptors, const int32_t* d_query_counts,
                                    int query_capacity, const orb_search_params* params, void* d_scratch,
                                    int32_t* d_match_of_keypoint, int32_t* d_match_of_query, int32_t* d_nmatches,
                                    void* stream);

/* Host-memory form of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1710-1860) for ONE
 * frame pair - what a drop-in body of that method calls once per tracked frame (compat/orb_b200_matcher.cpp). All pointers are
 * HOST pointers; one packed upload, orb_assign_features_to_grid_device + orb_project_last_frame_device +
 * orb_search_by_projection_device(ORB_SEARCH_BEST), one packed download, one synchronisation (staging is cached per host
 * thread). match_of_keypoint (n_cur) = index of the last-frame map point given to each current keypoint or -1
 * (CurrentFrame.mvpMapPoints), *nmatches = the method's return value. */
typedef struct orb_last_frame_search {
  int32_t n_cur;
  const orb_keypoint* cur_keypoints_un;  /* CurrentFrame.mvKeysUn */
  const uint8_t* cur_descriptors;        /* CurrentFrame.mDescriptors (n_cur x 32) */
  const float* cur_uright;               /* CurrentFrame.mvuRight, NULL = monocular */
  const uint8_t* cur_occupied;           /* 1 = CurrentFrame.mvpMapPoints[i] && Observations() > 0 (:1795), NULL = none */
  const float* bounds4;                  /* mnMinX, mnMaxX, mnMinY, mnMaxY */
  int32_t n_last;
  const orb_keypoint* last_keypoints;    /* LastFrame.mvKeys (octave, angle) */
  const float* last_world_pos;           /* n_last x 3: LastFrame.mvpMapPoints[i]->GetWorldPos() */
  const uint8_t* last_mp_flags;          /* bit 0: map point exists and !mvbOutlier[i]; bit 1: its Observations() > 0 */
  const uint8_t* last_mp_descriptors;    /* n_last x 32: GetDescriptor() */
  const float* Tcw;                      /* CurrentFrame.mTcw, row-major 4x4 */
  int32_t direction;                     /* 0 neither, 1 bForward, 2 bBackward (:1723-1730) */
  const float* cam4;                     /* fx, fy, cx, cy */
  float mbf, th;                         /* CurrentFrame.mbf; the search radius factor th */
  const float* scale_factors; int32_t nlevels;   /* CurrentFrame.mvScaleFactors */
  int32_t th_dist;                       /* TH_HIGH */
  float nn_ratio;                        /* mfNNratio (unused by ORB_SEARCH_BEST, kept for symmetry) */
  int32_t check_orientation;             /* mbCheckOrientation */
} orb_last_frame_search;
int orb_search_by_projection_last_frame(int device, const orb_last_frame_search* args, int32_t* match_of_keypoint, int* nmatches);

/* Host-memory form of ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (src/ORBmatcher.cc:72-169,
 * the local-map search of Tracking::SearchLocalPoints) for ONE frame: `queries` are built by the caller from the MapPoint track
 * fields Frame::isInFrustum wrote (mTrackProjX / Y / XR, mnTrackScaleLevel, mTrackViewCos, mbTrackInView && !isBad(); radius as
 * :88-100), in map-point order; params = {ORB_SEARCH_RATIO_LEVEL, TH_HIGH, mfNNratio, 0}. Same staging and outputs as the
 * last-frame form. */
int orb_search_by_projection_host(int device, int n_cur, const orb_keypoint* cur_keypoints_un, const uint8_t* cur_descriptors,
                                  const float* cur_uright, const uint8_t* cur_occupied, const float* bounds4, int n_queries,
                                  const orb_proj_query* queries, const uint8_t* query_descriptors, const orb_search_params* params,
                                  int32_t* match_of_keypoint, int* nmatches);

/* Replaces ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (src/ORBmatcher.cc:247-420).
 * The DBoW2 FeatureVectors (node id -> feature indices) are passed as the node id of every feature
 * (d_node1 / d_node2, -1 = none). Keyframe side: (batch, query_capacity) features with d_usable1 = 1
 * where the feature has a good map point; frame side: `frames` (grid and uright unused; `occupied`, if given,
 * marks keypoints that are not candidates).
 * d_match_of_keypoint (batch, capacity) = keyframe feature index per frame keypoint
 * (vpMapPointMatches) or -1; d_match_of_query (batch, query_capacity) = frame keypoint or -1.
 * The same call is ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:729-880):
 * frames = keyframe 2 with occupied[i] = !vpMapPoints2[i] || isBad(), th = TH_LOW - 1 (that variant tests
 * bestDist1 < TH_LOW), vpMatches12 = d_match_of_query. */
int orb_search_by_bow_device(int device, const orb_keypoint* d_keypoints1, const uint8_t* d_descriptors1,
                             const int32_t* d_node1, const uint8_t* d_usable1, const int32_t* d_counts1,
                             int query_capacity, const orb_device_frames* frames, const int32_t* d_node2,
                             const orb_search_params* params, void* d_scratch, int32_t* d_match_of_keypoint,
                             int32_t* d_match_of_query, int32_t* d_nmatches, void* stream);

/* Host-memory form of orb_search_by_bow_device for ONE pair (all pointers HOST pointers; one packed upload, the kernels, one
 * packed download, one synchronisation): what the drop-in bodies of ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)
 * (src/ORBmatcher.cc:247-420: occupied2 = NULL, th = TH_LOW, vpMapPointMatches from match_of_keypoint2) and
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (:729-880: occupied2[i] = !vpMapPoints2[i] || isBad(), th = TH_LOW - 1,
 * vpMatches12 from match_of_query1) call. node1 / node2 = DBoW2 node id of every feature (mFeatVec inverted), -1 = none. */
int orb_search_by_bow_host(int device, int n1, const orb_keypoint* keypoints1_un, const uint8_t* descriptors1, const int32_t* node1,
                           const uint8_t* usable1, int n2, const orb_keypoint* keypoints2_un, const uint8_t* descriptors2,
                           const int32_t* node2, const uint8_t* occupied2, const orb_search_params* params, int32_t* match_of_keypoint2,
                           int32_t* match_of_query1, int* nmatches);

/* Per keyframe pair of SearchForTriangulation. */
typedef struct orb_triangulation_pair {
  float F12[9];        /* fundamental matrix, row-major (F12.at<float>(r, c) = F12[3*r + c]) */
  float ex, ey;        /* epipole: keyframe 1's camera centre projected into keyframe 2 (src/ORBmatcher.cc:897-901) */
  int32_t only_stereo; /* bOnlyStereo */
} orb_triangulation_pair;

/* Replaces ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
 * (src/ORBmatcher.cc:884-1100) for `batch` keyframe pairs: features of keyframe 1 WITHOUT a map point
 * (d_has_mappoint1 == 0) are matched inside their vocabulary node against keyframe-2 features without a map
 * point (frames2->occupied = has a map point), best distance <= TH_LOW, later candidate wins ties, epipole test
 * for monocular pairs, CheckDistEpipolarLine (:205-227) with level_sigma2 of keyframe 2's octave;
 * frames2->uright / d_uright1 = mvuRight (>= 0: stereo observation; NULL = monocular). A matched keyframe-2
 * feature is taken for later features (vbMatched2), rotation consistency as in the other searches.
 * d_matches12 (batch, query_capacity) = vMatches12 (keyframe-2 index or -1), d_nmatches = return value. */
int orb_search_for_triangulation_device(int device, const orb_keypoint* d_keypoints1, const uint8_t* d_descriptors1,
                                        const int32_t* d_node1, const uint8_t* d_has_mappoint1, const float* d_uright1,
                                        const int32_t* d_counts1, int query_capacity, const orb_device_frames* frames2,
                                        const int32_t* d_node2, const orb_triangulation_pair* d_pairs,
                                        const float* scale_factors, const float* level_sigma2, int nlevels,
                                        int check_orientation, void* d_scratch, int32_t* d_matches12, int32_t* d_nmatches,
                                        void* stream);

/* Host-memory form of orb_search_for_triangulation_device for ONE keyframe pair (the drop-in body of
 * ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:884-1100); pair / scale_factors / level_sigma2 are host pointers too. */
int orb_search_for_triangulation_host(int device, int n1, const orb_keypoint* keypoints1_un, const uint8_t* descriptors1,
                                      const int32_t* node1, const uint8_t* has_mappoint1, const float* uright1, int n2,
                                      const orb_keypoint* keypoints2_un, const uint8_t* descriptors2, const int32_t* node2,
                                      const uint8_t* has_mappoint2, const float* uright2, const orb_triangulation_pair* pair,
                                      const float* scale_factors, const float* level_sigma2, int nlevels, int check_orientation,
                                      int32_t* matches12, int* nmatches);

/* ---- Input stage and map-point descriptors (the callers either side of the path) ------ */

enum { ORB_RGB2GRAY = 0, ORB_BGR2GRAY = 1, ORB_RGBA2GRAY = 2, ORB_BGRA2GRAY = 3 };

/* Replaces cv::cvtColor(im, im, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) of
 * Tracking::GrabImage{Stereo,RGBD,Monocular} (src/Tracking.cc:250-276, 310-324, 369-383) for `batch`
 * interleaved 8-bit frames on the device (OpenCV 4.x 15-bit coefficients, bit-exact vs cv2). */
int orb_cvt_color_gray_device(int device, const uint8_t* d_src, int width, int height, size_t src_step,
                              size_t src_frame_stride, int batch, int code, uint8_t* d_gray, size_t gray_step,
                              size_t gray_frame_stride, void* stream);

/* Replaces cv::remap(src, dst, M1, M2, cv::INTER_LINEAR) with CV_32FC1 maps and BORDER_CONSTANT(0), the
 * stereo rectification of Examples/Stereo/stereo_euroc.cc:181-188 (maps from initUndistortRectifyMap, computed
 * once by the caller: dst_height x dst_width floats each, shared by all frames). Bit-exact vs cv2.remap. */
int orb_remap_linear_device(int device, const uint8_t* d_src, int src_width, int src_height, size_t src_step,
                            size_t src_frame_stride, int batch, const float* d_map_x, const float* d_map_y,
                            int dst_width, int dst_height, uint8_t* d_dst, size_t dst_step, size_t dst_frame_stride,
                            void* stream);

/* Replaces MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:365-448) for n_points map points at once:
 * the descriptors of map point p's observations are rows d_offsets[p] .. d_offsets[p+1]-1 of d_descriptors
 * (at most max_observations each). d_best_index[p] = the observation with the smallest median distance to all
 * observations (first wins), -1 for a map point without observations, -2 if it exceeds max_observations;
 * d_best_descriptor (n_points x 32, may be NULL) receives that descriptor (MapPoint::mDescriptor). */
int orb_distinctive_descriptors_device(int device, const uint8_t* d_descriptors, const int32_t* d_offsets, int n_points,
                                       int max_observations, int32_t* d_best_index, uint8_t* d_best_descriptor,
                                       void* stream);

/* Integer-pipe microbenchmark used for the matching roofline: independent chains of POPC
 * (what=0), LOP3 (what=1) or the matcher's own mix of 1 POPC per 4 LOP3 (what=2, reported in
 * units of (1 POPC + 4 LOP3) per second) on a full grid; returns ops/s in *ops_per_s. */
int orb_int_pipe_peak(int device, int what, double* ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H_ */
