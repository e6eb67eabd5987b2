/* include/orbfe.h -- C ABI of the B200-native ORB front-end (libORBfe_b200.so).
 *
 * Drop-in boundary for ONE hot path of ORB-SLAM3 (reference = LY-zhang-yi-hao/ORB-SLAM3_byZyh):
 * ORB extraction and Hamming matching.  The reference has no FFI/plugin layer; these entry
 * points are what a header-only adapter (orb-slam3_byzyh_b200/host/ORBextractor.h,
 * ORBmatcher_b200.h) binds so that Tracking / Frame / LocalMapping link unchanged.  Each entry
 * point cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions: plain pointers + sizes, no C++/torch types; `int` status (ORBFE_OK == 0,
 * errors < -1) unless the reference function itself returns a count; every function is
 * synchronous on return unless it takes an explicit stream.  There is NO CPU fallback: when
 * no CUDA device is usable every call fails with ORBFE_ERR_CUDA.
 */
#ifndef ORBFE_H
#define ORBFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBFE_OK 0
#define ORBFE_EMPTY_IMAGE (-1)   /* operator() returns -1 on an empty image (ORBextractor.cc:1561) */
#define ORBFE_ERR_INVALID (-2)   /* bad argument / unsupported geometry */
#define ORBFE_ERR_CUDA (-3)      /* CUDA runtime error or no device */
#define ORBFE_ERR_CAPACITY (-4)  /* caller buffer too small */

#define ORBFE_MAX_LEVELS 16
#define ORBFE_EDGE 19            /* EDGE_THRESHOLD, ORBextractor.cc:78 */

/* Byte-compatible with cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id). */
typedef struct OrbfeKeyPoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} OrbfeKeyPoint;

typedef struct OrbfeExtractor OrbfeExtractor; /* opaque; one instance == one ORBextractor */

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * include/ORBextractor.h:49-50, src/ORBextractor.cc:468-571.  `device` = CUDA ordinal. */
int orbfe_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                           int minThFAST, int device, OrbfeExtractor** out);
void orbfe_extractor_destroy(OrbfeExtractor* h);
const char* orbfe_last_error(void);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares, include/ORBextractor.h:61-81.  Arrays hold nlevels floats. */
int orbfe_get_levels(const OrbfeExtractor* h);
float orbfe_get_scale_factor(const OrbfeExtractor* h);
int orbfe_scale_tables(const OrbfeExtractor* h, float* scale, float* inv_scale, float* sigma2,
                       float* inv_sigma2);
int orbfe_features_per_level(const OrbfeExtractor* h, int* n_per_level);
/* Upper bound of keypoints one frame can return (octree may overshoot the per-level target); holds for frames whose
 * levels start DistributeOctTree from at most 8 roots (ORBextractor.cc:718: round(width / height) of the level), i.e.
 * aspect ratios up to 8.5 : 1.  orbfe_max_keypoints_for is the bound for a rows x cols frame whatever its shape (the
 * rectified size when rectification maps are set); a call whose `capacity` is smaller than the keypoint count of a
 * frame returns ORBFE_ERR_CAPACITY. */
int orbfe_max_keypoints(const OrbfeExtractor* h);
int orbfe_max_keypoints_for(const OrbfeExtractor* h, int rows, int cols);

/* int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)
 * include/ORBextractor.h:57-59, src/ORBextractor.cc:1557-1682.  image = rows x cols CV_8UC1,
 * `step` bytes per row (host memory).  Writes *n_out keypoints (cv::KeyPoint layout) and
 * n_out x 32 descriptor bytes; returns monoIndex (>= 0), ORBFE_EMPTY_IMAGE, or an error. */
int orbfe_extract(OrbfeExtractor* h, const uint8_t* image, int rows, int cols, size_t step,
                  int lap0, int lap1, OrbfeKeyPoint* keypoints, uint8_t* descriptors,
                  int capacity, int* n_out);

/* Batched form of the same call: B frames of identical size, frame b at images +
 * b*frame_stride (host memory, pinned for best throughput).  Outputs are per-frame slabs of
 * `capacity` entries: keypoints[b*capacity + i], descriptors[(b*capacity + i)*32].
 * n_out[b] / mono_out[b] = keypoint count / monoIndex of frame b.  Frames are independent
 * (reference: one operator() call per frame, src/Frame.cc:513-523). */
int orbfe_extract_batch(OrbfeExtractor* h, const uint8_t* images, int B, int rows, int cols,
                        size_t step, size_t frame_stride, int lap0, int lap1,
                        OrbfeKeyPoint* keypoints, uint8_t* descriptors, int capacity,
                        int* n_out, int* mono_out);

/* Streaming form of orbfe_extract_batch for a caller that produces batch after batch (a dataset
 * replay, a multi-camera rig): _submit enqueues the whole batch (H2D, kernels, D2H) and returns at
 * once; _wait blocks until the OLDEST submitted batch has its results in host memory (FIFO) and
 * returns its status (ORBFE_ERR_CAPACITY as for the synchronous call).  With a second batch
 * submitted before the first is waited for, its host->device copies run under the kernels of the
 * first, so the pipeline fill of the synchronous call is paid once instead of once per batch.
 * Host buffers (pinned, for the copies to be asynchronous) must stay valid and untouched until
 * the matching _wait; at most 4 batches in flight (ORBFE_ERR_INVALID beyond that).  The
 * reference has no counterpart: Frame::ExtractORB (src/Frame.cc:513-523) is one blocking call per
 * frame; orbfe_extract_batch(...) == _submit(...) followed by _wait() until nothing is in flight. */
int orbfe_extract_batch_submit(OrbfeExtractor* h, const uint8_t* images, int B, int rows, int cols,
                               size_t step, size_t frame_stride, int lap0, int lap1,
                               OrbfeKeyPoint* keypoints, uint8_t* descriptors, int capacity,
                               int* n_out, int* mono_out);
int orbfe_extract_batch_wait(OrbfeExtractor* h);

/* Same, with inputs and outputs already resident in device memory (device pointers) and the
 * work enqueued on `stream` (a cudaStream_t; NULL = the extractor's own stream).  Does not
 * synchronise: results are valid after the stream is synchronised. */
int orbfe_extract_batch_device(OrbfeExtractor* h, const uint8_t* d_images, int B, int rows,
                               int cols, size_t step, size_t frame_stride, int lap0, int lap1,
                               OrbfeKeyPoint* d_keypoints, uint8_t* d_descriptors, int capacity,
                               int* d_n_out, int* d_mono_out, void* stream);

/* std::vector<cv::Mat> ORBextractor::mvImagePyramid (include/ORBextractor.h:83): level `level`
 * of frame `frame` of the last extract call.  with_border = 0 copies the w x h ROI, 1 copies
 * the (w+38) x (h+38) bordered buffer the ROI lives in.  dst is host memory. */
int orbfe_level_size(const OrbfeExtractor* h, int rows, int cols, int level, int* w, int* hgt);
int orbfe_pyramid_level(OrbfeExtractor* h, int frame, int level, int with_border, uint8_t* dst,
                        size_t dst_step);

/* Stage taps for parity tests (values of the last extract call, frame `frame`). */
int orbfe_debug_candidates(OrbfeExtractor* h, int frame, int level, int32_t* xys, int capacity,
                           int* n_out); /* FAST candidates (x,y,score) window coords, emission order */
int orbfe_debug_level_keypoints(OrbfeExtractor* h, int frame, int level, int32_t* xys,
                                int capacity, int* n_out); /* octree-retained, list order */
int orbfe_debug_blurred(OrbfeExtractor* h, int frame, int level, uint8_t* dst, size_t dst_step);
/* Run only DistributeOctTree (src/ORBextractor.cc:711-1057) on caller-supplied candidates. */
int orbfe_debug_octree(OrbfeExtractor* h, const int32_t* xys, int n, int minX, int maxX, int minY,
                       int maxY, int N, int32_t* keep_idx, int capacity, int* n_out);
/* Per-kernel device milliseconds of the last chunk of the last batch call (CUDA events on the
 * launching stream; needs orbfe_set_profiling(h,1)).  Order: h2d, pyramid, fast (per-cell FAST + NMS +
 * retry + emission, one kernel), octree, layout, blur, describe, d2h; h2d/d2h are reported as 0 (they
 * overlap neighbouring chunks on their own streams). */
#define ORBFE_NUM_STAGES 8
int orbfe_set_profiling(OrbfeExtractor* h, int enable);
int orbfe_stage_ms(OrbfeExtractor* h, float* ms /*[ORBFE_NUM_STAGES]*/);
/* Number of kernel launches issued by this extractor since creation. */
long long orbfe_launch_count(const OrbfeExtractor* h);
/* Stereo rectification fused into the extractor: System::TrackStereo remaps both images before it hands them to
 * Tracking (cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR), src/System.cc:286-293).  With maps set
 * (CV_32FC1 x / y maps of the rectified size rows x cols, copied to the device), every orbfe_extract* call takes the RAW
 * camera frames (rows / cols of the call = raw size) and computes pyramid level 0 as the rectified image, bit-exact
 * with cv::remap + copyMakeBorder; all outputs are those of ORBextractor run on the rectified image, and
 * orbfe_pyramid_level(h, frame, 0, ...) returns the rectified image itself.  NULL maps switch it off. */
int orbfe_extractor_set_rectification(OrbfeExtractor* h, const float* map_x, const float* map_y,
                                      int rows, int cols);
/* Device-memory budget of ONE chunk of intermediates (default 6 GiB, env ORBFE_MAX_BYTES); batches larger than
 * one chunk are processed chunk after chunk.  orbfe_extract_batch (host pointers) keeps two chunks in flight on two
 * compute streams, so it allocates up to twice this budget. */
int orbfe_set_max_bytes(OrbfeExtractor* h, unsigned long long bytes);
/* Geometry of the last image size seen: FAST cells, candidate slots and keypoint slots per
 * frame, bytes of one frame's padded pyramid, device bytes of all intermediates per frame. */
int orbfe_frame_geometry(const OrbfeExtractor* h, int* cells, int* slots, int* kpcap,
                         unsigned long long* pyr_stride, unsigned long long* per_frame_bytes);

/* ------------------------------- Hamming matching ---------------------------------------- */

/* static int ORBmatcher::DescriptorDistance(a, b)  include/ORBmatcher.h:43,
 * src/ORBmatcher.cc:2384-2404, batched: out[i] = hamming256(a[i], b[i]) for n pairs (host). */
int orbfe_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* out, int device);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) + the `d0 < d1*0.7` ratio test of
 * Frame::ComputeStereoFishEyeMatches, src/Frame.cc:47,1553,1562.  idx2/dist2 are nq x 2
 * (best, second; -1 when train has < 2 rows); match[i] = accepted train index or -1.
 * `train_offset` is added to every returned index (map sharding). Host pointers.
 * nt < 2^23 per call and rows 16-byte aligned.  From 2^18 descriptor pairs on the search runs on the tensor cores
 * (tcgen05.mma on one signed byte per descriptor bit, csrc/knn_umma.cu) with the same result bit for bit. */
int orbfe_knn2(const uint8_t* query, int nq, const uint8_t* train, int nt, int train_offset,
               int32_t* idx2, int32_t* dist2, int32_t* match, int device);
/* Device-pointer form on `stream`, no synchronisation. */
int orbfe_knn2_device(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt,
                      int train_offset, int32_t* d_idx2, int32_t* d_dist2, void* stream);
/* Merge G per-shard (idx2, dist2) tables [G][nq][2] (global indices, e.g. the all-gathered
 * results of a map sharded over G GPUs) into the global best two (ties -> lower
 * index) and apply the ratio test.  Host-pointer and device-pointer forms. */
int orbfe_knn2_merge(const int32_t* idx2_shards, const int32_t* dist2_shards, int G, int nq,
                     int32_t* idx2, int32_t* dist2, int32_t* match, int device);
int orbfe_knn2_merge_device(const int32_t* d_idx2_shards, const int32_t* d_dist2_shards, int G,
                            int nq, int32_t* d_idx2, int32_t* d_dist2, int32_t* d_match,
                            void* stream);
/* The merge with the exchange fused in: peer_tabs[s] (host array of G <= 16 DEVICE pointers) addresses shard s's packed
 * table { idx2[nq][2], dist2[nq][2] } where GPU s wrote it -- this GPU's own buffer or a peer mapping over NVLink
 * (torch symmetric memory, cudaIpcOpenMemHandle).  The merge kernel gathers with peer loads, so no all-gather runs;
 * the caller orders "all shards written" before and "all ranks merged" after (one device barrier per step with
 * double-buffered tables, orbfe/dist.py). */
int orbfe_knn2_merge_peers_device(const int32_t* const* peer_tabs, int G, int nq, int32_t* d_idx2,
                                  int32_t* d_dist2, int32_t* d_match, void* stream);
/* Same for tables gathered as ONE buffer: packed[s] = { idx2[nq][2], dist2[nq][2] } of shard s, so
 * that a single all-gather moves both tables of every shard. */
int orbfe_knn2_merge_packed_device(const int32_t* d_packed, int G, int nq, int32_t* d_idx2,
                                   int32_t* d_dist2, int32_t* d_match, void* stream);

/* The part of ORB_SLAM3::Frame the projection matchers read (Nleft == -1 layout):
 * mvKeysUn, mvuRight, mDescriptors, mnMinX..mnMaxY, mfGridElement{Width,Height}Inv,
 * mvScaleFactors (include/Frame.h), plus the 64x48 grid that AssignFeaturesToGrid builds
 * (src/Frame.cc:469-504) -- rebuilt on the device from `keys`. */
typedef struct OrbfeFrameView {
    int32_t n;
    const OrbfeKeyPoint* keys;
    const float* uright;      /* may be NULL (monocular) */
    const uint8_t* desc;      /* n x 32 */
    float min_x, min_y, max_x, max_y;
    float grid_w_inv, grid_h_inv;
} OrbfeFrameView;

/* One entry per candidate map point, structure of arrays, all of length m.  The caller has
 * already projected the points (camera model + pose stay on the host side of the boundary,
 * SURVEY 8c) and evaluated the reference's early `continue`s into `valid`. */
typedef struct OrbfeProjPoints {
    int32_t m;
    const float* u;          /* mTrackProjX / uv(0) */
    const float* v;          /* mTrackProjY / uv(1) */
    const float* ur;         /* mTrackProjXR / uv(0)-mbf*invz (stereo gate) */
    const float* radius;     /* r * mvScaleFactors[level]   (ORBmatcher.cc:76-83, 2010) */
    const int32_t* min_level;
    const int32_t* max_level;
    const float* angle;      /* source keypoint angle (rotation histogram); may be NULL */
    const uint8_t* valid;
    const uint8_t* blocks;   /* MapPoint::Observations() > 0 */
    const uint8_t* desc;     /* m x 32, MapPoint::GetDescriptor() */
} OrbfeProjPoints;

#define ORBFE_SEARCH_MAPPOINTS 0 /* SearchByProjection(Frame&, vector<MapPoint*>&, th, ...)   ORBmatcher.cc:46   */
#define ORBFE_SEARCH_LASTFRAME 1 /* SearchByProjection(Frame&, const Frame&, th, bMono)       ORBmatcher.cc:1951 */
#define ORBFE_SEARCH_KEYFRAME 2  /* SearchByProjection(Frame&, KeyFrame*, set<>&, th, ORBdist) ORBmatcher.cc:2197 */

typedef struct OrbfeSearchParams {
    int32_t mode;
    int32_t th_accept;         /* TH_HIGH (modes 0,1) or ORBdist (mode 2) */
    float nnratio;             /* ORBmatcher::mfNNratio */
    int32_t check_orientation; /* ORBmatcher::mbCheckOrientation */
} OrbfeSearchParams;

/* claimed[i] != 0 <=> F.mvpMapPoints[i] already holds a point that blocks (see modes).
 * assigned[i] (in/out, length F.n): index of the map point now held by keypoint i (-1 = the
 * reference would store NULL; entries never touched keep their input value).
 * best_idx/best_dist (length m, may be NULL): per map point accepted keypoint (-1 = none)
 * and best distance.  Returns nmatches (>= 0) exactly as the reference function does, or an
 * error (< -1).  Host pointers. */
int orbfe_search_by_projection(const OrbfeFrameView* frame, const OrbfeProjPoints* pts,
                               const OrbfeSearchParams* prm, const uint8_t* claimed,
                               int32_t* assigned, int32_t* best_idx, int32_t* best_dist,
                               int device);

/* The same search against a map that is SHARDED over several GPUs by contiguous index ranges (BASELINE config 5:
 * "SearchByProjection ... map sharded over 8 GPUs with NVLink best-match gather").  The reference walks the map points
 * in vector order and a keypoint accepted by point j is skipped by every later point (src/ORBmatcher.cc:54, 103-105,
 * 156), so the shards cannot simply be searched independently.  A shard holds the projected points [j0, j0+m) in HBM
 * (orbfe_map_shard_create: host arrays of THIS shard, uploaded once); the frame is set on every shard
 * (orbfe_map_shard_set_frame).  One search = passes until the claim table stops changing:
 *   claims (device int32[frame n]): -1 = keypoint held a blocking point on entry, INT_MAX = free, else the GLOBAL index
 *   of the first accepting point (orbfe_claims_init_device builds the static table from F.mvpMapPoints' flags);
 *   orbfe_map_shard_pass evaluates the shard's points against claims_in and lowers claims_out (pre-set to the static
 *   table) to its own first acceptors; the shards' claims_out are combined by an elementwise minimum -- 4 bytes per
 *   frame keypoint per pass, the only exchange: an all-reduce(MIN), or orbfe_claims_min_peers_device reading the other
 *   shards' tables with peer loads over NVLink; the result is the next claims_in.
 * orbfe_map_shard_finish then raises assigned[k] to the (global) index of the last accepting point and adds the
 * shard's match count; assigned tables combine by elementwise maximum, counts by sum.  Mode ORBFE_SEARCH_MAPPOINTS (and
 * ORBFE_SEARCH_KEYFRAME without orientation check).  Results equal orbfe_search_by_projection on the whole map. */
typedef struct OrbfeMapShard OrbfeMapShard;
int orbfe_map_shard_create(const OrbfeProjPoints* pts_shard, int j0, int device, OrbfeMapShard** out);
void orbfe_map_shard_destroy(OrbfeMapShard* h);
int orbfe_map_shard_set_frame(OrbfeMapShard* h, const OrbfeFrameView* frame, void* stream);
int orbfe_claims_init_device(const uint8_t* d_claimed, int n, int32_t* d_claims, void* stream);
int orbfe_map_shard_pass(OrbfeMapShard* h, const OrbfeSearchParams* prm, const int32_t* d_claim_in,
                         int32_t* d_claim_out, void* stream);
int orbfe_claims_min_peers_device(const int32_t* const* peer_tabs, int n_shards, int n, int32_t* d_out,
                                  const int32_t* d_prev, int32_t* d_changed, void* stream);
int orbfe_map_shard_finish(OrbfeMapShard* h, int32_t* d_assigned, int32_t* d_nmatches, void* stream);
int orbfe_map_shard_results(OrbfeMapShard* h, int32_t* best_idx, int32_t* best_dist, void* stream);

/* The same search on a fisheye stereo frame (F.Nleft != -1): modes ORBFE_SEARCH_MAPPOINTS
 * (src/ORBmatcher.cc:46-240 incl. the right-camera branch :171-237 and the mvLeftToRightMatch /
 * mvRightToLeftMatch partner writes :159-163, :215-219) and ORBFE_SEARCH_LASTFRAME (:1951-2185
 * incl. :2090-2155).  `left` = mvKeys + descriptor rows [0,Nleft), `right` = mvKeysRight + rows
 * [Nleft,N) (both with the frame's grid bounds); l2r / r2l = mvLeftToRightMatch /
 * mvRightToLeftMatch (-1 = none).  pts_left carries the left projections plus the shared angle /
 * blocks / desc arrays, pts_right the right-camera projections (u, v, radius, levels, valid =
 * mbTrackInViewR && mnTrackScaleLevelR != -1 for mode 0).  claimed / assigned are indexed like
 * F.mvpMapPoints: [0,Nleft) left, [Nleft,N) right. */
int orbfe_search_by_projection_fisheye(const OrbfeFrameView* left, const OrbfeFrameView* right,
                                       const int32_t* l2r, const int32_t* r2l,
                                       const OrbfeProjPoints* pts_left, const OrbfeProjPoints* pts_right,
                                       const OrbfeSearchParams* prm, const uint8_t* claimed,
                                       int32_t* assigned, int32_t* best_idx_left,
                                       int32_t* best_idx_right, int device);

/* int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 * vector<int>& vnMatches12, int windowSize)   include/ORBmatcher.h:67, src/ORBmatcher.cc:735-891
 * (monocular initialisation; first of the SURVEY 8(f) "next" matchers).  f1 / f2 = mvKeysUn +
 * mDescriptors of the two frames (f2 with its grid bounds); prev_matched = vbPrevMatched as n1 x 2
 * floats (in/out); matches12 = vnMatches12 (out, -1 = none).  Returns nmatches. */
int orbfe_search_for_initialization(const OrbfeFrameView* f1, const OrbfeFrameView* f2,
                                    float* prev_matched, int window_size, float nnratio,
                                    int check_orientation, int32_t* matches12, int device);

/* Keyframe-side searches.  ORBmatcher::Fuse (include/ORBmatcher.h:83,86;
 * src/ORBmatcher.cc:1326-1534, 1536-1688), SearchBySim3 (:1690-1940) and the Sim3 SearchByProjection
 * overloads (:496-733) run, per projected map point, KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:843-892)
 * + the `[nPredictedLevel-1, nPredictedLevel]` level filter + a strict-`<` Hamming argmin, with no state
 * carried from point to point.  `kf` = the keyframe's mvKeysUn (mvKeys / mvKeysRight for fisheye), mvuRight,
 * descriptor rows and grid bounds (KeyFrame::mnMinX.. are the frame's bounds truncated to int);
 * `pts` as for orbfe_search_by_projection (angle / blocks unused; min_level = nPredictedLevel-1,
 * max_level = nPredictedLevel).  ORBFE_GATE_FUSE adds Fuse's reprojection gate (:1436-1461): chi2 7.8
 * on (ex,ey,er) when the keypoint has a stereo coordinate, 5.99 on (ex,ey) otherwise, scaled by
 * mvInvLevelSigma2[octave]; it reads pts->ur (= uv(0) - bf*invz).
 * best_idx[j] = keypoint index, or -1 when nothing is within th_accept (TH_LOW for Fuse, TH_HIGH for
 * SearchBySim3); best_dist[j] (may be NULL) = smallest distance (256 = no candidate).  Returns the number
 * of points with a match.  The pointer-graph updates (AddObservation / Replace, vpReplacePoint) stay with
 * the caller: host/ORBmatcher_b200.h shows them.  The Sim3 SearchByProjection overloads additionally skip
 * keypoints already in vpMatched: that is orbfe_search_by_projection with ORBFE_SEARCH_KEYFRAME,
 * claimed = (vpMatched[i] != NULL), th_accept = floor(TH_LOW * ratioHamming), check_orientation = 0. */
#define ORBFE_GATE_NONE 0
#define ORBFE_GATE_FUSE 1
typedef struct OrbfeWindowParams {
    int32_t th_accept;
    int32_t gate;
    const float* inv_level_sigma2; /* [n_levels], ORBFE_GATE_FUSE only */
    int32_t n_levels;
} OrbfeWindowParams;
int orbfe_search_window(const OrbfeFrameView* kf, const OrbfeProjPoints* pts,
                        const OrbfeWindowParams* prm, int32_t* best_idx, int32_t* best_dist,
                        int device);

/* int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12,
 * const Sophus::Sim3f& S12, const float th)   include/ORBmatcher.h:80, src/ORBmatcher.cc:1690-1940.
 * pts12: one entry per KF1 slot = its map point transformed by S21 and projected into KF2 (valid = the
 * slot has a good, not yet matched point that passes :1736-1765); pts21: the reverse (:1826-1862).
 * match12[i1] = KF2 keypoint index when both directions agree (:1922-1937), else -1; the caller stores
 * vpMatches12[i1] = vpMapPoints2[match12[i1]].  Returns nFound. */
int orbfe_search_by_sim3(const OrbfeFrameView* kf1, const OrbfeFrameView* kf2,
                         const OrbfeProjPoints* pts12, const OrbfeProjPoints* pts21, int th_accept,
                         int32_t* match12, int device);

/* ---- Bag of words ------------------------------------------------------------
 * ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h:30-31).
 * orbfe_vocabulary_create uploads the tree: node 0 is the root, node i > 0 has parent[i], descriptor
 * desc[32*i..] and weight[i] (the leaf weight; ORBvoc.txt columns, TemplatedVocabulary.h:1379-1417).  A
 * node's children are ordered by ascending id, as loadFromTextFile (:1390) and HKmeansStep push them;
 * leaves (= nodes without children) get word ids in node order (:1409-1416). */
typedef struct OrbfeVocabulary OrbfeVocabulary;
int orbfe_vocabulary_create(int k, int L, int n_nodes, const int32_t* parent, const uint8_t* desc,
                            const double* weight, int device, OrbfeVocabulary** out);
void orbfe_vocabulary_destroy(OrbfeVocabulary* voc);

/* Per feature, TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1226-1258, distance = FORB::distance, FORB.cpp:81-101):
 * the word the descriptor falls into, its weight, and the node on level L - levelsup of the path
 * (0 when that level is <= 0; -1 when the leaf lies above it, where the reference leaves it
 * uninitialised).  The caller folds the three arrays into DBoW2::BowVector / FeatureVector exactly like
 * transform(features, v, fv, levelsup) (:1125-1197): for i in order, if weight[i] > 0:
 * v.addWeight(word_id[i], weight[i]); fv.addFeature(node_id[i], i); then v.normalize(L1)
 * (host/ORBmatcher_b200.h: ComputeBoW).  Replaces the descriptor loop of Frame::ComputeBoW
 * (src/Frame.cc:984-998) and KeyFrame::ComputeBoW (src/KeyFrame.cc:101-111).  Host pointers. */
int orbfe_bow_transform(OrbfeVocabulary* voc, const uint8_t* desc, int n, int levelsup,
                        int32_t* word_id, double* weight, int32_t* node_id);
/* Same with device pointers on `stream` (cudaStream_t), no synchronisation. */
int orbfe_bow_transform_device(OrbfeVocabulary* voc, const uint8_t* d_desc, int n, int levelsup,
                               int32_t* d_word_id, double* d_weight, int32_t* d_node_id,
                               void* stream);

/* The fold itself on the device, for batches of frames whose descriptors stay in HBM (TF_IDF weighting, L1_NORM scoring:
 * the configuration of ORBvoc.txt).  Frame b owns the features [frame_start[b], frame_start[b+1]) of the per-feature arrays
 * (at most `capacity` <= 4096 of them) and the output slabs [b * capacity ...): bow_word / bow_value = the BowVector in map
 * order (n_bow[b] entries; values are DBoW2's doubles bit for bit: weights added in feature order, L1 norm added in key
 * order), fv_node / fv_feat = the FeatureVector in map order (n_fv[b] nodes, NodeId compared as unsigned; node j holds
 * fv_feat[fv_start[b * (capacity + 1) + j] .. fv_start[.. + j + 1]), feature indices relative to the frame). */
int orbfe_bow_fold_device(const int32_t* d_word_id, const double* d_weight, const int32_t* d_node_id,
                          const int32_t* d_frame_start, int B, int capacity, uint32_t* d_bow_word,
                          double* d_bow_value, int32_t* d_n_bow, uint32_t* d_fv_node, int32_t* d_fv_start,
                          int32_t* d_fv_feat, int32_t* d_n_fv, void* stream);

/* DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) flattened in map order: node ids ascending,
 * features of node j = feat[start[j] .. start[j+1]) in push_back order. */
typedef struct OrbfeFeatureVector {
    int32_t n_nodes;
    const int32_t* node_id;
    const int32_t* start; /* n_nodes + 1 */
    const int32_t* feat;
} OrbfeFeatureVector;

/* One side of a BoW search: mDescriptors, keypoint angles (rotation histogram; may be NULL when
 * check_orientation == 0), valid[i] (may be NULL = all) and mFeatVec. */
typedef struct OrbfeBowSide {
    int32_t n;
    const uint8_t* desc;
    const float* angle;
    const uint8_t* valid;
    OrbfeFeatureVector fv;
} OrbfeBowSide;

/* int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 *     include/ORBmatcher.h:58, src/ORBmatcher.cc:260-494:  a = pKF (valid[i] = vpMapPointsKF[i] is a good point),
 *     b = F (valid NULL), th_low = TH_LOW, strict = 0, n_left_b = F.Nleft;
 * int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 *     include/ORBmatcher.h:59, src/ORBmatcher.cc:893-1044:  a = pKF1, b = pKF2 (valid = good map point, and
 *     index < mvKeysUn.size() for fisheye keyframes), strict = 1 (`bestDist1 < TH_LOW`, :973), n_left_b = -1.
 * match_a[ia] = index in b matched to a's feature ia after the rotation-histogram cull, else -1;
 * match_a_right[ia] = the right-camera match of the fisheye branch (:374-407), required iff n_left_b != -1.
 * The caller stores vpMapPointMatches[match_a[ia]] = vpMapPointsKF[ia] (and the right one), resp.
 * vpMatches12[ia] = vpMapPoints2[match_a[ia]].  Returns nmatches. */
int orbfe_search_by_bow(const OrbfeBowSide* a, const OrbfeBowSide* b, int th_low, int strict,
                        float nnratio, int check_orientation, int n_left_b, int32_t* match_a,
                        int32_t* match_a_right, int device);

/* int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t,size_t>>& vMatchedPairs,
 * const bool bOnlyStereo, const bool bCoarse)   include/ORBmatcher.h:76-77, src/ORBmatcher.cc:1046-1324; pinhole
 * keyframes (mpCamera2 == NULL) with prm->rig == NULL, fisheye stereo rigs with prm->rig set (below).  Each side: mvKeysUn, mDescriptors, mvuRight (NULL = monocular), has_map_point[i] = (GetMapPoint(i) != NULL),
 * mFeatVec.  f12 = the fundamental matrix of Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:191-194:
 * K1^-T * hat(t12) * R12 * K2^-1, row major; it is constant per keyframe pair and stays on the host with Eigen),
 * epipole = pKF2->mpCamera->project(T2w * Cw) (:1063), scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors /
 * mvLevelSigma2, th_low = TH_LOW.  matches12[i1] = i2 or -1; vMatchedPairs = those pairs in ascending i1.
 * Returns nmatches. */
typedef struct OrbfeTriSide {
    int32_t n;
    const OrbfeKeyPoint* keys;
    const uint8_t* desc;
    const float* uright;
    const uint8_t* has_map_point;
    OrbfeFeatureVector fv;
} OrbfeTriSide;
/* The two-camera branch (pKF1->mpCamera2 && pKF2->mpCamera2: KannalaBrandt8 stereo rigs, src/ORBmatcher.cc:1071-1095,
 * 1160-1241): each side's keys / desc rows are [mvKeys (left camera, n_left of them) | mvKeysRight], no keypoint is
 * "stereo" (:1121, :1175), the epipole-distance gate is skipped (:1196) and the epipolar gate is
 * KannalaBrandt8::epipolarConstrain (src/CameraModels/KannalaBrandt8.cpp:322-328: TriangulateMatches(...) > 0.0001f) with
 * the cameras and the relative pose of the (bRight1, bRight2) combination: pair[2 * bRight1 + bRight2] = {camera of kp1,
 * camera of kp2, R12, t12} = {ll, lr, rl, rr} (:1205-1239).  level_sigma2_1 = pKF1->mvLevelSigma2. */
typedef struct OrbfeTriCameraPair {
    float params1[8], params2[8];
    float precision1, precision2;
    float R12[9], t12[3];
} OrbfeTriCameraPair;
typedef struct OrbfeTriRig {
    int32_t n_left1, n_left2;
    const float* level_sigma2_1;
    OrbfeTriCameraPair pair[4];
} OrbfeTriRig;
typedef struct OrbfeTriParams {
    float f12[9];
    float epipole[2];
    const float* scale_factors2;
    const float* level_sigma2_2;
    int32_t n_levels;
    int32_t only_stereo, coarse, check_orientation, th_low;
    const OrbfeTriRig* rig; /* NULL: pinhole keyframes (mpCamera2 == NULL) */
} OrbfeTriParams;
int orbfe_search_for_triangulation(const OrbfeTriSide* kf1, const OrbfeTriSide* kf2,
                                   const OrbfeTriParams* prm, int32_t* matches12, int device);

/* Batches of stereo pairs that never leave the device (the frame-pair shards of BASELINE configs 2 and 3).
 * orbfe_stereo_match_batch_device: Frame::ComputeStereoMatches (src/Frame.cc:1102-1358) for pairs 0 .. B-1 on the
 * pyramids both extractors hold from their last orbfe_extract_batch_device call (one chunk) and on that call's output
 * slabs (`capacity` rows per frame, d_n_l[b] / d_n_r[b] valid rows); writes mvuRight / mvDepth into [B][capacity] float
 * slabs.  Enqueued on `stream` (0 = the left extractor's stream); the caller orders it after both extractions.
 * orbfe_knn2_batch_device: the brute-force kNN-2 + 0.7 ratio test of Frame::ComputeStereoFishEyeMatches
 * (src/Frame.cc:1545-1562: BFMatcher(NORM_HAMMING).knnMatch(left rows [monoLeft, n), right rows [monoRight, n), 2)) for
 * B pairs: per pair the row ranges [begin, end) of the two descriptor slabs; idx2 / dist2 ([B][capacity][2]) and match
 * ([B][capacity], -1 = ratio test failed) are indexed by query row - begin, train indices are relative to t_begin
 * (the row numbers BFMatcher reports). */
int orbfe_stereo_match_batch_device(OrbfeExtractor* left, OrbfeExtractor* right, int B,
                                    const OrbfeKeyPoint* d_keys_l, const uint8_t* d_desc_l, const int* d_n_l,
                                    const OrbfeKeyPoint* d_keys_r, const uint8_t* d_desc_r, const int* d_n_r,
                                    int capacity, float mbf, float mb, float* d_u_right, float* d_depth,
                                    void* stream);
int orbfe_knn2_batch_device(const uint8_t* d_desc_q, const int* d_q_begin, const int* d_q_end,
                            const uint8_t* d_desc_t, const int* d_t_begin, const int* d_t_end, int B,
                            int capacity, int32_t* d_idx2, int32_t* d_dist2, int32_t* d_match, void* stream);

/* void Frame::ComputeStereoMatches()  include/Frame.h:116, src/Frame.cc:1102-1358.  Uses the
 * device-resident pyramids (frame `frame` of each extractor's last call) of the left/right
 * extractors, as the reference reads mpORBextractor{Left,Right}->mvImagePyramid.
 * Writes mvuRight / mvDepth (length nl). Host pointers. */
int orbfe_stereo_match(OrbfeExtractor* left, OrbfeExtractor* right, int frame,
                       const OrbfeKeyPoint* keys_l, const uint8_t* desc_l, int nl,
                       const OrbfeKeyPoint* keys_r, const uint8_t* desc_r, int nr, float mbf,
                       float mb, float* u_right, float* depth);

/* ---- Frame intake: the OpenCV calls ORB-SLAM3 makes on an image before ORBextractor ----
 * cv::cvtColor(im, im, cv::COLOR_{RGB,BGR,RGBA,BGRA}2GRAY)   src/Tracking.cc:1563-1590, 1623-1636, 1702-1716
 * channels = 3 or 4, rgb_order != 0 when the first channel is red (mbRGB).  8-bit, OpenCV 4.x fixed point. */
int orbfe_cvt_gray(const uint8_t* src, int rows, int cols, size_t src_step, int channels,
                   int rgb_order, uint8_t* dst, size_t dst_step, int device);
int orbfe_cvt_gray_device(const uint8_t* d_src, int rows, int cols, size_t src_step, int channels,
                          int rgb_order, uint8_t* d_dst, size_t dst_step, void* stream);
/* cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR)   src/System.cc:286-293: 8-bit single channel,
 * CV_32FC1 maps (x and y, dst_rows x dst_cols, as Settings builds them with initUndistortRectifyMap),
 * BORDER_CONSTANT with value 0 (cv::remap's defaults). */
int orbfe_remap_linear(const uint8_t* src, int src_rows, int src_cols, size_t src_step,
                       const float* map_x, const float* map_y, int dst_rows, int dst_cols,
                       uint8_t* dst, size_t dst_step, int device);
int orbfe_remap_linear_device(const uint8_t* d_src, int src_rows, int src_cols, size_t src_step,
                              const float* d_map_x, const float* d_map_y, size_t map_step_floats,
                              int dst_rows, int dst_cols, uint8_t* d_dst, size_t dst_step,
                              void* stream);
/* cv::resize(im, imToFeed, settings_->newImSize())   src/System.cc:295-297, 371-376, 457-459 (INTER_LINEAR, 8-bit). */
int orbfe_resize_linear(const uint8_t* src, int src_rows, int src_cols, size_t src_step,
                        int dst_rows, int dst_cols, uint8_t* dst, size_t dst_step, int device);

/* void Frame::UndistortKeyPoints()   include/Frame.h:309, src/Frame.cc:1003-1051: cv::undistortPoints(mat, mat, K,
 * mDistCoef, cv::Mat(), mK) on the keypoint coordinates (five iterations in double, OpenCV 4.x), every other
 * KeyPoint field copied; dist_coef = mDistCoef (k1 k2 p1 p2 [k3 ...], up to 14), a zero first coefficient means
 * "no distortion" and copies the keypoints (:1005-1009).  fx..cy = the float entries of mK. */
int orbfe_undistort_keypoints(const OrbfeKeyPoint* keys, int n, float fx, float fy, float cx,
                              float cy, const float* dist_coef, int n_dist, OrbfeKeyPoint* keys_un,
                              int device);

/* void MapPoint::ComputeDistinctiveDescriptors()   include/MapPoint.h:138, src/MapPoint.cc:438-529, batched over map
 * points: the descriptors observed for map point p, in the order the reference pushes them into
 * vDescriptors (:449-471), are rows start[p] .. start[p+1]-1 of `desc`; best_idx[p] = the row (relative to start[p]) with
 * the smallest median distance to all rows of the point (:507-521; median = sorted[(N-1)/2] with the row's own 0,
 * first row wins ties), -1 for a point without descriptors.  The caller stores mDescriptor = that row. */
int orbfe_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int n_points,
                                  int32_t* best_idx, int device);

/* KannalaBrandt8 geometry behind the fisheye stereo matcher, batched over points / matches.
 * params = mvParameters {fx, fy, cx, cy, k1, k2, k3, k4} (include/CameraModels/GeometricCamera.h), precision = the
 * Newton stop of KannalaBrandt8 (include/CameraModels/KannalaBrandt8.h:42-59, 102; 1e-6 by default).  Host arrays.
 *   cv::Point2f KannalaBrandt8::project(const cv::Point3f&) / Eigen::Vector2f project(const Eigen::Vector3f&)
 *       include/CameraModels/KannalaBrandt8.h:63-65, src/CameraModels/KannalaBrandt8.cpp:40-55, 84-101
 *   cv::Point3f KannalaBrandt8::unproject(const cv::Point2f&)   KannalaBrandt8.h:69, .cpp:180-217: rays (x, y, 1) */
int orbfe_kb8_project(const float* params, const float* p3d, int n, float* uv, int device);
int orbfe_kb8_unproject(const float* params, float precision, const float* uv, int n, float* rays,
                        int device);
/* float KannalaBrandt8::TriangulateMatches(pCamera2, kp1, kp2, R12, t12, sigmaLevel, unc, p3D)
 *   KannalaBrandt8.h:91-93, .cpp:439-515 (+ ::Triangulate :553-565), once per match: the loop body of
 *   Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1560-1587, sigma1 / unc2 = mvLevelSigma2 of the two keypoints'
 *   octaves, R12 / t12 = mRlr / mtlr) and KannalaBrandt8::epipolarConstrain (.cpp:322-328: depth > 0.0001f).
 * R12 row-major 3x3, pt1 / pt2 = n x 2 keypoint coordinates in camera 1 / camera 2.  depth[i] = the reference's return
 * value (z of the point in camera 1, or -1 low parallax, -2 / -3 behind a camera, -4 / -5 reprojection error in
 * camera 1 / 2); p3d[i] = the triangulated point in camera 1 where depth[i] > 0, NaN otherwise.  fp32 with a
 * double-precision null-vector solve where the reference calls Eigen::JacobiSVD: tolerance-based parity. */
int orbfe_kb8_triangulate_matches(const float* params1, float precision1, const float* params2,
                                  float precision2, const float* R12, const float* t12,
                                  const float* pt1, const float* pt2, const float* sigma1,
                                  const float* unc2, int n, float* depth, float* p3d, int device);

/* Stage tap of the triangulation: the homogeneous solutions x (4 doubles each, the right singular vector of the smallest
 * singular value) of n row-major 4 x 4 float systems -- the step where the reference calls Eigen::JacobiSVD<Matrix4f>
 * (KannalaBrandt8.cpp:566-568); the oracle's restatement of that step is compared with it bit for bit. */
int orbfe_debug_kb8_null_vectors(const float* A, int n, double* x, int device);

/* Library/build identification: "orbfe-b200 sm_100a <git-describe-or-date>" */
const char* orbfe_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ORBFE_H */
