/*
 * include/b200slam.h — C ABI of libb200slam.so: the B200-native (sm_100a) replacement for the hot path of
 * DreamWaterFound/self_commit_ORB-SLAM2.
 *
 * The reference has no FFI layer: its boundary is three C++ classes.  The entry points below
 * are what a C++ shim with the reference's own signatures binds to (the shim lives in
 * self_commit_orb-slam2_b200/host/, the reference-side patch in INTEGRATION.md):
 *
 *   ORBextractor::ORBextractor(...)            /root/reference/include/ORBextractor.h:92,  src/ORBextractor.cc:492
 *   ORBextractor::operator()(...)              include/ORBextractor.h:110, src/ORBextractor.cc:1544
 *   ORBextractor::Get*()                       include/ORBextractor.h:118-161
 *   ORBmatcher::DescriptorDistance             include/ORBmatcher.h,  src/ORBmatcher.cc:1913
 *   ORBmatcher::SearchByBoW (2 overloads)      src/ORBmatcher.cc:230, :656
 *   ORBmatcher::SearchByProjection(F,F,th,mono) src/ORBmatcher.cc:1569
 *   Optimizer::LocalBundleAdjustment           include/Optimizer.h:112, src/Optimizer.cc:629
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an int status
 * (B2S_OK = 0) and never throws; buffers are caller-owned, device memory is owned by handles; handles are
 * independent (one CUDA stream each, no global mutable state) so left/right extractors may run on two
 * threads like src/Frame.cc:159-167.  There is NO CPU fallback: without a CUDA device every call fails
 * with B2S_ERR_NO_DEVICE.
 */
#ifndef B200SLAM_H
#define B200SLAM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  B2S_OK = 0,
  B2S_ERR_NO_DEVICE = 1,    /* no CUDA device / driver: the product does not fall back to the CPU */
  B2S_ERR_BAD_ARG = 2,
  B2S_ERR_CUDA = 3,         /* a CUDA call failed; b2s_last_error() has the text */
  B2S_ERR_CAPACITY = 4,     /* caller buffer (cap) or an internal candidate buffer too small */
  B2S_ERR_ABORTED = 5       /* LocalBA: stop flag was set before the first round (no write-back) */
};

const char* b2s_last_error(void);   /* thread-local text of the last failure */
int b2s_version(void);
int b2s_device_count(void);
/* Issue-rate micro-benchmarks of `device` (about 20 ms): out8[0] packed u16x2 min/max (ALU pipe) and [1] IMAD (FMA pipe) in
 * G warp-instructions/s, [2] FP64 DFMA TFLOP/s, [3] POPC G warp-instructions/s, [4] ALU + FMA interleaved, [5] IMMA.16832.U8.U8 (the integer
 * tensor pipe as mma.sync reaches it) in G warp-instructions/s; the rest 0.
 * bench.py reports the kernels' figures against these measured bounds. */
int b2s_measure_peaks(int device, double* out8);

/* Layout-identical to cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} b2s_keypoint;

/* ------------------------------------------------------------------ extractor */
typedef struct b2s_extractor b2s_extractor;

/* ctor: src/ORBextractor.cc:492-609. max_width/max_height/max_batch size the device buffers. */
int b2s_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int max_width,
                         int max_height, int max_batch, int device, b2s_extractor** out);
void b2s_extractor_destroy(b2s_extractor* h);
/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares + per-level quotas;
 * any pointer may be NULL; arrays hold nlevels entries. */
int b2s_extractor_tables(const b2s_extractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                         int32_t* nfeatures_per_level);
int b2s_extractor_max_keypoints(const b2s_extractor* h); /* upper bound of N for one image (nfeatures + 2*nlevels..) */

/* operator(): one image, HOST buffers (H2D/D2H inside). kps: cap entries; desc: cap x 32 bytes.
 * pyr_out: optional array of nlevels host pointers that receive mvImagePyramid level images (tightly packed
 * w_l x h_l) — needed by Frame::ComputeStereoMatches (src/Frame.cc:1044); NULL to skip the D2H.
 * Empty image (w<=0||h<=0||img==NULL) -> *n_out = 0, B2S_OK (src/ORBextractor.cc:1553). */
int b2s_extract(b2s_extractor* h, const uint8_t* img, int width, int height, int stride, b2s_keypoint* kps, uint8_t* desc,
                int cap, int* n_out, uint8_t* const* pyr_out);

/* Batched: B same-sized images (e.g. L/R pairs of several frames), HOST buffers; imgs = B pointers;
 * kps: B*cap, desc: B*cap*32, n_out: B. */
int b2s_extract_batch(b2s_extractor* h, const uint8_t* const* imgs, int batch, int width, int height, int stride,
                      b2s_keypoint* kps, uint8_t* desc, int cap, int* n_out);

/* Batched, DEVICE-resident: d_imgs = batch images at d_imgs + b*img_pitch_bytes (row stride `stride`), results stay
 * in HBM as fixed-size records (the layout that is all-gathered over NVLink):
 *   d_kps[b*cap + i], d_desc[(b*cap + i)*32], d_counts[b].   stream: a cudaStream_t (NULL = the handle's own stream);
 * the call is asynchronous on that stream. */
int b2s_extract_batch_device(b2s_extractor* h, const uint8_t* d_imgs, size_t img_pitch_bytes, int batch, int width,
                             int height, int stride, b2s_keypoint* d_kps, uint8_t* d_desc, int32_t* d_counts, int cap,
                             void* stream);
/* status of the last device batch (candidate-buffer overflow etc.); synchronises the handle's stream. */
int b2s_extractor_check(b2s_extractor* h);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
long long b2s_extractor_launch_count(const b2s_extractor* h);

/* per-stage CUDA-event timing on the launching stream (bench.py roofline): enable, run, then read the accumulated
 * device milliseconds of {pyramid resize chain, FAST cells, quadtree, blur, orient+describe} and the call count. */
int b2s_extractor_set_timing(b2s_extractor* h, int enable);
int b2s_extractor_get_timing(b2s_extractor* h, double* stage_ms5, long long* calls);

/* test hooks: device -> host copies of intermediates of image `b` of the last call */
int b2s_extractor_debug_level(b2s_extractor* h, int b, int level, int blurred, uint8_t* out, int* w, int* hgt);
int b2s_extractor_debug_candidates(b2s_extractor* h, int b, int level, int32_t* xy /*2*cap*/, int32_t* resp, int cap,
                                   int* n);

/* ------------------------------------------------------------------ stereo matching */
/* Frame::ComputeStereoMatches (src/Frame.cc:1026-1421) for the stereo pairs of the batch this handle extracted last; the
 * image pyramids (mvImagePyramid, the function's only other input) are still resident, so no image leaves the device.
 * Pair p = (image first_left+p, image first_right+p) of that batch, p in [0, n_pairs).
 * bf = Frame::mbf; mb = the baseline the reference sees when the function runs (this fork assigns mb AFTER the call,
 * src/Frame.cc:125,186,1108, so it sees 0 and the disparity bound maxD = mbf/mb is +inf; pass the real baseline to get
 * the upstream behaviour).  Outputs per left feature: mvuRight / mvDepth (-1 = no stereo match); n_matched[p] = matches
 * that survive the median cull (:1395-1415).
 *
 * _device: d_kps/d_desc/d_counts are the device record arrays of that batch (e.g. the buffers given to
 * b2s_extract_batch_device, `cap` records per image, cap <= 4096); d_uright/d_depth: device float [n_pairs][cap];
 * asynchronous on `stream`.  The host variant uses the records b2s_extract_batch left on the device and writes
 * [n_pairs][cap_out] host arrays (cap_out >= the cap of that call). */
int b2s_stereo_match_device(b2s_extractor* h, int first_left, int first_right, int n_pairs, const b2s_keypoint* d_kps,
                            const uint8_t* d_desc, const int32_t* d_counts, int cap, float bf, float mb, float* d_uright,
                            float* d_depth, int32_t* d_nmatched, void* stream);
int b2s_stereo_match(b2s_extractor* h, int first_left, int first_right, int n_pairs, float bf, float mb, float* uright,
                     float* depth, int cap_out, int32_t* n_matched);

/* ------------------------------------------------------------------ matcher */
typedef struct b2s_matcher b2s_matcher;
int b2s_matcher_create(int max_features, int max_batch, int device, b2s_matcher** out);
void b2s_matcher_destroy(b2s_matcher* h);
long long b2s_matcher_launch_count(const b2s_matcher* h);

/* DescriptorDistance (src/ORBmatcher.cc:1913): n pairs a[i] vs b[i], HOST buffers -> dist[n] */
int b2s_descriptor_distance(b2s_matcher* h, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* SearchByBoW on flattened arrays (HOST buffers).  A = keyframe side, B = frame side.
 * nodeA/nodeB: DBoW2 FeatureVector node id per feature (features with equal ids are candidates);
 * validA/validB: 1 = has a usable MapPoint (validB may be NULL: SearchByBoW(KF,F) does not need it);
 * strict_lt: 0 -> best<=th_low (KF,F variant :308), 1 -> best<th_low (KF,KF variant :741).
 * matchB[j] = index into A, or -1. */
int b2s_search_by_bow(b2s_matcher* h, const uint8_t* descA, const int32_t* nodeA, const uint8_t* validA, const float* angA,
                      int nA, const uint8_t* descB, const int32_t* nodeB, const uint8_t* validB, const float* angB, int nB,
                      int th_low, float nnratio, int strict_lt, int check_ori, int32_t* matchB, int* nmatches);

/* Batched HOST-buffer variant: `batch` independent (A,B) pairs laid out with strides capA/capB features;
 * nA/nB: host int arrays [batch]; matchB: batch*capB; nmatches: batch. One H2D per array, one D2H per result. */
int b2s_search_by_bow_batch(b2s_matcher* h, int batch, const uint8_t* descA, const int32_t* nodeA, const uint8_t* validA,
                            const float* angA, const int32_t* nA, int capA, const uint8_t* descB, const int32_t* nodeB,
                            const uint8_t* validB, const float* angB, const int32_t* nB, int capB, int th_low,
                            float nnratio, int strict_lt, int check_ori, int32_t* matchB, int32_t* nmatches);

/* Test hook (host only, no GPU needed): the distance from which a frame-side candidate cannot influence SearchByBoW for
 * (TH_LOW, ratio) — the K-list stage lists and counts only candidates below it (257 = no cut).  tests/test_bow_cut_logic.py
 * checks by enumeration that no decision of src/ORBmatcher.cc:284-310 can depend on a candidate at or beyond it. */
int b2s_debug_bow_distance_cut(int th_low, float nnratio);

/* Batched device-resident variant: `batch` independent (A,B) pairs with strides capA/capB features. All pointers are
 * DEVICE pointers; nA/nB are device int arrays [batch]. Asynchronous on `stream`. */
int b2s_search_by_bow_device(b2s_matcher* h, int batch, const uint8_t* d_descA, const int32_t* d_nodeA,
                             const uint8_t* d_validA, const float* d_angA, const int32_t* d_nA, int capA,
                             const uint8_t* d_descB, const int32_t* d_nodeB, const uint8_t* d_validB,
                             const float* d_angB, const int32_t* d_nB, int capB, int th_low, float nnratio, int strict_lt,
                             int check_ori, int32_t* d_matchB, int32_t* d_nmatches, void* stream);

/* SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1569) after projection. */
typedef struct {
  float u, v;      /* projection of the last frame's map point into the current frame (:1620-1621) */
  float invz;      /* 1/zc (:1614) */
  float angle;     /* LastFrame.mvKeysUn[i].angle */
  int32_t octave;  /* LastFrame.mvKeys[i].octave */
  int32_t has_obs; /* pMP->Observations()>0 */
  uint8_t desc[32];/* pMP->GetDescriptor() */
} b2s_proj_query;

typedef struct {
  float mnMinX, mnMinY, mnMaxX, mnMaxY; /* Frame::mnMinX.. (src/Frame.cc:193-214) */
  float bf;                             /* Frame::mbf */
  const float* scale_factors;           /* mvScaleFactors, nlevels entries (host pointer) */
  int nlevels;
} b2s_frame_geom;

/* mode: 0 levels [oct-1,oct+1]; 1 forward (>=oct); 2 backward ([0,oct]) (:1637-1642).
 * match_cur[j] = query index or -1; *nmatches as the reference counts them. HOST buffers. */
int b2s_search_by_projection_last(b2s_matcher* h, const b2s_proj_query* q, int nq, const float* kpx, const float* kpy,
                                  const int32_t* octave, const float* angle, const float* uright,
                                  const uint8_t* occupied, const uint8_t* desc, int nf, const b2s_frame_geom* g, float th,
                                  int mode, int th_high, int check_ori, int32_t* match_cur, int* nmatches);

/* Batched, device-resident form for a sequence (the stream mode of SURVEY.md 8d / 8e): pair b matches the queries
 * d_q[b*capQ .. +d_nq[b]) against the current frame b given as the extractor's device records (d_kps, d_desc: stride capF
 * entries per pair, d_nf[b] valid) and the stereo matcher's d_uright.  Same candidate order, thresholds, greedy resolution and
 * rotation culling as b2s_search_by_projection_last (occupied = none: the current frame holds no MapPoint yet,
 * src/Tracking.cc:881-883).  d_match_cur: [batch][capF], d_nmatches: [batch].  Asynchronous on `stream` (NULL = the
 * matcher's own stream); needs batch <= max_batch and capQ, capF <= max_features of b2s_matcher_create. */
int b2s_search_by_projection_last_device(b2s_matcher* h, int batch, const b2s_proj_query* d_q, const int32_t* d_nq, int capQ,
                                         const b2s_keypoint* d_kps, const float* d_uright, const uint8_t* d_desc,
                                         const int32_t* d_nf, int capF, const b2s_frame_geom* g, float th, int mode,
                                         int th_high, int check_ori, int32_t* d_match_cur, int32_t* d_nmatches, void* stream);
/* The same with HOST buffers (one upload, one download, synchronous): q [batch][capQ], kps / uright / desc [batch][capF]. */
int b2s_search_by_projection_last_batch(b2s_matcher* h, int batch, const b2s_proj_query* q, const int32_t* nq, int capQ,
                                        const b2s_keypoint* kps, const float* uright, const uint8_t* desc, const int32_t* nf,
                                        int capF, const b2s_frame_geom* g, float th, int mode, int th_high, int check_ori,
                                        int32_t* match_cur, int32_t* nmatches);
/* HOST-buffer form for a frame sequence: kps / desc / depth hold batch + 1 consecutive frames ([batch + 1][cap]), uright the
 * mvuRight of frames 1 .. batch; frame b + 1 is matched against frame b, the queries being formed on the device exactly as
 * b2s_track_queries_device forms them (Tcl: [batch][12]).  match_cur [batch][cap], nmatches [batch].  Synchronous. */
int b2s_search_by_projection_sequence(b2s_matcher* h, int batch, const b2s_keypoint* kps, const uint8_t* desc,
                                      const float* depth, const float* uright, const int32_t* n, int cap, const float* Tcl,
                                      float fx, float fy, float cx, float cy, int has_obs, const b2s_frame_geom* g, float th,
                                      int mode, int th_high, int check_ori, int32_t* match_cur, int32_t* nmatches);
/* The projection half of the same function (src/ORBmatcher.cc:1600-1626) for a stereo sequence, on the device: feature i of
 * last frame b with stereo depth d_depth_last > 0 stands for the map point Frame::UnprojectStereo (src/Frame.cc:679-696)
 * gives it, is moved into the current camera by d_Tcl[b] = [R | t] (3 x 4 row-major floats) and projected with fx, fy, cx, cy;
 * query i belongs to feature i (features without depth become skipped queries), d_nq[b] = d_n_last[b].  has_obs = what
 * pMP->Observations() > 0 is for these points (0 for the temporal points of Tracking::UpdateLastFrame). */
int b2s_track_queries_device(b2s_matcher* h, int batch, const b2s_keypoint* d_kps_last, const uint8_t* d_desc_last,
                             const float* d_depth_last, const int32_t* d_n_last, int cap, const float* d_Tcl, float fx,
                             float fy, float cx, float cy, int has_obs, b2s_proj_query* d_q, int32_t* d_nq, void* stream);

/* SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, float th) (src/ORBmatcher.cc:70-175), the local-map
 * matcher of Tracking::SearchLocalPoints; queries are the map points after Frame::isInFrustum (src/Frame.cc:608-735). */
typedef struct {
  float u, v, ur;    /* mTrackProjX, mTrackProjY, mTrackProjXR */
  float view_cos;    /* mTrackViewCos (RadiusByViewingCos, :178-185) */
  int32_t level;     /* mnTrackScaleLevel */
  uint8_t in_view;   /* mbTrackInView && !isBad() (:83-87) */
  uint8_t has_obs;   /* pMP->Observations()>0 */
  uint8_t pad[2];
  uint8_t desc[32];  /* pMP->GetDescriptor() */
} b2s_map_query;

/* Levels [level-1, level], radius RadiusByViewingCos(view_cos) (x th when th != 1) x scale[level], stereo gate on `ur`,
 * best <= th_high (TH_HIGH), ratio test `best > nnratio*second` only when both candidates share the pyramid level.
 * occupied[j] = the feature already holds a MapPoint with Observations()>0 (:123-125; may be NULL).
 * match_cur[j] = query index or -1 (last writer wins); *nmatches as the reference counts them. HOST buffers. */
int b2s_search_by_projection_map(b2s_matcher* h, const b2s_map_query* q, int nq, const float* kpx, const float* kpy,
                                 const int32_t* octave, const float* uright, const uint8_t* occupied, const uint8_t* desc,
                                 int nf, const b2s_frame_geom* g, float th, int th_high, float nnratio, int32_t* match_cur,
                                 int* nmatches);

/* Windowed best-match search on a KeyFrame's feature grid: the search core shared by
 *   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th)                        src/ORBmatcher.cc:1020-1174 (B2S_WIN_CHI2)
 *   ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint)          :1179-1310                  (no flag)
 *   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) :388-512                    (B2S_WIN_GREEDY)
 * The caller keeps the projection (Rcw/tcw or Sim3), the visibility / distance / normal tests and the map bookkeeping
 * (Replace / AddObservation) and passes one query per surviving map point, in the reference's iteration order. */
typedef struct {
  float u, v, ur;      /* projection into the keyframe (ur = u - bf*invz; only the chi-square gate reads it) */
  float radius;        /* th * pKF->mvScaleFactors[nPredictedLevel] */
  int32_t min_level;   /* nPredictedLevel-1 */
  int32_t max_level;   /* nPredictedLevel   */
  uint8_t valid;       /* 0: skipped by the reference before the search (bad / behind the camera / out of range ...) */
  uint8_t pad[3];
  uint8_t desc[32];    /* pMP->GetDescriptor() */
} b2s_win_query;
#define B2S_WIN_CHI2 1   /* reprojection gate e2*mvInvLevelSigma2[level] <= 7.8 (mvuRight >= 0) / 5.99 (:1097-1124) */
#define B2S_WIN_GREEDY 2 /* features with occupied[j] != 0 or chosen by an earlier query are skipped (:462-463,498-502) */

/* best_idx[i] = keyframe feature with the smallest descriptor distance (first minimum in GetFeaturesInArea order,
 * src/KeyFrame.cc:752-796) if that distance <= th_dist (TH_LOW), else -1; best_dist may be NULL; occupied is only
 * read with B2S_WIN_GREEDY (may be NULL); inv_level_sigma2 only with B2S_WIN_CHI2.  HOST buffers. */
int b2s_search_windows(b2s_matcher* h, const b2s_win_query* q, int nq, const float* kpx, const float* kpy,
                       const int32_t* octave, const float* uright, const float* inv_level_sigma2, const uint8_t* occupied,
                       const uint8_t* desc, int nf, const b2s_frame_geom* g, int flags, int th_dist, int32_t* best_idx,
                       int32_t* best_dist, int* n_accepted);

/* ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vMatchedPairs, bOnlyStereo)
 * (src/ORBmatcher.cc:810-1009, CheckDistEpipolarLine :186-215) on flattened keyframes (LocalMapping::CreateNewMapPoints). */
typedef struct {
  const uint8_t* desc;    /* mDescriptors, n x 32 */
  const int32_t* node;    /* DBoW2 FeatureVector node id per feature (mFeatVec) */
  const uint8_t* has_mp;  /* GetMapPoint(idx) != NULL */
  const uint8_t* stereo;  /* mvuRight[idx] >= 0 */
  const float* x;         /* mvKeysUn[idx].pt.x */
  const float* y;
  const int32_t* octave;  /* mvKeysUn[idx].octave */
  const float* angle;     /* mvKeysUn[idx].angle */
  int32_t n;
} b2s_kf_features;

/* ---- persistent Frame feature grid -------------------------------------------------------------------
 * Frame::AssignFeaturesToGrid (src/Frame.cc:461-491) once per frame; the frame's keypoints, descriptors and the 64 x 48
 * cell index then stay on the device, so the consecutive matchers Tracking runs on one frame (SearchByProjection against
 * the last frame, SearchLocalPoints, relocalisation, Fuse / loop searches on a keyframe) upload only their queries.
 * The grid belongs to the matcher handle it was created with.  inv_level_sigma2 (mvInvLevelSigma2) may be NULL unless
 * B2S_WIN_CHI2 searches are made.  _create_device takes the extractor's device-resident records (b2s_extract_batch_device)
 * and an optional device array of mvuRight (b2s_stereo_match_device); NULL = monocular (-1). */
typedef struct b2s_frame_grid b2s_frame_grid;
int b2s_frame_grid_create(b2s_matcher* h, const float* kpx, const float* kpy, const int32_t* octave, const float* angle,
                          const float* uright, const uint8_t* desc, int nf, const b2s_frame_geom* g,
                          const float* inv_level_sigma2, b2s_frame_grid** out);
int b2s_frame_grid_create_device(b2s_matcher* h, const b2s_keypoint* d_kps, const uint8_t* d_desc, int nf,
                                 const float* d_uright, const b2s_frame_geom* g, const float* inv_level_sigma2, void* stream,
                                 b2s_frame_grid** out);
void b2s_frame_grid_destroy(b2s_frame_grid* grid);
int b2s_frame_grid_size(const b2s_frame_grid* grid);
/* Frame::GetFeaturesInArea (src/Frame.cc:741-852): indices in the reference's order; *n = count (B2S_ERR_CAPACITY if > cap) */
int b2s_frame_grid_features_in_area(b2s_frame_grid* grid, float x, float y, float r, int min_level, int max_level,
                                    int32_t* out, int cap, int* n);
/* b2s_search_by_projection_last / _map / b2s_search_windows on a resident grid (same semantics, same results) */
int b2s_search_by_projection_last_grid(b2s_matcher* h, b2s_frame_grid* grid, const b2s_proj_query* q, int nq,
                                       const uint8_t* occupied, float th, int mode, int th_high, int check_ori,
                                       int32_t* match_cur, int* nmatches);
int b2s_search_by_projection_map_grid(b2s_matcher* h, b2s_frame_grid* grid, const b2s_map_query* q, int nq,
                                      const uint8_t* occupied, float th, int th_high, float nnratio, int32_t* match_cur,
                                      int* nmatches);
int b2s_search_windows_grid(b2s_matcher* h, b2s_frame_grid* grid, const b2s_win_query* q, int nq, const uint8_t* occupied,
                            int flags, int th_dist, int32_t* best_idx, int32_t* best_dist, int* n_accepted);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — src/ORBmatcher.cc:515-643
 * (Tracking::MonocularInitialization, src/Tracking.cc:937-944).  prevx / prevy = vbPrevMatched; octave1 / angle1 / desc1 =
 * F1.mvKeysUn / F1.mDescriptors (only level-0 keypoints search, :537); the F2 side as for the other projection matchers;
 * window = windowSize; th_low = TH_LOW; nnratio / check_ori = the matcher's constructor arguments.
 * match12[i1] = F2 feature or -1 (vnMatches12); the caller then sets vbPrevMatched[i1] = F2.mvKeysUn[match12[i1]].pt
 * (:636-638). */
int b2s_search_for_initialization(b2s_matcher* h, const float* prevx, const float* prevy, const int32_t* octave1,
                                  const float* angle1, const uint8_t* desc1, int n1, const float* kpx2, const float* kpy2,
                                  const int32_t* octave2, const float* angle2, const uint8_t* desc2, int n2,
                                  const b2s_frame_geom* g, int window, int th_low, float nnratio, int check_ori,
                                  int32_t* match12, int* nmatches);

/* F12: 3x3 row-major float; (ex, ey): epipole of camera 1 in image 2 (:815-823); scale_factors / level_sigma2:
 * pKF2->mvScaleFactors / mvLevelSigma2.  match12[idx1] = idx2 or -1 (vMatchedPairs in ascending idx1). HOST buffers. */
int b2s_search_for_triangulation(b2s_matcher* h, const b2s_kf_features* kf1, const b2s_kf_features* kf2, const float* F12,
                                 float ex, float ey, const float* scale_factors, const float* level_sigma2, int nlevels,
                                 int only_stereo, int check_ori, int32_t* match12, int* nmatches);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-440) for a batch of map points:
 * point p owns descriptors [offsets[p], offsets[p+1]) (its observations in map order, bad keyframes removed);
 * best_idx[p] = descriptor (relative index) with the least median distance to the others, -1 without descriptors. */
int b2s_distinctive_descriptors(b2s_matcher* h, const uint8_t* desc, const int32_t* offsets, int n_points,
                                int32_t* best_idx);

/* ------------------------------------------------------------------ DBoW2 vocabulary */
/* TemplatedVocabulary<FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1256) as used by
 * Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:880-896, src/KeyFrame.cc, levelsup = 4): it produces the word
 * ids / weights of mBowVec and the node ids of mFeatVec that SearchByBoW and SearchForTriangulation consume.
 * The vocabulary is passed flattened: node 0 = root, nodes 1..n-1 in the order of the text file (loadFromTextFile
 * :1338-1425: "parent isLeaf d0..d31 weight" per line); children keep the file order, word ids go to the flagged leaves in
 * file order. */
typedef struct {
  int32_t k, L, n_nodes;
  const int32_t* parent;    /* [n_nodes], parent[0] = -1; parents precede their children */
  const uint8_t* leaf_flag; /* [n_nodes] */
  const uint8_t* desc;      /* [n_nodes][32] */
  const double* weight;     /* [n_nodes] */
} b2s_vocabulary_desc;
typedef struct b2s_vocabulary b2s_vocabulary;
int b2s_vocabulary_create(const b2s_vocabulary_desc* d, int device, b2s_vocabulary** out);
void b2s_vocabulary_destroy(b2s_vocabulary* v);
int b2s_vocabulary_words(const b2s_vocabulary* v);
/* Per feature: word id, word weight (0 = stopped word, skipped by DBoW2; may be NULL), node id `levelsup` levels above the
 * leaves (0 = root when L - levelsup <= 0).  The host composes BowVector (sum of weights per word, L1-normalised) and
 * FeatureVector (features per node) from these arrays; the device variant feeds b2s_search_by_bow_device directly. */
int b2s_bow_transform(b2s_vocabulary* v, const uint8_t* features, int n, int levelsup, int32_t* word_id, double* weight,
                      int32_t* node_id);
int b2s_bow_transform_device(b2s_vocabulary* v, const uint8_t* d_features, int n, int levelsup, int32_t* d_word_id,
                             double* d_weight, int32_t* d_node_id, void* stream);

/* ------------------------------------------------------------------ LocalBA */
typedef struct {
  int32_t kf;       /* index into Tcw[] */
  int32_t mp;       /* index into points[] */
  float obs[3];     /* kpUn.pt.x, kpUn.pt.y, mvuRight (<0: monocular edge, src/Optimizer.cc:794) */
  float inv_sigma2; /* mvInvLevelSigma2[octave] */
} b2s_ba_edge;

typedef struct {
  int n_kf;             /* local keyframes first [0,n_local), then fixed cameras */
  int n_local;
  const float* Tcw;     /* n_kf x 16, KeyFrame::GetPose() row-major 4x4 float */
  const uint8_t* fixed; /* n_kf: setFixed(true) (mnId==0 or lFixedCameras) */
  int n_mp;
  const float* points;  /* n_mp x 3, MapPoint::GetWorldPos() */
  int n_edges;
  const b2s_ba_edge* edges; /* insertion order of src/Optimizer.cc:770-853 */
  float fx, fy, cx, cy, bf;
  int its1, its2;       /* optimize(5), optimize(10) */
} b2s_ba_problem;

typedef struct {
  float* Tcw_out;        /* n_local x 16 (SetPose values) */
  float* points_out;     /* n_mp x 3 (SetWorldPos values) */
  uint8_t* edge_outlier; /* n_edges: 1 -> (KF,MP) goes to vToErase (:927-958) */
  int32_t* trace;        /* optional (>=256): accept(1)/reject(0) per LM trial, -1 terminated */
  double chi2_final;
  int n_trials;
} b2s_ba_result;

typedef struct b2s_ba_solver b2s_ba_solver;
int b2s_ba_create(int max_kf, int max_mp, int max_edges, int max_batch, int device, b2s_ba_solver** out);
void b2s_ba_destroy(b2s_ba_solver* h);
long long b2s_ba_launch_count(const b2s_ba_solver* h);
/* How many SMs (= thread blocks of the persistent LM kernel) one batched call may occupy; 0 = all of them (default).  A batch
 * of several windows deals this budget to the windows by estimated cost (the kernel ends with its slowest window).  All SMs
 * give the shortest LocalBA latency; a pipeline that runs other kernels next to LocalBA gets more total throughput from about
 * half of them, because a window's serial part (the reduced-system factorisation) idles fewer blocks.  One-window calls are
 * not affected (16 blocks). */
int b2s_ba_set_sm_budget(b2s_ba_solver* h, int sms);
/* Duration (CUDA events on the solver's stream) of the persistent LM kernel of the last b2s_local_ba(_batch) call and the
 * number of LM trials (accepted + rejected, all windows) it ran — measurement hook for bench.py's roofline. */
float b2s_ba_last_kernel_ms(const b2s_ba_solver* h, long long* lm_trials);
/* Optimizer::LocalBundleAdjustment from graph construction to write-back values. stop: pbStopFlag (may be NULL). */
int b2s_local_ba(b2s_ba_solver* h, const b2s_ba_problem* p, const volatile uint8_t* stop, b2s_ba_result* r);
/* `batch` independent windows solved concurrently (replicas). */
int b2s_local_ba_batch(b2s_ba_solver* h, int batch, const b2s_ba_problem* p, b2s_ba_result* r);

/* ------------------------------------------------------------------ PoseOptimization */
/* Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:363-605) from graph construction to SetPose / mvbOutlier. */
typedef struct {
  const float* Tcw;        /* pFrame->mTcw, 16 floats row-major */
  int32_t n;               /* pFrame->N */
  const uint8_t* has_mp;   /* pFrame->mvpMapPoints[i] != NULL */
  const float* Xw;         /* pMP->GetWorldPos(), n x 3 (ignored where has_mp == 0) */
  const float* kpx;        /* mvKeysUn[i].pt.x */
  const float* kpy;
  const float* uright;     /* mvuRight[i]; < 0: monocular edge (:418) */
  const float* inv_sigma2; /* mvInvLevelSigma2[mvKeysUn[i].octave] */
  float fx, fy, cx, cy, bf;
} b2s_pose_problem;
typedef struct {
  float* Tcw_out;     /* 16 floats: the pose handed to pFrame->SetPose (:600-603) */
  uint8_t* outlier;   /* n: pFrame->mvbOutlier */
  int32_t* trace;     /* optional (>= 256): accept(1)/reject(0) per LM trial, -1 terminated */
  int32_t n_inliers;  /* return value: nInitialCorrespondences - nBad (:605) */
  int32_t n_trials;
} b2s_pose_result;
/* The handle is the LocalBA solver handle (b2s_ba_create); frames of a batch are independent (one CTA each). */
int b2s_pose_optimization(b2s_ba_solver* h, const b2s_pose_problem* p, b2s_pose_result* r);
int b2s_pose_optimization_batch(b2s_ba_solver* h, int batch, const b2s_pose_problem* p, b2s_pose_result* r);

#ifdef __cplusplus
}
#endif
#endif
