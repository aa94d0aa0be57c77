/*
 * m3d.h — C ABI of libm3d.so: the B200 (sm_100a) multi-view 3D reconstruction hot path.
 *
 * The reference (sidd-bme/macaque-3d-pose-estimation) has no FFI / plugin registry: its
 * boundary for this path is the Python object API of
 *   src/third_party/aniposelib/cameras.py   (Camera*, CameraGroup)
 *   src/pipeline/step2_crossviewmatching.py (geometry_affinity2, matchSVT, calc_3dpose)
 *   src/utils/multicam_toolbox.py           (undistortPoints, triangulatePoints)
 * Each entry point below names the reference function it replaces (file:line).  The
 * Python host layer (macaque_3d_pose_estimation_b200/) binds these with ctypes and keeps
 * the reference's signatures; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - every array is C-contiguous float64 unless stated, missing observation = NaN
 *     (validity is decided on the x coordinate only, cameras.py:630,659);
 *   - "dev" pointers are device pointers on the rig's GPU, owned by the caller;
 *     `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous on that stream;
 *   - "*_host" entry points take HOST pointers, run a chunked, double-buffered
 *     H2D -> kernel -> D2H pipeline and return when the outputs are in host memory;
 *   - return value: 0 = ok, negative = error (m3d_last_error() gives the message of the
 *     last failure on the calling thread);
 *   - a rig handle is immutable after creation and may be used from several threads and
 *     streams at once.  No global state.
 */
#ifndef M3D_H_
#define M3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3D_MAX_CAMS 16
#define M3D_MAX_JOINTS 32   /* keypoints per detection in m3d_ray_affinity */
#define M3D_MAX_DETS 128    /* detections per frame in m3d_ray_affinity */
#define M3D_PEER_HANDLE_BYTES 64 /* size of the window handle of m3d_peer_alloc (a CUDA IPC memory handle) */
#define M3D_MAX_DETS_SVT 112 /* detections per frame in m3d_match_svt (shared-memory limit of its Jacobi SVD) */

#define M3D_OK 0
#define M3D_ERR_INVALID -1  /* bad argument (shape, NULL, unsupported parameter)        */
#define M3D_ERR_CUDA -2     /* CUDA runtime error                                       */
#define M3D_ERR_NO_GPU -3   /* no usable sm_100 device: there is NO CPU fallback        */

enum { M3D_MODEL_PINHOLE = 0, M3D_MODEL_FISHEYE = 1, M3D_MODEL_OMNIDIR = 2 };

/* One camera, the fields of the reference's Camera / FisheyeCamera / OmnidirCamera
 * (cameras.py:173-189, 339-354, 429-472).
 *   pinhole : K = `matrix`, dist = `dist` (4, 5, 8 or 12 OpenCV coefficients
 *             k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]); tilt (tauX, tauY) unsupported
 *   fisheye : K = `matrix`, dist = `dist` (k1..k4, Kannala-Brandt)
 *   omnidir : K = `K` (skew K[1] is used), dist = `D` (k1 k2 p1 p2), xi = `xi[0]`      */
typedef struct m3d_cam {
  int32_t model;
  int32_t n_dist;
  double K[9];     /* row-major 3x3 */
  double dist[14];
  double rvec[3];  /* Rodrigues rotation vector (cameras.py:240) */
  double tvec[3];
  double xi;
} m3d_cam;

typedef struct m3d_rig m3d_rig; /* opaque camera group (CameraGroup, cameras.py:558-561) */

int m3d_version(void);
const char* m3d_last_error(void);
int m3d_device_count(void); /* number of visible CUDA devices (<=0: product unusable) */

/* CameraGroup(cameras) on GPU `device`.  Rotation matrices are built with the
 * cv2.Rodrigues formula (utils.py:9-15 make_M). */
int m3d_rig_create(const m3d_cam* cams, int32_t n_cams, int32_t device, m3d_rig** out);
void m3d_rig_destroy(m3d_rig* rig);
int32_t m3d_rig_num_cams(const m3d_rig* rig);
int32_t m3d_rig_device(const m3d_rig* rig);
/* Camera.get_extrinsics_mat / make_M for every camera -> host array (C,4,4)
 * (cameras.py:252, utils.py:9-15). */
int m3d_rig_extrinsics(const m3d_rig* rig, double* M_host);

/* Subset-search strategy of m3d_triangulate_ransac on this rig.  The pruned search
 * (pair certificates) gives the same selection as the exhaustive one by
 * construction; the switch exists so that tests can check exactly that.
 *   M3D_RANSAC_AUTO        pruned search when every camera of the rig is certified, else exhaustive
 *   M3D_RANSAC_EXHAUSTIVE  always solve every subset in front of the selected one */
enum { M3D_RANSAC_AUTO = 0, M3D_RANSAC_EXHAUSTIVE = 1 };
int m3d_rig_set_ransac_mode(m3d_rig* rig, int32_t mode);
/* Bit c set: camera c's distortion model is certified for the pruned search (plain pinhole
 * k1 k2 p1 p2 k3 whose distortion map is strongly monotone on the whole plane). */
int32_t m3d_rig_certified_mask(const m3d_rig* rig);

/* ---- per-camera maps ------------------------------------------------------------- */
/* Camera.undistort_points / FisheyeCamera.~ / OmnidirCamera.~ (cameras.py:310,376,498):
 * xy_dev (n,2) pixels of camera `cam` -> out_dev (n,2) normalised coordinates. */
int m3d_undistort_cam(const m3d_rig* rig, int32_t cam, const double* xy_dev, int64_t n,
                      double* out_dev, void* stream);
/* Camera.project (cameras.py:318,384,509): p3d_dev (n,3) -> out_dev (n,2) pixels. */
int m3d_project_cam(const m3d_rig* rig, int32_t cam, const double* p3d_dev, int64_t n,
                    double* out_dev, void* stream);
/* Camera.distort_points (cameras.py:301,366,487): normalised (n,2) -> pixels (n,2). */
int m3d_distort_cam(const m3d_rig* rig, int32_t cam, const double* xy_dev, int64_t n,
                    double* out_dev, void* stream);

/* ---- camera-group maps ------------------------------------------------------------ */
/* Batched form of the per-camera loop at cameras.py:608-614: xy_dev (C,N,2) -> (C,N,2). */
int m3d_undistort(const m3d_rig* rig, const double* xy_dev, int64_t N, double* out_dev,
                  void* stream);
/* CameraGroup.project (cameras.py:580-591): p3d_dev (N,3) -> out_dev (C,N,2). */
int m3d_project(const m3d_rig* rig, const double* p3d_dev, int64_t N, double* out_dev,
                void* stream);
/* Values of the `undistort` argument of m3d_triangulate / m3d_triangulate_error(_host):
 *   0                    points are already undistorted (cameras.py:615 `undistort=False`)
 *   M3D_UNDISTORT        the reference's per-camera undistortion, float64 throughout (default)
 *   M3D_UNDISTORT_FAST   opt-in: the first three of OpenCV's five fixed-point iterations in float32.
 *                        Results stay inside the tolerances BASELINE.json states (1e-4 relative or
 *                        0.01 mm, 1e-3 px; measured ~1e-6 mm / 3e-5 px) but are no longer the
 *                        float64 values bit for bit; plain pinhole rigs only (others run the strict
 *                        path); the subset RANSAC always runs strict. */
#define M3D_UNDISTORT 1
#define M3D_UNDISTORT_FAST 3
/* CameraGroup.triangulate (cameras.py:593-637): xy_dev (C,N,2) -> p3d_dev (N,3).
 * undistort != 0 applies the per-camera undistortion first.  A camera is used for a
 * point iff its (undistorted) x is not NaN; fewer than two -> (NaN,NaN,NaN). */
int m3d_triangulate(const m3d_rig* rig, const double* xy_dev, int64_t N, int32_t undistort,
                    double* p3d_dev, void* stream);
/* CameraGroup.reprojection_error (cameras.py:746-783): p3d_dev (N,3), xy_dev (C,N,2) raw
 * pixels.  mean == 0: out_dev (C,N,2) residuals xy - project(p3d).  mean != 0: out_dev (N)
 * mean residual norm over cameras with a non-NaN residual, NaN when fewer than two. */
int m3d_reproj_error(const m3d_rig* rig, const double* p3d_dev, const double* xy_dev,
                     int64_t N, int32_t mean, double* out_dev, void* stream);
/* Fused triangulate + reprojection_error(mean=True): the plain branch of the 3D stage
 * (step4_aniposefiltering.py:306-309) in one pass over the input.
 * err_dev may be NULL. */
int m3d_triangulate_error(const m3d_rig* rig, const double* xy_dev, int64_t N,
                          int32_t undistort, double* p3d_dev, double* err_dev, void* stream);
/* CameraGroup.triangulate_ransac == triangulate_possible with one candidate per camera
 * (cameras.py:639-743): exhaustive search over camera subsets in itertools.product order
 * (camera V[j] of the k valid ones is dropped iff bit k-1-j of the step index s is set),
 * subsets smaller than min_cams skipped unless they are the full valid set, accept when
 * err < best (best starts at init_best = 200), stop when best < threshold (= 0.5).
 *   p3d_dev (N,3)          NaN when nothing was selected
 *   picked_dev (C,N) u8    1 for the cameras of the selected subset      [may be NULL]
 *   xy_picked_dev (C,N,2)  raw pixels of the selected cameras, else NaN  [may be NULL]
 *   err_dev (N)            mean reprojection error of the selection, 0.0 when none
 *   subset_dev (N) i32     step index s of the selection, -1 when none   [may be NULL]
 *   neval_dev (N) i32      subsets the reference would have triangulated [may be NULL] */
int m3d_triangulate_ransac(const m3d_rig* rig, const double* xy_dev, int64_t N,
                           int32_t undistort, int32_t min_cams, double threshold,
                           double init_best, double* p3d_dev, uint8_t* picked_dev,
                           double* xy_picked_dev, double* err_dev, int32_t* subset_dev,
                           int32_t* neval_dev, void* stream);

/* CameraGroup.triangulate_possible with P candidates per camera (cameras.py:639-724): per
 * point, itertools.product over the cameras that have a valid candidate (ascending), each
 * offering its valid candidates (ascending) and then "none"; same skip / accept / stop rules
 * as above.  xy_dev (C,N,P,2) raw pixels (NaN x = no candidate); picked_dev (C,N,P) u8;
 * xy_picked_dev (C,N,2) the chosen candidates; index_dev (N) i32 position of the accepted
 * combination in product order, -1 when none; neval_dev (N) i32.  Limits: C * P <= 32, P <= 15.
 * (The reference only ever calls this with P == 1, which is m3d_triangulate_ransac.) */
int m3d_triangulate_possible(const m3d_rig* rig, const double* xy_dev, int64_t N, int32_t P,
                             int32_t undistort, int32_t min_cams, double threshold,
                             double init_best, double* p3d_dev, uint8_t* picked_dev,
                             double* xy_picked_dev, double* err_dev, int32_t* index_dev,
                             int32_t* neval_dev, void* stream);

/* ---- host-buffer pipelines (H2D and D2H inside the call) --------------------------- */
/* Same results as the _dev calls above; xy_host (C,N,2) etc. live in host memory
 * (page-locked memory gives full PCIe speed; pageable memory works but is slower). */
int m3d_triangulate_error_host(const m3d_rig* rig, const double* xy_host, int64_t N,
                               int32_t undistort, double* p3d_host, double* err_host);
int m3d_triangulate_ransac_host(const m3d_rig* rig, const double* xy_host, int64_t N,
                                int32_t undistort, int32_t min_cams, double threshold,
                                double init_best, double* p3d_host, uint8_t* picked_host,
                                double* xy_picked_host, double* err_host,
                                int32_t* subset_host, int32_t* neval_host);
/* The same pipelines for float32 observations (the 2D pose network's native output): xy_host
 * (C,N,2) float32 is widened to float64 on the device, every result is what the float64 entry point
 * returns for the same (float32-representable) values; xy_picked_host (C,N,2) float32 is exact for
 * the same reason.  Halves the host-to-device bytes of the call (reference callers hold float64
 * arrays: kp2d.pickle after step 1's smoothing — they use the entry points above). */
int m3d_triangulate_error_host_f32(const m3d_rig* rig, const float* xy_host, int64_t N,
                                   int32_t undistort, double* p3d_host, double* err_host);
int m3d_triangulate_ransac_host_f32(const m3d_rig* rig, const float* xy_host, int64_t N,
                                    int32_t undistort, int32_t min_cams, double threshold,
                                    double init_best, double* p3d_host, uint8_t* picked_host,
                                    float* xy_picked_host, double* err_host, int32_t* subset_host,
                                    int32_t* neval_host);
/* The same pipelines on a SPAN of the arrays: every array is laid out for N_total points, the call
 * processes points [first, first + count) and touches nothing else.  Rigs of the same cameras on
 * different GPUs, each called from its own thread on a disjoint span, spread one host array over all
 * GPUs of the box in ONE process (CameraGroup does this for large numpy inputs). */
int m3d_triangulate_error_host_span(const m3d_rig* rig, const double* xy_host, int64_t N_total,
                                    int64_t first, int64_t count, int32_t undistort,
                                    double* p3d_host, double* err_host);
int m3d_triangulate_ransac_host_span(const m3d_rig* rig, const double* xy_host, int64_t N_total,
                                     int64_t first, int64_t count, int32_t undistort,
                                     int32_t min_cams, double threshold, double init_best,
                                     double* p3d_host, uint8_t* picked_host, double* xy_picked_host,
                                     double* err_host, int32_t* subset_host, int32_t* neval_host);
/* cudaHostRegister / cudaHostUnregister for caller-owned numpy buffers. */
int m3d_host_register(void* ptr, int64_t bytes);
int m3d_host_unregister(void* ptr);

/* ---- cross-view association (step2) ------------------------------------------------ */
/* geometry_affinity2 (step2_crossviewmatching.py:373-432) for F frames at once.
 *   kp_dev (F,M,J,3)   undistorted x, y and score of every detection, M <= M3D_MAX_DETS
 *   dim_dev (F,C+1) i32 cumulative detection counts per camera (dimGroup); detections
 *                      beyond dim[f][C] are padding and give distance 300 / affinity rows
 *                      that the caller ignores
 *   aff_dev (F,M,M)    affinity; dist_dev (F,M,M) mean ray distance [may be NULL]
 * thr_kp = 0.1 (THR_KP, step2:21). */
int m3d_ray_affinity(const m3d_rig* rig, const double* kp_dev, const int32_t* dim_dev,
                     int32_t F, int32_t M, int32_t J, double thr_kp, double* aff_dev,
                     double* dist_dev, void* stream);
/* mct.triangulatePoints (multicam_toolbox.py:433-486): inhomogeneous least squares
 * X = -pinv(A[:, :3]) A[:, 3] on already-undistorted points.
 *   xy_dev (C,N,2), use_dev (C,N) u8 (frame_use transposed), p3d_dev (N,3). */
int m3d_triangulate_ls(const m3d_rig* rig, const double* xy_dev, const uint8_t* use_dev,
                       int64_t N, double* p3d_dev, void* stream);
/* matchSVT (step2_crossviewmatching.py:130-216) with pselect = 1 and
 * dual_stochastic_SVT = False, F frames at once: W_dev (F,M,M) affinity,
 * dim_dev (F,C+1) i32, match_dev (F,M,M) u8, iters_dev (F) i32 [may be NULL]. */
int m3d_match_svt(const double* W_dev, const int32_t* dim_dev, int32_t F, int32_t M,
                  int32_t C, double alpha, double lambda, double mu, double tol,
                  int32_t max_iter, uint8_t* match_dev, int32_t* iters_dev, int32_t device,
                  void* stream);

/* Association weights of MultiEstimator.predict_data (step2_crossviewmatching.py:557-575) for F
 * keyframes: W = alpha_id * [same identity in different cameras] + (1 - alpha_id) * aff, zero where
 * aff <= 0 or NaN.  aff_dev (F,M,M), cid_dev (F,M) i32 identity label or -1, W_dev (F,M,M). */
int m3d_association_weights(const double* aff_dev, const int32_t* cid_dev, const int32_t* dim_dev,
                            int32_t F, int32_t M, int32_t C, double alpha_id, double* W_dev,
                            int32_t device, void* stream);
/* Person clusters of the match matrices (step2_crossviewmatching.py:598-607): a column with sum > 1.9
 * is a person, a detection belongs to the first person column it is matched to.
 * label_dev (F,M) i32: that column index, -1 = unmatched (or padding). */
int m3d_match_clusters(const uint8_t* match_dev, const int32_t* dim_dev, int32_t F, int32_t M, int32_t C,
                       int32_t* label_dev, int32_t device, void* stream);

/* ---- 2D keypoint filter of the step-4 stage ---------------------------------------- */
/* anipose filter_pose.viterbi_path (src/third_party/anipose/filter_pose.py:48-120, with
 * remove_dups :26-46 and the score threshold of filter_pose_viterbi :157) for S independent
 * series of F frames at once; caller: step4_aniposefiltering.py:144-170 (one series per
 * (animal, camera, joint)).
 *   cand_dev (S,F,P,3)  x, y, score of the P candidate detections per frame (not modified)
 *   n_back              frames a detection stays a particle (config n_back, step4: 3)
 *   thres_dist          scale of the transition model (config offset_threshold, step4: 25)
 *   score_threshold     candidates with score < threshold are dropped (step4: 0.3)
 *   dup_thres           candidates within this distance of an earlier one of the same frame
 *                       are dropped (the reference passes 5)
 *   out_dev (S,F,3)     chosen particle per frame: x, y, score * 2^-age; (-1,-1,0.001) = missing
 *   choice_dev (S,F) i32 age * P + candidate index of the choice, -1 = missing   [may be NULL]
 * Limits: n_back <= 8 and n_back * P <= 32. */
int m3d_viterbi_filter(const double* cand_dev, int64_t S, int64_t F, int32_t P, int32_t n_back,
                       double thres_dist, double score_threshold, double dup_thres,
                       double* out_dev, int32_t* choice_dev, int32_t device, void* stream);

/* ---- temporal / skeletal refinement of the 3D stage -------------------------------- */
/* CameraGroup.optim_points and optim_points_jointlenfix (cameras.py:1116-1270; the default
 * `optim = true` branch of step4_aniposefiltering.py:247-271) for ONE animal: minimise
 *   rho(|score (p2d - project(P))|)  +  scale_smooth diff_n(P)  +  limb-length terms
 * (the residual vector of _error_fun_triangulation, cameras.py:1560-1620) over the 3D points P (F,J,3)
 * and, unless fix_lengths, the limb lengths.
 *   p2d_dev (C,F,J,2) raw pixels, NaN = missing;  scores_dev (C,F,J) or NULL
 *   constraints (K,2) / constraints_weak (Kw,2): HOST arrays of joint index pairs
 *   loss: 0 linear, 1 soft_l1, 2 huber (reproj_loss);  n_deriv_smooth: order of the temporal difference
 *   params_dev: [P (F*J*3) | L (K) | Lw (Kw)], the start point x0 on entry (the reference's
 *               _initialize_params_triangulation), the solution on return (L | Lw are read-only when
 *               fix_lengths != 0)
 *   mode 0  solve: Levenberg-Marquardt with exact Jacobian blocks and a block-preconditioned CG inner
 *           solve, until the relative cost decrease of an accepted step is < ftol or max_iter steps;
 *           info_host[0..5] = final cost (0.5 |r|^2), initial cost, LM steps, CG iterations, residual
 *           evaluations, status (1 stationary, 2 ftol reached, 3 no further decrease, 0 max_iter)
 *   mode 1  out_dev <- residual vector at params: [r1 (C,F,J,2), NaN where p2d is NaN | r2 ((F-n),J,3) |
 *           r3 (K,F) | r4 (Kw,F)] — the reference's vector is this with the NaN entries dropped
 *   mode 2  out_dev holds a parameter-space vector v on entry and J v (same dense layout) on return */
int m3d_optim_points(const m3d_rig* rig, const double* p2d_dev, const double* scores_dev, int32_t F,
                     int32_t J, const int32_t* constraints, int32_t K, const int32_t* constraints_weak,
                     int32_t Kw, double scale_smooth, double scale_length, double scale_length_weak,
                     double reproj_error_threshold, int32_t loss, int32_t n_deriv_smooth,
                     int32_t fix_lengths, double ftol, int32_t max_iter, int32_t mode,
                     double* params_dev, double* out_dev, double* info_host, void* stream);

/* Batched keyframe association (crossview.associate_batch): the device stages around the kernels above.
 *   m3d_undistort_detections   kp_raw_dev (F,M,J,3) raw pixels + score -> kp_und_dev (F,M,J,3) undistorted
 *                              x, y (NaN in padding slots) + score; the camera of a slot follows from
 *                              dim_dev (F,C+1) (step2_crossviewmatching.py:306-325, 520-530)
 *   m3d_cluster_members        label_dev (F,M) from m3d_match_clusters -> persons as member tables
 *                              (step2:598-607, 697-698).  Pass 1 (offsets_dev NULL): count_dev (F) persons per
 *                              frame, dup_dev (F) u8 = 1 where a cluster holds two detections of one camera
 *                              (count 0: the caller resolves such frames, step2:610-657).  Pass 2 (offsets_dev
 *                              (F) i64 = exclusive prefix sums of count): frame_dev (P), column_dev (P),
 *                              members_dev (P,C) detection index per camera or -1, in (frame, column) order
 *   m3d_triangulate_ls_members calc_3dpose (step2:436-461) of P persons: camera c contributes keypoint j of
 *                              detection members[p][c] of frame[p] when x is not NaN and score >= thr_kp;
 *                              p3d_dev (P,J,3), NaN with fewer than two contributing cameras */
int m3d_undistort_detections(const m3d_rig* rig, const double* kp_raw_dev, const int32_t* dim_dev, int64_t F,
                             int32_t M, int32_t J, double* kp_und_dev, void* stream);
int m3d_cluster_members(const int32_t* label_dev, const int32_t* dim_dev, int32_t F, int32_t M, int32_t C,
                        const int64_t* offsets_dev, int32_t* count_dev, uint8_t* dup_dev, int32_t* frame_dev,
                        int32_t* column_dev, int32_t* members_dev, int32_t device, void* stream);
int m3d_triangulate_ls_members(const m3d_rig* rig, const double* kp_und_dev, const int32_t* frame_dev,
                               const int32_t* members_dev, int64_t P, int32_t M, int32_t J, double thr_kp,
                               double* p3d_dev, void* stream);

/* ---- multi-GPU result window ---------------------------------------- */
/* The reference is one NumPy process (cameras.py:639-743 returns one array for the whole recording);
 * the frame-sharded run keeps that contract by letting ONE rank own the frame-ordered result arrays
 * and every other rank (one process per GPU) write its finished tiles straight into them over
 * NVLink with its copy engine - no gather collective, no SM on either side.
 *   m3d_peer_alloc   owner: cudaMalloc `bytes` on `device`, export the M3D_PEER_HANDLE_BYTES handle
 *                    (send it to the other ranks by any means, e.g. torch.distributed object broadcast)
 *   m3d_peer_open    other ranks: map the owner's allocation for use from `device` (peer access to the
 *                    owner's GPU is enabled on demand); the pointer is valid for kernels and copies
 *   m3d_peer_push    asynchronous device-to-device copy of `bytes` into the window, ordered on `stream`
 *   m3d_peer_close / m3d_peer_free   unmap / release
 * Completion is the caller's: order a collective (or any cross-rank signal) after the pushes of every
 * rank before the owner reads (sharding.PeerResults.finish). */
int m3d_peer_alloc(int32_t device, int64_t bytes, void** dptr_out, uint8_t* handle_out);
int m3d_peer_open(int32_t device, const uint8_t* handle, void** dptr_out);
int m3d_peer_push(void* dst_window, const void* src_dev, int64_t bytes, void* stream);
int m3d_peer_close(int32_t device, void* dptr);
int m3d_peer_free(int32_t device, void* dptr);

/* ---- measurement helpers ----------------------------------------------------------- */
/* Number of kernels this library has launched on the calling process (bench.py's
 * gpu_launches). */
int64_t m3d_launch_count(void);
/* Per-kernel device times for bench.py's roofline: while enabled, the launches of the triangulation
 * and subset-search kernels are bracketed by CUDA events on their own stream.  m3d_profile_read
 * waits for the recorded launches, writes {"kernel": {"launches": n, "ms": total}, ...} (JSON,
 * NUL-terminated) into buf and clears the records.  Off by default; process-wide. */
int m3d_profile_enable(int32_t on);
int m3d_profile_read(char* buf, int64_t cap);
/* fp64 FMA peak probe: runs a dependent-chain DFMA kernel and returns TFLOP/s. */
int m3d_probe_fp64_tflops(int32_t device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* M3D_H_ */
