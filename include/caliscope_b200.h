/*
 * caliscope_b200 -- C ABI of the B200 bundle-adjustment engine.
 *
 * Drop-in boundary for the bundle-adjustment hot path of mprib/caliscope.  The
 * reference has no FFI; its seam is the Python call
 *     scipy.optimize.least_squares(joint_residuals, x0, args=(...), jac=joint_jacobian, ...)
 * at /root/reference/src/caliscope/core/capture_volume.py:387-411 (inside
 * CaptureVolume.optimize, :322-444).  The entry points below are what a ctypes
 * binding behind that seam calls; INTEGRATION.md shows the stub.
 *
 * Conventions: plain pointers and sizes, fp64, row-major, no exceptions cross the
 * ABI.  Every function returns 0 on success or a negative CB_E_* code;
 * cb_ba_error_string() gives the text, cb_ba_last_error() the detail (CUDA error
 * string) of the most recent failure on the calling thread.
 *
 * Parameter vector layout == BundleParameterization.pack
 * (/root/reference/src/caliscope/core/bundle_parameterization.py:127-149):
 *   x = [block_0 | ... | block_{n_cams-1} | X_0 | X_1 | ...],
 *   block_i = [rvec(3), tvec(3)] (+ [s, k1, k2] iff cam_flags[i] & CB_CAM_FREE_INTRINSICS),
 *   X_j = xyz of world point j.
 */
#ifndef CALISCOPE_B200_H
#define CALISCOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_BA_ABI_VERSION 2

/* cam_flags bits (CameraBlock.free_intrinsics / .fisheye, bundle_parameterization.py:36-51) */
#define CB_CAM_FREE_INTRINSICS 1
#define CB_CAM_FISHEYE 2

/* loss ids: scipy.optimize.least_squares(loss=...) as forwarded by capture_volume.py:405 */
#define CB_LOSS_LINEAR 0
#define CB_LOSS_SOFT_L1 1
#define CB_LOSS_HUBER 2
#define CB_LOSS_CAUCHY 3
#define CB_LOSS_ARCTAN 4

/* error codes */
#define CB_OK 0
#define CB_E_INVALID (-1)      /* bad argument (null pointer, index out of range, ...) */
#define CB_E_CUDA (-2)         /* CUDA runtime failure; see cb_ba_last_error() */
#define CB_E_NO_DEVICE (-3)    /* no usable CUDA device */
#define CB_E_UNSUPPORTED (-4)  /* feature not implemented by this build */
#define CB_E_CALLBACK (-5)     /* the all-reduce callback reported failure */
#define CB_E_NOMEM (-6)

typedef struct CbBaProblem CbBaProblem; /* opaque, device resident */

/*
 * Problem description == the args tuple CaptureVolume.optimize hands to scipy
 * (capture_volume.py:390-399) plus BundleParameterization.blocks flattened.
 *   cam_const[i*9 + 0..8] = fx_initial, fy_initial, cx, cy, c4..c8 where
 *     Brown-Conrady: (c4..c8) = (k1_initial, k2_initial, p1, p2, k3)
 *     fisheye      : (c4..c7) = (k1, k2, k3, k4), c8 unused
 *   obs_cam  = camera_indices          (capture_volume.py:353-355; int16 there, int32 here)
 *   obs_pt   = image_to_world_indices  (capture_volume.py:358)
 *   obs_xy   = image_coords            (capture_volume.py:357), n_obs x 2
 * obs_* may be host pointers (obs_on_device = 0; copied inside the call) or device
 * pointers on `device` (obs_on_device = 1).  Inputs are never modified.
 */
typedef struct {
  int32_t n_cams;
  int32_t n_pts;
  int64_t n_obs;
  const int32_t* cam_flags; /* host, n_cams */
  const double* cam_const;  /* host, n_cams*9 */
  const int32_t* obs_cam;
  const int32_t* obs_pt;
  const double* obs_xy;
  int32_t obs_on_device;
  int32_t obs_cam_bits; /* 0 or 32: obs_cam is int32; 16: obs_cam is int16 (what capture_volume.py:353-355 builds) and is
                           widened on the device -- halves that upload and saves the caller a conversion pass */
  /* Optional (NULL: the engine decides).  Internal camera order, host, n_cams entries: cam_order[slot] = camera index.
   * Only the layout of the reduced camera system depends on it (cameras that see the same points should be neighbours,
   * so that whole 96-column tile pairs of the Schur product are empty); no input or output of the ABI is reordered.
   * Every rank of a sharded solve MUST pass the same order (caliscope_b200.distributed does). */
  const int32_t* cam_order;
} CbBaProblemDesc;

/*
 * Sum-all-reduce hook for observation sharding (one process per GPU).  Called on
 * the host thread inside cb_ba_solve with a DEVICE buffer of n doubles that must be
 * summed element-wise across ranks in place, ordered after all work already queued
 * on `stream` and before any work queued afterwards (torch.distributed.all_reduce on
 * the current stream satisfies this).  Return 0 on success.
 */
typedef int (*CbAllReduceSum)(void* user, double* device_buf, int64_t n, void* stream);

/* least_squares keyword arguments the reference passes (capture_volume.py:403-410)
 * plus scipy's defaults for the ones it leaves out (xtol = gtol = 1e-8). */
typedef struct {
  double ftol;
  double xtol;
  double gtol;
  int64_t max_nfev; /* <= 0: scipy's default 100 * n */
  int32_t loss;     /* CB_LOSS_* */
  double f_scale;
  int32_t verbose;     /* 0 silent, 1 summary, 2 per-iteration table on stderr */
  int32_t use_bounds;  /* 1: s in [0.5,2], k1 in [-1,1], k2 in [-2,2] (bundle_parameterization.py:151-164) */
  double lambda0;      /* initial LM damping; <= 0: 1e-4 */
  double pcg_tol;      /* relative PCG tolerance for the reduced camera system; <= 0: 1e-6 */
  int32_t pcg_max_iter; /* <= 0: 4 * n_camera_params */
  CbAllReduceSum allreduce; /* NULL: single GPU (unless nccl_comm is set) */
  void* allreduce_user;
  void* nccl_comm;          /* ncclComm_t from cb_nccl_comm_create: the engine calls ncclAllReduce itself on the solve stream */
  void* peer_group;         /* CbPeerGroup* from cb_peer_create/_connect: all-reduce over NVLink peer memory, fused into
                               the Schur finalize kernel (takes precedence over nccl_comm / allreduce) */
  int32_t rank;       /* informational (verbose output only on rank 0) */
  int32_t world_size; /* 1 if allreduce is NULL */
  int32_t time_kernels; /* diagnostic: launch every trial directly with CUDA events around the point pass and the Schur
                           product (CbBaResult.rj_ms / syrk_ms); 0 (default): the loop is replayed from CUDA graphs, where
                           events cannot be read back */
  int32_t pad_;
} CbBaOptions;

/* mirrors scipy.optimize.OptimizeResult fields the reference reads
 * (capture_volume.py:413-433): status, nfev, cost; plus njev/nit/optimality. */
typedef struct {
  int32_t status; /* scipy codes: 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 ftol&xtol */
  int64_t nfev;
  int64_t njev;
  int64_t nit;       /* LM iterations (linearisations solved) */
  double cost;       /* 0.5 * sum rho(f^2), scipy's definition */
  double initial_cost;
  double optimality; /* inf-norm of the gradient at the solution */
  double lambda_final;
  int64_t pcg_iterations; /* total PCG iterations over the solve */
  int64_t kernel_launches;
  double solve_ms; /* device time of the LM loop (CUDA events on the solve stream) */
  double rj_ms;    /* total device time spent in the residual+Jacobian (point pass) kernel, CUDA events around each launch */
  int64_t rj_launches;
  double syrk_ms;  /* total device time spent in the Schur product kernel */
  int64_t syrk_launches;
  int64_t trials_queued; /* LM trials handed to the device (the last one runs predicated-off) */
  int32_t used_graph;    /* 0: trials launched directly, 1: one CUDA graph per trial, 2: device loop (one WHILE-conditional graph per solve) */
  int32_t pad_;
} CbBaResult;

int cb_ba_abi_version(void);
const char* cb_ba_error_string(int code);
const char* cb_ba_last_error(void);
void cb_ba_default_options(CbBaOptions* opt);

/* Upload + index build (sort by camera / by point, chunk and pair tables). */
int cb_ba_problem_create(const CbBaProblemDesc* desc, int device, void* stream, CbBaProblem** out);
int cb_ba_problem_destroy(CbBaProblem* p);
int64_t cb_ba_problem_n_params(const CbBaProblem* p);
/* Facts about how the engine laid the problem out (measurement / diagnostics): what = 0: 1 if the Schur product walks
 * compacted row lists (sparse visibility), 1: floating-point operations one Schur-product launch issues, 2: 1 if the reduced
 * system is solved directly (<= 96 camera parameters), 3: CTAs of the Schur product.  -1 for an unknown key. */
double cb_ba_problem_stat(const CbBaProblem* p, int what);

/* Rigid-distance constraint rows (reprojection.py:112-117 and :207-226): groups_a / groups_b are n_c x 4 world-point
 * row indices, distances and weights n_c doubles -- the arrays CaptureVolume._build_constraint_arrays produces
 * (capture_volume.py:446-516) with weights = (pixel_sigma / f_median) / sigma (:377-381).  Host pointers, copied.
 * Call at most once, after cb_ba_problem_create.  Under observation sharding every point a row touches must belong
 * to the same rank (shard by connected component of the constraint graph). */
int cb_ba_problem_set_constraints(CbBaProblem* p, int64_t n_c, const int32_t* groups_a, const int32_t* groups_b,
                                  const double* distances, const double* weights, void* stream);
int64_t cb_ba_problem_n_constraints(const CbBaProblem* p);

/* Constraint rows at x: r_out (n_c) == the tail of joint_residuals; dir_out (n_c x 3, nullable) = weight * unit vector
 * between the two endpoint means: the Jacobian entry of a group-a (group-b) member is +(-) dir / 4, repeats summed. */
int cb_ba_constraint_rows(CbBaProblem* p, const double* x, double* r_out, double* dir_out, void* stream);

/* Replaces least_squares(joint_residuals, x0, jac=joint_jacobian, method="trf", ...)
 * (capture_volume.py:387-411).  x_inout: host, n_params doubles, overwritten with result.x. */
int cb_ba_solve(CbBaProblem* p, const CbBaOptions* opt, double* x_inout, CbBaResult* result, void* stream);

/* == joint_residuals (reprojection.py:75-119), reprojection rows only: r_out host, 2*n_obs,
 * interleaved (x, y) / fx_initial in the caller's observation order. */
int cb_ba_residuals(CbBaProblem* p, const double* x, double* r_out, void* stream);

/* Dense blocks of joint_jacobian (reprojection.py:171-205) in the caller's observation order:
 * Jc host n_obs*2*9 (columns rvec3, tvec3, s, k1, k2; zeros where a block is locked),
 * Jp host n_obs*2*3; both already divided by fx_initial. */
int cb_ba_jacobian_blocks(CbBaProblem* p, const double* x, double* Jc, double* Jp, void* stream);

/* == reprojection_errors (reprojection.py:35-72): pixel errors with the intrinsics in x. err_xy host n_obs*2. */
int cb_ba_reproj_errors_px(CbBaProblem* p, const double* x, double* err_xy, void* stream);

/* Test/diagnostic access to one damped linearisation (all host outputs, any may be NULL):
 * U n_cams*P*P, gc n_cams*P, V n_pts*9, gp n_pts*3, S (n_cams*P)^2, b n_cams*P,
 * dc n_cams*P (PCG solution of S dc = -b), dp n_pts*3 (back-substituted), with P = cb_ba_cam_stride(). */
int cb_ba_cam_stride(const CbBaProblem* p);
int cb_ba_normal_equations(CbBaProblem* p, const double* x, double lambda, int32_t loss, double f_scale,
                           double* cost, double* U, double* gc, double* V, double* gp, double* S, double* b,
                           double* dc, double* dp, void* stream);

/* Per-camera order statistics of the pixel error norm for the percentile filter
 * (capture_volume.py:709-753): for camera c with n_c observations, lo[c] / hi[c] are the
 * floor / ceil order statistics of rank (n_c - 1) * q / 100 (numpy's linear interpolation
 * nodes), count[c] = n_c; err host n_obs (euclidean error per observation, caller order). */
int cb_ba_error_order_stats(CbBaProblem* p, const double* x, double q_percent, double* err, double* lo,
                            double* hi, int64_t* count, void* stream);

/* Overall and per-camera (nullable, n_cams) RMS pixel error, reduced on the device
 * (ReprojectionReport.overall_rmse / by_camera, capture_volume.py:197-202). */
int cb_ba_rmse_px(CbBaProblem* p, const double* x, double* overall, double* per_camera, void* stream);

/* Device-side observation cull == _filter_by_reprojection_thresholds (capture_volume.py:607-646) at array level:
 * keep error <= thresholds[camera] (host, n_cams), restore lowest-error observations up to min_per_camera, compact
 * the observation list on the device and build the filtered problem (same cameras and point numbering) from it.
 * keep_mask: host, n_obs bytes in the caller's observation order, or NULL.  *out must be destroyed by the caller.
 * min_per_camera = 0 applies the thresholds only (sharded solves: the floor is a global property and is folded into the
 * thresholds by the caller, caliscope_b200/distributed.global_cull_thresholds). */
int cb_ba_cull(CbBaProblem* p, const double* x, const double* thresholds, int32_t min_per_camera, CbBaProblem** out,
               int64_t* n_kept, uint8_t* keep_mask, void* stream);

/* Diagnostic: mean milliseconds of one PCG-kernel launch forced to run exactly max_iter iterations on the
 * system left by the last cb_ba_normal_equations call. */
int cb_ba_debug_pcg_time(CbBaProblem* p, int max_iter, int reps, double* ms_per_launch, void* stream);

/* Measured fp64 throughput of the device (TFLOP/s): `mma.sync.m8n8k4.f64` (DMMA, the tensor path the Schur product
 * runs on) and plain DFMA register chains, 8 warps per SM, a few milliseconds each.  The roofline denominator for the
 * Schur product; there is no driver-measured fp64 figure in MEASURED_PEAKS.json. */
int cb_debug_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops);

/* ---- the step in front of bundle adjustment ------------------------------------------- */

/* == CameraData.undistort_points (reference cameras/camera_array.py:135-174): cv2.undistortPoints (5 fixed-point
 * iterations) / cv2.fisheye.undistortPoints (Newton on theta) on float32 copies of the points, float32 results,
 * here for the observations of all cameras in one launch.
 *   cam_fisheye[n_cams]; cam_k[n_cams][5] = fx,fy,cx,cy,skew; cam_dist[n_cams][12] = k1 k2 p1 p2 k3 k4 k5 k6 s1..s4
 *   (zero-padded; fisheye: k1..k4); obs_cam[n] camera row per point (NULL with a single camera);
 *   to_pixels: 0 = normalised image plane (P = I), 1 = pixels (P = K).
 *   on_device: xy_in / obs_cam / xy_out are device pointers. */
int cb_undistort_points(int32_t n_cams, const int32_t* cam_fisheye, const double* cam_k, const double* cam_dist,
                        int64_t n, const int32_t* obs_cam, const double* xy_in, int on_device, int to_pixels,
                        double* xy_out, int device, void* stream);

typedef struct CbTriStats {
  double group_ms;  /* upload + radix sort + group boundaries */
  double dlt_ms;    /* the DLT kernel alone (CUDA events on the launch stream) */
  double total_ms;
  int32_t kernel_launches;
  int32_t pad_;
} CbTriStats;

/* == triangulate_image_points (reference core/point_data.py:122-229).  Observations with equal obs_key (a
 * non-negative 63-bit packing of (sync_index, object_id, keypoint_id) made by the caller) form one group; every
 * group gets the DLT point of its rows  [x*P2 - P0 ; y*P2 - P1]  (smallest right singular vector, via the 4x4 normal
 * matrix), groups in ascending key order.  proj = host [n_cams][3][4] normalised projection matrices.
 * Outputs (host, room for max_groups): xyz[g][3] (NaN when the group has < 2 rows), count[g], rep_row[g] = caller
 * row of the group's first observation, camset_sig[g][2] = order-independent 128-bit signature of the group's
 * camera multiset (lets the host mirror reproduce the reference's by-camera-set output order). */
int cb_triangulate_dlt(int32_t n_cams, const double* proj, int64_t n_obs, const int32_t* obs_cam,
                       const int64_t* obs_key, const double* obs_xy, int obs_on_device, int32_t max_groups,
                       int32_t* n_groups_out, double* xyz_out, int32_t* count_out, int32_t* rep_row_out,
                       uint64_t* camset_sig_out, CbTriStats* stats, int device, void* stream);

/* cb_undistort_points (normalised output) + cb_triangulate_dlt in one call: obs_px are raw pixels, the undistorted
 * coordinates stay in HBM between the two kernels (what ImagePoints.triangulate does, point_data.py:416-559). */
int cb_undistort_triangulate(int32_t n_cams, const int32_t* cam_fisheye, const double* cam_k, const double* cam_dist,
                             const double* proj, int64_t n_obs, const int32_t* obs_cam, const int64_t* obs_key,
                             const double* obs_px, int obs_on_device, int32_t max_groups, int32_t* n_groups_out,
                             double* xyz_out, int32_t* count_out, int32_t* rep_row_out, uint64_t* camset_sig_out,
                             CbTriStats* stats, int device, void* stream);

/* Optional NCCL transport owned by the engine (no host callback per all-reduce).  NCCL is resolved at run time from
 * the libnccl the process already has loaded (PyTorch's).  Rank 0 calls cb_nccl_unique_id and distributes the 128
 * bytes (e.g. torch.distributed.broadcast); every rank then calls cb_nccl_comm_create (collective). */
int cb_nccl_unique_id(char id_out[128]);
int cb_nccl_comm_create(const char id[128], int rank, int world_size, int device, void** comm_out);
int cb_nccl_comm_destroy(void* comm);

/* Peer-memory transport (one node, NVLink/NVSwitch, one process per GPU).  Every rank creates its symmetric buffer
 * (capacity_doubles >= (n_cams*P)^2 + 3*n_cams*P + 65 for the problems it will solve, P = 9 with free intrinsics
 * else 6), the 64-byte CUDA IPC handles are exchanged by any means, then every rank connects with all handles in
 * rank order.  During cb_ba_solve the reduced camera system is then summed by schur_finalize_peer_kernel reading
 * the peers' buffers directly; no NCCL call is made.  world_size <= 16.  Destroy only after every rank is done. */
typedef struct CbPeerGroup CbPeerGroup;
int cb_peer_create(int rank, int world_size, int device, int64_t capacity_doubles, CbPeerGroup** out,
                   char handle_out[64]);
int cb_peer_connect(CbPeerGroup* group, const char* handles /* world_size x 64 bytes */);
int cb_peer_destroy(CbPeerGroup* group);

/* Number of kernel launches issued by this library in the calling process so far. */
int64_t cb_ba_launch_count(void);


/* ------------------------------------------------------------------------------------------------------------------
 * Extrinsic bootstrap (what produces bundle adjustment's start vector).
 *
 * cb_pnp_ippe == compute_camera_to_object_poses_pnp (reference core/bootstrap_pose/pose_network_builder.py:211-330):
 *   obs_px are raw pixels of camera obs_cam (undistorted on the device with float32 rounding exactly as
 *   CameraData.undistort_points does, cameras/camera_array.py:135-174), obs_obj the object-frame coordinates (n_obs x 3,
 *   NaN z = 0), obs_key >= 0 packs (camera, sync_index, object) -- rows with equal key form one PnP group, groups come
 *   back in ascending key order (== the reference's groupby order when the key is packed camera-major).  Per group:
 *   R (3x3 row-major) and t of the object in the camera frame (cv2.solvePnP(SOLVEPNP_IPPE) + cv2.Rodrigues), the
 *   reprojection RMSE in the normalised plane (:317-318), status (0 ok, 1 fewer than min_points rows, 2 non-planar
 *   target -- the reference switches to SQPNP there, which this build does not implement --, 3 degenerate), row count and
 *   one representative caller row.  All pointers host.
 *
 * cb_stereo_rmse == calculate_stereo_rmse_for_pair for every pair at once (:638-685, with the common observations of
 *   _precompute_common_observations :576-603): pair p = cameras (pair_a[p] < pair_b[p]) with pose pair_Rt[p] = [R (9) | t (3)]
 *   (camera a at the origin).  obs_key >= 0 packs (sync_index, object, keypoint).  Outputs per pair: rmse (NaN when the
 *   pair has fewer than min_common common observations, the reference returns None there) and the number of common
 *   observations.
 * ------------------------------------------------------------------------------------------------------------------ */
int cb_pnp_ippe(int32_t n_cams, const int32_t* cam_fisheye, const double* cam_k, const double* cam_dist, int64_t n_obs,
                const int32_t* obs_cam, const int64_t* obs_key, const double* obs_px, const double* obs_obj,
                int32_t min_points, int32_t max_groups, int32_t* n_groups_out, double* R_out, double* t_out,
                double* rmse_out, int32_t* status_out, int32_t* count_out, int32_t* rep_row_out, CbTriStats* stats,
                int device, void* stream);

int cb_stereo_rmse(int32_t n_cams, const int32_t* cam_fisheye, const double* cam_k, const double* cam_dist,
                   int32_t n_pairs, const int32_t* pair_a, const int32_t* pair_b, const double* pair_Rt, int64_t n_obs,
                   const int32_t* obs_cam, const int64_t* obs_key, const double* obs_px, int32_t min_common,
                   double* rmse_out, int64_t* count_out, CbTriStats* stats, int device, void* stream);

/* cb_relative_pose_network == compute_relative_poses (pose_network_builder.py:484-531) -> reject_outliers (:331-413) ->
 *   aggregate_poses (:533-573, quaternion_average :416-438) for every camera pair in one device pass; the relative poses
 *   never leave the device.
 *   Input: the PnP poses (camera <- object) of the cameras the array does not ignore, sorted by (sync_index, object_id,
 *   camera id): group g has camera id cam_id[g], position cam_pos[g] in the reference's camera dict, pose R[9g..], t[3g..]
 *   (NaN poses allowed: cv2's degenerate groups); frame_start[f] .. frame_start[f+1] are the groups of one (sync_index,
 *   object_id).  Every combination (i < j) of a frame's groups is one candidate relative pose; the reference forms it only if
 *   cam_pos[i] < cam_pos[j] (its `combinations(dict order) if a < b` quirk).
 *   rot_mult / tr_mult: IQR multipliers of the rotation-angle / translation-magnitude rules (1.5 in the reference).
 *   Output, camera pairs in ascending (a, b), at most max_pairs: ids, aggregated R (9) and t (3), number of samples kept.
 *   Optional (n_rel = sum over frames of s(s-1)/2, frame-major, np.triu_indices order inside a frame): rel_valid[m] = 1 when
 *   the combination was formed, rel_keep[m] = 1 when it passed the NaN filter and the IQR rule. */
int cb_relative_pose_network(int32_t n_groups, int32_t n_frames, const int32_t* frame_start, const int32_t* cam_id,
                             const int32_t* cam_pos, const double* R, const double* t, double rot_mult, double tr_mult,
                             int32_t max_pairs, int32_t* n_pairs_out, int32_t* pair_a, int32_t* pair_b, double* R_out,
                             double* t_out, int64_t* count_out, int64_t n_rel, uint8_t* rel_valid, uint8_t* rel_keep,
                             CbTriStats* stats, int device, void* stream);


/* Host-side shard selection of a sharded solve (distributed.shard_points): the observations whose point index lies in
 * [pt_lo, pt_hi), in the caller's order, point index made local (pt - pt_lo).  All host threads; no device work.
 * With every output pointer null it only counts (*n_sel_out).  sel_index = positions in the caller's list. */
int cb_shard_select(int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt, const double* obs_xy, int32_t pt_lo,
                    int32_t pt_hi, int64_t capacity, int64_t* n_sel_out, int64_t* sel_index, int32_t* cam_out,
                    int32_t* pt_out, double* xy_out, int32_t n_threads);

/* ------------------------------------------------------------------------------------------------------------------
 * Numeric CSV tables at the boundary of the path: xy_<TRACKER>.csv / xyz_<TRACKER>.csv as written
 * by ImagePoints.to_csv / WorldPoints.to_csv (reference core/point_data.py:358-373, 662-677 through
 * persistence._safe_write_csv, persistence.py:27-41: `to_csv(index=False, float_format="%.6f")`, temp file + fsync +
 * rename) and read by pd.read_csv.  Byte-identical files, all host cores; no device work.
 *   col_kind[c]: 0 = int64 column, 1 = float64 column (NaN -> empty field).  header: the column names joined by ','.
 *   cb_csv_parse_numeric: out is column-major float64 [n_cols][n_rows] (sizes from cb_csv_scan); col_all_int / col_has_empty
 *   let the caller rebuild pandas' dtypes (int64 iff every field is an integer literal).
 * ------------------------------------------------------------------------------------------------------------------ */
int cb_csv_write_numeric(const char* path, const char* header, int64_t n_rows, int32_t n_cols, const int32_t* col_kind,
                         const void* const* col_data, int32_t n_threads);
int cb_csv_scan(const char* path, int64_t* n_rows, int32_t* n_cols);
int cb_csv_parse_numeric(const char* path, int64_t n_rows, int32_t n_cols, double* out, int32_t* col_all_int,
                         int32_t* col_has_empty, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* CALISCOPE_B200_H */
