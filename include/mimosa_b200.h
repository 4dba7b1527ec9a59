/* mimosa_b200.h — C ABI of the B200-native LiDAR geometric-factor path (scan-to-map point-to-plane ICP).
 *
 * Drop-in boundary for ntnu-arl/mimosa's  lidar::Geometric -> ICPFactor -> IncrementalVoxelMapPCL  path.
 * Every entry point names the reference interface it replaces (paths relative to the mimosa repo).
 * Plain pointers and sizes only; every call returns an int status (0 = MB_OK); no exception and no
 * CUDA/torch type crosses this boundary.  All host buffers are caller-owned.  There is NO CPU fallback:
 * without a usable sm_100 device mb_init() fails with MB_ERR_NO_DEVICE.
 *
 * Threading: ONE CALLER AT A TIME PER CONTEXT.  Every map, factor and scan handle created from an mb_ctx shares
 * that context's stream, device-block pool, page-locked staging buffers and completion flag, so calls on any handles
 * of one context must be serialised by the caller (the reference already serialises them: ICPFactor::linearize is
 * `const` but mutates per-point caches, mimosa/include/mimosa/lidar/geometric_factor.hpp:79-116, and instances in the
 * smoother are linearised under graph_mutex_, mimosa/src/graph/manager.cpp:574; the one call outside that mutex,
 * geometric.cpp:196, runs on the LiDAR callback thread, which an adapter guards with the same lock).  Handles of
 * DIFFERENT contexts (one per thread, or one per GPU) may be used concurrently.
 */
#ifndef MIMOSA_B200_H_
#define MIMOSA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_API __attribute__((visibility("default")))

enum mb_status {
  MB_OK = 0,
  MB_ERR_INVALID_ARG = 1,
  MB_ERR_NO_DEVICE = 2,   /* no CUDA device / not sm_100: the product never falls back to the CPU */
  MB_ERR_CUDA = 3,        /* a CUDA runtime call failed; see mb_last_error() */
  MB_ERR_UNSUPPORTED = 4, /* project_on_degneneracy=true, binary factor, k>MB_MAX_K, coordinate out of range */
  MB_ERR_NCCL = 5,
  MB_ERR_CAPACITY = 6
};

/* Per-point state of ICPFactor::RejectStatus, mimosa/include/mimosa/lidar/geometric_factor.hpp:35-46. */
enum mb_reject_status {
  MB_UNPROCESSED = 0,
  MB_INSUFFICIENT_CORRES_POINTS = 1,
  MB_CORRES_MAX_DIST = 2,
  MB_EIGEN_SOLVER_FAIL = 3,
  MB_MIN_EIGEN_VALUE_LOW = 4,
  MB_LINE = 5,
  MB_CORRES_PLANE_INVALID = 6,
  MB_MAX_ERROR = 7,
  MB_VALID = 8
};

#define MB_MAX_K 8         /* num_corres_points supported by the device path (reference default 5) */
#define MB_POINT_STRIDE 32 /* sizeof(mimosa::lidar::Point), mimosa/include/mimosa/lidar/point.hpp:18-39 */

/* mimosa::lidar::RegistrationConfig, mimosa/include/mimosa/lidar/geometric_config.hpp:17-33.
 * Same fields, same order, reals kept as float (they are promoted to double where the reference
 * promotes them: geometric_factor.hpp:283,299,333-339). */
typedef struct mb_icp_config {
  float source_voxel_grid_filter_leaf_size;
  float source_voxel_grid_min_dist_in_voxel;
  float target_ivox_map_leaf_size;
  float target_ivox_map_min_dist_in_voxel;
  uint64_t num_corres_points;
  float max_corres_distance;
  float plane_validity_distance;
  float lidar_point_noise_std_dev;
  int32_t use_huber;
  float huber_threshold;
  int32_t reg_4_dof;
  int32_t project_on_degneneracy; /* must be 0 (every shipped config); 1 -> MB_ERR_UNSUPPORTED */
  float degen_thresh_rot;
  float degen_thresh_trans;
} mb_icp_config;

/* What ICPFactor::linearize hands to GTSAM plus what Geometric::getFactors reads back from the factor:
 *   HessianFactor(key, G = H, g, f)               geometric_factor.hpp:559-560
 *   getLocalizabilities()                          geometric_factor.hpp:52-62  (used at geometric.cpp:208-228)
 *   getDegenInfo()                                 geometric_factor.hpp:64-70
 *   status histogram                               geometric.cpp:280-323 (LidarGeometricDebug.msg:11-19)
 *   getLinearizeCount()                            geometric_factor.hpp:72
 * Matrices are row-major; column j of an eigenvector matrix belongs to eigenvalue j (ascending). */
typedef struct mb_linearization {
  double H[36];
  double g[6]; /* = -J^T e */
  double f;    /* = sum e^2 */
  int64_t counts[9];
  double loc_trans_comp[3], loc_rot_comp[3], loc_trans_final[3], loc_rot_final[3];
  double eigvec_trans[9], eigvec_rot[9];
  double degen_rot[3], degen_trans[3], degen_eigvec_rot[9], degen_eigvec_trans[9];
  int32_t linearize_count;
  int32_t n_searched; /* points that redid data association in this call (telemetry, not in the reference) */
} mb_linearization;

/* One iteration of the stand-alone Gauss-Newton harness (stands in for ISAM2's update(),
 * mimosa/src/graph/manager.cpp:585-588):  delta = (H + lambda I)^-1 g,  T <- T * Pose3::Expmap(delta). */
typedef struct mb_icp_trace {
  double H[36], g[6], f, delta[6];
  double R[9], t[3]; /* pose after this iteration's retract */
  int64_t counts[9];
  int32_t n_searched;
  int32_t solve_ok;
  double loc_trans_comp[3], loc_rot_comp[3]; /* component localizabilities of this iteration's linearisation */
} mb_icp_trace;

/* Where the fields of one point sit inside a sensor_msgs/PointCloud2 record (one descriptor covers the nine
 * vendor structs of mimosa/include/mimosa/lidar/point.hpp:41-178).  Offsets in bytes, -1 = field absent. */
typedef struct mb_cloud_layout {
  uint32_t point_step;
  int32_t off_x, off_y, off_z; /* float32 */
  int32_t off_intensity;
  int32_t intensity_type; /* 0: float32 intensity; 1: uint16 reflectivity (PointOusterOdyssey) */
  int32_t off_time;
  int32_t time_type; /* 0: uint32 ns since header (Ouster t, LivoxFromCustom2 t); 1: float32 s since header (Velodyne time);
                        2: float64 absolute s (Hesai / Rslidar timestamp); 3: float64 absolute ns (Livox timestamp) */
  int32_t off_ring;  /* -1: no ring field (Livox, OusterOdyssey) */
  int32_t ring_type; /* 0: uint16; 1: uint8 (PointOusterR8); 2: float32 (PointVelodyneAnybotics) */
  int32_t off_tag;   /* Livox tag byte, -1 otherwise */
  int32_t ring_filter; /* 1: drop geometric points with ring % ring_skip_divisor != 0 (manager.cpp:318-330); 0: the ring
                          field only serves organize_pointcloud_by_ring — the reference skips the filter for
                          PointVelodyneAnybotics ("unreliable ring numbers") although it re-orders that cloud by ring */
} mb_cloud_layout;

/* lidar::ManagerConfig's input filters (mimosa/include/mimosa/lidar/manager.hpp:30-34) + the two skip divisors of
 * GeometricConfig (geometric_config.hpp:48-49); float fields stay float like the reference's. */
typedef struct mb_input_filter {
  float intensity_min, intensity_max, range_min, range_max;
  float ns_max; /* compared as (float)t_ns > ns_max, manager.cpp:306 */
  int32_t point_skip_divisor, ring_skip_divisor, create_full_res_pointcloud;
  float z_offset;   /* manager.cpp:18 */
  double header_ts; /* msg.header.stamp in seconds */
} mb_input_filter;

/* The two message re-orderings lidar::Manager::prepareInput applies before its filter loop (manager.cpp:179-243;
 * ManagerConfig::transpose_pointcloud / organize_pointcloud_by_ring, manager.hpp:28-29).  width x height is the
 * PointCloud2's organisation (height == 1: unorganised).  The reference transposes only PointRslidar and
 * PointVelodyneAnybotics clouds and never re-orders Livox / OusterOdyssey clouds by ring (`if constexpr` on the
 * point type): the adapter passes the flags accordingly. */
typedef struct mb_cloud_order {
  uint32_t width, height;
  int32_t transpose_pointcloud;        /* index i of the transposed cloud = old row i % height, old column i / height */
  int32_t organize_pointcloud_by_ring; /* stable re-ordering by ring number; applied only when the (transposed) cloud's
                                          height is 1, as in the reference.  Ring numbers >= 128 index out of range in
                                          the reference (manager.cpp:213-219); here any 16-bit ring is ordered */
} mb_cloud_order;

typedef struct mb_ctx mb_ctx;
typedef struct mb_map mb_map;
typedef struct mb_factor mb_factor;
typedef struct mb_scan mb_scan;

/* ---- context ---------------------------------------------------------------------------------------
 * One context per process and GPU (one process per GPU; `device` is the CUDA ordinal, normally
 * LOCAL_RANK).  Replaces nothing in the reference (it has no device) — the analogue of constructing
 * lidar::Geometric, mimosa/src/lidar/geometric.cpp:13-45. */
MB_API int mb_init(int device, mb_ctx** out);
MB_API int mb_shutdown(mb_ctx* ctx);
MB_API const char* mb_last_error(void);
MB_API int mb_version(void);
/* sizeof of the ABI structs as compiled (0: mb_icp_config, 1: mb_linearization, 2: mb_icp_trace) so a foreign
 * binding can verify its layout. */
MB_API size_t mb_sizeof(int which);
/* Synchronise the context's stream. */
MB_API int mb_sync(mb_ctx* ctx);
/* CUDA-event timer on the context's stream (device time of everything enqueued between the calls). */
MB_API int mb_timer_begin(mb_ctx* ctx);
MB_API int mb_timer_end(mb_ctx* ctx, float* ms);
/* Number of kernel launches this context has issued since creation (bench.py's gpu_launches). */
MB_API int mb_launch_count(mb_ctx* ctx, uint64_t* out);
/* Residency window of the linearisation kernel.  ICPFactor::linearize is called several times in a row on
 * one factor (ISAM2's update and its additional iterations, mimosa/src/graph/manager.cpp:585-588; the reference has
 * no analogue of a launch).  After mb_factor_linearize has handed over its result the kernel stays on the device for
 * `microseconds` and a call on the same factor that arrives inside the window only posts its pose through mapped
 * memory instead of launching (41 -> 30 us per call measured); anything else enqueued on the context meanwhile simply
 * waits for the window to close, and every other mb_factor_ / mb_sync / mb_timer_ call closes it at once.  0 turns it off.
 * Default 30, or the environment variable MB_RESIDENT_US at mb_init. */
MB_API int mb_set_resident_window(mb_ctx* ctx, unsigned microseconds);
/* Write `bytes` of device memory (L2 flush between timed iterations). */
MB_API int mb_flush_l2(mb_ctx* ctx, size_t bytes);
/* Same purpose by READING `bytes` of device memory: the L2 ends up full of clean lines (no write-back traffic during
 * the timed kernel).  bench.py reports the k-NN roofline under both protocols. */
MB_API int mb_flush_l2_read(mb_ctx* ctx, size_t bytes);
/* Page-lock a caller-owned host buffer (cudaHostRegister) so that scans handed to mb_factor_create /
 * mb_scan_upload from it are fetched by DMA without a CPU staging pass; e.g. the point buffer a LiDAR driver
 * re-uses for every scan.  Purely an optimisation: pageable buffers work everywhere. */
MB_API int mb_host_register(void* p, size_t bytes);
MB_API int mb_host_unregister(void* p);

/* Multi-GPU: scan blocks are sharded across ranks, the map is replicated, and the packed normal equations
 * are all-reduced once per iteration (new in this implementation; the reference is single-process).
 * mb_comm_unique_id fills a 128-byte NCCL id on rank 0; the caller ships it to the other ranks
 * (torch.distributed / MPI / anything) and every rank calls mb_comm_init. */
MB_API int mb_comm_unique_id(void* id128);
MB_API int mb_comm_init(mb_ctx* ctx, int rank, int world, const void* id128);
/* Single node: exchange the per-iteration packet through peer memory (NVLink / NVSwitch) inside the persistent ICP
 * kernel instead of an NCCL all-reduce between kernels.  After mb_comm_init every rank calls mb_comm_ipc_handle
 * (64-byte CUDA IPC handle of its mailbox), the caller all-gathers the handles in rank order, and every rank calls
 * mb_comm_ipc_open with all `world` handles.  From then on block 0 of the loop kernel stores its packet into every
 * rank's mailbox (flag-in-data words: no fence, no flag), sums the mailbox in rank order and publishes the sum to the
 * other blocks — no collective call on that path.  Without this set-up the library runs one kernel chain per
 * linearisation with ncclAllReduce in between. */
MB_API int mb_comm_ipc_handle(mb_ctx* ctx, void* handle64);
MB_API int mb_comm_ipc_open(mb_ctx* ctx, const void* handles /* world x 64 bytes */);
/* Device-side barrier of all ranks on the context's stream (after mb_comm_ipc_open): a one-block kernel that exchanges
 * a word through the peer mailboxes, so that what every rank enqueues next starts within a few microseconds of the
 * others.  bench.py puts it in front of each timed step: a host-side barrier alone releases eight processes hundreds of
 * microseconds apart, and the early ranks would count that wait inside their first in-kernel exchange. */
MB_API int mb_comm_barrier(mb_ctx* ctx);

/* ---- map: mimosa::lidar::IncrementalVoxelMapPCL over gtsam_points::iVox ------------------------------
 * mb_map_create   <- IncrementalVoxelMapPCL(leaf) + set_lru_horizon + set_neighbor_voxel_mode +
 *                    voxel_insertion_setting().set_min_dist_in_cell,  mimosa/src/lidar/geometric.cpp:23-28
 * mb_map_insert   <- IncrementalVoxelMapPCL::insert, mimosa/src/lidar/incremental_voxel_map.cpp:19-24
 *                    (xyz = first three floats of each record; stride 12 for packed V3F as at
 *                    geometric.cpp:487-495, 32 for lidar::Point)
 * mb_map_snapshot <- the deep copy-constructor, incremental_voxel_map.hpp:34-43, used at geometric.cpp:494
 * mb_map_knn      <- IncrementalVoxelMapPCL::knn_search, incremental_voxel_map.cpp:26-32
 *                    (idx = (voxel_id << 32) | point_id, d2 ascending, ok = found == k)
 * mb_map_points   <- iVox::point(i), geometric_factor.hpp:184
 * mb_map_download <- IncrementalVoxelMapPCL::getCloud, incremental_voxel_map.cpp:34-38 (plus voxel layout)
 * mb_map_upload   <- restore of a downloaded map (no reference analogue; the reference cannot checkpoint) */
MB_API int mb_map_create(mb_ctx* ctx, float leaf, float min_dist, int cap, int nbr_mode, uint64_t lru_horizon,
                         mb_map** out);
MB_API int mb_map_release(mb_map* map);
MB_API int mb_map_insert(mb_map* map, const float* xyz, size_t n, size_t stride_bytes);
MB_API int mb_map_snapshot(mb_map* map, mb_map** out);
MB_API int mb_map_size(mb_map* map, size_t* n_voxels, size_t* n_points, uint64_t* lru_counter);
MB_API int mb_map_knn(mb_map* map, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok);
MB_API int mb_map_points(mb_map* map, const uint64_t* idx, size_t n, double* xyz);
/* coords n_vox*3 i32, counts n_vox i32, lru n_vox u32 (may be NULL), pts n_vox*cap*3 f32 (zero padded). */
MB_API int mb_map_download(mb_map* map, int32_t* coords, int32_t* counts, uint32_t* lru, float* pts);
MB_API int mb_map_upload(mb_map* map, const int32_t* coords, const int32_t* counts, const uint32_t* lru,
                         const float* pts, size_t n_vox, uint64_t lru_counter);
/* Device-resident k-NN over `nq` queries already in HBM (float64 xyz triples uploaded once with
 * mb_map_knn_stage); runs only the search kernel, results stay on the device.  For roofline timing. */
MB_API int mb_map_knn_stage(mb_map* map, const double* q, size_t nq, int k);
MB_API int mb_map_knn_staged_run(mb_map* map);
MB_API int mb_map_knn_staged_run_prefix(mb_map* map, size_t nq); /* only the first nq staged queries (bench: code warm-up) */
MB_API int mb_map_knn_staged_fetch(mb_map* map, uint64_t* idx, double* d2, uint8_t* ok);

/* ---- factor: mimosa::lidar::ICPFactor (unary) -------------------------------------------------------
 * mb_factor_create     <- ICPFactor(key, ivox_target, cloud_source, config) + commonConstructor,
 *                         geometric_factor.hpp:119-156.  The scan is copied (H2D); the map handle is
 *                         retained (shared_ptr semantics) and must not be inserted into while a factor
 *                         references it — take mb_map_snapshot first, as geometric.cpp:494 does.
 *                         [shard_begin, shard_end) selects this rank's block of the scan (0, n for one GPU).
 * mb_factor_linearize  <- ICPFactor::linearize(values), geometric_factor.hpp:231-562, with
 *                         R,t = values.at<Pose3>(X(key)) (row-major R) and gravity_unit =
 *                         values.at<Unit3>(G(0)).unitVector() (read unconditionally at :257)
 * mb_factor_download_state <- getStatuses / getCorresMeansTarget / getCorresNormalsTarget,
 *                         geometric_factor.hpp:48-50 (+ the DA anchor and localizability vectors, :79-106)
 * mb_icp_run           <- the smoother's repeated update(), mimosa/src/graph/manager.cpp:585-588,
 *                         as a device-resident Gauss-Newton loop (see mb_icp_trace) */
MB_API int mb_factor_create(mb_ctx* ctx, mb_map* map, const void* pts, size_t n, size_t stride_bytes,
                            const mb_icp_config* cfg, size_t shard_begin, size_t shard_end, mb_factor** out);
MB_API int mb_factor_release(mb_factor* f);
MB_API int mb_factor_reset(mb_factor* f); /* back to the freshly constructed state */
MB_API int mb_factor_linearize(mb_factor* f, const double R[9], const double t[3], const double gravity_unit[3],
                               mb_linearization* out);
/* Arrays cover this rank's shard only (shard_end - shard_begin points); any pointer may be NULL. */
MB_API int mb_factor_download_state(mb_factor* f, uint8_t* status, double* p_da, double* mean, double* normal,
                                    double* loc_rot, double* loc_trans, uint64_t* knn_idx);
/* R,t updated in place; trace may be NULL or hold `iters` entries. */
MB_API int mb_icp_run(mb_factor* f, double R[9], double t[3], int iters, double lambda, mb_icp_trace* trace);
/* Host-side Gauss-Newton step of the harness (what mb_icp_run does on the device between two linearisations,
 * for callers that drive mb_factor_linearize themselves):  delta = (H + lambda I)^-1 g,  T <- T * Expmap(delta).
 * Returns MB_OK and leaves T unchanged (delta = 0, *solve_ok = 0) when a pivot is not positive. */
MB_API int mb_gn_step(const double H[36], const double g[6], double lambda, double R[9], double t[3], double delta[6],
                      int* solve_ok);
/* Debug/ablation switches: bit0 = disable the data-association cache ("forced" search every call);
 * bit1 = replay mb_icp_run as one captured CUDA graph (only the NCCL all-reduce mode launches several kernels per
 * iteration; otherwise the whole loop is one kernel and the bit has no effect). */
MB_API int mb_factor_set_flags(mb_factor* f, uint32_t flags);

/* ---- scan preparation and map update: the steps either side of the factor in the LiDAR callback ---------
 * A device-resident scan (records of `stride_bytes`, xyz = first three floats; 32 for lidar::Point) lets them
 * run back to back without host round-trips.  All arithmetic is FLOAT, like the reference.
 * mb_downsample       <- Geometric::downsample, mimosa/src/lidar/geometric.cpp:55-126 (greedy per-voxel thinning;
 *                        output = kept input indices in voxel-creation order then in-voxel order); host in/out
 * mb_scan_upload      <- the PCL cloud handed to lidar::Manager::prepareInput / deskewPoints
 * mb_scan_deskew      <- lidar::Manager::deskewPoints, per-point part, mimosa/src/lidar/manager.cpp:494-509:
 *                        p <- R_Le_Lt p + t_Le_Lt with the pose of the point's unique timestamp; `poses` holds
 *                        n_poses x 12 floats (R row-major, then t), `pose_index[i]` selects the pose of point i.
 *                        The IMU propagation that produces the poses (:459-492) stays on the host.
 * mb_scan_transform   <- Geometric::preprocess, p <- R_B_L p + t_B_L, geometric.cpp:153-160
 * mb_scan_downsample  <- Geometric::downsample on the device scan; returns a new scan (kept records, in order)
 * mb_factor_create_from_scan <- ICPFactor(key, map, sm_Be_cloud_ds_, config), geometric.cpp:194
 * mb_map_insert_scan  <- Geometric::updateMap, geometric.cpp:483-495: W = R_W_Be p + t_W_Be in float for the
 *                        scan, then IncrementalVoxelMapPCL::insert; the caller snapshots first (geometric.cpp:494) */
MB_API int mb_downsample(mb_ctx* ctx, const float* xyz, size_t n, size_t stride_bytes, float leaf, size_t cap,
                         float min_dist, uint32_t* out_idx, size_t* n_out);
MB_API int mb_scan_upload(mb_ctx* ctx, const void* pts, size_t n, size_t stride_bytes, mb_scan** out);
/* lidar::Manager::prepareInput, mimosa/src/lidar/manager.cpp:149-383: decode the PointCloud2 payload, apply the
 * NaN / Livox-tag / intensity / range / ns_max filters and the point / ring skip divisors, and group the kept
 * points by timestamp.  points_full <- lidar::Point records (x, y, z + z_offset, intensity, t_ns, original index,
 * range); geometric_idx[0..*n_geometric) indexes points_full (geometric_point_idxs_); unique_ns[0..*n_unique) are
 * the distinct timestamps ascending and pose_index[j] the timestamp group of points_full[j] — what
 * mb_scan_deskew takes once the host has propagated one pose per timestamp.  Host arrays need room for n_points
 * entries. */
MB_API int mb_scan_from_cloud(mb_ctx* ctx, const void* data, size_t n_points, const mb_cloud_layout* layout,
                              const mb_input_filter* filter, mb_scan** points_full, uint32_t* geometric_idx,
                              size_t* n_geometric, uint32_t* pose_index, uint32_t* unique_ns, size_t* n_unique,
                              uint32_t* last_point_ns);
/* Same with the message re-orderings of manager.cpp:179-243 in front (order == NULL: none).  The re-ordered cloud is
 * never materialised: the decode kernel fetches record src(i) of the message for index i of the re-ordered cloud, and
 * `idx` of the output points is i, as in the reference. */
MB_API int mb_scan_from_cloud_ordered(mb_ctx* ctx, const void* data, size_t n_points, const mb_cloud_layout* layout,
                                      const mb_input_filter* filter, const mb_cloud_order* order, mb_scan** points_full,
                                      uint32_t* geometric_idx, size_t* n_geometric, uint32_t* pose_index,
                                      uint32_t* unique_ns, size_t* n_unique, uint32_t* last_point_ns);
MB_API int mb_scan_release(mb_scan* scan);
MB_API int mb_scan_size(mb_scan* scan, size_t* n, size_t* stride_bytes);
MB_API int mb_scan_download(mb_scan* scan, void* pts);
/* Records idx[0..n) of `scan`, in that order, as a new scan: Geometric::preprocess works on
 * points_deskewed[geometric_point_idxs] (mimosa/src/lidar/geometric.cpp:151-158). */
MB_API int mb_scan_gather(mb_scan* scan, const uint32_t* idx, size_t n, mb_scan** out);
MB_API int mb_scan_deskew(mb_scan* scan, const uint32_t* pose_index, const float* poses, size_t n_poses);
MB_API int mb_scan_transform(mb_scan* scan, const float R[9], const float t[3]);
MB_API int mb_scan_downsample(mb_scan* scan, float leaf, size_t cap, float min_dist, mb_scan** out);
MB_API int mb_factor_create_from_scan(mb_ctx* ctx, mb_map* map, mb_scan* scan, const mb_icp_config* cfg,
                                      size_t shard_begin, size_t shard_end, mb_factor** out);
MB_API int mb_map_insert_scan(mb_map* map, mb_scan* scan, const float R[9], const float t[3]);

#ifdef __cplusplus
}
#endif
#endif /* MIMOSA_B200_H_ */
