/* ngicp_b200.h — C ABI of the B200-native nano_gicp scan-to-map hot path.
 *
 * This is the drop-in boundary: the reference has no FFI for this path, its boundary is the C++
 * class nano_gicp::NanoGICP<PointSource,PointTarget> (reference src/dlio/include/nano_gicp/nano_gicp.h:63-150)
 * statically linked into DLIO's odom node. include/nano_gicp/nano_gicp.h in this repo keeps that
 * class surface and forwards every call to the entry points below; INTEGRATION.md shows the wiring.
 * Each entry point names the reference member it replaces (paths relative to reference src/dlio/).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; every function returns an int status
 *    (NGICP_OK = 0). No exceptions, no abort(), nothing printed. ngicp_last_error() gives text.
 *  - host point clouds: any float AoS with xyz at floats 0..2 of every record, `stride_bytes` apart
 *    (the reference's dlio::Point is 32 bytes: include/dlio/dlio.h:85-108).
 *  - 4x4 matrices are COLUMN-MAJOR (Eigen's default storage): element (r,c) at [4*c+r].
 *    6x6 H is symmetric, written in full.
 *  - host covariance lists are the reference's CovarianceList layout: n x Matrix4d, 16 doubles each,
 *    column-major, only the upper-left 3x3 block non-zero (nano_gicp.h:59).
 *  - one CUDA stream per handle; handles are independent and may be driven from different host
 *    threads concurrently (DLIO does: gicp on the lidar thread, gicp_temp on the submap thread,
 *    src/dlio/odom.cc:798-801,1005,1737). A single handle is not thread-safe (neither is the reference).
 *  - indices ("trees") are reference-counted so one handle can build a submap index and another
 *    adopt it (odom.cc:1737-1738 -> :995).
 *  - every compute entry point runs hand-written sm_100a CUDA kernels; there is no CPU fallback.
 */
#ifndef NGICP_B200_H_
#define NGICP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngicp_handle ngicp_handle; /* one NanoGICP object                          */
typedef struct ngicp_index ngicp_index;   /* one nanoflann::KdTreeFLANN<PointT> (cloud+index) */

enum {
  NGICP_OK = 0,
  NGICP_ERR_INVALID = 1,   /* bad argument / missing cloud, index or covariances        */
  NGICP_ERR_CUDA = 2,      /* CUDA runtime error (text in ngicp_last_error)              */
  NGICP_ERR_NO_DEVICE = 3, /* no usable sm_100 device: the library never falls back to CPU */
  NGICP_ERR_UNSUPPORTED = 4,
  NGICP_ERR_LM_NOT_CONVERGED = 5 /* LM inner loop exhausted ("lm not converged!!", lsq_registration.cc:124-127);
                                    outputs are still written, converged = 0              */
};

/* nano_gicp.h:61 — same order as the reference enum */
enum { NGICP_REG_NONE = 0, NGICP_REG_MIN_EIG = 1, NGICP_REG_NORMALIZED_MIN_EIG = 2, NGICP_REG_PLANE = 3, NGICP_REG_FROBENIUS = 4 };

enum { NGICP_SOURCE = 0, NGICP_TARGET = 1 };

/* Configuration = the reference's setters. Defaults (ngicp_default_params) are the reference's
 * constructor defaults: nano_gicp.cc:53-66, lsq_registration.cc:53-67. */
typedef struct ngicp_params {
  int k_correspondences;        /* setCorrespondenceRandomness   nano_gicp.cc:82-84   (default 20)      */
  double max_corr_dist;         /* setMaxCorrespondenceDistance  nano_gicp.cc:87-89   (default FLT_MAX) */
  int regularization;           /* setRegularizationMethod       nano_gicp.cc:92-94   (default PLANE)   */
  int max_iterations;           /* setMaximumIterations          lsq_registration.cc:83-85  (64)        */
  double rotation_epsilon;      /* setRotationEpsilon            lsq_registration.cc:73-75  (2e-3)      */
  double transformation_epsilon;/* setTransformationEpsilon      lsq_registration.cc:78-80  (5e-4)      */
  double lm_init_lambda_factor; /* setInitialLambdaFactor        lsq_registration.cc:88-90  (1e-9)      */
  int lm_max_iterations;        /* lm_max_iterations_            lsq_registration.cc:62     (10)        */
  int use_gauss_newton;         /* lsq_optimizer_type_           lsq_registration.cc:60     (0 = LM)    */
} ngicp_params;

void ngicp_default_params(ngicp_params* p);
const char* ngicp_version(void);
/* text of the last error on this handle (or, with h == NULL, of the calling thread) */
const char* ngicp_last_error(const ngicp_handle* h);

/* NanoGICP::NanoGICP / ~NanoGICP (nano_gicp.cc:53-69). device = CUDA ordinal. */
int ngicp_create(int device, ngicp_handle** out);
int ngicp_destroy(ngicp_handle* h);
int ngicp_set_params(ngicp_handle* h, const ngicp_params* p);
int ngicp_get_params(const ngicp_handle* h, ngicp_params* p);
/* Host clouds and their lifetime. Every entry point that takes a host cloud (ngicp_set_input, ngicp_set_input_batch,
 * ngicp_index_build, ngicp_knn, ngicp_batch_covariances, ngicp_filter_scan, ngicp_scan_ingest) is done with the caller's
 * buffer when it returns: pageable memory is packed into the handle's own staging buffer, page-locked memory
 * (cudaHostAlloc / cudaHostRegister; stride <= 32 bytes) is copied straight out of the caller's buffer and the call
 * waits for that copy (not for the kernels behind it). ngicp_set_async_input(h, 1) drops that wait for page-locked clouds:
 * ngicp_set_input / ngicp_filter_scan then return with the copy in flight and the buffer must stay unchanged until the
 * next call on this handle that synchronises (ngicp_compute_covariances with a density pointer, ngicp_align,
 * ngicp_synchronize) — the reference keeps a pointer to the caller's cloud for just as long (nano_gicp.cc:141). */
int ngicp_set_async_input(ngicp_handle* h, int on);
/* the handle's cudaStream_t (as void*), for callers that want to order their own work after it */
void* ngicp_stream(ngicp_handle* h);
int ngicp_synchronize(ngicp_handle* h);

/* ---- index = nanoflann::KdTreeFLANN<PointT> (nanoflann_adaptor.h:57-152) ----------------------
 * ngicp_index_build replaces KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex
 * (nanoflann_adaptor.h:132-138, nanoflann.h:1405-1417): uploads the cloud, computes voxel keys,
 * Morton-sorts them with a radix sort and builds the multi-level voxel hash. The new index holds
 * one reference owned by the caller. `h` supplies device, stream and scratch memory. */
int ngicp_index_build(ngicp_handle* h, const void* points, size_t n, size_t stride_bytes, ngicp_index** out);
/* same, cloud already in device memory as packed float4 (x,y,z,unused) */
int ngicp_index_build_device(ngicp_handle* h, const void* d_points_f4, size_t n, ngicp_index** out);
int ngicp_index_retain(ngicp_index* idx);
int ngicp_index_release(ngicp_index* idx);
size_t ngicp_index_size(const ngicp_index* idx);
/* KdTreeFLANN::nearestKSearch (nanoflann_adaptor.h:141-152), batched over nq queries. Exact k-NN,
 * fp32 distance ((dx*dx)+(dy*dy))+(dz*dz); every row ordered by (distance, index) ascending — the
 * documented tie-break of this build (the reference keeps KD-visit order on ties, nanoflann.h:207-240).
 * out_idx / out_sqd: nq x k, original point indices; rows with fewer than k hits padded with -1 / +inf. */
int ngicp_knn(ngicp_handle* h, const ngicp_index* idx, const void* queries, size_t nq, size_t stride_bytes,
              int k, int32_t* out_idx, float* out_sqd);
/* Inspection: the neighbour sets the covariance stage works on — the k_indices of nearestKSearch(cloud[i], k) inside
 * calculate_covariances (nano_gicp.cc:343) — produced by the production search (K2, leaf-scheduled). out_idx: n x k
 * original indices, row i = point i: the point itself first, then its other k-1 nearest neighbours in unspecified
 * order (the SET is exact under the (distance, index) tie-break). out_density_terms (optional, n doubles):
 * sum_{j>=1} d2_j / ((k-1)(k+2)/2) per point, the summand of nano_gicp.cc:345-347,389. */
int ngicp_self_neighbours(ngicp_handle* h, int which, int k, int32_t* out_idx, double* out_density_terms);
/* the 64-bit voxel key of every point in ORIGINAL order (spec in DESIGN.md §keys; bit-exact vs oracle) */
int ngicp_index_keys(ngicp_handle* h, const ngicp_index* idx, uint64_t* out_keys, float origin_h0[4]);

/* ---- cloud / tree / covariance bookkeeping (nano_gicp.cc:97-171) ------------------------------ */
/* setInputSource / setInputTarget (nano_gicp.cc:135-161): build an index over the cloud, attach it,
 * DROP existing covariances of that side. The pointer-identity early-out of the reference
 * (:136,:151) lives in the C++ wrapper, which owns the shared_ptr. */
int ngicp_set_input(ngicp_handle* h, int which, const void* points, size_t n, size_t stride_bytes);
/* same as ngicp_set_input for a cloud that already lives in device memory as packed float4
 * (x,y,z,unused); stream-ordered on the handle's stream, no host copy. */
int ngicp_set_input_device(ngicp_handle* h, int which, const void* d_points_f4, size_t n);
/* attach an existing index (the `target_kdtree_ = submap_kdtree` assignment, odom.cc:995, and
 * registerInputSource/Target, nano_gicp.cc:119-132). Keeps covariances. Takes its own reference. */
int ngicp_attach_index(ngicp_handle* h, int which, ngicp_index* idx);
/* borrow the attached index (no reference taken) — `gicp_temp.target_kdtree_` read at odom.cc:1738 */
ngicp_index* ngicp_get_index(ngicp_handle* h, int which);
int ngicp_swap_source_and_target(ngicp_handle* h); /* nano_gicp.cc:97-104 */
int ngicp_clear(ngicp_handle* h, int which);       /* clearSource / clearTarget, nano_gicp.cc:107-116 */

/* calculateSourceCovariances / calculateTargetCovariances (nano_gicp.cc:174-191,330-392):
 * k-NN (k = k_correspondences) + 3x3 covariance (/k) + regularisation; *density = source_density_. */
int ngicp_compute_covariances(ngicp_handle* h, int which, float* density);
/* getSourceCovariances / getTargetCovariances (nano_gicp.h:106-112): n x 16 doubles, original order */
int ngicp_get_covariances(ngicp_handle* h, int which, double* out_4x4, size_t n);
/* setSourceCovariances / setTargetCovariances (nano_gicp.cc:164-171). The device keeps the upper-left 3x3 block as six
 * fp32 values (xx, xy, xz, yy, yz, zz): covariances handed in as fp64 are rounded to fp32 (the reference keeps fp64;
 * its own covariances are fp64 images of fp32 point differences, for which the tests measure < 1e-7 absolute). */
int ngicp_set_covariances(ngicp_handle* h, int which, const double* in_4x4, size_t n);
int ngicp_has_covariances(const ngicp_handle* h, int which, size_t* n);

/* ---- registration (nano_gicp.cc:194-326, lsq_registration.cc:108-229) -------------------------- */
/* update_correspondences (nano_gicp.cc:206-245). Optional outputs, original source order:
 * corr[n_src] (target index or -1), sqd[n_src] (valid where corr >= 0), mahal n_src x 16 doubles. */
int ngicp_update_correspondences(ngicp_handle* h, const double T[16], int32_t* corr, float* sqd, double* mahal, int* num_correspondences);
/* linearize (nano_gicp.cc:248-302): re-associates, then H (6x6), b (6), error. H/b may be NULL. */
int ngicp_linearize(ngicp_handle* h, const double T[16], double H[36], double b[6], double* error, int* num_correspondences);
/* compute_error (nano_gicp.cc:305-326): cached correspondences / Mahalanobis of the last linearize */
int ngicp_compute_error(ngicp_handle* h, const double T[16], double* error);
/* pcl::Registration::align -> NanoGICP::computeTransformation -> LsqRegistration::computeTransformation
 * (nano_gicp.cc:194-203, lsq_registration.cc:108-134). guess may be NULL (= identity, what DLIO passes,
 * odom.cc:1005). Outputs: final_transformation_ (float 4x4), nr_iterations_, converged_,
 * getFinalHessian(), getFinalError(). The 6x6 LM solve runs on the host. */
int ngicp_align(ngicp_handle* h, const float guess[16], float T_out[16], int* nr_iterations, int* converged,
                double H_final[36], double* final_error);
/* pcl::transformPointCloud(*input_, output, final_transformation_) (lsq_registration.cc:133): fp32
 * R*p+t of the source xyz into `out` (same stride; other fields untouched). DLIO discards it. */
int ngicp_transform_source(ngicp_handle* h, const float T[16], void* out_points, size_t n, size_t stride_bytes);

/* ---- batched / sharded work units (BASELINE configs 3 and 5; independent units, no collective) --
 * Bulk covariance build over many keyframes in ONE pass: `points` holds all keyframes back to back,
 * seg_offsets[n_seg+1] are the keyframe boundaries. Every keyframe gets its own index and its
 * covariances come from neighbours inside that keyframe only (the reference computes covariances
 * per scan and concatenates them per submap: nano_gicp.cc:174-181, odom.cc:1719-1729).
 * out_4x4 (n x 16 doubles) and/or out_cov6 (n x 6 floats: xx,xy,xz,yy,yz,zz) may be NULL. */
int ngicp_batch_covariances(ngicp_handle* h, const void* points, size_t n, size_t stride_bytes,
                            const int64_t* seg_offsets, int n_seg, double* out_4x4, float* out_cov6, float* seg_density);

/* Batched registration units (BASELINE config 5 shape): ngicp_set_input_batch makes `which` a cloud of n_seg scans
 * stored back to back (seg_offsets[n_seg+1]); covariances are then computed per scan by ngicp_compute_covariances.
 * ngicp_batch_linearize runs correspondence search + fused linearisation for ALL scans in two launches, scan s at pose
 * T16s[16*s..] against the (single-cloud) target; outputs are per scan (H36s n_seg x 36, b6s n_seg x 6, errs, ncorrs). */
int ngicp_set_input_batch(ngicp_handle* h, int which, const void* points, size_t n, size_t stride_bytes,
                          const int64_t* seg_offsets, int n_seg);
int ngicp_batch_linearize(ngicp_handle* h, int n_scans, const double* T16s, double* H36s, double* b6s, double* errs, int* ncorrs);

/* ---- device-resident keyframe store + submap assembly (additive, not in the reference) -------
 * Replaces the host round trip DLIO makes for every submap: keyframe cloud + covariance list kept on the host
 * (odom.cc:1592), transformed there (pcl::transformPointCloud + cov <- Td cov Td^T, odom.cc:1757-1762), concatenated
 * (odom.cc:1719-1729) and handed back through setInputTarget / setTargetCovariances (odom.cc:1737, :998).
 *   capture   : snapshot the handle's SOURCE cloud and covariances (already in HBM) as a keyframe, original point order
 *   transform : in place, points fp32 R p + t, covariances Td C Td^T with Td = T.cast<double>()
 *   assemble  : concatenate keyframes in the given order into the handle's TARGET cloud, build its index and attach
 *               the concatenated covariances — no PCIe traffic. */
typedef struct ngicp_keyframe ngicp_keyframe;
int ngicp_keyframe_capture(ngicp_handle* h, ngicp_keyframe** out);
int ngicp_keyframe_transform(ngicp_handle* h, ngicp_keyframe* kf, const float T[16]);
size_t ngicp_keyframe_size(const ngicp_keyframe* kf);
int ngicp_keyframe_release(ngicp_handle* h, ngicp_keyframe* kf);
/* copy a keyframe to the host for inspection: xyz n x 3 floats and/or n x 16 doubles (either may be NULL) */
int ngicp_keyframe_download(ngicp_handle* h, const ngicp_keyframe* kf, float* xyz, double* cov_4x4);
int ngicp_submap_assemble(ngicp_handle* h, ngicp_keyframe* const* kfs, int n_kfs);

/* ---- scan pre-filters on the device --------------------------------------------------------
 * The two PCL filters DLIO runs right before setInputSource, in DLIO's order: pcl::CropBox (reference
 * src/dlio/src/dlio/odom.cc:114-116 configure, :500-502 apply) then pcl::VoxelGrid (odom.cc:118, :575-584), xyz only.
 *   crop_min / crop_max  box corners, both NULL = no crop; crop_negative != 0 keeps the points OUTSIDE the box (DLIO)
 *   leaf                 voxel size per axis, NULL = no voxel grid. Output = fp32 centroid of every occupied voxel in
 *                        ascending PCL voxel-index order (sum in ascending input order / count)
 *   set_as               0 / 1: the filtered cloud also becomes the handle's SOURCE / TARGET cloud without leaving the
 *                        device (replaces setInputSource(filtered cloud)); -1: filter only
 *   out_xyz              optional host buffer, capacity n x 3 floats; *n_out = points that survived
 * Non-finite points are dropped (pcl::removeNaNFromPointCloud, odom.cc:496-498). */
int ngicp_filter_scan(ngicp_handle* h, const void* points, size_t n, size_t stride_bytes, const float crop_min[3], const float crop_max[3],
                      int crop_negative, const float leaf[3], int set_as, float* out_xyz, size_t* n_out);

/* ---- deskew on the device (reference dlio::OdomNode::deskewPointcloud, odom.cc:588-706) ----------
 * The reference sorts the scan by per-point time stamp (:634-635), lists the unique stamps (:638-650), integrates the
 * IMU to one pose per unique stamp on the host (:671-672, stays on the host), and moves every point by the pose of its
 * stamp (:690-701, pt = (frames[i] * baselink2lidar_T) * pt). Two calls, the scan stays in HBM in between:
 *   ngicp_scan_ingest  removes non-finite points (:496-498), applies the optional CropBox (:500-502), sorts by time
 *                      stamp (stable: ties keep input order; std::partial_sort_copy leaves them unspecified) and
 *                      returns the unique stamps, ascending, as raw field values converted to double (the caller adds
 *                      its sweep reference time as in :612/:622/:632). time_type 0 = uint32 (Ouster `t`), 1 = float
 *                      (Velodyne `time`), 2 = double (Hesai `timestamp`), at byte offset time_offset_bytes of a record.
 *   ngicp_scan_deskew  frames16 = n_unique column-major fp32 4x4 matrices (already multiplied by the extrinsic), or ONE
 *                      matrix for the no-IMU paths (:659, :681: pcl::transformPointCloud of the whole scan);
 *                      then the optional VoxelGrid (:575-584) and, with set_as 0/1, setInputSource/Target. Output
 *                      (optional, capacity n_kept x 3 floats) is in time order (no voxel grid) or voxel order. */
int ngicp_scan_ingest(ngicp_handle* h, const void* points, size_t n, size_t stride_bytes, size_t time_offset_bytes, int time_type,
                      const float crop_min[3], const float crop_max[3], int crop_negative, double* unique_stamps, size_t* n_unique, size_t* n_kept);
int ngicp_scan_deskew(ngicp_handle* h, const float* frames16, size_t n_frames, const float leaf[3], int set_as, float* out_xyz, size_t* n_out);

/* ---- host loop of one odometry sequence (BASELINE configs 4 and 5) --------------------------------
 * The GICP-relevant part of dlio::OdomNode's per-scan callback, in C++ over the calls above: callbackPointCloud
 * (odom.cc:737-837), computeSpaciousness / computeDensity / setAdaptiveParams (:1398-1436, :1600-1626), getNextPose
 * (:984-1018), propagateGICP (:1230-1246), updateKeyframes (:1517-1598), pushSubmapIndices / buildSubmap /
 * buildKeyframesAndSubmap (:1628-1780), computeConvexHull / computeConcaveHull (:1438-1515). IMU integration and the
 * geometric observer stay with the caller: one scan is TWO calls so that the caller can integrate its IMU buffer at the
 * unique time stamps in between (odom.cc:671-672).
 *   scan_begin   removeNaN + CropBox (negative, +-crop_size) + time sort on the device; returns the unique stamps
 *                (capacity n doubles) and keeps the planar ranges of the cropped scan for computeSpaciousness
 *   scan_finish  frames16 = n_unique column-major fp32 world poses of the sensor (or one, or NULL = "no IMU": the current
 *                pose, odom.cc:656-664); deskew + VoxelGrid + setInputSource, metrics + adaptive parameters, source
 *                covariances, align against the current submap, pose propagation, keyframe decision, keyframe capture /
 *                transform and submap re-assembly, all without the scan or a covariance leaving HBM.
 * res->valid == 0: the scan was dropped (empty after the crop, or <= gicp_min_num_points points, odom.cc:764-767).
 * Planar keyframe sets (pcl::ConvexHull's 2-D case) and spatial ones are handled inside (2-D / 3-D hull and alpha shape);
 * a caller that wants PCL's own hulls for spatial sets installs the two hull callbacks (indices of the points on the hull
 * of n xyz doubles; return the count or < 0). A degenerate spatial set without callbacks fails with NGICP_ERR_UNSUPPORTED. One ngicp_odom owns its handle's source, target and keyframes for its lifetime. */
typedef struct ngicp_odom ngicp_odom;
typedef struct ngicp_odom_params {
  float crop_size;             /* preprocessing/cropBoxFilter/size   (cfg/params.yaml:43) */
  float voxel_res;             /* preprocessing/voxelFilter/res      (:45); <= 0: no voxel grid */
  float keyframe_thresh_dist;  /* keyframe/threshD                   (:48) */
  float keyframe_thresh_rot;   /* keyframe/threshR, degrees          (:49) */
  int submap_knn, submap_kcv, submap_kcc; /* submap/keyframe/{knn,kcv,kcc} (:53-55) */
  int gicp_min_num_points;     /* gicp/minNumPoints                  (:57) */
  float gicp_max_corr_dist;    /* gicp/maxCorrespondenceDistance     (:59) */
  int adaptive;                /* adaptive                           (cfg/dlio.yaml:17) */
  int time_offset_bytes;       /* byte offset and type (as ngicp_scan_ingest) of the per-point time stamp */
  int time_type;
} ngicp_odom_params;
typedef struct ngicp_odom_result {
  int valid;
  float T[16];        /* lidar pose after the scan, column-major */
  float T_corr[16];   /* the alignment's correction, column-major */
  int converged, iterations, n_points;
  int new_keyframe, submap_changed; /* submap_changed: the submap this scan was aligned against differs from the previous scan's */
  int n_keyframes, n_submap;
} ngicp_odom_result;
typedef int (*ngicp_hull_fn)(const double* xyz, int n, double alpha, int* out_indices, void* user);
void ngicp_odom_default_params(ngicp_odom_params* p);
int ngicp_odom_create(ngicp_handle* h, const ngicp_odom_params* p, ngicp_odom** out);
int ngicp_odom_destroy(ngicp_odom* o);
const char* ngicp_odom_last_error(const ngicp_odom* o);
int ngicp_odom_set_hull_callbacks(ngicp_odom* o, ngicp_hull_fn convex, ngicp_hull_fn concave, void* user);
int ngicp_odom_set_pose(ngicp_odom* o, const float T_colmajor[16]);
int ngicp_odom_scan_begin(ngicp_odom* o, const void* records, size_t n, size_t stride_bytes, double* unique_stamps, size_t* n_unique, size_t* n_kept);
/* host wall clock per stage of the loop, seconds summed over `scans` valid scans: 0 ingest (crop + time sort, waits for the
 * stamps), 1 planar ranges, 2 deskew + VoxelGrid + index build (waits for the point count), 3 median + adaptive parameters,
 * 4 source covariances (waits for the density), 5 align, 6 keyframe decision + capture, 7 keyframe transform + hulls + submap */
#define NGICP_ODOM_STAGES 8
int ngicp_odom_get_profile(ngicp_odom* o, double seconds[NGICP_ODOM_STAGES], long* scans, int reset);
/* the planar hull used inside (exported for tests): indices (ascending) of the points on the convex hull (concave == 0)
 * or on the alpha shape of n xyz doubles; returns the count, -3 when the set is spatial (callback case), -1 on bad input */
int ngicp_hull_planar(const double* xyz, int n, int concave, double alpha, int* out_indices);
/* the spatial counterpart, used when no hull callbacks are installed: extreme points of the 3-D convex hull (incremental
 * insertion) or vertices of the 3-D alpha shape (Delaunay tetrahedra with circumradius <= alpha, faces of exactly one kept
 * tetrahedron); returns the count, -2 for a degenerate set (fewer than 4 points, all coplanar, numerical trouble) */
int ngicp_hull_spatial(const double* xyz, int n, int concave, double alpha, int* out_indices);
int ngicp_odom_scan_finish(ngicp_odom* o, const float* frames16, size_t n_frames, ngicp_odom_result* res, int* submap_ids, int submap_cap);

/* ---- timing hooks used by bench.py (device time of the last call's stages, milliseconds) ------- */
typedef struct ngicp_timings {
  float index_ms;       /* K1: keys + radix sort + reorder + voxel hash   */
  float knn_ms;         /* K2: exact k-NN                                  */
  float covariance_ms;  /* K3: covariance + regularisation                 */
  float linearize_ms;   /* K4: sum over the align's linearize launches     */
  float error_ms;       /* K5: sum over the align's compute_error launches */
  int linearize_calls;
  int error_calls;
  int kernel_launches;  /* kernels launched by this handle since the last reset */
  float correspond_ms;  /* K4a: correspondence search launches of the align (linearize_ms holds K4b, the fused linearisation) */
} ngicp_timings;
int ngicp_enable_timing(ngicp_handle* h, int on);
int ngicp_get_timings(ngicp_handle* h, ngicp_timings* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* NGICP_B200_H_ */
