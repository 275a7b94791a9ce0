/*
 * pmgpu.h — C ABI of the B200-native ICP hot path (libpmgpu.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * names the libpointmatcher (v1.3.1) interface it replaces; paths are relative to the reference
 * tree.  The C++ module classes in libpointmatcher_b200/host/ (same names, parameters and
 * exceptions as the reference's) and the ctypes binding in libpointmatcher_b200/capi.py are thin
 * marshalling layers over these functions.  INTEGRATION.md shows the binding a libpointmatcher
 * maintainer would add to Registry.cpp.
 *
 * Conventions
 *  - clouds are `rows x n` column-major float matrices exactly as `DataPoints::features`
 *    (pointmatcher/PointMatcher.h:169,331): rows == 4 (x, y, z, w) for 3-D clouds, rows == 3 (x, y, w) for
 *    2-D clouds — every module of the reference branches on it (PointToPlane.cpp:294-310,
 *    ErrorMinimizer.cpp:308-313, TransformationCheckersImpl.cpp:114-131).  A 2-D cloud lives on the device as
 *    (x, y, 0, w), where every z term of the 3-D kernels is an exact zero; transforms of a 2-D context cross
 *    this ABI as 3 x 3 matrices, normals have 2 rows, eigenvalues 2, eigenvectors 4.  2-D supports the
 *    matchers, the distance outlier filters, PointToPoint / PointToPlane, Counter / Differential (with the
 *    reference's 3 x 3 quaternion quirk) and SurfaceNormal; the covariance, similarity and force2D / 4DOF
 *    variants return PMGPU_ERR_UNSUPPORTED.  Only float is implemented; anything else returns
 *    PMGPU_ERR_UNSUPPORTED (there is no CPU fallback by design).
 *  - match results are `k x n` column-major (`Matches::ids/dists`, PointMatcher.h:373-374):
 *    ids int32 (-1 = Matches::InvalidId), dists float SQUARED distances (+inf = InvalidDist).
 *  - transforms are `rows x rows` column-major (Eigen default): float[16], or float[9] for 2-D clouds.
 *  - every pointer argument may be a host pointer (pageable or pinned) or a device pointer
 *    (copies use cudaMemcpyDefault); NULL for an optional output skips that download.
 *  - all functions return a PMGPU_* status; pmgpu_last_error(ctx) gives the message.  The
 *    status -> reference exception mapping is listed beside each code.
 *  - a context owns one CUDA stream and all device buffers; one context per thread / per
 *    concurrent registration (the reference's ICP object is not re-entrant either,
 *    evaluations/eval_solution.cpp:628-630).
 */
#ifndef PMGPU_H
#define PMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmgpu_ctx pmgpu_ctx;

enum pmgpu_status {
    PMGPU_OK = 0,
    PMGPU_ERR_CUDA = 1,                 /* std::runtime_error (CUDA failure) */
    PMGPU_ERR_BAD_ARG = 2,              /* std::runtime_error / InvalidParameter */
    PMGPU_ERR_UNSUPPORTED = 3,          /* ConfigurationError("GPU module: only float/3-D ...") */
    PMGPU_ERR_NO_REFERENCE = 4,         /* std::runtime_error: init() was not called */
    PMGPU_ERR_NO_READING = 5,
    PMGPU_ERR_NO_MATCHES = 6,           /* weights/minimize called before knn */
    PMGPU_ERR_NO_OUTLIER_TO_FILTER = 7, /* ConvergenceError("no outlier to filter")       Matches.cpp:76-77 */
    PMGPU_ERR_BAD_QUANTILE = 8,         /* ConvergenceError("quantile must be ...")       Matches.cpp:79-80 */
    PMGPU_ERR_NO_POINT_TO_MINIMIZE = 9, /* ConvergenceError("ErrorMnimizer: no point ...") ErrorMinimizer.cpp:76-77 */
    PMGPU_ERR_NO_NORMALS = 10,          /* InvalidField("Field normals not found")        DataPoints.cpp:941 */
    PMGPU_ERR_NOT_ORTHOGONAL = 11,      /* TransformationError                            TransformationsImpl.cpp:62-63 */
    PMGPU_ERR_KNN_TOO_LARGE = 12,       /* libnabo: "knn larger than the number of points" */
    PMGPU_ERR_NAN = 13,                 /* ConvergenceError("abs rotation norm not a number") TransformationCheckersImpl.cpp:154-157 */
    PMGPU_ERR_COMM = 14                 /* NCCL failure, or the peers of a sharded registration did not arrive */
};

/* OutlierFiltersImpl.h: the distance filters of the hot path and the two row-8f-3 filters */
enum pmgpu_filter_type {
    PMGPU_FILTER_MAXDIST = 0,    /* MaxDistOutlierFilter     param = maxDist (un-squared)  OutlierFiltersImpl.cpp:66-81   */
    PMGPU_FILTER_MEDIANDIST = 1, /* MedianDistOutlierFilter  param = factor                OutlierFiltersImpl.cpp:109-125 */
    PMGPU_FILTER_TRIMMEDDIST = 2,/* TrimmedDistOutlierFilter param = ratio                 OutlierFiltersImpl.cpp:132-147 */
    PMGPU_FILTER_ROBUST = 3,     /* RobustOutlierFilter      param = tuning                OutlierFiltersImpl.cpp:420-598 */
    PMGPU_FILTER_VARTRIMMEDDIST = 5, /* VarTrimmedDistOutlierFilter param = lambda; minRatio / maxRatio through
                                     pmgpu_set_var_trimmed_ratios       OutlierFiltersImpl.cpp:152-218 */
    PMGPU_FILTER_SURFACENORMAL = 4,/* SurfaceNormalOutlierFilter param = maxAngle: weight 0 where |n_reading . n_reference| <
                                      cos(maxAngle), both normalised (OutlierFiltersImpl.cpp:222-285); needs the reference
                                      normals and pmgpu_reading_set_normals, otherwise all ones like the reference */
    PMGPU_FILTER_MINDIST = 6     /* MinDistOutlierFilter     param = minDist (un-squared): weight 0 below it
                                                                                            OutlierFiltersImpl.cpp:87-101.
                                    A match without a neighbour (dist = +inf) always reads weight 0 here; the reference gives it
                                    1 under a chain of MinDist / Null filters only and drops it in ErrorElements either way
                                    (ErrorMinimizer.cpp:103-106), so nothing computed from the weights differs */
};
/* RobustOutlierFilter: M-estimator weights w(e^2), e^2 = dist / scale^2.  Its discrete
 * parameters travel in the filter word: bits 0-7 PMGPU_FILTER_ROBUST | bits 8-15 robustFct | bits 16-19
 * scaleEstimator | bits 20-27 nbIterationForScale.  scaleEstimator "mad" = sqrt(median |d - median d|)
 * (Matches.cpp:88-122) is two more exact radix selects on the device; "berg" (OutlierFiltersImpl.cpp:420-432, 523-537:
 * 1.9 sqrt(median d) at the first iteration, then 0.85 (scale - target) + target with target = the tuning given and the
 * tuning constant of Bergstrom for cauchy / tukey / huber) one select at the first iteration; "std" =
 * sqrt(Matches::getStandardDeviation()) (Matches.cpp:124-129) two fp64 reductions; "none" = 1.
 * `approximation` (weight 0 where e^2 >= approximation^2) through pmgpu_set_robust_approximation.  distanceType
 * point2plane = PMGPU_ROBUST_P2PLANE or-ed into the word: the weight function then sees dot(n / |n|, p - q)^2 instead of the
 * match distance (computePointToPlaneDistance, OutlierFiltersImpl.cpp:468-500; needs the reference normals, 3-D clouds; the
 * scale estimators keep reading the match distances).  At most one robust filter per chain, not with a sharded reading.
 * limits_out[f] of pmgpu_weights returns the scale. */
enum {
    PMGPU_ROBUST_CAUCHY = 0, PMGPU_ROBUST_WELSCH, PMGPU_ROBUST_SC, PMGPU_ROBUST_GM, PMGPU_ROBUST_TUKEY, PMGPU_ROBUST_HUBER, PMGPU_ROBUST_L1,
    PMGPU_ROBUST_STUDENT
};
enum { PMGPU_SCALE_NONE = 0, PMGPU_SCALE_MAD = 1, PMGPU_SCALE_BERG = 2, PMGPU_SCALE_STD = 3 };
#define PMGPU_ROBUST_P2PLANE (1 << 28)
#define PMGPU_ROBUST_WORD(fct, scale, nb_iter) (PMGPU_FILTER_ROBUST | ((fct) << 8) | ((scale) << 16) | ((nb_iter) << 20))


enum pmgpu_minimizer {
    PMGPU_MIN_P2POINT = 0,     /* PointToPointErrorMinimizer          ErrorMinimizers/PointToPoint.cpp:61-101 */
    PMGPU_MIN_P2PLANE = 1,     /* PointToPlaneErrorMinimizer          ErrorMinimizers/PointToPlane.cpp:171-312 */
    PMGPU_MIN_P2POINT_COV = 2, /* PointToPointWithCovErrorMinimizer   ErrorMinimizers/PointToPointWithCov.cpp:49-145 */
    PMGPU_MIN_P2PLANE_COV = 3, /* PointToPlaneWithCovErrorMinimizer   ErrorMinimizers/PointToPlaneWithCov.cpp:60-162 */
    PMGPU_MIN_P2POINT_SIM = 4  /* PointToPointSimilarityErrorMinimizer ErrorMinimizers/PointToPointSimilarity.cpp:49-101:
                                  rotation + translation + one scale; T_iter is then a similarity, applied as it is
                                  (SimilarityTransformation, TransformationsImpl.cpp:156-210) — fused loop and
                                  pmgpu_minimize; pmgpu_knn still insists on a rigid T */
};
/* or-ed into PMGPU_MIN_P2PLANE[_COV]: PointToPlaneErrorMinimizer force4DOF (PointToPlane.cpp:203-214,
 * 266-281) — the unknowns are the rotation about z and the translation: the 4x4 sub-system
 * (cross_z, n) of the same sums, T = AngleAxis(x0, unitZ) + translation. */
#define PMGPU_MIN_FORCE4DOF 0x100
/* or-ed into PMGPU_MIN_P2PLANE: force2D on 3-D clouds (PointToPlane.cpp:177-186, 294-310) — the unknowns are
 * the rotation about z and the x/y translation: the 3x3 sub-system (cross_z, nx, ny) with the residual
 * n . (p - q) taken over x and y only; T = identity with Rotation2D(x0) and (x1, x2) in its xy block.
 * Not combinable with force4DOF (the reference throws, PointToPlane.cpp:59-64) nor, here, with the covariance. */
#define PMGPU_MIN_FORCE2D 0x200

/* SurfaceNormalDataPointsFilter keep* flags (DataPointsFilters/SurfaceNormal.h:65-80) */
enum pmgpu_normals_flags {
    PMGPU_NORMALS_SORT_EIGEN = 1,
    PMGPU_NORMALS_SMOOTH = 2
};

/* ---- context ------------------------------------------------------------------------- */
int pmgpu_ctx_create(int device, pmgpu_ctx** ctx_out);
void pmgpu_ctx_destroy(pmgpu_ctx* ctx);
const char* pmgpu_last_error(const pmgpu_ctx* ctx);
const char* pmgpu_status_string(int status);
/* CUDA stream of the context as an opaque pointer (cudaStream_t), for event timing */
void* pmgpu_ctx_stream(pmgpu_ctx* ctx);
int pmgpu_sync(pmgpu_ctx* ctx);
/* number of kernel launches issued by this context since creation (bench `gpu_launches`) */
uint64_t pmgpu_launch_count(const pmgpu_ctx* ctx);
/* Per-stage device timing with CUDA events recorded on the context's stream around the kernels
 * of each stage (0 = kNN match, 1 = select + weights, 2 = minimise / compose / check,
 * 3 = covariance).  pmgpu_timing_collect synchronises, adds the elapsed milliseconds of all
 * recorded intervals into ms_out[4] and their number into count_out[4], and clears them. */
int pmgpu_timing_enable(pmgpu_ctx* ctx, int on);
int pmgpu_timing_collect(pmgpu_ctx* ctx, double* ms_out, int* count_out);

/* ---- K1: KDTreeMatcher::init (MatchersImpl.cpp:77-83) ----------------------------------
 * Uploads the reference features and builds the search structure over the first rows-1
 * coordinates.  `normals`: optional 3 x n block with column stride `normals_ld` floats
 * (the "normals" descriptor rows inside DataPoints::descriptors, DataPoints.cpp:917-942);
 * needed by the point-to-plane minimizers.  Replaces any previous reference. */
int pmgpu_ref_set(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld);
/* ICP::compute's preamble fused with init (ICP.cpp:291-302): computes the float mean of the first
 * rows-1 coordinates exactly as the reference does on the host (sequential row sums / n), uploads
 * the cloud, subtracts the mean on the device (the same float subtraction) and builds the
 * structure over the centred cloud.  mean_out[rows] receives the mean (last entry unused = 1). */
int pmgpu_ref_set_centered(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld, float* mean_out);
/* (re)attach normals to the current reference without rebuilding the tree */
int pmgpu_ref_set_normals(pmgpu_ctx* ctx, const float* normals, int normals_ld);

/* ---- reading (the `filteredReading` argument of findClosests / compute) ---------------- */
/* the reference's normals as resident (3 x nr, the caller's column order) — those pmgpu_ref_compute_normals made on the
 * device never visit the host otherwise; for ErrorElements / residual-error consumers (PointToPlane.cpp:314-352) */
int pmgpu_ref_get_normals(pmgpu_ctx* ctx, float* normals_out);
int pmgpu_reading_set(pmgpu_ctx* ctx, const float* features, int rows, int n);
/* Sharded registration: uploads THIS rank's share of the whole reading `features` (4 x n, host) — the chunks of `chunk` consecutive
 * columns c with c mod nranks == rank, in order — straight from the caller's matrix (one strided copy, no host-side gather).
 * Equivalent to pmgpu_reading_set on those columns.  3-D clouds. */
int pmgpu_reading_set_sharded(pmgpu_ctx* ctx, const float* features, int rows, int n, int rank, int nranks, int chunk);

/* RigidTransformation::compute on the resident reading, in place (TransformationsImpl.cpp:49-87;
 * the `transformations.apply(reading, T_refMean_dataIn)` of ICP.cpp:345-347).  Returns
 * PMGPU_ERR_NOT_ORTHOGONAL if |1 - det R| > 1e-3. */
/* KDTreeVarDistMatcher (MatchersImpl.cpp:105-150): one maximum search distance per reading point — the reading's `maxDistField`
 * descriptor (1 row, column stride `ld`).  After this call a NEGATIVE max_dist in pmgpu_knn / pmgpu_icp_params means "use the per-point
 * distances" (a match farther than its point's distance is missing: id -1, dist inf).  Capped matching stays off in that mode. */
int pmgpu_reading_set_max_dists(pmgpu_ctx* ctx, const float* max_dists, int ld);
/* the reading's "normals" descriptor (3 rows of a descriptor matrix with column stride `ld`), after pmgpu_reading_set; they turn with
 * the reading (pmgpu_reading_apply_transform, T_iter) as RigidTransformation::compute turns them (TransformationsImpl.cpp:71-84) */
int pmgpu_reading_set_normals(pmgpu_ctx* ctx, const float* normals, int ld);
int pmgpu_reading_apply_transform(pmgpu_ctx* ctx, const float* T);
/* optional download of the resident reading (4 x n) */
int pmgpu_reading_get(pmgpu_ctx* ctx, float* features_out);

/* ---- K2: RigidTransformation::compute + KDTreeMatcher::findClosests ---------------------
 * (TransformationsImpl.cpp:49-87, MatchersImpl.cpp:85-101).  Applies T (NULL = identity) to
 * the resident reading and finds, for every transformed point, the k nearest reference
 * points: exact search, squared float distances accumulated x,y,z each op rounded once,
 * candidates kept iff dist <= maxDist^2, result ascending in (dist, index) — the result
 * libnabo's brute-force search returns.  epsilon > 0 is accepted and answered exactly.
 * Returns PMGPU_ERR_NOT_ORTHOGONAL if |1 - det R| > 1e-3 (TransformationsImpl.cpp:90-105).
 * ids_out / dists_out: optional k x n downloads; visit_out: optional number of reference
 * points examined (Matcher::visitCounter, MatchersImpl.cpp:98). */
int pmgpu_knn(pmgpu_ctx* ctx, const float* T, int k, float epsilon, float max_dist, int32_t* ids_out, float* dists_out, uint64_t* visit_out);

/* ---- K3: OutlierFilters::compute (OutlierFilter.cpp:63-103) -----------------------------
 * Evaluates the filter chain on the resident matches.  nfilters == 0 is the empty chain
 * (weight = dist != inf).  weights_out: optional k x n download; limits_out: optional
 * per-filter squared-distance limits (nfilters floats). */
int pmgpu_weights(pmgpu_ctx* ctx, int nfilters, const int* types, const float* params, float* weights_out, float* limits_out);
/* The matches and weights resident on the device, for what the reference derives from ErrorElements after the fact
 * (ErrorMinimizer::getErrorElements, getResidualError, getOverlap; ErrorMinimizer.cpp:58-193, PointToPlane.cpp:314-470,
 * PointToPoint.cpp:101-165): ids / dists / weights (k x nq, column-major like Matches, any may be null) of the last
 * evaluation — staged calls or the last executed iteration of the fused loop — and T_match (4 x 4 column-major), the
 * transform of the reading those matches were made with.  After a capped fused loop the matches the filters rejected beyond
 * the search radius read id -2 / dist FLT_MAX / weight 0: ErrorElements drops them like any other rejected match. */
int pmgpu_matches_get(pmgpu_ctx* ctx, int32_t* ids_out, float* dists_out, float* weights_out, float* T_match_out);

/* VarTrimmedDistOutlierFilter (OutlierFiltersImpl.h:147-172, OutlierFiltersImpl.cpp:152-218): TrimmedDist with the ratio
 * that minimises FRMS(i) = cumsum(sorted dists)[i] / i / (i / N)^(2 lambda) over minRatio N <= i < maxRatio N.  The
 * distances are sorted on the device, the running sum is taken in float one element after the other exactly like the
 * reference's std::partial_sum (one thread walks the chain, four warps feed it: ~4 ns per match), FRMS and its first minimum in parallel, and the limit is
 * read from the sorted array.  One such filter per chain; it switches capped matching off (every distance counts).
 * pmgpu_set_var_trimmed_ratios sets minRatio / maxRatio (defaults 0.05 / 0.99) of the context's VarTrimmedDist filter;
 * PMGPU_ERR_BAD_ARG unless 0 < min_ratio < max_ratio <= 1.  pmgpu_var_trimmed_ratio: the ratio the last evaluation chose. */
int pmgpu_set_var_trimmed_ratios(pmgpu_ctx* ctx, float min_ratio, float max_ratio);
/* RobustOutlierFilter `approximation` (OutlierFiltersImpl.cpp:400, 591-595; metres, un-squared; +inf = none, the default):
 * of the context's robust filter. */
int pmgpu_set_robust_approximation(pmgpu_ctx* ctx, float approximation);
int pmgpu_var_trimmed_ratio(pmgpu_ctx* ctx, float* ratio_out);

/* ---- K4-K7: ErrorMinimizer::compute (ErrorMinimizer.cpp:217-232) ------------------------
 * Uses the resident reading (transformed by the T of the last pmgpu_knn), matches and
 * weights.  T_out: the incremental 4x4; cov_out: 6x6 (WithCov variants, optional);
 * stats_out[5] (optional) = {pointUsedRatio, weightedPointUsedRatio, nbRejectedMatches,
 * nbRejectedPoints, nbKept} (ErrorMinimizer.cpp:92-140). */
int pmgpu_minimize(pmgpu_ctx* ctx, int minimizer, float sensor_std_dev, float* T_out, float* cov_out, float* stats_out);

/* ---- K8: SurfaceNormalDataPointsFilter::inPlaceFilter (SurfaceNormal.cpp:82-290) --------
 * Self-kNN on `features` (rows x n) + per-point 3x3 eigen-solve.  Outputs are optional and
 * written with column stride `ld` floats each: normals 3 x n, densities 1 x n, eig_values
 * 3 x n, eig_vectors 9 x n (row-major serialisation, utils.h:89-103), matched_ids knn x n
 * (as float, SurfaceNormal.cpp:254-257), mean_dists 1 x n.  If `attach` != 0 and the cloud
 * is the current reference (same n), the normals also become the reference normals.
 * epsilon > 0 is accepted and answered exactly (a valid, different answer from libnabo's approximate search).  Eigenvalues /
 * eigenvectors are always reported in ascending order of the eigenvalue (the reference's `sortEigen` order; without sortEigen
 * it reports them in Eigen::EigenSolver's unspecified order), and the rank test of SurfaceNormal.cpp:190-232 is applied
 * whichever outputs are asked for: a degenerate neighbourhood reads normal 0, density 0. */
typedef struct pmgpu_normals_out {
    float* normals;     int normals_ld;
    float* densities;   int densities_ld;
    float* eig_values;  int eig_values_ld;
    float* eig_vectors; int eig_vectors_ld;
    float* matched_ids; int matched_ids_ld;
    float* mean_dists;  int mean_dists_ld;
} pmgpu_normals_out;
int pmgpu_normals(pmgpu_ctx* ctx, const float* features, int rows, int n, int knn, float epsilon, float max_dist, int flags,
                  const pmgpu_normals_out* out, int* degenerate_out);
/* same, on the resident reference; attaches the normals to it (config 3 pre-step) */
int pmgpu_ref_compute_normals(pmgpu_ctx* ctx, int knn, float epsilon, float max_dist, int flags);
/* Centres the RESIDENT reference on the mean of `features` (the host cloud it was set from) and returns the mean — the second
 * half of pmgpu_ref_set_centered as its own step, so that a chain whose last reference filter is SurfaceNormalDataPointsFilter
 * can run  pmgpu_ref_set -> pmgpu_ref_compute_normals -> pmgpu_ref_center  on one upload and one structure (normals do not
 * change under a translation; they are computed on the caller's coordinates, exactly like the filter's own output). */
int pmgpu_ref_center(pmgpu_ctx* ctx, const float* features, int rows, int n, float* mean_out);

/* ---- fused ICP loop: ICP::computeWithTransformedReference (ICP.cpp:371-430) -------------
 * Runs iterations entirely on the device against the resident reference and reading:
 * T_iter <- dT * T_iter with the Counter / Differential checkers evaluated on the device
 * (TransformationCheckersImpl.cpp:45-158).  The reading is the one given to
 * pmgpu_reading_set (already expressed in the reference frame).  T_iter_out: final T_iter. */
typedef struct pmgpu_icp_params {
    int knn;
    float epsilon;
    float max_dist;
    int nfilters;
    int filter_type[8];
    float filter_param[8];
    int minimizer;
    float sensor_std_dev;
    int max_iterations;       /* CounterTransformationChecker.maxIterationCount */
    int use_differential;     /* DifferentialTransformationChecker on/off */
    float min_diff_rot_err;
    float min_diff_trans_err;
    int smooth_length;
} pmgpu_icp_params;
int pmgpu_icp_run(pmgpu_ctx* ctx, const pmgpu_icp_params* params, const float* T_iter_init, float* T_iter_out, int* iterations_out,
                  float* cov_out, float* stats_out);
/* enqueue exactly `n_iterations` iteration slots without synchronising (bench inner loop).
 *
 * Capped matching.  Inside the fused loop nothing outside can see the matches the outlier
 * filters reject, so with maxDist = inf and a non-empty filter chain the matcher of iteration
 * i+1 stops at a squared radius of 1.5 x the largest distance the filters of iteration i had to
 * know exactly (their order statistics and limits).  The select kernels verify afterwards that
 * the new order statistics and limits lie below the radius used; if they do, T is bit-identical
 * to the uncapped loop.  If not, the slot is void (T_iter and the iteration count stay as they
 * were) and the next slot matches without a cap.  pmgpu_icp_run tops up voided slots itself;
 * a caller of pmgpu_icp_enqueue compares iterations_out with the slots it enqueued.
 * pmgpu_icp_cap_redos: voided slots since pmgpu_icp_reset (valid after pmgpu_icp_result). */
int pmgpu_icp_enqueue(pmgpu_ctx* ctx, const pmgpu_icp_params* params, int n_iterations);
int pmgpu_icp_reset(pmgpu_ctx* ctx, const float* T_iter_init);
int pmgpu_icp_result(pmgpu_ctx* ctx, float* T_iter_out, int* iterations_out, float* cov_out, float* stats_out);
int pmgpu_icp_cap_redos(const pmgpu_ctx* ctx);
/* One iteration of the loop for a host that looks at every iteration (Inspector::dumpIteration, ICP.cpp:403-405; a host-side
 * TransformationChecker such as Bound, ICP.cpp:414-427): exactly one slot, matched WITHOUT the cap — pmgpu_matches_get then shows
 * what the reference's inspector is shown, every match with its weight — then the state as pmgpu_icp_result returns it; with a
 * WithCov minimiser, the covariance of this iteration.  iterations_out unchanged from the previous call = the device checkers
 * (Counter, Differential) had already stopped the loop.  Call pmgpu_icp_reset first. */
int pmgpu_icp_step(pmgpu_ctx* ctx, const pmgpu_icp_params* params, float* T_iter_out, int* iterations_out, float* cov_out, float* stats_out);

/* ---- host-side pre-filters of the default chain (CPU, once per cloud) ----
 * ICPChainBase::setDefault / examples/data/default.yaml put RandomSamplingDataPointsFilter on
 * the reading and SamplingSurfaceNormalDataPointsFilter on the reference (ICP.cpp:100-113).
 * Both draw from std::rand() in point order, and the second one's bins come from a recursive
 * std::nth_element, so they run on the host exactly as in the reference; nothing here touches
 * the device.
 * pmgpu_host_random_sampling (RandomSampling.cpp:58-75): keep_out[0..return) = kept columns.
 * pmgpu_host_sampling_surface_normal (SamplingSurfaceNormal.cpp:80-342): `features` (4 x n) and
 * `descriptors` (desc_rows x n, may be null) are modified in place like the reference's cloud
 * (samplingMethod 1 stores the bin mean in the kept column); outputs are written at the kept
 * columns' ORIGINAL positions (normals 3 x n, densities n, eig_values 3 x n, eig_vectors 9 x n),
 * the caller compacts by keep_out (sorted ascending).  Returns the number of kept points,
 * -1 on a bad argument; unfit_out: points dropped for lack of a normal. */
enum {
    PMGPU_KEEP_NORMALS = 1,
    PMGPU_KEEP_DENSITIES = 2,
    PMGPU_KEEP_EIGEN_VALUES = 4,
    PMGPU_KEEP_EIGEN_VECTORS = 8
};
void pmgpu_host_srand(unsigned seed);
int pmgpu_host_random_sampling(int n, float prob, int32_t* keep_out);
/* The other std::rand() consumers among the reference's DataPointsFilters (same libc stream):
 * pmgpu_host_rand = std::rand() (FixStepSampling.cpp:82 draws its phase from it);
 * pmgpu_host_max_point_count (MaxPointCount.cpp:71-110): srand(seed), order_out[0..return) = the
 *   ORIGINAL column each kept column holds (the reference's swap through an Eigen view copies
 *   column idx onto column j and leaves idx in place, so duplicates are possible); the identity
 *   and n when max_count > n - 1;
 * pmgpu_host_max_density (MaxDensity.cpp:60-105): densities[i * stride]; keep_out = kept columns. */
int pmgpu_host_rand(void);
/* Page-locks / releases a host buffer the caller keeps handing to pmgpu_*_set (cudaHostRegister): uploads from registered
 * memory run at the link's full rate and asynchronously (the reference has no counterpart; Eigen owns its matrices). */
int pmgpu_host_pin(void* ptr, size_t bytes);
int pmgpu_host_unpin(void* ptr);
int pmgpu_host_max_point_count(int n, uint64_t seed, uint64_t max_count, int32_t* order_out);
int pmgpu_host_max_density(const float* densities, int stride, int n, float max_density, int32_t* keep_out);
int pmgpu_host_sampling_surface_normal(float* features, int rows, int n, float* descriptors, int desc_rows, float ratio, int knn,
                                       int sampling_method, float max_box_dim, int average_descriptors, int flags, int32_t* keep_out,
                                       float* normals_out, float* densities_out, float* eig_values_out, float* eig_vectors_out,
                                       int* unfit_out);

/* ---- multi-GPU: queries sharded over ranks, replicated reference ------------------------
 * After pmgpu_comm_init the quantile histograms and the normal-equation sums of
 * pmgpu_weights / pmgpu_minimize / pmgpu_icp_* are all-reduced over the communicator, so
 * every rank computes the same limit and the same T.  `unique_id` is the 128-byte
 * ncclUniqueId produced by pmgpu_comm_unique_id on rank 0 and distributed by the caller. */
int pmgpu_comm_unique_id(void* unique_id_128);
int pmgpu_comm_init(pmgpu_ctx* ctx, const void* unique_id_128, int rank, int nranks);
/* Peer mailboxes: the per-iteration exchanges without any library collective.  Each rank's context owns a ~1 MB mailbox in
 * its GPU's memory; pmgpu_comm_peer_handle allocates it and writes a 128-byte handle, the caller distributes the handles
 * (rank order, nranks x 128 bytes) and pmgpu_comm_peer_init maps every peer's mailbox (cudaIpcOpenMemHandle across
 * processes; direct pointers + peer access between contexts of one process).  From then on the last block of the
 * histogram / accumulate / covariance kernels stores its rank's histograms or sums into every peer's mailbox over NVLink,
 * raises a flag and sums the ranks' contributions in rank order itself — one kernel per stage as on one GPU, every rank
 * bit-identical (libpointmatcher_b200/csrc/comm.cuh).  At most 8 ranks.  Both inits may be combined: the mailboxes take
 * the per-iteration exchanges, NCCL the one-off all-gather of pmgpu_ref_compute_normals, which under a communicator
 * computes only this rank's slice of the map's normals.  A rank whose peers do not arrive within 2 s
 * stops with PMGPU_ERR_COMM.  All ranks must issue the same sequence of calls; a rank's reading slice may be empty. */
int pmgpu_comm_peer_handle(pmgpu_ctx* ctx, void* handle_128);
int pmgpu_comm_peer_init(pmgpu_ctx* ctx, const void* handles, int rank, int nranks);
int pmgpu_comm_destroy(pmgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PMGPU_H */
