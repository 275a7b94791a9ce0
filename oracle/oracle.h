/*
 * oracle/oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The oracle is a dependency-free C++17 restatement of the libpointmatcher ICP hot path
 * (reference files cited next to every function in oracle.cpp).  The reference itself cannot
 * be compiled in this image (Eigen3, Boost and libnabo are absent), so:
 *   - kNN match indices: PARITY UNPINNED by the reference (libnabo is an un-vendored external
 *     dependency, >= 1.0.7, and no reference test inspects Matches.ids).  The oracle restates
 *     libnabo's published algorithms (brute force; bucketed kd-tree, linear heap).
 *   - final transforms ARE pinned by the reference's own fixtures: five golden
 *     examples/data/icp_data/*.ref_trans (criterion of utest/utest.cpp:146-158, < 3 %), validT3d
 *     (utest/utest.h:66-84) — tests/test_reference_goldens.py — and the known-answer tests
 *     icpSingular / icpIdentity (utest/utest.cpp:162-220) — tests/test_oracle_golden.py.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* filter types for the outlier chain (OutlierFiltersImpl.cpp) */
enum { ORC_FILTER_MAXDIST = 0, ORC_FILTER_MEDIANDIST = 1, ORC_FILTER_TRIMMEDDIST = 2, ORC_FILTER_ROBUST = 3, ORC_FILTER_SURFACENORMAL = 4,
       ORC_FILTER_VARTRIMMEDDIST = 5 /* param = lambda; minRatio / maxRatio through orc_set_var_trimmed_ratios */,
       ORC_FILTER_MINDIST = 6 /* MinDistOutlierFilter, param = minDist (un-squared), OutlierFiltersImpl.cpp:87-101 */ };
/* RobustOutlierFilter (OutlierFiltersImpl.cpp:420-598): the filter word carries its discrete parameters:
 *   bits 0-7  ORC_FILTER_ROBUST | bits 8-15 robust function | bits 16-19 scale estimator |
 *   bits 20-27 nbIterationForScale;   filter_param = tuning.  distanceType point2point, approximation inf. */
enum { ORC_ROBUST_CAUCHY = 0, ORC_ROBUST_WELSCH, ORC_ROBUST_SC, ORC_ROBUST_GM, ORC_ROBUST_TUKEY, ORC_ROBUST_HUBER, ORC_ROBUST_L1,
       ORC_ROBUST_STUDENT };
enum { ORC_SCALE_NONE = 0, ORC_SCALE_MAD = 1, ORC_SCALE_BERG = 2, ORC_SCALE_STD = 3 };
#define ORC_ROBUST_P2PLANE (1 << 28) /* distanceType point2plane, or-ed into the word */
#define ORC_ROBUST_WORD(fct, scale, nb_iter) (ORC_FILTER_ROBUST | ((fct) << 8) | ((scale) << 16) | ((nb_iter) << 20))
/* error minimizers */
enum { ORC_MIN_P2POINT = 0, ORC_MIN_P2PLANE = 1, ORC_MIN_P2POINT_COV = 2, ORC_MIN_P2PLANE_COV = 3, ORC_MIN_P2POINT_SIM = 4 };
#define ORC_MIN_FORCE4DOF 0x100 /* or-ed into a point-to-plane minimizer id: PointToPlaneErrorMinimizer force4DOF */
#define ORC_MIN_FORCE2D 0x200  /* likewise: force2D (rotation about z + x/y translation on 3-D clouds) */
/* status codes */
enum {
    ORC_OK = 0,
    ORC_ERR_NO_OUTLIER_TO_FILTER = 1, /* ConvergenceError("no outlier to filter")  Matches.cpp:76 */
    ORC_ERR_NO_POINT_TO_MINIMIZE = 2, /* ConvergenceError  ErrorMinimizer.cpp:76 */
    ORC_ERR_NOT_ORTHOGONAL = 3,       /* TransformationError  TransformationsImpl.cpp:62 */
    ORC_ERR_BAD_QUANTILE = 4,         /* ConvergenceError  Matches.cpp:79 */
    ORC_ERR_NAN = 5,                  /* ConvergenceError  TransformationCheckersImpl.cpp:154 */
    ORC_ERR_BAD_ARG = 6,
    ORC_ERR_KNN_TOO_LARGE = 7
};

typedef struct orc_icp_config {
    int knn;
    float epsilon;
    float max_dist;          /* KDTreeMatcher maxDist (not squared) */
    int search_type;         /* 0 brute force, 1 kd-tree linear heap */
    int nfilters;
    int filter_type[8];
    float filter_param[8];   /* maxDist (un-squared) | factor | ratio */
    int minimizer;
    float sensor_std_dev;
    int max_iterations;      /* CounterTransformationChecker */
    int use_differential;    /* DifferentialTransformationChecker on/off */
    float min_diff_rot_err;
    float min_diff_trans_err;
    int smooth_length;
    int acc_double;          /* 0: float sums as the reference; 1: double sums ("truth") */
    int nthreads;
} orc_icp_config;

/* --- kNN (MatchersImpl.cpp:77-101 + libnabo semantics) -------------------------------- */
void* orc_kdtree_create(const float* feat, int rows, int n);
void orc_kdtree_destroy(void* tree);
/* query: rows x nq column-major; ids/dists: k x nq column-major.  Returns the visit count. */
long orc_kdtree_knn(void* tree, const float* query, int rows, int nq, int k, float eps,
                    float max_radius, int32_t* ids, float* dists, int nthreads);
long orc_bruteforce_knn(const float* ref, int rows, int nr, const float* query, int nq, int k,
                        float max_radius, int32_t* ids, float* dists, int nthreads);

/* KDTreeVarDistMatcher (MatchersImpl.cpp:132-150): one maximum radius per query (the reading's maxDistField descriptor) */
long orc_bruteforce_knn_var(const float* ref, int rows, int nr, const float* query, int nq, int k,
                            const float* max_radii, int32_t* ids, float* dists, int nthreads);

/* --- RigidTransformation::compute (TransformationsImpl.cpp:49-105) -------------------- */
int orc_rigid_transform(const float* T16, const float* in, int n, float* out);
int orc_rotate_normals(const float* T16, const float* in3, int n, float* out3);

/* --- Matches::getDistsQuantile / outlier chain ----------------------------------------- */
/* RobustOutlierFilter `approximation` (metres; +inf: none) of the filters evaluated from here on */
void orc_set_robust_approximation(float approximation);
int orc_dists_quantile(const float* dists, long n, float quantile, float* out);
/* VarTrimmedDistOutlierFilter (OutlierFiltersImpl.cpp:152-218): minRatio / maxRatio of the filters evaluated from now on
 * (defaults 0.05 / 0.99), and optimizeInlierRatio on its own */
void orc_set_var_trimmed_ratios(float min_ratio, float max_ratio);
int orc_var_trimmed_ratio(const float* dists, long n, float min_ratio, float max_ratio, float lambda, float* ratio_out);
/* the chain with what RobustOutlierFilter distanceType point2plane reads: ids, the clouds (4 x n) and the reference normals */
int orc_outlier_weights_geom(const float* dists, const int32_t* ids, int knn, int n, int nfilters, const int* types, const float* params,
                             const float* reading4xn, const float* reference4xnr, const float* ref_normals, float* weights, float* limits_out);
int orc_outlier_weights(const float* dists, int knn, int n, int nfilters, const int* types,
                        const float* params, float* weights, float* limits_out);
/* SurfaceNormalOutlierFilter (OutlierFiltersImpl.cpp:222-285, type ORC_FILTER_SURFACENORMAL, param = maxAngle)
 * needs the matches and both clouds' normals: the reading's (3 x n, already rotated like the reading) and the
 * reference's (3 x nr).  orc_outlier_weights_sn evaluates a chain that may contain it; orc_icp takes the
 * un-rotated reading normals through orc_set_reading_normals (NULL switches the filter to "all ones"). */
int orc_outlier_weights_sn(const float* dists, const int32_t* ids, int knn, int n, int nfilters, const int* types, const float* params,
                           const float* reading_normals, const float* ref_normals, float* weights, float* limits_out);
void orc_set_reading_normals(const float* normals3xn);

/* --- ErrorElements + minimizers -------------------------------------------------------- */
/* reading: 4 x nq (already transformed), reference 4 x nr, ref_normals 3 x nr or NULL.
 * T_out 4x4 col-major; cov_out 6x6 col-major (WithCov variants, else untouched);
 * stats_out = {pointUsedRatio, weightedPointUsedRatio, nbRejectedMatches, nbRejectedPoints,
 *              nbKept}. */
int orc_minimize(int minimizer, const float* reading, int nq, const float* reference, int nr,
                 const float* ref_normals, const int32_t* ids, const float* dists,
                 const float* weights, int knn, float sensor_std_dev, int acc_double,
                 float* T_out, float* cov_out, float* stats_out);

/* --- SurfaceNormalDataPointsFilter (SurfaceNormal.cpp:82-290) -------------------------- */
/* any output pointer may be NULL.  normals 3 x n, densities n, eig_values 3 x n,
 * eig_vectors 9 x n, matched_ids knn x n (float), mean_dists n, ids_out/dists_out knn x n.
 * degenerate_out: number of degenerate points; gap_out[n]: relative eigen-gap
 * (lambda_mid - lambda_min) / trace for tolerance gating in tests. */
int orc_surface_normals(const float* feat, int rows, int n, int knn, float eps, float max_dist,
                        int sort_eigen, int smooth_normals, int nthreads, float* normals,
                        float* densities, float* eig_values, float* eig_vectors,
                        float* matched_ids, float* mean_dists, int32_t* ids_out,
                        float* dists_out, float* gap_out, int* degenerate_out);

/* --- ICP::compute (ICP.cpp:264-449) ----------------------------------------------------- */
/* reading 4 x nq, reference 4 x nr, ref_normals 3 x nr or NULL, T_init 4x4.
 * T_out: final 4x4.  T_iters_out: optional [max_iterations][16] per-iteration T_iter.
 * iterations_out: iterations executed. */
int orc_icp(const float* reading, int nq, const float* reference, int nr,
            const float* ref_normals, const float* T_init, const orc_icp_config* cfg,
            float* T_out, float* T_iters_out, int* iterations_out, float* cov_out,
            float* stats_out);

/* wall-clock of the last orc_icp call: {structure build s, iteration loop s, matching s, iterations} */
void orc_last_timings(double* out4);

/* misc helpers used by the tests */
void orc_quaternion_angular_distance(const float* Ta16, const float* Tb16, float* out);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
