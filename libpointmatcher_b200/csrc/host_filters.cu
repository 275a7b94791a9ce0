// host_filters.cu — the CPU pre-filters either side of the hot path: what
// BASELINE config 1 (examples/data/default.yaml, ICPChainBase::setDefault, ICP.cpp:100-113) runs
// ONCE per cloud before the first iteration.  They stay on the host, as in the reference: both
// draw from std::rand() in point order (RandomSampling.cpp:66, SamplingSurfaceNormal.cpp:273), and
// SamplingSurfaceNormal's bins come out of a recursive std::nth_element whose within-bin order
// decides which points survive — reproducing the reference means running the same serial
// algorithm on the same C++ runtime, not re-deriving it for a GPU.  No device code in this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/pmgpu.h"
#include "core/linalg.h"

namespace {

struct BinBuild {
    float* features;      // rows x n, column-major, modified in place by samplingMethod 1
    int rows, n;
    float* descriptors;   // desc_rows x n or null
    int desc_rows;
    float ratio;
    int knn, sampling_method;
    float max_box_dim;
    bool average_descriptors;
    int flags;
    std::vector<int> indices;
    std::vector<int> keep;
    int unfit = 0;
    float* normals;       // 3 x n, written at the kept columns
    float* densities;     // n
    float* eig_values;    // 3 x n
    float* eig_vectors;   // 9 x n, row-major serialisation (utils.h:90-103)

    float f(int dim, int col) const { return features[(size_t)col * rows + dim]; }
};

// SamplingSurfaceNormal.cpp:232-342
void fuse_range(BinBuild& d, int first, int last) {
    const int count = last - first;
    const int dimN = d.rows - 1;  // 3, or 2 for 2-D clouds
    float lo[3], hi[3], sum[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < dimN; ++a) { lo[a] = std::numeric_limits<float>::max(); hi[a] = -std::numeric_limits<float>::max(); }
    for (int i = 0; i < count; ++i)
        for (int a = 0; a < dimN; ++a) {
            const float v = d.f(a, d.indices[first + i]);
            lo[a] = std::min(lo[a], v);
            hi[a] = std::max(hi[a], v);
            sum[a] += v;  // rowwise().sum(): column after column
        }
    float box_dim = hi[0] - lo[0];
    for (int a = 1; a < dimN; ++a) box_dim = std::max(box_dim, hi[a] - lo[a]);
    if (box_dim > d.max_box_dim) { d.unfit += count; return; }
    float mean[3];
    for (int a = 0; a < dimN; ++a) mean[a] = sum[a] / (float)count;
    // C = NN NN^T in float, every entry a dot product over the bin's points
    float C[9] = {0};
    float max_norm = 0.f;
    for (int i = 0; i < count; ++i) {
        float v[3] = {0.f, 0.f, 0.f};
        for (int a = 0; a < dimN; ++a) v[a] = d.f(a, d.indices[first + i]) - mean[a];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) C[r + 3 * c] += v[r] * v[c];
        max_norm = std::max(max_norm, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    }
    float eig_va[3] = {1.f, 0.f, 0.f};  // Vector::Identity(3, 1)
    float eig_ve[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    const bool want_eig = (d.flags & (PMGPU_KEEP_NORMALS | PMGPU_KEEP_EIGEN_VALUES | PMGPU_KEEP_EIGEN_VECTORS)) != 0;
    if (want_eig) {
        float Cq[9];
        std::memcpy(Cq, C, sizeof(Cq));
        if (pm::fullpiv_qr_rank3(Cq, dimN) + 1 >= dimN) {  // SamplingSurfaceNormal.cpp:258
            if (dimN == 2) {
                double w[2], V[4];
                pm::sym_eig2((double)C[0], (double)C[1], (double)C[4], w, V);
                eig_va[0] = (float)w[0]; eig_va[1] = (float)w[1];
                eig_ve[0] = (float)V[0]; eig_ve[1] = (float)V[1]; eig_ve[3] = (float)V[2]; eig_ve[4] = (float)V[3];
            } else {
                double A[9], w[3], V[9];
                for (int i = 0; i < 9; ++i) A[i] = (double)C[i];
                pm::jacobi_eig3(A, w, V);
                for (int i = 0; i < 3; ++i) eig_va[i] = (float)w[i];
                for (int i = 0; i < 9; ++i) eig_ve[i] = (float)V[i];
            }
        } else {
            d.unfit += count;
            return;
        }
    }
    float normal[3] = {0.f, 0.f, 0.f};
    if (d.flags & PMGPU_KEEP_NORMALS) {  // computeNormal, utils.h:122-139: first smallest eigenvalue
        int smallest = 0;
        float value = std::numeric_limits<float>::max();
        for (int j = 0; j < dimN; ++j)
            if (eig_va[j] < value) { smallest = j; value = eig_va[j]; }
        for (int a = 0; a < dimN; ++a) normal[a] = eig_ve[a + 3 * smallest];
    }
    float density = 0.f;
    if (d.flags & PMGPU_KEEP_DENSITIES) {  // computeDensity, utils.h:105-120
        const float volume = (float)((4. / 3.) * M_PI * std::pow((double)max_norm, 3));
        density = (float)count / volume;
    }
    auto write = [&](int k) {
        // spans follow the cloud's dimension: normals dimN, eigValues dimN, eigVectors dimN * dimN
        if (d.flags & PMGPU_KEEP_NORMALS) std::memcpy(d.normals + dimN * (size_t)k, normal, dimN * sizeof(float));
        if (d.flags & PMGPU_KEEP_DENSITIES) d.densities[k] = density;
        if (d.flags & PMGPU_KEEP_EIGEN_VALUES) std::memcpy(d.eig_values + dimN * (size_t)k, eig_va, dimN * sizeof(float));
        if (d.flags & PMGPU_KEEP_EIGEN_VECTORS)
            for (int r = 0; r < dimN; ++r)
                for (int c = 0; c < dimN; ++c) d.eig_vectors[dimN * dimN * (size_t)k + dimN * r + c] = eig_ve[r + 3 * c];
    };
    if (d.sampling_method == 0) {
        for (int i = 0; i < count; ++i) {
            const float r = (float)std::rand() / (float)RAND_MAX;
            if (r < d.ratio) {
                const int k = d.indices[first + i];
                d.keep.push_back(k);
                write(k);
            }
        }
    } else {
        const int k = d.indices[first];
        d.keep.push_back(k);
        for (int a = 0; a < dimN; ++a) d.features[(size_t)k * d.rows + a] = mean[a];
        d.features[(size_t)k * d.rows + dimN] = 1.f;
        if (d.descriptors && d.desc_rows > 0 && d.average_descriptors) {
            std::vector<float> merged(d.desc_rows, 0.f);
            for (int i = 0; i < count; ++i)
                for (int r = 0; r < d.desc_rows; ++r) merged[r] += d.descriptors[(size_t)d.indices[first + i] * d.desc_rows + r];
            for (int r = 0; r < d.desc_rows; ++r) d.descriptors[(size_t)k * d.desc_rows + r] = merged[r] / (float)count;
        }
        write(k);
    }
}

// SamplingSurfaceNormal.cpp:177-230: split the widest dimension of the CELL at the median
void build_bins(BinBuild& d, int first, int last, float* min_values, float* max_values) {
    const int count = last - first;
    if (count <= d.knn) { fuse_range(d, first, last); return; }
    const int dimN = d.rows;  // the bounds vectors have `rows` entries (homogeneous row included, extent 0)
    int cut_dim = 0;
    float best = max_values[0] - min_values[0];
    for (int a = 1; a < dimN; ++a)
        if (max_values[a] - min_values[a] > best) { best = max_values[a] - min_values[a]; cut_dim = a; }  // argMax: first maximum
    const int right_count = count / 2, left_count = count - right_count;
    std::nth_element(d.indices.begin() + first, d.indices.begin() + first + left_count, d.indices.begin() + last,
                     [&](int a, int b) { return d.f(cut_dim, a) < d.f(cut_dim, b); });
    const float cut_val = d.f(cut_dim, d.indices[first + left_count]);
    std::vector<float> left_max(max_values, max_values + dimN), right_min(min_values, min_values + dimN);
    left_max[cut_dim] = cut_val;
    right_min[cut_dim] = cut_val;
    build_bins(d, first, first + left_count, min_values, left_max.data());
    build_bins(d, first + left_count, last, right_min.data(), max_values);
}

}  // namespace

extern "C" {

void pmgpu_host_srand(unsigned seed) { std::srand(seed); }

int pmgpu_host_random_sampling(int n, float prob, int32_t* keep_out) {
    if (n < 0 || !keep_out) return -1;
    const double p = (double)prob;  // `const double prob` holding the float parameter (RandomSampling.h:65)
    int j = 0;
    for (int i = 0; i < n; ++i) {
        const float r = (float)std::rand() / (float)RAND_MAX;
        if (r < p) keep_out[j++] = i;
    }
    return j;
}

int pmgpu_host_rand(void) { return std::rand(); }

// MaxPointCount.cpp:71-110.  The reference means to swap columns j and idx, but `const auto feat =
// cloud.features.col(j)` is an Eigen view, not a copy: column j takes column idx and column idx stays
// as it is.  idx >= j and only positions < j have been overwritten, so kept column j is ORIGINAL
// column idx_j (duplicates possible) — restated as that.
int pmgpu_host_max_point_count(int n, uint64_t seed, uint64_t max_count, int32_t* order_out) {
    if (n < 0 || !order_out) return -1;
    if (n == 0) return 0;
    const size_t N = (size_t)(n - 1);  // as the reference: cols() - 1 converted to size_t
    if (!(max_count <= N)) {
        for (int i = 0; i < n; ++i) order_out[i] = i;
        return n;
    }
    std::srand((unsigned)seed);
    for (size_t j = 0; j < max_count; ++j) {
        const size_t idx = j + static_cast<size_t>((N - j) * (static_cast<float>(std::rand() / static_cast<float>(RAND_MAX))));
        order_out[j] = (int32_t)idx;
    }
    return (int)max_count;
}

// MaxDensity.cpp:60-105: points denser than max_density survive with probability max_density / density
int pmgpu_host_max_density(const float* densities, int stride, int n, float max_density, int32_t* keep_out) {
    if (n < 0 || !keep_out || (n > 0 && !densities) || stride < 1) return -1;
    float last = -std::numeric_limits<float>::infinity();
    for (int i = 0; i < n; ++i) last = std::max(last, densities[(size_t)i * stride]);
    int saturated = 0;
    for (int i = 0; i < n; ++i) saturated += densities[(size_t)i * stride] == last;
    int j = 0;
    for (int i = 0; i < n; ++i) {
        const float density = densities[(size_t)i * stride];
        if (density > max_density) {
            const float r = (float)std::rand() / (float)RAND_MAX;
            float accept = max_density / density;
            if (density == last) accept = accept * (1 - saturated / n);  // integer division, as the reference
            if (r < accept) keep_out[j++] = i;
        } else
            keep_out[j++] = i;
    }
    return j;
}

int pmgpu_host_sampling_surface_normal(float* features, int rows, int n, float* descriptors, int desc_rows, float ratio, int knn, int sampling_method,
                                       float max_box_dim, int average_descriptors, int flags, int32_t* keep_out, float* normals_out,
                                       float* densities_out, float* eig_values_out, float* eig_vectors_out, int* unfit_out) {
    if (!features || (rows != 4 && rows != 3) || n < 0 || !keep_out || knn < 1) return -1;
    if ((flags & PMGPU_KEEP_NORMALS) && !normals_out) return -1;
    if ((flags & PMGPU_KEEP_DENSITIES) && !densities_out) return -1;
    if ((flags & PMGPU_KEEP_EIGEN_VALUES) && !eig_values_out) return -1;
    if ((flags & PMGPU_KEEP_EIGEN_VECTORS) && !eig_vectors_out) return -1;
    BinBuild d;
    d.features = features; d.rows = rows; d.n = n;
    d.descriptors = descriptors; d.desc_rows = desc_rows;
    d.ratio = ratio; d.knn = knn; d.sampling_method = sampling_method; d.max_box_dim = max_box_dim;
    d.average_descriptors = average_descriptors != 0;
    d.flags = flags;
    d.normals = normals_out; d.densities = densities_out; d.eig_values = eig_values_out; d.eig_vectors = eig_vectors_out;
    d.indices.resize(n);
    for (int i = 0; i < n; ++i) d.indices[i] = i;
    if (n > 0) {
        std::vector<float> lo(rows, std::numeric_limits<float>::max()), hi(rows, -std::numeric_limits<float>::max());
        for (int i = 0; i < n; ++i)
            for (int a = 0; a < rows; ++a) { lo[a] = std::min(lo[a], d.f(a, i)); hi[a] = std::max(hi[a], d.f(a, i)); }
        build_bins(d, 0, n, lo.data(), hi.data());
    }
    std::sort(d.keep.begin(), d.keep.end());
    for (size_t i = 0; i < d.keep.size(); ++i) keep_out[i] = d.keep[i];
    if (unfit_out) *unfit_out = d.unfit;
    return (int)d.keep.size();
}

}  // extern "C"
