// host/PointMatcher.h — the reference's public header (pointmatcher/PointMatcher.h, v1.3.1) for the
// ICP hot path, with the GPU modules registered under the reference's names.
//
// Same class names, method signatures, parameter tables and exception types as the reference, so
// code and YAML written against libpointmatcher compiles / loads unchanged for this path:
//   PointMatcher<T>::DataPoints / Matches / OutlierWeights / TransformationParameters
//   Matcher (KDTreeMatcher) . OutlierFilter(s) (MaxDist, MedianDist, TrimmedDist, Null)
//   ErrorMinimizer (PointToPoint[WithCov], PointToPlane[WithCov]) . DataPointsFilter(s)
//   (SurfaceNormal, Identity) . Transformation (Rigid) . TransformationChecker(s) (Counter,
//   Differential, Bound) . Inspector (Null) . ICP (setDefault, loadFromYaml, operator())
// Every module marshals to the C ABI of include/pmgpu.h; there is no CPU implementation of the hot
// path behind these classes (no Eigen either: Matrix below is a minimal column-major container
// with the memory layout of Eigen::Matrix<T, Dynamic, Dynamic>).
//
// Only T = float, 3-D (features.rows() == 4) runs on the GPU; PointMatcher<double> and 2-D clouds
// are accepted by the interfaces and answered with ConfigurationError, never silently.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cctype>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <memory>
#include <future>
#include <iostream>

#include "../../include/pmgpu.h"
#include "Support.h"

namespace pmb {

// minimal column-major dense matrix, layout-compatible with Eigen's default
template <typename T>
class Matrix {
public:
    Matrix() : r_(0), c_(0) {}
    Matrix(int rows, int cols) : r_(rows), c_(cols), d_((size_t)rows * cols) {}
    static Matrix Zero(int rows, int cols) { Matrix m(rows, cols); std::fill(m.d_.begin(), m.d_.end(), T(0)); return m; }
    static Matrix Constant(int rows, int cols, T v) { Matrix m(rows, cols); std::fill(m.d_.begin(), m.d_.end(), v); return m; }
    static Matrix Identity(int rows, int cols) {
        Matrix m = Zero(rows, cols);
        for (int i = 0; i < std::min(rows, cols); ++i) m(i, i) = T(1);
        return m;
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    size_t size() const { return d_.size(); }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    T& operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
    const T& operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
    T& operator()(int i) { return d_[i]; }
    const T& operator()(int i) const { return d_[i]; }
    void resize(int rows, int cols) { r_ = rows; c_ = cols; d_.assign((size_t)rows * cols, T(0)); }
    // keeps the leading block, like Eigen's conservativeResize
    void conservativeResize(int rows, int cols) {
        Matrix m = Zero(rows, cols);
        for (int j = 0; j < std::min(cols, c_); ++j)
            for (int i = 0; i < std::min(rows, r_); ++i) m(i, j) = (*this)(i, j);
        *this = m;
    }
    // dense product with a single accumulator per entry, k ascending and no FMA contraction
    // assumed (the depth-4 GEMM order of the reference's 4x4 bookkeeping, ICP.cpp:411-412,448)
    Matrix operator*(const Matrix& o) const {
        Matrix out(r_, o.c_);
        for (int j = 0; j < o.c_; ++j)
            for (int i = 0; i < r_; ++i) {
                volatile T acc = (*this)(i, 0) * o(0, j);
                for (int k = 1; k < c_; ++k) {
                    volatile T prod = (*this)(i, k) * o(k, j);
                    acc = acc + prod;
                }
                out(i, j) = acc;
            }
        return out;
    }
    bool operator==(const Matrix& o) const { return r_ == o.r_ && c_ == o.c_ && d_ == o.d_; }

private:
    int r_, c_;
    std::vector<T> d_;
};

}  // namespace pmb

template <typename T>
struct PointMatcher {
    // ---------------------------------------------------------------------------------------------
    // basic types (PointMatcher.h:160-200, 371-397)
    // ---------------------------------------------------------------------------------------------
    typedef T ScalarType;
    typedef pmb::Matrix<T> Matrix;
    typedef pmb::Matrix<T> Vector;
    typedef pmb::Matrix<int> IntMatrix;
    typedef Matrix TransformationParameters;
    typedef Matrix OutlierWeights;
    typedef PointMatcherSupport::Parametrizable Parametrizable;
    typedef Parametrizable::Parameters Parameters;
    typedef Parametrizable::ParameterDoc ParameterDoc;
    typedef Parametrizable::ParametersDoc ParametersDoc;
    typedef Parametrizable::InvalidParameter InvalidParameter;
    typedef PointMatcherSupport::InvalidModuleType InvalidModuleType;
    typedef PointMatcherSupport::TransformationError TransformationError;
    typedef PointMatcherSupport::ConfigurationError ConfigurationError;
    typedef PointMatcherSupport::InvalidElement InvalidElement;

    struct ConvergenceError : std::runtime_error {  // PointMatcher.h:148-151
        explicit ConvergenceError(const std::string& reason) : std::runtime_error(reason) {}
    };

    // ---- DataPoints (PointMatcher.h:207-358) -------------------------------------------------------
    struct DataPoints {
        struct Label {
            std::string text;
            size_t span;
            Label(const std::string& text = "", const size_t span = 0) : text(text), span(span) {}
            bool operator==(const Label& that) const { return text == that.text && span == that.span; }
        };
        struct Labels : std::vector<Label> {
            bool contains(const std::string& text) const {
                for (const Label& l : *this)
                    if (l.text == text) return true;
                return false;
            }
        };
        struct InvalidField : std::runtime_error {  // PointMatcher.h:250-253
            explicit InvalidField(const std::string& reason) : std::runtime_error(reason) {}
        };

        Matrix features;          // (dim + 1) x N, homogeneous row last
        Labels featureLabels;
        Matrix descriptors;       // (sum of spans) x N
        Labels descriptorLabels;

        DataPoints() {}
        DataPoints(const Matrix& features, const Labels& featureLabels) : features(features), featureLabels(featureLabels) {}
        DataPoints(const Matrix& features, const Labels& featureLabels, const Matrix& descriptors, const Labels& descriptorLabels)
            : features(features), featureLabels(featureLabels), descriptors(descriptors), descriptorLabels(descriptorLabels) {}

        unsigned getNbPoints() const { return features.cols(); }
        unsigned getEuclideanDim() const { return features.rows() - 1; }
        unsigned getHomogeneousDim() const { return features.rows(); }
        unsigned getDescriptorDim() const { return descriptors.rows(); }
        bool descriptorExists(const std::string& name) const { return descriptorLabels.contains(name); }
        bool descriptorExists(const std::string& name, unsigned dim) const {
            for (const Label& l : descriptorLabels)
                if (l.text == name) return l.span == dim;
            return false;
        }
        unsigned getDescriptorDimension(const std::string& name) const {
            for (const Label& l : descriptorLabels)
                if (l.text == name) return l.span;
            return 0;
        }
        // DataPoints.cpp:917-942: row where the named descriptor starts
        unsigned getDescriptorStartingRow(const std::string& name) const {
            unsigned row = 0;
            for (const Label& l : descriptorLabels) {
                if (l.text == name) return row;
                row += l.span;
            }
            throw InvalidField("Field " + name + " not found");
        }
        Matrix getDescriptorCopyByName(const std::string& name) const {
            const unsigned row = getDescriptorStartingRow(name), span = getDescriptorDimension(name);
            Matrix out(span, descriptors.cols());
            for (int j = 0; j < descriptors.cols(); ++j)
                for (unsigned i = 0; i < span; ++i) out(i, j) = descriptors(row + i, j);
            return out;
        }
        // DataPoints.cpp:783-822: re-use an existing same-span descriptor or append rows
        void allocateDescriptor(const std::string& name, const unsigned dim) {
            for (const Label& l : descriptorLabels) {
                if (l.text == name) {
                    if (l.span == dim) return;
                    throw InvalidField("Field " + name + " already exists but has a different dimension");
                }
            }
            const int oldRows = descriptors.rows(), n = features.cols();
            if (descriptors.cols() != n && oldRows > 0) throw InvalidField("descriptors and features have different point counts");
            Matrix grown = Matrix::Zero(oldRows + dim, n);
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < oldRows; ++i) grown(i, j) = descriptors(i, j);
            descriptors = grown;
            descriptorLabels.push_back(Label(name, dim));
        }
        // DataPoints::load (IO.cpp:376-392): format from the extension.  The CSV reader (IO.cpp:535-760:
        // header names x y z pad -> features, nx ny nz -> "normals", any other name -> its own
        // descriptor; no header: 2 or 3 columns) and the ASCII legacy-VTK reader (IO.cpp:949-1250:
        // POINTS, NORMALS, VECTORS, SCALARS) — the formats of the reference's example data.
        static DataPoints load(const std::string& fileName) {
            std::ifstream ifs(fileName.c_str());
            if (!ifs.good()) throw std::runtime_error("Cannot open file " + fileName);
            const size_t dot = fileName.rfind('.');
            std::string ext = dot == std::string::npos ? "" : fileName.substr(dot);
            for (auto& c : ext) c = (char)std::tolower(c);
            if (ext == ".csv") return loadCSV(ifs);
            if (ext == ".vtk") return loadVTK(ifs, fileName);
            throw std::runtime_error("loadAnyFormat(): Unknown extension \"" + ext + "\" for file \"" + fileName +
                                     "\", extension must be either \".vtk\" or \".csv\"");
        }
        static DataPoints loadCSV(std::istream& is) {
            auto split = [](const std::string& line) {
                std::vector<std::string> out;
                std::string cur;
                for (char c : line) {
                    if (c == ' ' || c == '\t' || c == ',' || c == ';' || c == '\r') {
                        if (!cur.empty()) { out.push_back(cur); cur.clear(); }
                    } else cur.push_back(c);
                }
                if (!cur.empty()) out.push_back(cur);
                return out;
            };
            std::string line;
            if (!std::getline(is, line)) throw std::runtime_error("CSV parse error: empty file");
            const bool hasHeader = line.find_first_not_of(" ,+-.1234567890Ee\r") != std::string::npos;
            std::vector<std::string> header;
            std::vector<std::vector<T>> rows;
            auto pushRow = [&](const std::string& l) {
                const auto tok = split(l);
                if (tok.size() != header.size()) throw std::runtime_error("CSV parse error: a row does not have the width of the header");
                std::vector<T> r(tok.size());
                for (size_t i = 0; i < tok.size(); ++i) r[i] = (T)std::stod(tok[i]);
                rows.push_back(r);
            };
            if (hasHeader) header = split(line);
            else {
                const size_t dim = split(line).size();
                if (dim != 2 && dim != 3)
                    throw std::runtime_error("CSV parse error: " + std::to_string(dim) + " columns and no header: not obvious which columns to load for x, y or z");
                header = {"x", "y"};
                if (dim == 3) header.push_back("z");
                pushRow(line);
            }
            while (std::getline(is, line)) {
                if (split(line).empty()) break;  // the reader stops at the first empty line
                pushRow(line);
            }
            const int n = (int)rows.size();
            auto column = [&](const std::string& name) {
                for (size_t j = 0; j < header.size(); ++j)
                    if (header[j] == name) return (int)j;
                return -1;
            };
            if (column("x") < 0 || column("y") < 0) throw std::runtime_error("CSV parse error: no x / y column");
            DataPoints out;
            std::vector<int> featCols;
            for (const char* f : {"x", "y", "z", "pad"})
                if (column(f) >= 0) { featCols.push_back(column(f)); out.featureLabels.push_back(Label(f, 1)); }
            const bool hasPad = column("pad") >= 0;
            out.features = Matrix(featCols.size() + (hasPad ? 0 : 1), n);
            for (int i = 0; i < n; ++i) {
                for (size_t r = 0; r < featCols.size(); ++r) out.features(r, i) = rows[i][featCols[r]];
                if (!hasPad) out.features(featCols.size(), i) = T(1);
            }
            if (!hasPad) out.featureLabels.push_back(Label("pad", 1));
            // descriptors: columns with the same internal name are rows of one descriptor, in file order
            auto internal = [](const std::string& name) -> std::string {
                if (name == "nx" || name == "ny" || name == "nz" || name == "normal_x" || name == "normal_y" || name == "normal_z") return "normals";
                if (name == "red" || name == "green" || name == "blue" || name == "alpha") return "color";
                if (name.rfind("observationDirections", 0) == 0 && name.size() == 22) return "observationDirections";
                if (name.rfind("eigValues", 0) == 0 && name.size() == 10) return "eigValues";
                if (name.rfind("eigVectors", 0) == 0 && name.size() == 12) return "eigVectors";
                return name;
            };
            std::vector<std::pair<std::string, std::vector<int>>> groups;
            for (size_t j = 0; j < header.size(); ++j) {
                const std::string& name = header[j];
                if (name == "x" || name == "y" || name == "z" || name == "pad") continue;
                if (name == "time") throw std::runtime_error("CSV parse error: time columns are not supported by this reader");
                const std::string in = internal(name);
                bool found = false;
                for (auto& g : groups)
                    if (g.first == in) { g.second.push_back((int)j); found = true; }
                if (!found) groups.push_back({in, {(int)j}});
            }
            for (const auto& g : groups) {
                Matrix d(g.second.size(), n);
                for (int i = 0; i < n; ++i)
                    for (size_t r = 0; r < g.second.size(); ++r) d(r, i) = rows[i][g.second[r]];
                out.addDescriptor(g.first, d);
            }
            return out;
        }
        static DataPoints loadVTK(std::istream& is, const std::string& fileName = "") {
            std::string l1, l2, l3;
            std::getline(is, l1); std::getline(is, l2); std::getline(is, l3);
            if (l1.rfind("# vtk DataFile Version", 0) != 0) throw std::runtime_error("Header in VTK file " + fileName + " is not valid");
            while (!l3.empty() && (l3.back() == '\r' || l3.back() == ' ')) l3.pop_back();
            if (l3 != "ASCII") throw std::runtime_error("VTK reader: only ASCII legacy files are supported here, got " + l3);
            std::vector<std::string> tok;
            for (std::string t; is >> t;) tok.push_back(t);
            if (tok.size() < 2 || tok[0] != "DATASET" || (tok[1] != "POLYDATA" && tok[1] != "UNSTRUCTURED_GRID"))
                throw std::runtime_error("Invalid data type in VTK file: only POLYDATA and UNSTRUCTURED_GRID are supported");
            DataPoints out;
            int n = -1;
            size_t i = 2;
            auto block = [&](size_t first, int rows_) {
                if (first + (size_t)rows_ * n > tok.size()) throw std::runtime_error("VTK reader: truncated file");
                Matrix m(rows_, n);
                for (int c = 0; c < n; ++c)
                    for (int r = 0; r < rows_; ++r) m(r, c) = (T)std::stod(tok[first + (size_t)c * rows_ + r]);
                return m;
            };
            auto isNumber = [](const std::string& t) { return !t.empty() && t.find_first_not_of("0123456789") == std::string::npos; };
            while (i < tok.size()) {
                const std::string& key = tok[i];
                if (key == "POINTS") {
                    n = std::stoi(tok[i + 1]);
                    const Matrix p = block(i + 3, 3);
                    out.features = Matrix(4, n);
                    for (int c = 0; c < n; ++c) {
                        for (int r = 0; r < 3; ++r) out.features(r, c) = p(r, c);
                        out.features(3, c) = T(1);
                    }
                    for (const char* f : {"x", "y", "z", "pad"}) out.featureLabels.push_back(Label(f, 1));
                    i += 3 + 3 * (size_t)n;
                } else if (key == "VERTICES" || key == "LINES" || key == "POLYGONS" || key == "TRIANGLE_STRIPS" || key == "CELLS") {
                    i += 3 + (size_t)std::stol(tok[i + 2]);
                } else if (key == "CELL_TYPES") {
                    i += 2 + (size_t)std::stol(tok[i + 1]);
                } else if (key == "POINT_DATA") {
                    if (std::stoi(tok[i + 1]) != n) throw std::runtime_error("The number of points is greater than the amount of point data.");
                    i += 2;
                } else if (key == "NORMALS" || key == "VECTORS") {
                    out.addDescriptor(key == "NORMALS" ? "normals" : tok[i + 1], block(i + 3, 3));
                    i += 3 + 3 * (size_t)n;
                } else if (key == "SCALARS") {
                    const std::string name = tok[i + 1];
                    int comps = 1;
                    size_t j = i + 3;
                    if (isNumber(tok[j])) { comps = std::stoi(tok[j]); ++j; }
                    if (tok[j] == "LOOKUP_TABLE") j += 2;
                    out.addDescriptor(name, block(j, comps));
                    i = j + (size_t)comps * n;
                } else if (key == "COLOR_SCALARS") {
                    const int comps = std::stoi(tok[i + 2]);
                    out.addDescriptor("color", block(i + 3, comps));
                    i += 3 + (size_t)comps * n;
                } else {
                    throw std::runtime_error("VTK reader: unsupported section " + key);
                }
            }
            if (n < 0) throw std::runtime_error("VTK reader: no POINTS section");
            return out;
        }
        // keep the listed columns (ascending), in place: what the reference's filters do with
        // setColFrom + conservativeResize (e.g. RandomSampling.cpp:63-74)
        void keepColumns(const std::vector<int>& keep) {
            const int m = (int)keep.size();
            Matrix f(features.rows(), m);
            for (int j = 0; j < m; ++j)
                for (int i = 0; i < features.rows(); ++i) f(i, j) = features(i, keep[j]);
            features = f;
            if (descriptors.rows() > 0) {
                Matrix d(descriptors.rows(), m);
                for (int j = 0; j < m; ++j)
                    for (int i = 0; i < descriptors.rows(); ++i) d(i, j) = descriptors(i, keep[j]);
                descriptors = d;
            }
        }
        // DataPoints::concatenate (DataPoints.cpp:225-330): append dp's points; only the descriptors both clouds carry
        // (same name, same span) survive, in this cloud's order
        void concatenate(const DataPoints& dp) {
            const int n1 = features.cols(), n2 = dp.features.cols();
            if (features.rows() != dp.features.rows())
                throw InvalidField("Cannot concatenate DataPoints because the dimension of the features are not the same. Actual dimension: " +
                                   std::to_string(features.rows()) + " New dimension: " + std::to_string(dp.features.rows()));
            Labels kept;
            for (const Label& a : descriptorLabels)
                for (const Label& b : dp.descriptorLabels)
                    if (a.text == b.text) {
                        if (a.span != b.span)
                            throw InvalidField("The field " + a.text + " has dimension " + std::to_string(a.span) + " in this, different than dimension " +
                                               std::to_string(b.span) + " in that");
                        kept.push_back(a);
                        break;
                    }
            size_t rows = 0;
            for (const Label& l : kept) rows += l.span;
            Matrix f(features.rows(), n1 + n2), d((int)rows, rows ? n1 + n2 : 0);
            for (int j = 0; j < n1 + n2; ++j)
                for (int i = 0; i < features.rows(); ++i) f(i, j) = j < n1 ? features(i, j) : dp.features(i, j - n1);
            unsigned row = 0;
            for (const Label& l : kept) {
                const unsigned ra = getDescriptorStartingRow(l.text), rb = dp.getDescriptorStartingRow(l.text);
                for (int j = 0; j < n1 + n2; ++j)
                    for (unsigned i = 0; i < l.span; ++i) d(row + i, j) = j < n1 ? descriptors(ra + i, j) : dp.descriptors(rb + i, j - n1);
                row += l.span;
            }
            features = f;
            descriptors = d;
            descriptorLabels = kept;
        }
        void addDescriptor(const std::string& name, const Matrix& newDescriptor) {
            allocateDescriptor(name, newDescriptor.rows());
            const unsigned row = getDescriptorStartingRow(name);
            for (int j = 0; j < newDescriptor.cols(); ++j)
                for (int i = 0; i < newDescriptor.rows(); ++i) descriptors(row + i, j) = newDescriptor(i, j);
        }
    };

    // ---- Matches (PointMatcher.h:371-391, Matches.cpp) -----------------------------------------------
    struct Matches {
        typedef Matrix Dists;
        typedef IntMatrix Ids;
        static constexpr int InvalidId = -1;
        static T InvalidDist() { return std::numeric_limits<T>::infinity(); }
        Dists dists;  // knn x N squared distances
        Ids ids;      // knn x N reference columns
        Matches() {}
        Matches(const Dists& dists, const Ids ids) : dists(dists), ids(ids) {}
        Matches(const int knn, const int pointsCount) : dists(knn, pointsCount), ids(knn, pointsCount) {}
        // Matches.cpp:60-87 (host utility used outside the ICP loop, e.g. examples/icp_advance_api.cpp)
        T getDistsQuantile(const T quantile) const {
            std::vector<T> values;
            values.reserve(dists.size());
            for (size_t i = 0; i < dists.size(); ++i)
                if (dists(i) != std::numeric_limits<T>::infinity()) values.push_back(dists(i));
            if (values.size() == 0) throw ConvergenceError("no outlier to filter");
            if (quantile < 0.0 || quantile > 1.0) throw ConvergenceError("quantile must be between 0 and 1");
            if (quantile == 1.0) return *std::max_element(values.begin(), values.end());
            const size_t idx = std::min(values.size() - 1, (size_t)(T(values.size()) * quantile));
            std::nth_element(values.begin(), values.begin() + idx, values.end());
            return values[idx];
        }
    };

    // ---------------------------------------------------------------------------------------------
    // the device context shared by the modules of one chain
    // ---------------------------------------------------------------------------------------------
    struct GpuPipeline {
        pmgpu_ctx* ctx = nullptr;
        const void* readingKey = nullptr;  // host features pointer of the resident reading
        int readingCols = 0;
        explicit GpuPipeline(int device = 0) {
            const int rc = pmgpu_ctx_create(device, &ctx);
            if (rc != PMGPU_OK)
                throw std::runtime_error(std::string("GPU module: cannot create a pmgpu context (") + pmgpu_status_string(rc) +
                                         "); there is no CPU fallback");
        }
        ~GpuPipeline() { pmgpu_ctx_destroy(ctx); }
        GpuPipeline(const GpuPipeline&) = delete;
        GpuPipeline& operator=(const GpuPipeline&) = delete;
        // status -> the exception the reference throws at the same place (pmgpu.h)
        void check(int rc) const {
            if (rc == PMGPU_OK) return;
            const std::string msg = pmgpu_last_error(ctx);
            switch (rc) {
                case PMGPU_ERR_UNSUPPORTED: throw ConfigurationError(msg);
                case PMGPU_ERR_NO_OUTLIER_TO_FILTER:
                case PMGPU_ERR_BAD_QUANTILE:
                case PMGPU_ERR_NO_POINT_TO_MINIMIZE:
                case PMGPU_ERR_NAN: throw ConvergenceError(msg);
                case PMGPU_ERR_NO_NORMALS: throw typename DataPoints::InvalidField(msg);
                case PMGPU_ERR_NOT_ORTHOGONAL: throw TransformationError(msg);
                case PMGPU_ERR_BAD_ARG: throw InvalidParameter(msg);
                default: throw std::runtime_error(msg + " (pmgpu status " + std::to_string(rc) + ")");
            }
        }
        // modules created outside an ICP object share one pipeline per thread, so the stand-alone
        // sequence init -> findClosests -> outlierFilters.compute -> errorMinimizer.compute
        // (examples/icp_advance_api.cpp:164-217) works on resident data
        static std::shared_ptr<GpuPipeline> threadDefault() {
            static thread_local std::shared_ptr<GpuPipeline> p;
            if (!p) p = std::make_shared<GpuPipeline>(0);
            return p;
        }
    };
    struct GpuBound {
        std::shared_ptr<GpuPipeline> pipeline;
        virtual ~GpuBound() {}
        void bind(const std::shared_ptr<GpuPipeline>& p) { pipeline = p; }
        GpuPipeline& gpu() {
            if (!pipeline) pipeline = GpuPipeline::threadDefault();
            return *pipeline;
        }
    };
    // the GPU path computes in float on 3-D clouds (features.rows() == 4) and 2-D clouds (features.rows() == 3)
    static void requireFloat3D(int rows, const char* who) {
        if (!std::is_same<T, float>::value || (rows != 4 && rows != 3))
            throw ConfigurationError(std::string(who) + ": GPU module: only float clouds of 2 or 3 dimensions are supported");
    }

    // ---------------------------------------------------------------------------------------------
    // module interfaces (PointMatcher.h:404-641)
    // ---------------------------------------------------------------------------------------------
    struct Transformation : public Parametrizable {
        Transformation() {}
        Transformation(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~Transformation() {}
        virtual DataPoints compute(const DataPoints& input, const TransformationParameters& parameters) const = 0;
        virtual bool checkParameters(const TransformationParameters& parameters) const = 0;
        virtual TransformationParameters correctParameters(const TransformationParameters& parameters) const = 0;
    };
    struct Transformations : public std::vector<std::shared_ptr<Transformation>> {
        void apply(DataPoints& cloud, const TransformationParameters& parameters) const {
            DataPoints transformedCloud;
            for (const auto& t : *this) {
                transformedCloud = t->compute(cloud, parameters);
                std::swap(cloud, transformedCloud);
            }
        }
    };
    DEF_REGISTRAR(Transformation)

    struct DataPointsFilter : public Parametrizable {
        DataPointsFilter() {}
        DataPointsFilter(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~DataPointsFilter() {}
        virtual void init() {}
        virtual DataPoints filter(const DataPoints& input) = 0;
        virtual void inPlaceFilter(DataPoints& cloud) = 0;
    };
    struct DataPointsFilters : public std::vector<std::shared_ptr<DataPointsFilter>> {
        void init() { for (auto& f : *this) f->init(); }
        void apply(DataPoints& cloud) {
            for (auto& f : *this) f->inPlaceFilter(cloud);  // DataPointsFilter.cpp:106-131
        }
    };
    DEF_REGISTRAR(DataPointsFilter)

    struct Matcher : public Parametrizable {
        unsigned long visitCounter;
        Matcher() : visitCounter(0) {}
        Matcher(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params), visitCounter(0) {}
        virtual ~Matcher() {}
        void resetVisitCount() { visitCounter = 0; }
        unsigned long getVisitCount() const { return visitCounter; }
        virtual void init(const DataPoints& filteredReference) = 0;
        virtual Matches findClosests(const DataPoints& filteredReading) = 0;
    };
    DEF_REGISTRAR(Matcher)

    struct OutlierFilter : public Parametrizable {
        OutlierFilter() {}
        OutlierFilter(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~OutlierFilter() {}
        virtual OutlierWeights compute(const DataPoints& filteredReading, const DataPoints& filteredReference, const Matches& input) = 0;
    };
    struct OutlierFilters;  // below, needs the GPU filter base
    DEF_REGISTRAR(OutlierFilter)

    struct ErrorMinimizer : public Parametrizable {
        // The GPU minimisers never need the compacted clouds; they are built on request only (getErrorElements, getOverlap,
        // getResidualError), from the matches resident on the device (pmgpu_matches_get).
        struct ErrorElements {
            DataPoints reading, reference;
            OutlierWeights weights;
            Matches matches;
            int nbRejectedMatches = -1, nbRejectedPoints = -1;
            T pointUsedRatio = T(-1), weightedPointUsedRatio = T(-1);
            ErrorElements() {}
            // ErrorMinimizer.cpp:58-193: the kept (reading point, match) pairs in reading order, k innermost
            ErrorElements(const DataPoints& requestedPts, const DataPoints& sourcePts, const OutlierWeights& outlierWeights, const Matches& matchesIn) {
                const int knn = outlierWeights.rows(), n = requestedPts.features.cols();
                int pointsCount = 0;
                for (size_t i = 0; i < outlierWeights.size(); ++i) pointsCount += outlierWeights(i) != T(0) ? 1 : 0;
                if (pointsCount == 0) throw ConvergenceError("ErrorMnimizer: no point to minimize");
                std::vector<int> keptPoint, keptId;
                std::vector<T> keptDist, keptWeight;
                int rejectedMatchCount = 0, rejectedPointCount = 0;
                T weightSum = 0;
                for (int i = 0; i < n; ++i) {
                    bool matchExist = false;
                    for (int k = 0; k < knn; ++k) {
                        const T matchDist = matchesIn.dists(k, i);
                        if (matchDist == Matches::InvalidDist()) continue;
                        if (outlierWeights(k, i) != T(0)) {
                            keptPoint.push_back(i);
                            keptId.push_back(matchesIn.ids(k, i));
                            keptDist.push_back(matchDist);
                            keptWeight.push_back(outlierWeights(k, i));
                            weightSum += outlierWeights(k, i);
                            matchExist = true;
                        } else
                            ++rejectedMatchCount;
                    }
                    if (!matchExist) ++rejectedPointCount;
                }
                const int m = (int)keptPoint.size();
                pointUsedRatio = T(m) / T(knn * n);
                weightedPointUsedRatio = weightSum / T(knn * n);
                reading = requestedPts;
                reading.keepColumns(keptPoint);
                reference = sourcePts;
                reference.keepColumns(keptId);
                weights = OutlierWeights(1, m);
                matches = Matches(1, m);
                for (int j = 0; j < m; ++j) {
                    weights(0, j) = keptWeight[j];
                    matches.ids(0, j) = keptId[j];
                    matches.dists(0, j) = keptDist[j];
                }
                nbRejectedMatches = rejectedMatchCount;
                nbRejectedPoints = rejectedPointCount;
            }
        };
        ErrorMinimizer() {}
        ErrorMinimizer(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~ErrorMinimizer() {}
        T getPointUsedRatio() const { return lastErrorElements.pointUsedRatio; }
        T getWeightedPointUsedRatio() const { return lastErrorElements.weightedPointUsedRatio; }
        // ErrorMinimizer.cpp:232-236 — materialised on request: nothing is copied back from the device unless this is called
        ErrorElements getErrorElements() const {
            if (materialize) materialize(lastErrorElements);
            return lastErrorElements;
        }
        virtual T getOverlap() const { return lastErrorElements.weightedPointUsedRatio; }
        virtual T getResidualError(const DataPoints&, const DataPoints&, const OutlierWeights&, const Matches&) const {
            throw std::runtime_error("You must implement getResidualError() in your ErrorMinimizer");  // ErrorMinimizer.cpp:205-210
        }
        // set by ICP after a registration: fills lastErrorElements from the resident matches
        std::function<void(ErrorElements&)> materialize;
        virtual Matrix getCovariance() const { return Matrix::Zero(6, 6); }
        virtual TransformationParameters compute(const DataPoints& filteredReading, const DataPoints& filteredReference,
                                                 const OutlierWeights& outlierWeights, const Matches& matches) = 0;
    protected:
        mutable ErrorElements lastErrorElements;
    };
    DEF_REGISTRAR(ErrorMinimizer)

    struct TransformationChecker : public Parametrizable {
    protected:
        typedef std::vector<std::string> StringVector;
        Vector limits, conditionVariables;
        StringVector limitNames, conditionVariableNames;
    public:
        TransformationChecker() {}
        TransformationChecker(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~TransformationChecker() {}
        virtual void init(const TransformationParameters& parameters, bool& iterate) = 0;
        virtual void check(const TransformationParameters& parameters, bool& iterate) = 0;
        const Vector& getLimits() const { return limits; }
        const Vector& getConditionVariables() const { return conditionVariables; }
        const StringVector& getLimitNames() const { return limitNames; }
        const StringVector& getConditionVariableNames() const { return conditionVariableNames; }
    };
    struct TransformationCheckers : public std::vector<std::shared_ptr<TransformationChecker>> {
        void init(const TransformationParameters& parameters, bool& iterate) { for (auto& c : *this) c->init(parameters, iterate); }
        void check(const TransformationParameters& parameters, bool& iterate) { for (auto& c : *this) c->check(parameters, iterate); }
    };
    DEF_REGISTRAR(TransformationChecker)

    struct Inspector : public Parametrizable {
        Inspector() {}
        Inspector(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~Inspector() {}
        virtual void init() {}
        virtual void addStat(const std::string&, double) {}
        virtual void dumpIteration(const size_t, const TransformationParameters&, const DataPoints&, const DataPoints&, const Matches&,
                                   const OutlierWeights&, const TransformationCheckers&) {}
        virtual void finish(const size_t) {}
        virtual bool isNull() const { return false; }
        // does dumpIteration read its arguments?  (host copies of matches / weights / clouds are made only if it does)
        virtual bool needsIterationData() const { return !isNull(); }
    };
    DEF_REGISTRAR(Inspector)

    struct Logger : public Parametrizable {
        Logger() {}
        Logger(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params) : Parametrizable(className, paramsDoc, params) {}
        virtual ~Logger() {}
    };
    DEF_REGISTRAR(Logger)

#include "Modules.inl"

    // ---------------------------------------------------------------------------------------------
    // ICP (PointMatcher.h:652-764, ICP.cpp)
    // ---------------------------------------------------------------------------------------------
#include "ICP.inl"

    // registry (Registry.cpp:60-126): the GPU classes under the reference's names
    PointMatcher() {
        ADD_TO_REGISTRAR_NO_PARAM(Transformation, RigidTransformation, RigidTransformation)
        ADD_TO_REGISTRAR_NO_PARAM(DataPointsFilter, IdentityDataPointsFilter, IdentityDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, SurfaceNormalDataPointsFilter, SurfaceNormalDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, RandomSamplingDataPointsFilter, RandomSamplingDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, SamplingSurfaceNormalDataPointsFilter, SamplingSurfaceNormalDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, MinDistDataPointsFilter, MinDistDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, MaxDistDataPointsFilter, MaxDistDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, ObservationDirectionDataPointsFilter, ObservationDirectionDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, OrientNormalsDataPointsFilter, OrientNormalsDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, BoundingBoxDataPointsFilter, BoundingBoxDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, DistanceLimitDataPointsFilter, DistanceLimitDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, FixStepSamplingDataPointsFilter, FixStepSamplingDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, MaxPointCountDataPointsFilter, MaxPointCountDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, MaxQuantileOnAxisDataPointsFilter, MaxQuantileOnAxisDataPointsFilter)
        ADD_TO_REGISTRAR_NO_PARAM(DataPointsFilter, RemoveNaNDataPointsFilter, RemoveNaNDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, MaxDensityDataPointsFilter, MaxDensityDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, ShadowDataPointsFilter, ShadowDataPointsFilter)
        ADD_TO_REGISTRAR(DataPointsFilter, SimpleSensorNoiseDataPointsFilter, SimpleSensorNoiseDataPointsFilter)
        ADD_TO_REGISTRAR(Matcher, KDTreeMatcher, KDTreeMatcher)
        ADD_TO_REGISTRAR(Matcher, KDTreeVarDistMatcher, KDTreeVarDistMatcher)
        ADD_TO_REGISTRAR_NO_PARAM(OutlierFilter, NullOutlierFilter, NullOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, MaxDistOutlierFilter, MaxDistOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, MinDistOutlierFilter, MinDistOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, MedianDistOutlierFilter, MedianDistOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, TrimmedDistOutlierFilter, TrimmedDistOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, VarTrimmedDistOutlierFilter, VarTrimmedDistOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, RobustOutlierFilter, RobustOutlierFilter)
        ADD_TO_REGISTRAR(OutlierFilter, SurfaceNormalOutlierFilter, SurfaceNormalOutlierFilter)
        ADD_TO_REGISTRAR_NO_PARAM(ErrorMinimizer, PointToPointErrorMinimizer, PointToPointErrorMinimizer)
        ADD_TO_REGISTRAR(ErrorMinimizer, PointToPointWithCovErrorMinimizer, PointToPointWithCovErrorMinimizer)
        ADD_TO_REGISTRAR_NO_PARAM(ErrorMinimizer, PointToPointSimilarityErrorMinimizer, PointToPointSimilarityErrorMinimizer)
        ADD_TO_REGISTRAR(ErrorMinimizer, PointToPlaneErrorMinimizer, PointToPlaneErrorMinimizer)
        ADD_TO_REGISTRAR(ErrorMinimizer, PointToPlaneWithCovErrorMinimizer, PointToPlaneWithCovErrorMinimizer)
        ADD_TO_REGISTRAR(TransformationChecker, CounterTransformationChecker, CounterTransformationChecker)
        ADD_TO_REGISTRAR(TransformationChecker, DifferentialTransformationChecker, DifferentialTransformationChecker)
        ADD_TO_REGISTRAR(TransformationChecker, BoundTransformationChecker, BoundTransformationChecker)
        ADD_TO_REGISTRAR_NO_PARAM(Inspector, NullInspector, NullInspector)
        ADD_TO_REGISTRAR(Inspector, PerformanceInspector, PerformanceInspector)
        ADD_TO_REGISTRAR(Inspector, VTKFileInspector, VTKFileInspector)
        ADD_TO_REGISTRAR_NO_PARAM(Logger, NullLogger, NullLogger)
    }
    static const PointMatcher& get() {  // Registry.cpp:142-146
        static const PointMatcher pm;
        return pm;
    }
};
