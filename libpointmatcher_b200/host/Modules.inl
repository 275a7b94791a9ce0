// host/Modules.inl — module implementations, included inside `struct PointMatcher<T>`.
// Each class mirrors the reference class of the same name (file:line cited); the work is done by
// the C ABI (include/pmgpu.h) on the pipeline the module is bound to.

// ---- RigidTransformation (TransformationsImpl.{h,cpp}:49-151) ------------------------------------
// Host-side utility with the reference's float operation order.  Inside ICP the transform is fused
// into the kNN kernel (K2) and this class is only used for the 4x4 checks.
struct RigidTransformation : public Transformation {
    static const std::string description() { return "Rigid transformation."; }
    RigidTransformation() : Transformation("RigidTransformation", ParametersDoc(), Parameters()) {}
    static T det3(const TransformationParameters& p) {
        return p(0, 0) * (p(1, 1) * p(2, 2) - p(1, 2) * p(2, 1)) - p(0, 1) * (p(1, 0) * p(2, 2) - p(1, 2) * p(2, 0)) +
               p(0, 2) * (p(1, 0) * p(2, 1) - p(1, 1) * p(2, 0));
    }
    bool checkParameters(const TransformationParameters& parameters) const override {
        const T epsilon = T(0.001);
        if (parameters.rows() == 4) return !(std::fabs(T(1) - det3(parameters)) > epsilon);
        const T det2 = parameters(0, 0) * parameters(1, 1) - parameters(0, 1) * parameters(1, 0);
        return !(std::fabs(T(1) - det2) > epsilon);
    }
    DataPoints compute(const DataPoints& input, const TransformationParameters& parameters) const override {
        if (!checkParameters(parameters)) throw TransformationError("RigidTransformation: Error, rotation matrix is not orthogonal.");
        return apply(input, parameters);
    }
    static DataPoints apply(const DataPoints& input, const TransformationParameters& parameters) {
        DataPoints out = input;
        out.features = parameters * input.features;
        // rotate the descriptors named normals / observationDirections (TransformationsImpl.cpp:72-84)
        const int dim = parameters.rows() - 1;
        unsigned row = 0;
        for (const auto& label : input.descriptorLabels) {
            if ((label.text == "normals" || label.text == "observationDirections") && (int)label.span == dim) {
                for (int j = 0; j < input.descriptors.cols(); ++j)
                    for (int i = 0; i < dim; ++i) {
                        volatile T acc = parameters(i, 0) * input.descriptors(row, j);
                        for (int k = 1; k < dim; ++k) {
                            volatile T prod = parameters(i, k) * input.descriptors(row + k, j);
                            acc = acc + prod;
                        }
                        out.descriptors(row + i, j) = acc;
                    }
            }
            row += label.span;
        }
        return out;
    }
    TransformationParameters correctParameters(const TransformationParameters& parameters) const override {
        TransformationParameters ortho = parameters;
        if (ortho.cols() != 4) return ortho;
        auto col = [&](int c, T* v) { for (int i = 0; i < 3; ++i) v[i] = parameters(i, c); };
        auto normalize = [](T* v) { const T n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int i = 0; i < 3; ++i) v[i] /= n; };
        auto cross = [](const T* a, const T* b, T* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
        T c1[3], c2[3], n0[3], n1[3];
        col(1, c1); col(2, c2);
        normalize(c1); normalize(c2);
        cross(c1, c2, n0);
        cross(c2, n0, n1);
        for (int i = 0; i < 3; ++i) { ortho(i, 0) = n0[i]; ortho(i, 1) = n1[i]; ortho(i, 2) = c2[i]; }
        return ortho;
    }
};

// ---- IdentityDataPointsFilter (DataPointsFilters/Identity.{h,cpp}) ----------------------------------
struct IdentityDataPointsFilter : public DataPointsFilter {
    static const std::string description() { return "Does nothing."; }
    IdentityDataPointsFilter() : DataPointsFilter("IdentityDataPointsFilter", ParametersDoc(), Parameters()) {}
    DataPoints filter(const DataPoints& input) override { return input; }
    void inPlaceFilter(DataPoints&) override {}
};

// ---- host-side pre-filters of the default chain -----------------------------------
// CPU, once per cloud, like the reference; RandomSampling and SamplingSurfaceNormal run in
// csrc/host_filters.cu behind the C ABI (same std::rand stream and std::nth_element as the reference).
struct RandomSamplingDataPointsFilter : public DataPointsFilter {
    static const std::string description() { return "Subsampling. This filter reduces the size of the point cloud by randomly dropping points. Based on \\cite{Masuda1996Random}"; }
    static const ParametersDoc availableParameters() {
        return {{"prob", "probability to keep a point, one over decimation factor ", "0.75", "0", "1", &Parametrizable::Comp<T>}};
    }
    const double prob;
    RandomSamplingDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("RandomSamplingDataPointsFilter", availableParameters(), params), prob(Parametrizable::get<T>("prob")) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {  // RandomSampling.cpp:58-75
        const int n = cloud.features.cols();
        std::vector<int32_t> keep(n > 0 ? n : 1);
        const int m = pmgpu_host_random_sampling(n, (float)prob, keep.data());
        cloud.keepColumns(std::vector<int>(keep.begin(), keep.begin() + m));
    }
};

// MinDist.{h,cpp} / MaxDist.{h,cpp}: keep the points beyond / within a radius or an axis value
template <bool KEEP_BEYOND>
struct AxisThresholdDataPointsFilter : public DataPointsFilter {
    const int dim;
    const T limit;
    AxisThresholdDataPointsFilter(const std::string& name, const ParametersDoc& doc, const Parameters& params, const char* limitName)
        : DataPointsFilter(name, doc, params), dim(Parametrizable::get<int>("dim")), limit(Parametrizable::get<T>(limitName)) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {
        const int rows = cloud.features.rows(), n = cloud.features.cols();
        if (dim >= rows - 1)
            throw InvalidParameter(this->className + ": Error, filtering on dimension number " + std::to_string(dim) + ", larger than feature dimensionality " +
                                   std::to_string(rows - 2));
        std::vector<int> keep;
        const T absLimit = limit < 0 ? -limit : limit;
        for (int i = 0; i < n; ++i) {
            T v, lim;
            if (dim == -1) {  // Euclidean norm of the point
                T acc = 0;
                for (int r = 0; r < rows - 1; ++r) acc += cloud.features(r, i) * cloud.features(r, i);
                v = std::sqrt(acc);
                lim = absLimit;
            } else {
                v = cloud.features(dim, i);
                lim = limit;
            }
            if (KEEP_BEYOND ? (v > lim) : (v < lim)) keep.push_back(i);
        }
        cloud.keepColumns(keep);
    }
};
struct MinDistDataPointsFilter : public AxisThresholdDataPointsFilter<true> {
    static const std::string description() { return "Subsampling. Filter points before a minimum distance measured on a specific axis. If dim is set to -1, points are filtered based on a minimum radius."; }
    static const ParametersDoc availableParameters() {
        return {{"dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", &Parametrizable::Comp<int>},
                {"minDist", "minimum value authorized. If dim is set to -1 (radius), the absolute value of minDist will be used. All points before that will be filtered.", "1", "-inf", "inf", &Parametrizable::Comp<T>}};
    }
    MinDistDataPointsFilter(const Parameters& params = Parameters())
        : AxisThresholdDataPointsFilter<true>("MinDistDataPointsFilter", availableParameters(), params, "minDist") {}
};
struct MaxDistDataPointsFilter : public AxisThresholdDataPointsFilter<false> {
    static const std::string description() { return "Subsampling. Filter points beyond a maximum distance measured on a specific axis. If dim is set to -1, points are filtered based on a maximum radius."; }
    static const ParametersDoc availableParameters() {
        return {{"dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", &Parametrizable::Comp<int>},
                {"maxDist", "maximum distance authorized. If dim is set to -1 (radius), the absolute value of minDist will be used. All points beyond that will be filtered.", "1", "-inf", "inf", &Parametrizable::Comp<T>}};
    }
    MaxDistDataPointsFilter(const Parameters& params = Parameters())
        : AxisThresholdDataPointsFilter<false>("MaxDistDataPointsFilter", availableParameters(), params, "maxDist") {}
};

// ObservationDirection.{h,cpp}: descriptor observationDirections = sensor centre - point
struct ObservationDirectionDataPointsFilter : public DataPointsFilter {
    static const std::string description() { return "Observation direction. This filter extracts observation directions (vector from point to sensor), considering a sensor at position (x,y,z)."; }
    static const ParametersDoc availableParameters() {
        return {{"x", "x-coordinate of sensor", "0"}, {"y", "y-coordinate of sensor", "0"}, {"z", "z-coordinate of sensor", "0"}};
    }
    const T centerX, centerY, centerZ;
    ObservationDirectionDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("ObservationDirectionDataPointsFilter", availableParameters(), params), centerX(Parametrizable::get<T>("x")),
          centerY(Parametrizable::get<T>("y")), centerZ(Parametrizable::get<T>("z")) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {
        const int dim = cloud.features.rows() - 1, n = cloud.features.cols();
        if (dim != 2 && dim != 3)
            throw typename DataPoints::InvalidField("ObservationDirectionDataPointsFilter: Error, works only in 2 or 3 dimensions, cloud has " +
                                                    std::to_string(dim) + " dimensions.");
        const T center[3] = {centerX, centerY, centerZ};
        cloud.allocateDescriptor("observationDirections", dim);
        const unsigned row = cloud.getDescriptorStartingRow("observationDirections");
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < dim; ++r) cloud.descriptors(row + r, i) = center[r] - cloud.features(r, i);
    }
};

// OrientNormals.{h,cpp}: flip the normals toward (or away from) the observation point
struct OrientNormalsDataPointsFilter : public DataPointsFilter {
    static const std::string description() { return "Normals. Reorient normals so that they all point in the same direction, with respect to the observation points."; }
    static const ParametersDoc availableParameters() {
        return {{"towardCenter", "If set to true(1), all the normals will point inside the surface (i.e. toward the observation points).", "1", "0", "1", &Parametrizable::Comp<bool>}};
    }
    const bool towardCenter;
    OrientNormalsDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("OrientNormalsDataPointsFilter", availableParameters(), params), towardCenter(Parametrizable::get<bool>("towardCenter")) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {
        if (!cloud.descriptorExists("normals")) throw typename DataPoints::InvalidField("OrientNormalsDataPointsFilter: Error, cannot find normals in descriptors.");
        if (!cloud.descriptorExists("observationDirections"))
            throw typename DataPoints::InvalidField("OrientNormalsDataPointsFilter: Error, cannot find observation directions in descriptors.");
        const unsigned rn = cloud.getDescriptorStartingRow("normals"), ro = cloud.getDescriptorStartingRow("observationDirections");
        const int dim = cloud.getDescriptorDimension("normals"), n = cloud.features.cols();
        for (int i = 0; i < n; ++i) {
            T acc = 0;
            for (int r = 0; r < dim; ++r) acc += cloud.descriptors(ro + r, i) * cloud.descriptors(rn + r, i);
            const double scalar = acc;
            if (towardCenter ? (scalar < 0) : (scalar > 0))
                for (int r = 0; r < dim; ++r) cloud.descriptors(rn + r, i) = -cloud.descriptors(rn + r, i);
        }
    }
};

// SamplingSurfaceNormal.{h,cpp}: kd-split bins of <= knn points, one normal per bin
struct SamplingSurfaceNormalDataPointsFilter : public DataPointsFilter {
    static const std::string description() {
        return "Subsampling, Normals. This filter decomposes the point-cloud space in boxes, by recursively splitting the cloud through axis-aligned "
               "hyperplanes such as to maximize the evenness of the aspect ratio of the box. When the number of points in a box reaches a value knn or "
               "lower, the filter computes the center of mass of these points and its normal by taking the eigenvector corresponding to the smallest "
               "eigenvalue of all points in the box.";
    }
    static const ParametersDoc availableParameters() {
        return {
            {"ratio", "ratio of points to keep with random subsampling. Matrix (normal, density, etc.) will be associated to all points in the same bin.", "0.5", "0.0000001", "1.0", &Parametrizable::Comp<T>},
            {"knn", "determined how many points are used to compute the normals. Direct link with the rapidity of the computation (large = fast). Technically, limit over which a box is splitted in two", "7", "3", "2147483647", &Parametrizable::Comp<unsigned>},
            {"samplingMethod", "if set to 0, random subsampling using the parameter ratio. If set to 1, bin subsampling with the resulting number of points being 1/knn.", "0", "0", "1", &Parametrizable::Comp<unsigned>},
            {"maxBoxDim", "maximum length of a box above which the box is discarded", "inf"},
            {"averageExistingDescriptors", "whether the filter keep the existing point descriptors and average them or should it drop them", "1"},
            {"keepNormals", "whether the normals should be added as descriptors to the resulting cloud", "1"},
            {"keepDensities", "whether the point densities should be added as descriptors to the resulting cloud", "0"},
            {"keepEigenValues", "whether the eigen values should be added as descriptors to the resulting cloud", "0"},
            {"keepEigenVectors", "whether the eigen vectors should be added as descriptors to the resulting cloud", "0"}};
    }
    const T ratio;
    const unsigned knn, samplingMethod;
    const T maxBoxDim;
    const bool averageExistingDescriptors, keepNormals, keepDensities, keepEigenValues, keepEigenVectors;
    int unfitPointsCount = 0;
    SamplingSurfaceNormalDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("SamplingSurfaceNormalDataPointsFilter", availableParameters(), params),
          ratio(Parametrizable::get<T>("ratio")), knn(Parametrizable::get<unsigned>("knn")), samplingMethod(Parametrizable::get<unsigned>("samplingMethod")),
          maxBoxDim(Parametrizable::get<T>("maxBoxDim")), averageExistingDescriptors(Parametrizable::get<bool>("averageExistingDescriptors")),
          keepNormals(Parametrizable::get<bool>("keepNormals")), keepDensities(Parametrizable::get<bool>("keepDensities")),
          keepEigenValues(Parametrizable::get<bool>("keepEigenValues")), keepEigenVectors(Parametrizable::get<bool>("keepEigenVectors")) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {  // SamplingSurfaceNormal.cpp:80-170
        requireFloat3D(cloud.features.rows(), "SamplingSurfaceNormalDataPointsFilter");
        const int n = cloud.features.cols();
        if (averageExistingDescriptors) {
            unsigned insertDim = 0;
            for (const auto& l : cloud.descriptorLabels) insertDim += l.span;
            if (insertDim != cloud.getDescriptorDim())
                throw typename DataPoints::InvalidField("SamplingSurfaceNormalDataPointsFilter: Error, descriptor labels do not match descriptor data");
        }
        const int oldDescRows = cloud.descriptors.rows();
        const int dn = cloud.features.rows() - 1;  // spans follow the cloud's dimension (SamplingSurfaceNormal.cpp:100-112)
        if (keepNormals) cloud.allocateDescriptor("normals", dn);
        if (keepDensities) cloud.allocateDescriptor("densities", 1);
        if (keepEigenValues) cloud.allocateDescriptor("eigValues", dn);
        if (keepEigenVectors) cloud.allocateDescriptor("eigVectors", dn * dn);
        std::vector<float> normals(keepNormals ? dn * (size_t)n : 0), densities(keepDensities ? (size_t)n : 0), eigVa(keepEigenValues ? dn * (size_t)n : 0),
            eigVe(keepEigenVectors ? dn * dn * (size_t)n : 0);
        // the existing descriptors as their own (rows x n) block for the averaging of bin sampling
        std::vector<float> oldDesc((size_t)oldDescRows * n);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < oldDescRows; ++i) oldDesc[(size_t)j * oldDescRows + i] = (float)cloud.descriptors(i, j);
        std::vector<int32_t> keep(n > 0 ? n : 1);
        const int flags = (keepNormals ? PMGPU_KEEP_NORMALS : 0) | (keepDensities ? PMGPU_KEEP_DENSITIES : 0) | (keepEigenValues ? PMGPU_KEEP_EIGEN_VALUES : 0) |
                          (keepEigenVectors ? PMGPU_KEEP_EIGEN_VECTORS : 0);
        const int m = pmgpu_host_sampling_surface_normal(reinterpret_cast<float*>(cloud.features.data()), cloud.features.rows(), n,
                                                         oldDescRows ? oldDesc.data() : nullptr, oldDescRows, (float)ratio, (int)knn, (int)samplingMethod,
                                                         (float)maxBoxDim, averageExistingDescriptors ? 1 : 0, flags, keep.data(), normals.data(),
                                                         densities.data(), eigVa.data(), eigVe.data(), &unfitPointsCount);
        if (m < 0) throw std::runtime_error("SamplingSurfaceNormalDataPointsFilter: bad argument");
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < oldDescRows; ++i) cloud.descriptors(i, j) = T(oldDesc[(size_t)j * oldDescRows + i]);
        auto put = [&](const char* name, const std::vector<float>& src, int span) {
            const unsigned row = cloud.getDescriptorStartingRow(name);
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < span; ++i) cloud.descriptors(row + i, j) = T(src[(size_t)j * span + i]);
        };
        if (keepNormals) put("normals", normals, dn);
        if (keepDensities) put("densities", densities, 1);
        if (keepEigenValues) put("eigValues", eigVa, dn);
        if (keepEigenVectors) put("eigVectors", eigVe, dn * dn);
        cloud.keepColumns(std::vector<int>(keep.begin(), keep.begin() + m));
    }
};

// ---- the remaining per-cloud host filters of the reference's golden chain files --------------------
// (examples/data/icp_data/default*DataPointsFilter.yaml): O(N) host passes, once per cloud, as in the
// reference.  The std::rand() consumers go through pmgpu_host_* so both mirrors draw from one place.
#define PM_FILTER_COPY_THEN_IN_PLACE                  \
    DataPoints filter(const DataPoints& input) override { \
        DataPoints output(input);                     \
        inPlaceFilter(output);                        \
        return output;                                \
    }

struct BoundingBoxDataPointsFilter : public DataPointsFilter {  // BoundingBox.{h,cpp}
    static const std::string description() { return "Subsampling. Remove points laying in a bounding box which is axis aligned."; }
    static const ParametersDoc availableParameters() {
        return {{"xMin", "minimum value on x-axis defining one side of the bounding box", "-1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"xMax", "maximum value on x-axis defining one side of the bounding box", "1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"yMin", "minimum value on y-axis defining one side of the bounding box", "-1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"yMax", "maximum value on y-axis defining one side of the bounding box", "1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"zMin", "minimum value on z-axis defining one side of the bounding box", "-1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"zMax", "maximum value on z-axis defining one side of the bounding box", "1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"removeInside", "If set to true (1), remove points inside the bounding box; else (0), remove points outside the bounding box", "1", "0", "1", &Parametrizable::Comp<bool>}};
    }
    const T xMin, xMax, yMin, yMax, zMin, zMax;
    const bool removeInside;
    BoundingBoxDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("BoundingBoxDataPointsFilter", availableParameters(), params), xMin(Parametrizable::get<T>("xMin")),
          xMax(Parametrizable::get<T>("xMax")), yMin(Parametrizable::get<T>("yMin")), yMax(Parametrizable::get<T>("yMax")),
          zMin(Parametrizable::get<T>("zMin")), zMax(Parametrizable::get<T>("zMax")), removeInside(Parametrizable::get<bool>("removeInside")) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // BoundingBox.cpp:71-103
        const int n = cloud.features.cols(), rows = cloud.features.rows();
        std::vector<int> keep;
        for (int i = 0; i < n; ++i) {
            const bool x_in = cloud.features(0, i) > xMin && cloud.features(0, i) < xMax;
            const bool y_in = cloud.features(1, i) > yMin && cloud.features(1, i) < yMax;
            const bool z_in = rows == 3 || (cloud.features(2, i) > zMin && cloud.features(2, i) < zMax);
            const bool in_box = x_in && y_in && z_in;
            if (removeInside ? !in_box : in_box) keep.push_back(i);
        }
        cloud.keepColumns(keep);
    }
};

struct DistanceLimitDataPointsFilter : public DataPointsFilter {  // DistanceLimit.{h,cpp}
    static const std::string description() { return "Subsampling. Filter points based on distance measured on a specific axis. If dim is set to -1, points are filtered based on radius."; }
    static const ParametersDoc availableParameters() {
        return {{"dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", &Parametrizable::Comp<int>},
                {"dist", "distance limit of the filter. If dim is set to -1 (radius), the absolute value of dist will be used", "1", "-inf", "inf", &Parametrizable::Comp<T>},
                {"removeInside", "If set to true (1), remove points before the distance limit; else (0), remove points beyond the distance limit", "1", "0", "1", &Parametrizable::Comp<bool>}};
    }
    const int dim;
    const T dist;
    const bool removeInside;
    DistanceLimitDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("DistanceLimitDataPointsFilter", availableParameters(), params), dim(Parametrizable::get<int>("dim")),
          dist(Parametrizable::get<T>("dist")), removeInside(Parametrizable::get<bool>("removeInside")) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // DistanceLimit.cpp:66-127
        const int n = cloud.features.cols(), rows = cloud.features.rows();
        if (dim >= rows - 1)
            throw InvalidParameter("DistanceLimitDataPointsFilter: Error, filtering on dimension number " + std::to_string(dim) +
                                   ", larger than authorized axis id " + std::to_string(rows - 2));
        std::vector<int> keep;
        const T absDist = dist < 0 ? -dist : dist;
        for (int i = 0; i < n; ++i) {
            T v, lim;
            if (dim == -1) {
                T acc = 0;
                for (int r = 0; r < rows - 1; ++r) acc += cloud.features(r, i) * cloud.features(r, i);
                v = std::sqrt(acc);
                lim = absDist;
            } else {
                v = cloud.features(dim, i);
                lim = dist;
            }
            if (removeInside ? (v > lim) : (v < lim)) keep.push_back(i);
        }
        cloud.keepColumns(keep);
    }
};

struct FixStepSamplingDataPointsFilter : public DataPointsFilter {  // FixStepSampling.{h,cpp}
    static const std::string description() { return "Subsampling. This filter reduces the size of the point cloud by only keeping one point over step ones; with step varying in time from startStep to endStep, each iteration getting multiplied by stepMult. If use as prefilter (i.e. before the iterations), only startStep is used."; }
    static const ParametersDoc availableParameters() {
        return {{"startStep", "initial number of point to skip (initial decimation factor)", "10", "1", "2147483647", &Parametrizable::Comp<unsigned>},
                {"endStep", "maximal or minimal number of points to skip (final decimation factor)", "10", "1", "2147483647", &Parametrizable::Comp<unsigned>},
                {"stepMult", "multiplication factor to compute the new decimation factor for each iteration", "1", "0.0000001", "inf", &Parametrizable::Comp<double>}};
    }
    const unsigned startStep, endStep;
    const double stepMult;
    double step;
    FixStepSamplingDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("FixStepSamplingDataPointsFilter", availableParameters(), params), startStep(Parametrizable::get<unsigned>("startStep")),
          endStep(Parametrizable::get<unsigned>("endStep")), stepMult(Parametrizable::get<double>("stepMult")), step(startStep) {}
    void init() override { step = startStep; }
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // FixStepSampling.cpp:76-110
        const int iStep(step);
        const int n = cloud.features.cols();
        const int phase(pmgpu_host_rand() % iStep);
        std::vector<int> keep;
        for (int i = phase; i < n; i += iStep) keep.push_back(i);
        cloud.keepColumns(keep);
        const double deltaStep(startStep * stepMult - startStep);
        step *= stepMult;
        if (deltaStep < 0 && step < endStep) step = endStep;
        if (deltaStep > 0 && step > endStep) step = endStep;
    }
};

struct MaxPointCountDataPointsFilter : public DataPointsFilter {  // MaxPointCount.{h,cpp}
    static const std::string description() { return "Conditional subsampling. This filter reduces the size of the point cloud by randomly dropping points if their number is above maxCount. Based on \\cite{Masuda1996Random}"; }
    static const ParametersDoc availableParameters() {
        return {{"seed", "srand seed", "1", "0", "2147483647", &Parametrizable::Comp<size_t>},
                {"maxCount", "maximum number of points", "1000", "0", "2147483647", &Parametrizable::Comp<size_t>}};
    }
    const size_t maxCount;
    size_t seed;
    MaxPointCountDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("MaxPointCountDataPointsFilter", availableParameters(), params), maxCount(Parametrizable::get<size_t>("maxCount")),
          seed(Parametrizable::get<size_t>("seed")) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // MaxPointCount.cpp:71-110 (see pmgpu.h on the reference's view "swap")
        const int n = cloud.features.cols();
        if (n == 0 || !(maxCount <= (size_t)(n - 1))) return;
        std::vector<int32_t> order(n);
        const int m = pmgpu_host_max_point_count(n, seed, maxCount, order.data());
        cloud.keepColumns(std::vector<int>(order.begin(), order.begin() + m));
    }
};

struct MaxQuantileOnAxisDataPointsFilter : public DataPointsFilter {  // MaxQuantileOnAxis.{h,cpp}
    static const std::string description() { return "Subsampling. Filter points beyond a maximum quantile measured on a specific axis."; }
    static const ParametersDoc availableParameters() {
        return {{"dim", "dimension on which the filter will be applied. x=0, y=1, z=2", "0", "0", "2", &Parametrizable::Comp<unsigned>},
                {"ratio", "maximum quantile authorized. All points beyond that will be filtered.", "0.5", "0.0000001", "0.9999999", &Parametrizable::Comp<T>}};
    }
    const unsigned dim;
    const T ratio;
    MaxQuantileOnAxisDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("MaxQuantileOnAxisDataPointsFilter", availableParameters(), params), dim(Parametrizable::get<unsigned>("dim")),
          ratio(Parametrizable::get<T>("ratio")) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // MaxQuantileOnAxis.cpp:65-103
        if (int(dim) >= cloud.features.rows())
            throw InvalidParameter("MaxQuantileOnAxisDataPointsFilter: Error, filtering on dimension number " + std::to_string(dim) +
                                   ", larger than feature dimensionality " + std::to_string(cloud.features.rows()));
        const int n = cloud.features.cols();
        if (n == 0) return;
        const int nbPointsOut = n * ratio;
        std::vector<T> values;
        values.reserve(n);
        for (int x = 0; x < n; ++x) values.push_back(cloud.features(dim, x));
        std::nth_element(values.begin(), values.begin() + (std::ptrdiff_t)(values.size() * ratio), values.end());
        const T limit = values[nbPointsOut];
        std::vector<int> keep;
        for (int i = 0; i < n; ++i)
            if (cloud.features(dim, i) < limit) keep.push_back(i);
        cloud.keepColumns(keep);
    }
};

struct RemoveNaNDataPointsFilter : public DataPointsFilter {  // RemoveNaN.{h,cpp}
    static const std::string description() { return "Remove points having NaN as coordinate."; }
    RemoveNaNDataPointsFilter() : DataPointsFilter("RemoveNaNDataPointsFilter", ParametersDoc(), Parameters()) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // RemoveNaN.cpp:52-72
        const int n = cloud.features.cols(), rows = cloud.features.rows();
        std::vector<int> keep;
        for (int i = 0; i < n; ++i) {
            bool hasNaN = false;
            for (int r = 0; r < rows; ++r) hasNaN |= !(cloud.features(r, i) == cloud.features(r, i));
            if (!hasNaN) keep.push_back(i);
        }
        cloud.keepColumns(keep);
    }
};

struct MaxDensityDataPointsFilter : public DataPointsFilter {  // MaxDensity.{h,cpp}
    static const std::string description() { return "Subsampling. Reduce the points number by randomly removing points with a density highler than a treshold."; }
    static const ParametersDoc availableParameters() {
        return {{"maxDensity", "Maximum density of points to target. Unit: number of points per m^3.", "10", "0.0000001", "inf", &Parametrizable::Comp<T>}};
    }
    const T maxDensity;
    MaxDensityDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("MaxDensityDataPointsFilter", availableParameters(), params), maxDensity(Parametrizable::get<T>("maxDensity")) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // MaxDensity.cpp:60-105
        if (!cloud.descriptorExists("densities")) throw typename DataPoints::InvalidField("MaxDensityDataPointsFilter: Error, no densities found in descriptors.");
        const int n = cloud.features.cols();
        const unsigned row = cloud.getDescriptorStartingRow("densities");
        std::vector<float> dens(n > 0 ? n : 1);
        for (int i = 0; i < n; ++i) dens[i] = (float)cloud.descriptors(row, i);
        std::vector<int32_t> keep(n > 0 ? n : 1);
        const int m = pmgpu_host_max_density(dens.data(), 1, n, (float)maxDensity, keep.data());
        cloud.keepColumns(std::vector<int>(keep.begin(), keep.begin() + (m > 0 ? m : 0)));
    }
};

struct ShadowDataPointsFilter : public DataPointsFilter {  // Shadow.{h,cpp}
    static const std::string description() { return "Remove ghost points appearing on edge discontinuties. Assume that the origine of the point cloud is close to where the laser center was. Requires surface normal for every points"; }
    static const ParametersDoc availableParameters() {
        return {{"eps", "Small angle (in rad) around which a normal shoudn't be observable", "0.1", "0.0", "3.1416", &Parametrizable::Comp<T>}};
    }
    const T eps;  // sin of the parameter (Shadow.cpp:46)
    ShadowDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("ShadowDataPointsFilter", availableParameters(), params), eps(std::sin(Parametrizable::get<T>("eps"))) {}
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // Shadow.cpp:62-90
        if (!cloud.descriptorExists("normals")) throw typename DataPoints::InvalidField("ShadowDataPointsFilter, Error: cannot find normals in descriptors");
        const int dim = cloud.features.rows(), n = cloud.features.cols();
        const unsigned rn = cloud.getDescriptorStartingRow("normals");
        const int nd = cloud.getDescriptorDimension("normals");
        std::vector<int> keep;
        for (int i = 0; i < n; ++i) {
            T nn = 0, pn = 0;
            for (int r = 0; r < nd; ++r) nn += cloud.descriptors(rn + r, i) * cloud.descriptors(rn + r, i);
            for (int r = 0; r < dim - 1; ++r) pn += cloud.features(r, i) * cloud.features(r, i);
            nn = std::sqrt(nn);
            pn = std::sqrt(pn);
            T dot = 0;  // normalized() leaves a zero vector as it is
            for (int r = 0; r < nd && r < dim - 1; ++r)
                dot += (nn > 0 ? cloud.descriptors(rn + r, i) / nn : cloud.descriptors(rn + r, i)) * (pn > 0 ? cloud.features(r, i) / pn : cloud.features(r, i));
            if ((dot < 0 ? -dot : dot) > eps) keep.push_back(i);
        }
        cloud.keepColumns(keep);
    }
};

struct SimpleSensorNoiseDataPointsFilter : public DataPointsFilter {  // SimpleSensorNoise.{h,cpp}
    static const std::string description() { return "Add a 1D descriptor named <sensorNoise> that would represent the noise radius expressed in meter based on SICK LMS specifications \\cite{Pomerleau2012Noise}."; }
    static const ParametersDoc availableParameters() {
        return {{"sensorType", "Type of the sensor used. Choices: 0=Sick LMS-1xx, 1=Hokuyo URG-04LX, 2=Hokuyo UTM-30LX, 3=Kinect/Xtion", "0", "0", "2147483647", &Parametrizable::Comp<unsigned>},
                {"gain", "If the point cloud is coming from an untrusty source, you can use the gain to augment the uncertainty", "1", "1", "inf", &Parametrizable::Comp<T>}};
    }
    const unsigned sensorType;
    const T gain;  // read and never used, as in the reference
    SimpleSensorNoiseDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("SimpleSensorNoiseDataPointsFilter", availableParameters(), params), sensorType(Parametrizable::get<unsigned>("sensorType")),
          gain(Parametrizable::get<T>("gain")) {
        if (sensorType >= 5) throw InvalidParameter("SimpleSensorNoiseDataPointsFilter: Error, sensorType id " + std::to_string(sensorType) + " does not exist.");
    }
    PM_FILTER_COPY_THEN_IN_PLACE
    void inPlaceFilter(DataPoints& cloud) override {  // SimpleSensorNoise.cpp:75-140
        static const T laser[5][3] = {{T(0.012), T(0.0068), T(0.0008)}, {T(0.028), T(0.0013), T(0.0001)}, {T(0.018), T(0.0006), T(0.0015)},
                                      {0, 0, 0}, {T(0.004), T(0.0053), T(-0.0092)}};
        cloud.allocateDescriptor("simpleSensorNoise", 1);
        const unsigned row = cloud.getDescriptorStartingRow("simpleSensorNoise");
        const int dim = cloud.features.rows(), n = cloud.features.cols();
        for (int i = 0; i < n; ++i) {
            T acc = 0;
            for (int r = 0; r < dim - 1; ++r) acc += cloud.features(r, i) * cloud.features(r, i);
            const T norm = std::sqrt(acc);
            if (sensorType == 3)
                cloud.descriptors(row, i) = (norm * norm) * T(0.5 * 0.00285);
            else {
                const T v = laser[sensorType][1] * norm + laser[sensorType][2];
                cloud.descriptors(row, i) = v < laser[sensorType][0] ? laser[sensorType][0] : v;  // maxCoeff over (v, minRadius): a NaN stays
            }
        }
    }
};
#undef PM_FILTER_COPY_THEN_IN_PLACE

// ---- SurfaceNormalDataPointsFilter (DataPointsFilters/SurfaceNormal.{h,cpp}) — K8 -------------------
struct SurfaceNormalDataPointsFilter : public DataPointsFilter, public GpuBound {
    static const std::string description() {
        return "This filter extracts the surface normal vector and other statistics to each point by taking the eigenvector corresponding to "
               "the smallest eigenvalue of its nearest neighbors (GPU: exact kNN + per-point 3x3 eigen-solve).";
    }
    static const ParametersDoc availableParameters() {
        return {
            {"knn", "number of nearest neighbors to consider, including the point itself", "5", "3", "2147483647", &Parametrizable::Comp<unsigned>},
            {"maxDist", "maximum distance to consider for neighbors", "inf", "0", "inf", &Parametrizable::Comp<T>},
            {"epsilon", "approximation to use for the nearest-neighbor search", "0", "0", "inf", &Parametrizable::Comp<T>},
            {"keepNormals", "whether the normals should be added as descriptors to the resulting cloud", "1"},
            {"keepDensities", "whether the point densities should be added as descriptors to the resulting cloud", "0"},
            {"keepEigenValues", "whether the eigen values should be added as descriptors to the resulting cloud", "0"},
            {"keepEigenVectors", "whether the eigen vectors should be added as descriptors to the resulting cloud", "0"},
            {"keepMatchedIds", "whether the identifiers of matches points should be added as descriptors to the resulting cloud", "0"},
            {"keepMeanDist", "whether the distance to the nearest neighbor mean should be added as descriptors to the resulting cloud", "0"},
            {"sortEigen", "whether the eigenvalues and eigenvectors should be sorted (ascending) based on the eigenvalues", "0"},
            {"smoothNormals", "whether the normal vector should be average with the nearest neighbors", "0"}};
    }
    const unsigned knn;
    const T maxDist, epsilon;
    const bool keepNormals, keepDensities, keepEigenValues, keepEigenVectors, keepMatchedIds, keepMeanDist, sortEigen, smoothNormals;
    int degenerateCount = 0;

    SurfaceNormalDataPointsFilter(const Parameters& params = Parameters())
        : DataPointsFilter("SurfaceNormalDataPointsFilter", availableParameters(), params),
          knn(Parametrizable::get<unsigned>("knn")), maxDist(Parametrizable::get<T>("maxDist")), epsilon(Parametrizable::get<T>("epsilon")),
          keepNormals(Parametrizable::get<bool>("keepNormals")), keepDensities(Parametrizable::get<bool>("keepDensities")),
          keepEigenValues(Parametrizable::get<bool>("keepEigenValues")), keepEigenVectors(Parametrizable::get<bool>("keepEigenVectors")),
          keepMatchedIds(Parametrizable::get<bool>("keepMatchedIds")), keepMeanDist(Parametrizable::get<bool>("keepMeanDist")),
          sortEigen(Parametrizable::get<bool>("sortEigen")), smoothNormals(Parametrizable::get<bool>("smoothNormals")) {}
    DataPoints filter(const DataPoints& input) override {
        DataPoints output(input);
        inPlaceFilter(output);
        return output;
    }
    void inPlaceFilter(DataPoints& cloud) override {
        requireFloat3D(cloud.features.rows(), "SurfaceNormalDataPointsFilter");
        unsigned insertDim = 0;
        for (const auto& l : cloud.descriptorLabels) insertDim += l.span;
        if (insertDim != cloud.getDescriptorDim())
            throw typename DataPoints::InvalidField("SurfaceNormalDataPointsFilter: Error, descriptor labels do not match descriptor data");
        const int dn = cloud.features.rows() - 1;  // spans follow the cloud's dimension (SurfaceNormal.cpp:105-125)
        if (keepNormals) cloud.allocateDescriptor("normals", dn);
        if (keepDensities) cloud.allocateDescriptor("densities", 1);
        if (keepEigenValues) cloud.allocateDescriptor("eigValues", dn);
        if (keepEigenVectors) cloud.allocateDescriptor("eigVectors", dn * dn);
        if (keepMatchedIds) cloud.allocateDescriptor("matchedIds", knn);
        if (keepMeanDist) cloud.allocateDescriptor("meanDists", 1);
        // outputs land directly in the rows of the descriptor matrix (column stride = its row count)
        pmgpu_normals_out out;
        std::memset(&out, 0, sizeof(out));
        const int ld = cloud.descriptors.rows();
        float* base = reinterpret_cast<float*>(cloud.descriptors.data());
        auto at = [&](const char* name) { return base + cloud.getDescriptorStartingRow(name); };
        if (keepNormals) { out.normals = at("normals"); out.normals_ld = ld; }
        if (keepDensities) { out.densities = at("densities"); out.densities_ld = ld; }
        if (keepEigenValues) { out.eig_values = at("eigValues"); out.eig_values_ld = ld; }
        if (keepEigenVectors) { out.eig_vectors = at("eigVectors"); out.eig_vectors_ld = ld; }
        if (keepMatchedIds) { out.matched_ids = at("matchedIds"); out.matched_ids_ld = ld; }
        if (keepMeanDist) { out.mean_dists = at("meanDists"); out.mean_dists_ld = ld; }
        GpuPipeline& g = this->gpu();
        g.check(pmgpu_normals(g.ctx, reinterpret_cast<const float*>(cloud.features.data()), cloud.features.rows(), cloud.features.cols(), (int)knn,
                              (float)epsilon, (float)maxDist, (sortEigen ? PMGPU_NORMALS_SORT_EIGEN : 0) | (smoothNormals ? PMGPU_NORMALS_SMOOTH : 0), &out,
                              &degenerateCount));
    }
};

// ---- KDTreeMatcher (MatchersImpl.{h,cpp}:74-101) — K1 + K2 ------------------------------------------
struct KDTreeMatcher : public Matcher, public GpuBound {
    static const std::string description() {
        return "This matcher matches a point from the reading to its closest neighbors in the reference (GPU: exact search, answers of "
               "libnabo's brute-force search).";
    }
    static const ParametersDoc availableParameters() {
        return {{"knn", "number of nearest neighbors to consider it the reference", "1", "1", "2147483647", &Parametrizable::Comp<unsigned>},
                {"epsilon", "approximation to use for the nearest-neighbor search", "0", "0", "inf", &Parametrizable::Comp<T>},
                {"searchType", "Nabo search type. 0: brute force, check distance to every point in the data (very slow), 1: kd-tree with linear heap, good for small knn (~up to 30) and 2: kd-tree with tree heap, good for large knn (~from 30)", "1", "0", "2", &Parametrizable::Comp<unsigned>},
                {"maxDist", "maximum distance to consider for neighbors", "inf", "0", "inf", &Parametrizable::Comp<T>}};
    }
    const int knn;
    const T epsilon;
    const int searchType;
    const T maxDist;

    std::string maxDistField;  // KDTreeVarDistMatcher: the reading descriptor that holds one search distance per point

    KDTreeMatcher(const Parameters& params = Parameters())
        : Matcher("KDTreeMatcher", availableParameters(), params), knn(Parametrizable::get<int>("knn")), epsilon(Parametrizable::get<T>("epsilon")),
          searchType(Parametrizable::get<int>("searchType")), maxDist(Parametrizable::get<T>("maxDist")) {}
    // the per-point distances of `reading` for the resident reading (no-op for the plain matcher)
    void uploadMaxDists(GpuPipeline& g, const DataPoints& reading) const {
        if (maxDistField.empty()) return;
        const unsigned row = reading.getDescriptorStartingRow(maxDistField);  // throws InvalidField like getDescriptorViewByName
        g.check(pmgpu_reading_set_max_dists(g.ctx, reinterpret_cast<const float*>(reading.descriptors.data()) + row, reading.descriptors.rows()));
    }
protected:
    KDTreeMatcher(const std::string& className, const ParametersDoc& doc, const Parameters& params, const std::string& field)
        : Matcher(className, doc, params), knn(Parametrizable::get<int>("knn")), epsilon(Parametrizable::get<T>("epsilon")),
          searchType(Parametrizable::get<int>("searchType")), maxDist(T(-1)), maxDistField(field) {}
public:

    void init(const DataPoints& filteredReference) override { initImpl(filteredReference, nullptr); }
    // init() on the reference centred on its mean — the preamble of ICP::compute (ICP.cpp:291-302)
    // without a host copy of the cloud; mean4 receives the mean
    void initCentered(const DataPoints& filteredReference, float* mean4) { initImpl(filteredReference, mean4); }
    void initImpl(const DataPoints& filteredReference, float* mean4) {
        requireFloat3D(filteredReference.features.rows(), "KDTreeMatcher");
        GpuPipeline& g = this->gpu();
        const float* normals = nullptr;
        int ld = 0;
        if (filteredReference.descriptorExists("normals", filteredReference.features.rows() - 1) && filteredReference.descriptors.cols() == filteredReference.features.cols()) {
            normals = reinterpret_cast<const float*>(filteredReference.descriptors.data()) + filteredReference.getDescriptorStartingRow("normals");
            ld = filteredReference.descriptors.rows();
        }
        const float* feat = reinterpret_cast<const float*>(filteredReference.features.data());
        if (mean4) g.check(pmgpu_ref_set_centered(g.ctx, feat, filteredReference.features.rows(), filteredReference.features.cols(), normals, ld, mean4));
        else g.check(pmgpu_ref_set(g.ctx, feat, filteredReference.features.rows(), filteredReference.features.cols(), normals, ld));
        g.readingKey = nullptr;
    }
    // uploads the reading when it is not the resident one, then matches T * reading
    Matches findClosestsTransformed(const DataPoints& reading, const TransformationParameters* Tr) {
        requireFloat3D(reading.features.rows(), "KDTreeMatcher");
        GpuPipeline& g = this->gpu();
        if (g.readingKey != reading.features.data() || g.readingCols != reading.features.cols()) {
            g.check(pmgpu_reading_set(g.ctx, reinterpret_cast<const float*>(reading.features.data()), reading.features.rows(), reading.features.cols()));
            g.readingKey = reading.features.data();
            g.readingCols = reading.features.cols();
            uploadMaxDists(g, reading);
        }
        Matches matches(knn, reading.features.cols());
        uint64_t visits = 0;
        g.check(pmgpu_knn(g.ctx, Tr ? reinterpret_cast<const float*>(Tr->data()) : nullptr, knn, (float)epsilon, (float)maxDist, matches.ids.data(),
                          reinterpret_cast<float*>(matches.dists.data()), &visits));
        this->visitCounter += visits;
        return matches;
    }
    Matches findClosests(const DataPoints& filteredReading) override {
        this->gpu().readingKey = nullptr;  // the caller may have modified the cloud in place: always upload
        return findClosestsTransformed(filteredReading, nullptr);
    }
};

// KDTreeVarDistMatcher (MatchersImpl.h:105-127, MatchersImpl.cpp:105-150): one maximum search distance per reading point
struct KDTreeVarDistMatcher : public KDTreeMatcher {
    static const std::string description() {
        return "This matcher matches a point from the reading to its closest neighbors in the reference. A maximum search radius per point can be "
               "defined.";
    }
    static const ParametersDoc availableParameters() {
        return {{"knn", "number of nearest neighbors to consider it the reference", "1", "1", "2147483647", &Parametrizable::Comp<unsigned>},
                {"epsilon", "approximation to use for the nearest-neighbor search", "0", "0", "inf", &Parametrizable::Comp<T>},
                {"searchType", "Nabo search type. 0: brute force, check distance to every point in the data (very slow), 1: kd-tree with linear heap, good for small knn (~up to 30) and 2: kd-tree with tree heap, good for large knn (~from 30)", "1", "0", "2", &Parametrizable::Comp<unsigned>},
                {"maxDistField", "descriptor field name used to set a maximum distance to consider for neighbors per point", "maxSearchDist"}};
    }
    static std::string field(const Parameters& params) {
        const auto it = params.find("maxDistField");
        return it == params.end() ? std::string("maxSearchDist") : it->second;
    }
    KDTreeVarDistMatcher(const Parameters& params = Parameters()) : KDTreeMatcher("KDTreeVarDistMatcher", availableParameters(), params, field(params)) {
        Parametrizable::get<std::string>("maxDistField");
    }
};

// ---- outlier filters (OutlierFiltersImpl.{h,cpp}:66-147) — K3 ---------------------------------------
struct NullOutlierFilter : public OutlierFilter {
    static const std::string description() { return "Does nothing."; }
    NullOutlierFilter() : OutlierFilter("NullOutlierFilter", ParametersDoc(), Parameters()) {}
    OutlierWeights compute(const DataPoints&, const DataPoints&, const Matches& input) override {
        return OutlierWeights::Constant(input.ids.rows(), input.ids.cols(), 1);
    }
};
struct GpuDistOutlierFilter : public OutlierFilter, public GpuBound {
    int filterType;
    T value;
    GpuDistOutlierFilter(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params, int type, const char* paramName)
        : OutlierFilter(className, paramsDoc, params), filterType(type), value(Parametrizable::get<T>(paramName)) {}
    // parameters that do not fit the (type, value) pair go to the context before the chain is evaluated
    virtual void prepare(GpuPipeline&) const {}
    OutlierWeights compute(const DataPoints&, const DataPoints&, const Matches& input) override {
        GpuPipeline& g = this->gpu();
        prepare(g);
        OutlierWeights w(input.ids.rows(), input.ids.cols());
        const float p = (float)value;
        g.check(pmgpu_weights(g.ctx, 1, &filterType, &p, reinterpret_cast<float*>(w.data()), nullptr));
        return w;
    }
};
// OutlierFiltersImpl.h:147-172, OutlierFiltersImpl.cpp:152-218: TrimmedDist with the ratio minimising the fractional RMS distance
struct VarTrimmedDistOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() { return "Hard rejection threshold using quantile and variable ratio. Based on \\cite{Phillips2007VarTrimmed}."; }
    static const ParametersDoc availableParameters() {
        return {{"minRatio", "min ratio", "0.05", "0.0000001", "1", &Parametrizable::Comp<T>},
                {"maxRatio", "max ratio", "0.99", "0.0000001", "1", &Parametrizable::Comp<T>},
                {"lambda", "lambda (part of the term that balance the rmsd: 1/ratio^lambda", "2.35"}};
    }
    const T minRatio, maxRatio, lambda;
    VarTrimmedDistOutlierFilter(const Parameters& params = Parameters())
        : GpuDistOutlierFilter("VarTrimmedDistOutlierFilter", availableParameters(), params, PMGPU_FILTER_VARTRIMMEDDIST, "lambda"),
          minRatio(Parametrizable::get<T>("minRatio")), maxRatio(Parametrizable::get<T>("maxRatio")), lambda(Parametrizable::get<T>("lambda")) {
        if (this->minRatio >= this->maxRatio)
            throw InvalidParameter("VarTrimmedDistOutlierFilter: minRatio (" + std::to_string(minRatio) + ") should be smaller than maxRatio (" +
                                   std::to_string(maxRatio) + ")");
    }
    void prepare(GpuPipeline& g) const override { g.check(pmgpu_set_var_trimmed_ratios(g.ctx, (float)minRatio, (float)maxRatio)); }
};
struct MaxDistOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() { return "This filter considers as outlier links whose norms are above a fix threshold."; }
    static const ParametersDoc availableParameters() { return {{"maxDist", "threshold distance (Euclidean norm)", "1", "0.0000001", "inf", &Parametrizable::Comp<T>}}; }
    MaxDistOutlierFilter(const Parameters& params = Parameters()) : GpuDistOutlierFilter("MaxDistOutlierFilter", availableParameters(), params, PMGPU_FILTER_MAXDIST, "maxDist") {}
};
struct MinDistOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() { return "This filter considers as outlier links whose norms are below a threshold."; }
    static const ParametersDoc availableParameters() { return {{"minDist", "threshold distance (Euclidean norm)", "1", "0.0000001", "inf", &Parametrizable::Comp<T>}}; }
    MinDistOutlierFilter(const Parameters& params = Parameters()) : GpuDistOutlierFilter("MinDistOutlierFilter", availableParameters(), params, PMGPU_FILTER_MINDIST, "minDist") {}
};
struct MedianDistOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() { return "This filter considers as outlier links whose norms are above the median link norms times a factor."; }
    static const ParametersDoc availableParameters() { return {{"factor", "points farther away factor * median will be considered outliers.", "3", "0.0000001", "inf", &Parametrizable::Comp<T>}}; }
    MedianDistOutlierFilter(const Parameters& params = Parameters()) : GpuDistOutlierFilter("MedianDistOutlierFilter", availableParameters(), params, PMGPU_FILTER_MEDIANDIST, "factor") {}
};
struct TrimmedDistOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() { return "Hard rejection threshold using quantile."; }
    static const ParametersDoc availableParameters() { return {{"ratio", "percentage to keep", "0.85", "0.0000001", "1.0", &Parametrizable::Comp<T>}}; }
    TrimmedDistOutlierFilter(const Parameters& params = Parameters()) : GpuDistOutlierFilter("TrimmedDistOutlierFilter", availableParameters(), params, PMGPU_FILTER_TRIMMEDDIST, "ratio") {}
};

struct SurfaceNormalOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() {
        return "Hard rejection threshold using the angle between the surface normal vector of the reading and the reference. "
               "If normal vectors or not in the descriptor for both of the point clouds, does nothing.";
    }
    static const ParametersDoc availableParameters() {
        return {{"maxAngle", "Maximum authorised angle between the 2 surface normals (in radian)", "1.57", "0.0", "3.1416", &Parametrizable::Comp<T>}};
    }
    SurfaceNormalOutlierFilter(const Parameters& params = Parameters())
        : GpuDistOutlierFilter("SurfaceNormalOutlierFilter", availableParameters(), params, PMGPU_FILTER_SURFACENORMAL, "maxAngle") {}
};
// RobustOutlierFilter (OutlierFiltersImpl.h:199-262, .cpp:420-598): M-estimator weights on the GPU;
// the discrete parameters travel in the filter word (pmgpu.h)
struct RobustOutlierFilter : public GpuDistOutlierFilter {
    static const std::string description() {
        return "Robust weight function part of the M-Estimator familly. 8 robust functions to choose from (Cauchy, Welsch, Switchable Constraint, "
               "Geman-McClure, Tukey, Huber, L1 and Student). All the functions are M-Estimator (\\cite{RobustWeightFunctions}) except L1 and Student.";
    }
    static const ParametersDoc availableParameters() {
        return {
            {"robustFct", "Type of robust function used. Available fct: 'cauchy', 'welsch', 'sc'(aka Switchable-Constraint), 'gm' (aka Geman-McClure), 'tukey', 'huber' and 'L1'. (Default: cauchy)", "cauchy"},
            {"tuning", "Tuning parameter used to limit the influence of outliers.If the 'scaleEstimator' is 'mad' or 'none', this parameter acts as the tuning parameter.If the 'scaleEstimator' is 'berg' this parameter acts as the target scale (σ*).", "1.0", "0.0000001", "inf", &Parametrizable::Comp<T>},
            {"scaleEstimator", "The scale estimator is used to convert the error distance into a Mahalanobis distance. 3 estimators are available: 'none': no estimator (scale = 1), 'mad': use the median of absolute deviation (a kind of robust standard deviation), 'berg': an iterative exponentially decreasing estimator", "mad"},
            {"nbIterationForScale", "For how many iteration the 'scaleEstimator' is recalculated. After 'nbIterationForScale' iteration the previous scale is kept. A nbIterationForScale==0 means that the estiamtor is recalculated at each iteration.", "0", "0", "100", &Parametrizable::Comp<int>},
            {"distanceType", "Type of error distance used, either point to point ('point2point') or point to plane('point2plane'). Point to point gives better result normally.", "point2point"},
            {"approximation", "If the matched distance is larger than this threshold, its weight will be forced to zero. This can save computation as zero values are not minimized. If set to inf (default value), no approximation is done. The unit of this parameter is the same as the distance used, typically meters.", "inf", "0.0", "inf", &Parametrizable::Comp<T>}};
    }
    static int word(const Parameters& params) {
        auto get = [&](const char* name, const char* def) {
            const auto it = params.find(name);
            return it == params.end() ? std::string(def) : it->second;
        };
        static const std::map<std::string, int> fcts = {{"cauchy", PMGPU_ROBUST_CAUCHY}, {"welsch", PMGPU_ROBUST_WELSCH}, {"sc", PMGPU_ROBUST_SC},
                                                        {"gm", PMGPU_ROBUST_GM},         {"tukey", PMGPU_ROBUST_TUKEY},   {"huber", PMGPU_ROBUST_HUBER},
                                                        {"L1", PMGPU_ROBUST_L1},         {"student", PMGPU_ROBUST_STUDENT}};
        const auto f = fcts.find(get("robustFct", "cauchy"));
        if (f == fcts.end()) throw InvalidParameter("Invalid robust function name.");
        const std::string est = get("scaleEstimator", "mad");
        static const std::map<std::string, int> ests = {{"none", PMGPU_SCALE_NONE}, {"mad", PMGPU_SCALE_MAD}, {"berg", PMGPU_SCALE_BERG}, {"std", PMGPU_SCALE_STD}};
        const auto e = ests.find(est);
        if (e == ests.end()) throw InvalidParameter("Invalid scale estimator name.");
        const std::string dt = get("distanceType", "point2point");
        if (dt != "point2point" && dt != "point2plane") throw InvalidParameter("Invalid distance type name.");
        return PMGPU_ROBUST_WORD(f->second, e->second, std::stoi(get("nbIterationForScale", "0"))) | (dt == "point2plane" ? PMGPU_ROBUST_P2PLANE : 0);
    }
    T approximation;
    void prepare(GpuPipeline& g) const override { g.check(pmgpu_set_robust_approximation(g.ctx, (float)approximation)); }
    RobustOutlierFilter(const Parameters& params = Parameters())
        : GpuDistOutlierFilter("RobustOutlierFilter", availableParameters(), params, word(params), "tuning") {
        // every declared parameter is read (Registrar.h:103-110)
        Parametrizable::get<std::string>("robustFct"); Parametrizable::get<std::string>("scaleEstimator"); Parametrizable::get<int>("nbIterationForScale");
        Parametrizable::get<std::string>("distanceType");
        approximation = Parametrizable::get<T>("approximation");
    }
};

// chain (OutlierFilter.cpp:63-103): product of the filters' weights; empty chain -> dist != inf.
// A chain made of GPU distance filters is evaluated in one call (one collapsed threshold).
struct OutlierFilters : public std::vector<std::shared_ptr<OutlierFilter>>, public GpuBound {
    // a NullOutlierFilter is a factor of one (OutlierFiltersImpl.cpp:45-58): allowed in a GPU chain, never sent to the device
    bool allGpu() const {
        for (const auto& f : *this)
            if (!dynamic_cast<GpuDistOutlierFilter*>(f.get()) && !dynamic_cast<NullOutlierFilter*>(f.get())) return false;
        return true;
    }
    // fills the filter words of the chain, returns how many there are
    int spec(int* types, float* params, GpuPipeline& pipeline) const {
        int i = 0;
        for (const auto& f : *this) {
            const auto* g = dynamic_cast<const GpuDistOutlierFilter*>(f.get());
            if (!g) continue;
            g->prepare(pipeline);
            types[i] = g->filterType;
            params[i] = (float)g->value;
            ++i;
        }
        return i;
    }
    OutlierWeights compute(const DataPoints& filteredReading, const DataPoints& filteredReference, const Matches& input) {
        if (allGpu() && this->size() <= 8) {
            GpuPipeline& g = this->gpu();
            int types[8];
            float params[8];
            const int n = spec(types, params, g);
            OutlierWeights w(input.ids.rows(), input.ids.cols());
            g.check(pmgpu_weights(g.ctx, n, types, params, reinterpret_cast<float*>(w.data()), nullptr));
            return w;
        }
        OutlierWeights w = (*this->begin())->compute(filteredReading, filteredReference, input);
        for (auto it = this->begin() + 1; it != this->end(); ++it) {
            const OutlierWeights o = (*it)->compute(filteredReading, filteredReference, input);
            for (size_t i = 0; i < w.size(); ++i) w(i) = w(i) * o(i);
        }
        return w;
    }
};

// ---- error minimizers (ErrorMinimizers/*.cpp) — K4-K7 ------------------------------------------------
struct GpuErrorMinimizer : public ErrorMinimizer, public GpuBound {
    typedef typename ErrorMinimizer::ErrorElements ErrorElements;
    int kind;
    T sensorStdDev;
    Matrix covMatrix;
    GpuErrorMinimizer(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params, int kind)
        : ErrorMinimizer(className, paramsDoc, params), kind(kind), sensorStdDev(T(0.01)), covMatrix(Matrix::Zero(6, 6)) {}
    // uses the matches / weights resident on the device; the host arguments are what the reference
    // interface hands around (ErrorMinimizer.cpp:217-232)
    TransformationParameters compute(const DataPoints& filteredReading, const DataPoints&, const OutlierWeights&, const Matches&) override {
        requireFloat3D(filteredReading.features.rows(), "ErrorMinimizer");
        GpuPipeline& g = this->gpu();
        TransformationParameters out(filteredReading.features.rows(), filteredReading.features.rows());
        float cov[36], stats[5];
        g.check(pmgpu_minimize(g.ctx, kind, (float)sensorStdDev, reinterpret_cast<float*>(out.data()), cov, stats));
        setResults(cov, stats);
        return out;
    }
    void setResults(const float* cov, const float* stats) {
        if ((kind & 0xff) == PMGPU_MIN_P2POINT_COV || (kind & 0xff) == PMGPU_MIN_P2PLANE_COV)
            for (int i = 0; i < 36; ++i) covMatrix(i) = T(cov[i]);
        this->lastErrorElements.pointUsedRatio = T(stats[0]);
        this->lastErrorElements.weightedPointUsedRatio = T(stats[1]);
        this->lastErrorElements.nbRejectedMatches = (int)stats[2];
        this->lastErrorElements.nbRejectedPoints = (int)stats[3];
    }
    Matrix getCovariance() const override { return covMatrix; }
    bool isPlane() const { return (kind & 0xff) == PMGPU_MIN_P2PLANE || (kind & 0xff) == PMGPU_MIN_P2PLANE_COV; }
    static T deltaNorm(const ErrorElements& e, int i) {
        T acc = 0;
        for (int r = 0; r < e.reading.features.rows() - 1; ++r) {
            const T d = e.reading.features(r, i) - e.reference.features(r, i);
            acc += d * d;
        }
        return std::sqrt(acc);
    }
    // PointToPoint.cpp:153-163: sum |reading - reference|;  PointToPlane.cpp:314-352: sum w (n . (reading - reference))^2
    static T computeResidualError(const ErrorElements& e, bool plane, bool force2D) {
        const int m = e.reading.features.cols();
        double total = 0;
        if (!plane) {
            for (int i = 0; i < m; ++i) total += deltaNorm(e, i);
            return T(total);
        }
        const unsigned rn = e.reference.getDescriptorStartingRow("normals");
        const int dims = force2D ? 2 : e.reading.features.rows() - 1;
        for (int i = 0; i < m; ++i) {
            T dot = 0;
            for (int r = 0; r < dims; ++r) dot += (e.reading.features(r, i) - e.reference.features(r, i)) * e.reference.descriptors(rn + r, i);
            total += e.weights(0, i) * (dot * dot);
        }
        return T(total);
    }
    T getResidualError(const DataPoints& filteredReading, const DataPoints& filteredReference, const OutlierWeights& outlierWeights,
                       const Matches& matches) const override {
        const ErrorElements mPts(filteredReading, filteredReference, outlierWeights, matches);
        return computeResidualError(mPts, isPlane(), (kind & PMGPU_MIN_FORCE2D) != 0);
    }
    // the same number for the last iteration of a registration, from the resident matches
    T getResidualError() const { return computeResidualError(this->getErrorElements(), isPlane(), (kind & PMGPU_MIN_FORCE2D) != 0); }
    // PointToPoint.cpp:116-151, PointToPlane.cpp:369-466: the noise-based estimate when the clouds carry simpleSensorNoise,
    // else the weighted ratio of the outlier filters (ErrorMinimizer.cpp:227-230)
    T getOverlap() const override {
        if (!this->materialize || (kind & 0xff) == PMGPU_MIN_P2POINT_SIM) return this->lastErrorElements.weightedPointUsedRatio;
        const ErrorElements e = this->getErrorElements();
        const int m = e.reading.features.cols();
        if (m == 0) throw std::runtime_error("Error, last error element empty. Error minimizer needs to be called at least once before using this method.");
        const bool rNoise = e.reading.descriptorExists("simpleSensorNoise"), fNoise = e.reference.descriptorExists("simpleSensorNoise");
        if (!isPlane()) {
            if (!rNoise) return this->lastErrorElements.weightedPointUsedRatio;
            const unsigned row = e.reading.getDescriptorStartingRow("simpleSensorNoise");
            std::vector<T> dists(m);
            double sum = 0;
            for (int i = 0; i < m; ++i) sum += (dists[i] = deltaNorm(e, i));
            const T mean = T(sum) / T(m);
            int count = 0;
            for (int i = 0; i < m; ++i) count += dists[i] < (mean + e.reading.descriptors(row, i)) ? 1 : 0;
            return T(count) / T(m);
        }
        if (!rNoise && !fNoise) return this->lastErrorElements.weightedPointUsedRatio;
        const unsigned rr = rNoise ? e.reading.getDescriptorStartingRow("simpleSensorNoise") : 0;
        const unsigned fr = fNoise ? e.reference.getDescriptorStartingRow("simpleSensorNoise") : 0;
        T medianRadius = 0;
        const bool optimal = rNoise && fNoise && e.reference.descriptorExists("densities");
        if (optimal) {
            const unsigned dr = e.reference.getDescriptorStartingRow("densities");
            std::vector<T> values(m);
            for (int i = 0; i < m; ++i) values[i] = e.reference.descriptors(dr, i);
            std::nth_element(values.begin(), values.begin() + (std::ptrdiff_t)(values.size() * 0.5), values.end());
            const T medianDensity = values[(size_t)(values.size() * 0.5)];
            medianRadius = T(1.0 / std::pow((double)medianDensity, 1 / 3.0));
        }
        int count = 0, nbUniquePoint = 1;
        const int rows = e.reading.features.rows();
        std::vector<T> last(rows);
        for (int r = 0; r < rows; ++r) last[r] = e.reading.features(r, 0) * T(2);
        auto differs = [&](const T* a, int col) {
            for (int r = 0; r < rows; ++r)
                if (a[r] != e.reading.features(r, col)) return true;
            return false;
        };
        std::vector<T> prev(rows);
        for (int i = 0; i < m; ++i) {
            T unc;
            if (optimal) unc = (medianRadius + e.reading.descriptors(rr, i)) + e.reference.descriptors(fr, i);
            else if (rNoise && fNoise) unc = e.reading.descriptors(rr, i) + e.reference.descriptors(fr, i);
            else unc = rNoise ? e.reading.descriptors(rr, i) : e.reference.descriptors(fr, i);
            if (differs(last.data(), i) && deltaNorm(e, i) < unc) {
                for (int r = 0; r < rows; ++r) last[r] = e.reading.features(r, i);
                ++count;
            }
            if (i > 0 && differs(prev.data(), i)) ++nbUniquePoint;
            for (int r = 0; r < rows; ++r) prev[r] = e.reading.features(r, i);
        }
        return T(count) / T(nbUniquePoint + e.nbRejectedPoints);
    }
};
struct PointToPointErrorMinimizer : public GpuErrorMinimizer {
    static const std::string description() { return "Point-to-point error. Based on SVD decomposition."; }
    PointToPointErrorMinimizer() : GpuErrorMinimizer("PointToPointErrorMinimizer", ParametersDoc(), Parameters(), PMGPU_MIN_P2POINT) {}
};
struct PointToPointWithCovErrorMinimizer : public GpuErrorMinimizer {
    static const std::string description() { return "Point-to-point error. Additionally, it computes the covariance (Censi 2007)."; }
    static const ParametersDoc availableParameters() { return {{"sensorStdDev", "sensor standard deviation", "0.01", "0.", "inf", &Parametrizable::Comp<T>}}; }
    PointToPointWithCovErrorMinimizer(const Parameters& params = Parameters())
        : GpuErrorMinimizer("PointToPointWithCovErrorMinimizer", availableParameters(), params, PMGPU_MIN_P2POINT_COV) {
        this->sensorStdDev = Parametrizable::get<T>("sensorStdDev");
    }
};
struct PointToPointSimilarityErrorMinimizer : public GpuErrorMinimizer {
    static const std::string description() { return "Point-to-point similarity error (rotation + translation + scale). The scale is the same for all coordinates. Based on SVD decomposition."; }
    PointToPointSimilarityErrorMinimizer() : GpuErrorMinimizer("PointToPointSimilarityErrorMinimizer", ParametersDoc(), Parameters(), PMGPU_MIN_P2POINT_SIM) {}
};
struct PointToPlaneErrorMinimizer : public GpuErrorMinimizer {
    static const std::string description() { return "Point-to-plane error (or point-to-line in 2D)."; }
    static const ParametersDoc availableParameters() {
        return {{"force2D", "If set to true(1), the minimization will be forced to give a solution in 2D (i.e., on the XY-plane) even with 3D inputs.", "0", "0", "1", &Parametrizable::Comp<bool>},
                {"force4DOF", "If set to true(1), the minimization will optimize only yaw and translation, pitch and roll will follow the prior.", "0", "0", "1", &Parametrizable::Comp<bool>}};
    }
    PointToPlaneErrorMinimizer(const Parameters& params = Parameters()) : PointToPlaneErrorMinimizer("PointToPlaneErrorMinimizer", availableParameters(), params, PMGPU_MIN_P2PLANE) {}
protected:
    PointToPlaneErrorMinimizer(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params, int kind)
        : GpuErrorMinimizer(className, paramsDoc, params, kind) {
        const bool force2D = Parametrizable::get<bool>("force2D"), force4DOF = Parametrizable::get<bool>("force4DOF");
        if (force2D && force4DOF) throw ConfigurationError("Force 2D cannot be used together with force4DOF.");  // PointToPlane.cpp:59-64
        if (force2D && kind != PMGPU_MIN_P2PLANE) throw ConfigurationError(className + ": GPU module: force2D is not supported together with the covariance");
        if (force2D) this->kind |= PMGPU_MIN_FORCE2D;  // the 3x3 sub-system without z, PointToPlane.cpp:177-186
        if (force4DOF) this->kind |= PMGPU_MIN_FORCE4DOF;  // the 4x4 sub-system, PointToPlane.cpp:203-214
    }
};
struct PointToPlaneWithCovErrorMinimizer : public PointToPlaneErrorMinimizer {
    static const std::string description() { return "Point-to-plane error (or point-to-line in 2D). Additionally, it computes the covariance (Censi 2007)."; }
    static const ParametersDoc availableParameters() {
        ParametersDoc d = PointToPlaneErrorMinimizer::availableParameters();
        d.push_back({"sensorStdDev", "sensor standard deviation", "0.01", "0.", "inf", &Parametrizable::Comp<T>});
        return d;
    }
    PointToPlaneWithCovErrorMinimizer(const Parameters& params = Parameters())
        : PointToPlaneErrorMinimizer("PointToPlaneWithCovErrorMinimizer", availableParameters(), params, PMGPU_MIN_P2PLANE_COV) {
        this->sensorStdDev = Parametrizable::get<T>("sensorStdDev");
    }
};

// ---- transformation checkers (TransformationCheckersImpl.{h,cpp}) — host objects ------------------
struct Quat {
    T w, x, y, z;
    static Quat fromMatrix(const TransformationParameters& m) {  // Eigen's matrix -> quaternion
        Quat q;
        T t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1));
            q.w = T(0.5) * t;
            t = T(0.5) / t;
            q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1));
            T v[3];
            v[i] = T(0.5) * t;
            t = T(0.5) / t;
            q.w = (m(k, j) - m(j, k)) * t;
            v[j] = (m(j, i) + m(i, j)) * t;
            v[k] = (m(k, i) + m(i, k)) * t;
            q.x = v[0]; q.y = v[1]; q.z = v[2];
        }
        return q;
    }
    T angularDistance(const Quat& b) const {  // Eigen 3.3: d = a * conj(b); 2 atan2(|d.vec|, |d.w|)
        const T dw = w * b.w + x * b.x + y * b.y + z * b.z;
        const T dx = -w * b.x + x * b.w - y * b.z + z * b.y;
        const T dy = -w * b.y + y * b.w - z * b.x + x * b.z;
        const T dz = -w * b.z + z * b.w - x * b.y + y * b.x;
        return T(2) * std::atan2(std::sqrt(dx * dx + dy * dy + dz * dz), std::fabs(dw));
    }
};
struct CounterTransformationChecker : public TransformationChecker {
    struct MaxNumIterationsReached {};
    static const std::string description() { return "This checker stops the ICP loop after a certain number of iterations."; }
    static const ParametersDoc availableParameters() { return {{"maxIterationCount", "maximum number of iterations ", "40", "0", "2147483647", &Parametrizable::Comp<unsigned>}}; }
    const unsigned maxIterationCount;
    CounterTransformationChecker(const Parameters& params = Parameters())
        : TransformationChecker("CounterTransformationChecker", availableParameters(), params), maxIterationCount(Parametrizable::get<unsigned>("maxIterationCount")) {
        this->limits = Vector::Zero(1, 1);
        this->limits(0) = T(maxIterationCount);
        this->conditionVariableNames.push_back("Iteration");
        this->limitNames.push_back("Max iteration");
    }
    void init(const TransformationParameters&, bool&) override { this->conditionVariables = Vector::Zero(1, 1); }
    void check(const TransformationParameters&, bool& iterate) override {
        this->conditionVariables(0) += T(1);
        if (this->conditionVariables(0) >= this->limits(0)) {
            iterate = false;
            throw MaxNumIterationsReached();
        }
    }
};
struct DifferentialTransformationChecker : public TransformationChecker {
    static const std::string description() { return "This checker stops the ICP loop when the relative motions (i.e. abs(currentIter - lastIter)) of rotation and translation components are below a fix thresholds."; }
    static const ParametersDoc availableParameters() {
        return {{"minDiffRotErr", "threshold for rotation error (radian)", "0.001", "0.", "6.2831854", &Parametrizable::Comp<T>},
                {"minDiffTransErr", "threshold for translation error", "0.001", "0.", "inf", &Parametrizable::Comp<T>},
                {"smoothLength", "number of iterations over which to average the differencial error", "3", "0", "2147483647", &Parametrizable::Comp<unsigned>}};
    }
    const T minDiffRotErr, minDiffTransErr;
    const unsigned smoothLength;
    std::vector<Quat> rotations;
    std::vector<std::vector<T>> translations;
    DifferentialTransformationChecker(const Parameters& params = Parameters())
        : TransformationChecker("DifferentialTransformationChecker", availableParameters(), params), minDiffRotErr(Parametrizable::get<T>("minDiffRotErr")),
          minDiffTransErr(Parametrizable::get<T>("minDiffTransErr")), smoothLength(Parametrizable::get<unsigned>("smoothLength")) {
        this->limits = Vector::Zero(2, 1);
        this->limits(0) = minDiffRotErr;
        this->limits(1) = minDiffTransErr;
        this->conditionVariableNames.push_back("Mean abs differential rot err");
        this->conditionVariableNames.push_back("Mean abs differential trans err");
        this->limitNames.push_back("Min differential rotation err");
        this->limitNames.push_back("Min differential translation err");
    }
    void init(const TransformationParameters& parameters, bool&) override {
        this->conditionVariables = Vector::Zero(2, 1);
        rotations.clear();
        translations.clear();
        if (parameters.rows() == 4) {
            rotations.push_back(Quat::fromMatrix(parameters));
        } else {  // the 2-D case: the 2x2 rotation embedded in an identity (TransformationCheckersImpl.cpp:116-121)
            TransformationParameters m = Matrix::Identity(3, 3);
            for (int c = 0; c < 2; ++c)
                for (int r = 0; r < 2; ++r) m(r, c) = parameters(r, c);
            rotations.push_back(Quat::fromMatrix(m));
        }
        translations.push_back(translationOf(parameters));
    }
    static std::vector<T> translationOf(const TransformationParameters& p) {
        const int d = p.rows();
        return {p(0, d - 1), p(1, d - 1), d == 4 ? p(2, 3) : T(0)};
    }
    void check(const TransformationParameters& parameters, bool& iterate) override {
        // 2-D: the reference builds the quaternion from topLeftCorner(3, 3) of the 3x3 homogeneous matrix, translation column
        // included (TransformationCheckersImpl.cpp:131); fromMatrix reads exactly those nine entries
        rotations.push_back(Quat::fromMatrix(parameters));
        translations.push_back(translationOf(parameters));
        this->conditionVariables = Vector::Zero(2, 1);
        if (rotations.size() > smoothLength) {
            for (size_t i = rotations.size() - 1; i >= rotations.size() - smoothLength; --i) {
                this->conditionVariables(0) += std::fabs(rotations[i].angularDistance(rotations[i - 1]));
                const T dx = translations[i][0] - translations[i - 1][0], dy = translations[i][1] - translations[i - 1][1], dz = translations[i][2] - translations[i - 1][2];
                this->conditionVariables(1) += std::fabs(std::sqrt(dx * dx + dy * dy + dz * dz));
            }
            this->conditionVariables(0) /= T(smoothLength);
            this->conditionVariables(1) /= T(smoothLength);
            if (this->conditionVariables(0) < this->limits(0) && this->conditionVariables(1) < this->limits(1)) iterate = false;
        }
        if (std::isnan(this->conditionVariables(0))) throw ConvergenceError("abs rotation norm not a number");
        if (std::isnan(this->conditionVariables(1))) throw ConvergenceError("abs translation norm not a number");
    }
};
struct BoundTransformationChecker : public TransformationChecker {
    static const std::string description() { return "This checker stops the ICP loop with an exception when the transformation values exceed bounds."; }
    static const ParametersDoc availableParameters() {
        return {{"maxRotationNorm", "rotation bound", "1", "0", "inf", &Parametrizable::Comp<T>},
                {"maxTranslationNorm", "translation bound", "1", "0", "inf", &Parametrizable::Comp<T>}};
    }
    const T maxRotationNorm, maxTranslationNorm;
    Quat initialRotation3D;
    T initialRotation2D = T(0);
    std::vector<T> initialTranslation;
    BoundTransformationChecker(const Parameters& params = Parameters())
        : TransformationChecker("BoundTransformationChecker", availableParameters(), params), maxRotationNorm(Parametrizable::get<T>("maxRotationNorm")),
          maxTranslationNorm(Parametrizable::get<T>("maxTranslationNorm")) {
        this->limits = Vector::Zero(2, 1);
        this->limits(0) = maxRotationNorm;
        this->limits(1) = maxTranslationNorm;
        this->limitNames.push_back("Max rotation angle");
        this->limitNames.push_back("Max translation norm");
        this->conditionVariableNames.push_back("Rotation angle");
        this->conditionVariableNames.push_back("Translation norm");
    }
    void init(const TransformationParameters& parameters, bool&) override {
        this->conditionVariables = Vector::Zero(2, 1);
        if (parameters.rows() == 4) initialRotation3D = Quat::fromMatrix(parameters);
        else if (parameters.rows() == 3) initialRotation2D = std::acos(parameters(0, 0));
        else throw std::runtime_error("BoundTransformationChecker only works in 2D or 3D");
        initialTranslation = DifferentialTransformationChecker::translationOf(parameters);
    }
    void check(const TransformationParameters& parameters, bool&) override {
        if (parameters.rows() == 4) {
            this->conditionVariables(0) = Quat::fromMatrix(parameters).angularDistance(initialRotation3D);
        } else {  // TransformationCheckersImpl.cpp:207-211, normalizeAngle :229-236
            T a = std::acos(parameters(0, 0)) - initialRotation2D;
            while (a > T(M_PI)) a -= T(2 * M_PI);
            while (a < T(-M_PI)) a += T(2 * M_PI);
            this->conditionVariables(0) = a;
        }
        const std::vector<T> tr = DifferentialTransformationChecker::translationOf(parameters);
        const T dx = tr[0] - initialTranslation[0], dy = tr[1] - initialTranslation[1], dz = tr[2] - initialTranslation[2];
        this->conditionVariables(1) = std::sqrt(dx * dx + dy * dy + dz * dz);
        if (this->conditionVariables(0) > this->limits(0) || this->conditionVariables(1) > this->limits(1)) {
            std::ostringstream oss;
            oss << "limit out of bounds: rot: " << this->conditionVariables(0) << "/" << this->limits(0) << " tr: " << this->conditionVariables(1) << "/" << this->limits(1);
            throw ConvergenceError(oss.str());
        }
    }
};

struct NullInspector : public Inspector {
    static const std::string description() { return "Does nothing."; }
    NullInspector() : Inspector("NullInspector", ParametersDoc(), Parameters()) {}
    bool isNull() const override { return true; }
};
// PerformanceInspector (InspectorsImpl.h:66-98, InspectorsImpl.cpp:61-103): collects the statistics ICP reports through addStat; it
// has no use for the per-iteration data (Inspector::dumpIteration stays the empty base version), so the loop stays fused
struct PerformanceInspector : public Inspector {
    static const std::string description() { return "Keep statistics on performance."; }
    static const ParametersDoc availableParameters() {
        return {{"baseFileName", "base file name for the statistics files (if empty, disabled)", ""},
                {"dumpPerfOnExit", "dump performance statistics to stderr on exit", "0"},
                {"dumpStats", "dump the statistics on first and last step", "0"}};
    }
    const std::string baseFileName;
    const bool bDumpPerfOnExit, bDumpStats;
    std::map<std::string, std::vector<double>> stats;
    PerformanceInspector(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params)
        : Inspector(className, paramsDoc, params), baseFileName(Parametrizable::get<std::string>("baseFileName")),
          bDumpPerfOnExit(Parametrizable::get<bool>("dumpPerfOnExit")), bDumpStats(Parametrizable::get<bool>("dumpStats")) {}
    PerformanceInspector(const Parameters& params = Parameters()) : PerformanceInspector("PerformanceInspector", availableParameters(), params) {}
    void addStat(const std::string& name, double data) override {
        if (!bDumpStats) return;
        stats[name].push_back(data);
    }
    // one "name: count mean min max" group per statistic (the reference prints its Histogram class's bins; the dump format is
    // documented there as "will most probably change")
    virtual void dumpStats(std::ostream& stream) {
        bool first = true;
        for (const auto& kv : stats) {
            if (!first) stream << ", ";
            first = false;
            double sum = 0, lo = 0, hi = 0;
            for (size_t i = 0; i < kv.second.size(); ++i) {
                const double v = kv.second[i];
                sum += v;
                lo = i ? std::min(lo, v) : v;
                hi = i ? std::max(hi, v) : v;
            }
            stream << kv.first << ": " << kv.second.size() << " " << (kv.second.empty() ? 0.0 : sum / kv.second.size()) << " " << lo << " " << hi;
        }
    }
    ~PerformanceInspector() override {
        if (bDumpPerfOnExit && !stats.empty()) { dumpStats(std::cerr); std::cerr << "\n"; }
    }
    bool isNull() const override { return false; }
    bool needsIterationData() const override { return false; }
};

// VTKFileInspector (InspectorsImpl.h:100-193, InspectorsImpl.cpp:138-790), ASCII files: every iteration it may write the match
// links with their outlier weights (<base>-link-N.vtk), the reading as that iteration sees it (<base>-reading-N.vtk), the
// reference (<base>-reference-N.vtk) and one line of the checkers' condition variables and limits (<base>-iterationInfo.csv).
// Numbers are written the way Eigen's operator<< writes a matrix: stream precision, every coefficient right-aligned to the
// widest one, a blank between columns.
struct VTKFileInspector : public PerformanceInspector {
    static const std::string description() { return "Dump the different steps into VTK files."; }
    static const ParametersDoc availableParameters() {
        return {{"baseFileName", "base file name for the VTK files ", "point-matcher-output"},
                {"dumpPerfOnExit", "dump performance statistics to stderr on exit", "0"},
                {"dumpStats", "dump the statistics on first and last step", "0"},
                {"dumpIterationInfo", "dump iteration info", "0"},
                {"dumpDataLinks", "dump data links at each iteration", "0"},
                {"dumpReading", "dump the reading cloud at each iteration", "0"},
                {"dumpReference", "dump the reference cloud at each iteration", "0"},
                {"writeBinary", "write binary VTK files", "0"}};
    }
    const bool bDumpIterationInfo, bDumpDataLinks, bDumpReading, bDumpReference;
    std::unique_ptr<std::ofstream> streamIter;
    VTKFileInspector(const Parameters& params = Parameters())
        : PerformanceInspector("VTKFileInspector", availableParameters(), params), bDumpIterationInfo(Parametrizable::get<bool>("dumpIterationInfo")),
          bDumpDataLinks(Parametrizable::get<bool>("dumpDataLinks")), bDumpReading(Parametrizable::get<bool>("dumpReading")),
          bDumpReference(Parametrizable::get<bool>("dumpReference")) {
        if (Parametrizable::get<bool>("writeBinary")) throw ConfigurationError("VTKFileInspector: GPU host layer: writeBinary is not supported (ASCII files only)");
    }
    bool needsIterationData() const override { return bDumpIterationInfo || bDumpDataLinks || bDumpReading || bDumpReference; }

    // rows x cols block the way Eigen prints it: at(r, c), aligned columns
    template <typename F>
    static void writeMatrix(std::ostream& stream, int rows, int cols, F at) {
        std::vector<std::string> cells((size_t)rows * cols);
        size_t width = 0;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                std::ostringstream o;
                o.copyfmt(stream);
                o << at(r, c);
                cells[(size_t)r * cols + c] = o.str();
                width = std::max(width, o.str().size());
            }
        for (int r = 0; r < rows; ++r) {
            if (r) stream << "\n";
            for (int c = 0; c < cols; ++c) {
                if (c) stream << " ";
                const std::string& v = cells[(size_t)r * cols + c];
                stream << std::string(width - v.size(), ' ') << v;
            }
        }
    }
    static void writePoints(std::ostream& stream, const Matrix& features) {  // features.topLeftCorner(dim, n).transpose()
        const int dim = features.rows() == 4 ? 3 : features.rows();
        writeMatrix(stream, features.cols(), dim, [&](int r, int c) { return features(c, r); });
    }
    // InspectorsImpl.cpp:158-238 (ASCII branch)
    static void dumpDataPoints(const DataPoints& data, std::ostream& stream) {
        const Matrix& features = data.features;
        stream << "# vtk DataFile Version 3.0\nFile created by libpointmatcher\nASCII\nDATASET POLYDATA\n";
        stream << "POINTS " << features.cols() << " float\n";
        writePoints(stream, features);
        stream << "\n";
        stream << "VERTICES " << features.cols() << " " << features.cols() * 2 << "\n";
        for (int i = 0; i < features.cols(); ++i) stream << "1 " << i << "\n";
        stream << "POINT_DATA " << features.cols() << "\n";
        unsigned row = 0;
        for (const auto& label : data.descriptorLabels) {
            const int span = (int)label.span;
            const char* attribute = nullptr;
            int forced = 0;
            if (label.text == "normals") { attribute = "NORMALS"; forced = 3; }
            else if (label.text == "eigVectors") { attribute = "TENSORS"; forced = 9; }
            else if (label.text == "color") { attribute = "COLOR_SCALARS"; forced = 4; }
            else if (span == 1) { attribute = "SCALARS"; forced = 1; }
            else if (span == 3 || span == 2) { attribute = "VECTORS"; forced = 3; }
            if (attribute && span > 0 && span <= forced) {
                const bool color = std::string(attribute) == "COLOR_SCALARS";
                if (color) stream << attribute << " " << label.text << " " << forced << "\n";
                else {
                    stream << attribute << " " << label.text << " float\n";
                    if (std::string(attribute) == "SCALARS") stream << "LOOKUP_TABLE default\n";
                }
                const unsigned r0 = row;
                // padWithZeros / padWithOnes (colours) to the attribute's dimension, transposed: one point per line
                writeMatrix(stream, data.descriptors.cols(), forced,
                            [&](int r, int c) { return c < span ? data.descriptors(r0 + c, r) : (color ? T(1) : T(0)); });
                stream << "\n";
            }
            row += label.span;
        }
    }
    // InspectorsImpl.cpp:285-362
    static void dumpDataLinks(const DataPoints& ref, const DataPoints& reading, const Matches& matches, const OutlierWeights& w, std::ostream& stream) {
        const int refPtCount = ref.features.cols(), readingPtCount = reading.features.cols();
        stream << "# vtk DataFile Version 3.0\ncomment\nASCII\nDATASET POLYDATA\n";
        stream << "POINTS " << refPtCount + readingPtCount << " float\n";
        writePoints(stream, ref.features);
        stream << "\n";
        writePoints(stream, reading.features);
        stream << "\n";
        const int knn = matches.ids.rows();
        size_t matchCount = 0;
        for (int k = 0; k < knn; ++k)
            for (int i = 0; i < readingPtCount; ++i) matchCount += matches.ids(k, i) != Matches::InvalidId ? 1 : 0;
        stream << "LINES " << matchCount << " " << matchCount * 3 << "\n";
        for (int k = 0; k < knn; ++k)
            for (int i = 0; i < readingPtCount; ++i)
                if (matches.ids(k, i) != Matches::InvalidId) stream << "2 " << refPtCount + i << " " << matches.ids(k, i) << "\n";
        stream << "CELL_DATA " << matchCount << "\nSCALARS outlier float 1\nLOOKUP_TABLE default\n";
        for (int k = 0; k < knn; ++k)
            for (int i = 0; i < readingPtCount; ++i)
                if (matches.ids(k, i) != Matches::InvalidId) stream << w(k, i) << "\n";
    }
    std::string fileName(const std::string& role, size_t iteration) const {
        std::ostringstream oss;
        oss << this->baseFileName << "-" << role << "-" << iteration << ".vtk";
        return oss.str();
    }
    static std::unique_ptr<std::ofstream> open(const std::string& name) {
        std::unique_ptr<std::ofstream> f(new std::ofstream(name.c_str()));
        if (f->fail()) throw std::runtime_error("Couldn't open the file \"" + name + "\". Check if directory exist.");
        return f;
    }
    void init() override {
        if (!bDumpIterationInfo) return;
        streamIter = open(this->baseFileName + "-iterationInfo.csv");
    }
    // InspectorsImpl.cpp:384-450
    void dumpIteration(const size_t iterationNumber, const TransformationParameters&, const DataPoints& filteredReference, const DataPoints& reading,
                       const Matches& matches, const OutlierWeights& outlierWeights, const TransformationCheckers& transCheck) override {
        if (bDumpDataLinks) dumpDataLinks(filteredReference, reading, matches, outlierWeights, *open(fileName("link", iterationNumber)));
        if (bDumpReading) dumpDataPoints(reading, *open(fileName("reading", iterationNumber)));
        if (bDumpReference) dumpDataPoints(filteredReference, *open(fileName("reference", iterationNumber)));
        if (!bDumpIterationInfo || !streamIter) return;
        std::ostream& out = *streamIter;
        if (iterationNumber == 0) {
            for (size_t j = 0; j < transCheck.size(); ++j)
                for (size_t i = 0; i < transCheck[j]->getConditionVariableNames().size(); ++i) {
                    if (!(j == 0 && i == 0)) out << ", ";
                    out << transCheck[j]->getConditionVariableNames()[i] << ", " << transCheck[j]->getLimitNames()[i];
                }
            out << "\n";
        }
        for (size_t j = 0; j < transCheck.size(); ++j)
            for (int i = 0; i < (int)transCheck[j]->getConditionVariables().size(); ++i) {
                if (!(j == 0 && i == 0)) out << ", ";
                out << transCheck[j]->getConditionVariables()(i) << ", " << transCheck[j]->getLimits()(i);
            }
        out << "\n";
    }
    void finish(const size_t) override { streamIter.reset(); }
};

struct NullLogger : public Logger {
    static const std::string description() { return "Does nothing."; }
    NullLogger() : Logger("NullLogger", ParametersDoc(), Parameters()) {}
};
