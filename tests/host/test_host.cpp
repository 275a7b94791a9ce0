// tests/host/test_host.cpp — exercises the C++ boundary classes (libpointmatcher_b200/host/).
//   test_host cpu                                  : plugin runtime only (no GPU needed)
//   test_host icp <config.yaml> <reading.f32> <nq> <reference.f32> <nr> [normals.f32]
//                                                  : runs PM::ICP like examples/icp_simple.cpp and
//                                                    prints iterations + the 4x4 (row-major)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "PointMatcher.h"

typedef PointMatcher<float> PM;
typedef PM::DataPoints DP;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

template <typename E, typename F>
static bool throws(F f) {
    try {
        f();
    } catch (const E&) {
        return true;
    } catch (...) {
        return false;
    }
    return false;
}

static int run_cpu() {
    const PM& pm = PM::get();
    // registrar: names of the reference, parameter tables, bounds
    auto m = pm.MatcherRegistrar.create("KDTreeMatcher", {{"knn", "3"}, {"maxDist", "inf"}});
    CHECK(m->className == "KDTreeMatcher");
    CHECK(m->get<int>("knn") == 3);
    CHECK(std::isinf(m->get<float>("maxDist")));
    CHECK(throws<PM::InvalidParameter>([&] { pm.MatcherRegistrar.create("KDTreeMatcher", {{"knn", "0"}}); }));
    CHECK(throws<PM::InvalidParameter>([&] { pm.MatcherRegistrar.create("KDTreeMatcher", {{"bogus", "1"}}); }));
    CHECK(throws<PM::InvalidElement>([&] { pm.MatcherRegistrar.create("NoSuchMatcher"); }));
    CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("TrimmedDistOutlierFilter", {{"ratio", "1.5"}}); }));
    CHECK(throws<PM::InvalidParameter>([&] { pm.ErrorMinimizerRegistrar.create("PointToPointErrorMinimizer", {{"x", "1"}}); }));
    CHECK(throws<PM::ConfigurationError>([&] { pm.ErrorMinimizerRegistrar.create("PointToPlaneErrorMinimizer", {{"force2D", "1"}, {"force4DOF", "1"}}); }));
    CHECK(throws<PM::InvalidParameter>([&] { pm.DataPointsFilterRegistrar.create("SurfaceNormalDataPointsFilter", {{"knn", "2"}}); }));
    CHECK(pm.OutlierFilterRegistrar.getDescription("TrimmedDistOutlierFilter") == "Hard rejection threshold using quantile.");
    CHECK(pm.TransformationCheckerRegistrar.create("CounterTransformationChecker")->get<unsigned>("maxIterationCount") == 40);

    // YAML: the layout of examples/data/default.yaml
    const char* yaml =
        "readingDataPointsFilters:\n  - IdentityDataPointsFilter\n\n"
        "referenceDataPointsFilters:\n  - SurfaceNormalDataPointsFilter:\n      knn: 10\n      keepDensities: 1 # comment\n\n"
        "matcher:\n  KDTreeMatcher:\n    knn: 1\n    epsilon: 0\n\n"
        "outlierFilters:\n  - TrimmedDistOutlierFilter:\n      ratio: 0.75\n  - MaxDistOutlierFilter:\n      maxDist: 2.5\n\n"
        "errorMinimizer:\n  PointToPlaneErrorMinimizer\n\n"
        "transformationCheckers:\n  - CounterTransformationChecker:\n      maxIterationCount: 40\n  - DifferentialTransformationChecker:\n"
        "      minDiffRotErr: 0.001\n      minDiffTransErr: 0.01\n      smoothLength: 4\n\n"
        "inspector:\n  NullInspector\n\nlogger:\n  NullLogger\n";
    PM::ICP icp;
    std::istringstream in(yaml);
    icp.loadFromYaml(in);
    CHECK(icp.readingDataPointsFilters.size() == 1 && icp.referenceDataPointsFilters.size() == 1);
    CHECK(icp.matcher && icp.matcher->className == "KDTreeMatcher");
    CHECK(icp.outlierFilters.size() == 2 && icp.outlierFilters[1]->className == "MaxDistOutlierFilter");
    CHECK(icp.errorMinimizer->className == "PointToPlaneErrorMinimizer");
    CHECK(icp.transformationCheckers.size() == 2 && icp.transformationCheckers[1]->get<unsigned>("smoothLength") == 4);
    // unknown module type / unknown module / bad parameter (examples/data/unit_tests/badIcpConfig_*.yaml)
    {
        PM::ICP bad;
        std::istringstream b1("matcher:\n  KDTreeMatcher\nnotAModuleType:\n  Foo\n");
        CHECK(throws<PM::InvalidModuleType>([&] { bad.loadFromYaml(b1); }));
        std::istringstream b2("matcher:\n  NotAMatcher\n");
        CHECK(throws<PM::InvalidElement>([&] { bad.loadFromYaml(b2); }));
        std::istringstream b3("outlierFilters:\n  - TrimmedDistOutlierFilter:\n      ratio: 7\n");
        CHECK(throws<PM::InvalidParameter>([&] { bad.loadFromYaml(b3); }));
    }
    // host checkers: Counter throws at the limit, Differential needs smoothLength + 1 samples
    {
        PM::CounterTransformationChecker c(PM::Parameters{{"maxIterationCount", "2"}});
        bool it = true;
        PM::TransformationParameters I = PM::Matrix::Identity(4, 4);
        c.init(I, it);
        c.check(I, it);
        CHECK(it);
        CHECK(throws<PM::CounterTransformationChecker::MaxNumIterationsReached>([&] { c.check(I, it); }));
        PM::DifferentialTransformationChecker d(PM::Parameters{{"smoothLength", "2"}});
        it = true;
        d.init(I, it);
        d.check(I, it);
        CHECK(it);  // only 2 samples
        d.check(I, it);
        CHECK(!it);  // identical transforms -> below thresholds
        PM::BoundTransformationChecker b(PM::Parameters{{"maxTranslationNorm", "0.5"}});
        b.init(I, it);
        PM::TransformationParameters far = I;
        far(0, 3) = 1.f;
        CHECK(throws<PM::ConvergenceError>([&] { b.check(far, it); }));
    }
    // Matches::getDistsQuantile: float index arithmetic, infinite distances excluded
    {
        PM::Matches mt(1, 6);
        const float v[6] = {4, 5, 5, 5, 5, std::numeric_limits<float>::infinity()};
        for (int i = 0; i < 6; ++i) mt.dists(0, i) = v[i];
        CHECK(mt.getDistsQuantile(0.19f) == 4.f && mt.getDistsQuantile(0.5f) == 5.f && mt.getDistsQuantile(1.f) == 5.f);
    }
    // host pre-filters of the default chain: registered under the reference's names,
    // the default chain is the reference's (ICP.cpp:100-113), and the filters do on a small cloud what
    // RandomSampling.cpp / MinDist.cpp / SamplingSurfaceNormal.cpp do
    {
        PM::ICP def;
        def.setDefault();
        CHECK(def.readingDataPointsFilters.size() == 1 && def.readingDataPointsFilters[0]->className == "RandomSamplingDataPointsFilter");
        CHECK(def.referenceDataPointsFilters.size() == 1 && def.referenceDataPointsFilters[0]->className == "SamplingSurfaceNormalDataPointsFilter");
        CHECK(def.referenceDataPointsFilters[0]->get<unsigned>("knn") == 7);
        CHECK(throws<PM::InvalidParameter>([&] { pm.DataPointsFilterRegistrar.create("SamplingSurfaceNormalDataPointsFilter", {{"knn", "2"}}); }));
        CHECK(throws<PM::InvalidParameter>([&] { pm.DataPointsFilterRegistrar.create("RandomSamplingDataPointsFilter", {{"prob", "2"}}); }));
        const int n = 4000;
        DP cloud;
        cloud.features = PM::Matrix::Zero(4, n);
        unsigned s = 12345u;
        auto uni = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.f; };
        for (int i = 0; i < n; ++i) {  // a gently curved sheet: normals close to +-z
            const float x = 20.f * uni() - 10.f, y = 20.f * uni() - 10.f;
            cloud.features(0, i) = x; cloud.features(1, i) = y; cloud.features(2, i) = 0.01f * x * x + 0.002f * uni(); cloud.features(3, i) = 1.f;
        }
        cloud.featureLabels.push_back(DP::Label("x", 1)); cloud.featureLabels.push_back(DP::Label("y", 1));
        cloud.featureLabels.push_back(DP::Label("z", 1)); cloud.featureLabels.push_back(DP::Label("pad", 1));
        PM::Matrix tag(1, n);
        for (int i = 0; i < n; ++i) tag(0, i) = (float)i;
        cloud.addDescriptor("tag", tag);
        pmgpu_host_srand(1);
        DP half = pm.DataPointsFilterRegistrar.create("RandomSamplingDataPointsFilter", {{"prob", "0.5"}})->filter(cloud);
        CHECK(half.features.cols() > 0.45 * n && half.features.cols() < 0.55 * n && half.descriptors.cols() == half.features.cols());
        std::srand(1);  // the same std::rand stream decides (RandomSampling.cpp:66)
        int j = 0;
        for (int i = 0; i < n; ++i)
            if ((double)((float)std::rand() / (float)RAND_MAX) < (double)0.5f) { CHECK(half.descriptors(0, j) == (float)i); ++j; }
        CHECK(j == half.features.cols());
        DP far = pm.DataPointsFilterRegistrar.create("MinDistDataPointsFilter", {{"minDist", "5"}})->filter(cloud);
        DP near = pm.DataPointsFilterRegistrar.create("MaxDistDataPointsFilter", {{"maxDist", "5"}})->filter(cloud);
        CHECK(far.features.cols() + near.features.cols() <= n && far.features.cols() > 0 && near.features.cols() > 0);
        for (int i = 0; i < far.features.cols(); ++i)
            CHECK(std::sqrt(far.features(0, i) * far.features(0, i) + far.features(1, i) * far.features(1, i) + far.features(2, i) * far.features(2, i)) > 5.f);
        CHECK(throws<PM::InvalidParameter>([&] { pm.DataPointsFilterRegistrar.create("MinDistDataPointsFilter", {{"dim", "3"}}); }));
        auto ssn = pm.DataPointsFilterRegistrar.create("SamplingSurfaceNormalDataPointsFilter", {{"knn", "10"}, {"samplingMethod", "1"}, {"keepDensities", "1"}});
        DP bins = ssn->filter(cloud);
        CHECK(bins.features.cols() >= n / 10 && bins.features.cols() <= n / 5);
        CHECK(bins.descriptorExists("normals") && bins.descriptorExists("densities") && bins.descriptorExists("tag"));
        const unsigned nr = bins.getDescriptorStartingRow("normals");
        int upright = 0;
        for (int i = 0; i < bins.features.cols(); ++i) {
            const float nx = bins.descriptors(nr, i), ny = bins.descriptors(nr + 1, i), nz = bins.descriptors(nr + 2, i);
            CHECK(std::fabs(nx * nx + ny * ny + nz * nz - 1.f) < 1e-4f);
            if (std::fabs(nz) > 0.95f) ++upright;
            CHECK(bins.features(3, i) == 1.f);
        }
        CHECK(upright > 0.95 * bins.features.cols());
        pmgpu_host_srand(1);
        DP sub = pm.DataPointsFilterRegistrar.create("SamplingSurfaceNormalDataPointsFilter", {{"knn", "10"}, {"ratio", "0.7"}})->filter(cloud);
        CHECK(sub.features.cols() > 0.65 * n && sub.features.cols() < 0.75 * n);
        const unsigned tr = sub.getDescriptorStartingRow("tag");
        for (int i = 1; i < sub.features.cols(); ++i) CHECK(sub.descriptors(tr, i) > sub.descriptors(tr, i - 1));  // sorted by original index
    }
    // modules of SURVEY 8f rows 3-4: registered under the reference's names, parameters checked at construction
    {
        auto rob = pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"robustFct", "huber"}, {"tuning", "1.5"}, {"scaleEstimator", "mad"}, {"nbIterationForScale", "3"}});
        auto* g = dynamic_cast<PM::GpuDistOutlierFilter*>(rob.get());
        CHECK(g && g->filterType == PMGPU_ROBUST_WORD(PMGPU_ROBUST_HUBER, PMGPU_SCALE_MAD, 3) && g->value == 1.5f);
        CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"robustFct", "nope"}}); }));
        CHECK(!throws<PM::ConfigurationError>([&] { pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"scaleEstimator", "berg"}}); }));
        CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"scaleEstimator", "bogus"}}); }));
        CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"distanceType", "point2line"}}); }));
        {
            auto pp = pm.OutlierFilterRegistrar.create("RobustOutlierFilter", {{"distanceType", "point2plane"}, {"scaleEstimator", "berg"}});
            CHECK(dynamic_cast<PM::GpuDistOutlierFilter*>(pp.get())->filterType == (PMGPU_ROBUST_WORD(PMGPU_ROBUST_CAUCHY, PMGPU_SCALE_BERG, 0) | PMGPU_ROBUST_P2PLANE));
        }
        CHECK(pm.OutlierFilterRegistrar.create("SurfaceNormalOutlierFilter", {{"maxAngle", "0.42"}})->get<float>("maxAngle") == 0.42f);
        {   // a NullOutlierFilter is a factor of one: a chain that holds it is still a GPU chain
            PM::OutlierFilters chain;
            chain.push_back(pm.OutlierFilterRegistrar.create("NullOutlierFilter"));
            chain.push_back(pm.OutlierFilterRegistrar.create("TrimmedDistOutlierFilter", {{"ratio", "0.7"}}));
            CHECK(chain.allGpu());
            auto md = pm.OutlierFilterRegistrar.create("MinDistOutlierFilter", {{"minDist", "0.2"}});
            CHECK(dynamic_cast<PM::GpuDistOutlierFilter*>(md.get())->filterType == PMGPU_FILTER_MINDIST && md->get<float>("minDist") == 0.2f);
            CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("MinDistOutlierFilter", {{"minDist", "0"}}); }));
        }
        auto var = pm.MatcherRegistrar.create("KDTreeVarDistMatcher", {{"knn", "3"}, {"maxDistField", "radius"}});
        auto* vm = dynamic_cast<PM::KDTreeMatcher*>(var.get());
        CHECK(vm && vm->maxDistField == "radius" && vm->maxDist == -1.f && vm->knn == 3);
        CHECK(throws<PM::InvalidParameter>([&] { pm.MatcherRegistrar.create("KDTreeVarDistMatcher", {{"maxDist", "1"}}); }));
        CHECK(pm.ErrorMinimizerRegistrar.create("PointToPointSimilarityErrorMinimizer")->className == "PointToPointSimilarityErrorMinimizer");
        auto p4 = pm.ErrorMinimizerRegistrar.create("PointToPlaneErrorMinimizer", {{"force4DOF", "1"}});
        CHECK(dynamic_cast<PM::GpuErrorMinimizer*>(p4.get())->kind == (PMGPU_MIN_P2PLANE | PMGPU_MIN_FORCE4DOF));
        CHECK(pm.ErrorMinimizerRegistrar.create("PointToPlaneErrorMinimizer", {{"force2D", "1"}})->get<bool>("force2D"));
        CHECK(throws<PM::ConfigurationError>([&] { pm.ErrorMinimizerRegistrar.create("PointToPlaneWithCovErrorMinimizer", {{"force2D", "1"}}); }));
        {   // DataPoints::concatenate (DataPoints.cpp:225-330): common descriptors only, this cloud's order
            DP a, b;
            a.features = PM::Matrix::Constant(4, 2, 1);
            b.features = PM::Matrix::Constant(4, 3, 2);
            a.addDescriptor("normals", PM::Matrix::Constant(3, 2, 5));
            a.addDescriptor("densities", PM::Matrix::Constant(1, 2, 6));
            b.addDescriptor("densities", PM::Matrix::Constant(1, 3, 7));
            b.addDescriptor("color", PM::Matrix::Constant(4, 3, 8));
            a.concatenate(b);
            CHECK(a.features.cols() == 5 && a.features(0, 1) == 1 && a.features(3, 4) == 2);
            CHECK(a.descriptorLabels.size() == 1 && a.descriptorExists("densities") && !a.descriptorExists("normals"));
            CHECK(a.descriptors.rows() == 1 && a.descriptors(0, 1) == 6 && a.descriptors(0, 2) == 7);
            DP c2;
            c2.features = PM::Matrix::Zero(3, 1);
            CHECK(throws<DP::InvalidField>([&] { a.concatenate(c2); }));
            DP d1, d2;
            d1.features = PM::Matrix::Zero(4, 1);
            d2.features = PM::Matrix::Zero(4, 1);
            d1.addDescriptor("normals", PM::Matrix::Zero(3, 1));
            d2.addDescriptor("normals", PM::Matrix::Zero(2, 1));
            CHECK(throws<DP::InvalidField>([&] { d1.concatenate(d2); }));
        }
        {   // ErrorElements (ErrorMinimizer.cpp:58-193) and the residuals derived from them, host arguments only
            DP rd, rf;
            rd.features = PM::Matrix::Zero(4, 3);
            rf.features = PM::Matrix::Zero(4, 2);
            for (int j = 0; j < 3; ++j) { rd.features(0, j) = float(j); rd.features(3, j) = 1; }
            rf.features(0, 0) = 0.5f; rf.features(1, 0) = 2; rf.features(3, 0) = 1;
            rf.features(0, 1) = 5;    rf.features(2, 1) = 1; rf.features(3, 1) = 1;
            PM::Matrix nrm = PM::Matrix::Zero(3, 2);
            nrm(1, 0) = 1; nrm(2, 1) = 1;
            rf.addDescriptor("normals", nrm);
            PM::Matches m(2, 3);
            PM::OutlierWeights w(2, 3);
            const float inf = std::numeric_limits<float>::infinity();
            const int ids[6] = {0, 1, 1, 0, 0, -1};
            const float dd[6] = {1, 2, 3, 4, 5, inf}, ww[6] = {1, 0, 0.5f, 1, 0, 0};
            for (int i = 0; i < 6; ++i) { m.ids(i) = ids[i]; m.dists(i) = dd[i]; w(i) = ww[i]; }
            const PM::ErrorMinimizer::ErrorElements ee(rd, rf, w, m);
            CHECK(ee.reading.features.cols() == 3 && ee.nbRejectedMatches == 2 && ee.nbRejectedPoints == 1);
            CHECK(ee.matches.ids(0, 0) == 0 && ee.matches.ids(0, 1) == 1 && ee.matches.ids(0, 2) == 0 && ee.weights(0, 1) == 0.5f);
            CHECK(ee.reading.features(0, 1) == 1 && ee.reading.features(0, 2) == 1 && ee.reference.features(0, 1) == 5);
            CHECK(ee.reference.descriptors(2, 1) == 1 && ee.pointUsedRatio == 0.5f && std::fabs(ee.weightedPointUsedRatio - 2.5f / 6) < 1e-7f);
            auto p2p = pm.ErrorMinimizerRegistrar.create("PointToPointErrorMinimizer");
            const float r1 = std::sqrt(0.25f + 4.f), r2 = std::sqrt(16.f + 1.f), r3 = std::sqrt(0.25f + 4.f);
            CHECK(std::fabs(p2p->getResidualError(rd, rf, w, m) - (r1 + r2 + r3)) < 1e-5f);
            auto p2l = pm.ErrorMinimizerRegistrar.create("PointToPlaneErrorMinimizer");
            CHECK(std::fabs(p2l->getResidualError(rd, rf, w, m) - (1 * 4.f + 0.5f * 1.f + 1 * 4.f)) < 1e-5f);
            auto p2d = pm.ErrorMinimizerRegistrar.create("PointToPlaneErrorMinimizer", {{"force2D", "1"}});
            CHECK(std::fabs(p2d->getResidualError(rd, rf, w, m) - (4.f + 0.f + 4.f)) < 1e-5f);
            PM::OutlierWeights none = PM::OutlierWeights::Zero(2, 3);
            CHECK(throws<PM::ConvergenceError>([&] { PM::ErrorMinimizer::ErrorElements bad(rd, rf, none, m); }));
        }
        auto vt = pm.OutlierFilterRegistrar.create("VarTrimmedDistOutlierFilter", {{"minRatio", "0.1"}, {"lambda", "1.5"}});
        CHECK(vt->get<float>("lambda") == 1.5f && vt->get<float>("maxRatio") == 0.99f);
        CHECK(throws<PM::InvalidParameter>([&] { pm.OutlierFilterRegistrar.create("VarTrimmedDistOutlierFilter", {{"minRatio", "0.9"}, {"maxRatio", "0.5"}}); }));
        DP c;
        c.features = PM::Matrix::Zero(4, 2);
        c.features(0, 0) = 1.f; c.features(1, 1) = 2.f; c.features(3, 0) = c.features(3, 1) = 1.f;
        PM::Matrix nrm = PM::Matrix::Zero(3, 2);
        nrm(0, 0) = 1.f; nrm(1, 1) = 1.f;  // both point away from the sensor at the origin
        c.addDescriptor("normals", nrm);
        pm.DataPointsFilterRegistrar.create("ObservationDirectionDataPointsFilter")->inPlaceFilter(c);
        const unsigned ro = c.getDescriptorStartingRow("observationDirections"), rn = c.getDescriptorStartingRow("normals");
        CHECK(c.descriptors(ro, 0) == -1.f && c.descriptors(ro + 1, 1) == -2.f);
        pm.DataPointsFilterRegistrar.create("OrientNormalsDataPointsFilter")->inPlaceFilter(c);
        CHECK(c.descriptors(rn, 0) == -1.f && c.descriptors(rn + 1, 1) == -1.f);  // flipped toward the sensor
        pm.DataPointsFilterRegistrar.create("OrientNormalsDataPointsFilter", {{"towardCenter", "0"}})->inPlaceFilter(c);
        CHECK(c.descriptors(rn, 0) == 1.f && c.descriptors(rn + 1, 1) == 1.f);
        DP bare;
        bare.features = c.features;
        CHECK(throws<DP::InvalidField>([&] { pm.DataPointsFilterRegistrar.create("OrientNormalsDataPointsFilter")->inPlaceFilter(bare); }));
    }
    // loaders (IO.cpp:376-392, 535-760, 949-1250): the layouts of examples/data/*.csv and *.vtk
    {
        const std::string dir = std::getenv("PM_TEST_TMP") ? std::getenv("PM_TEST_TMP") : "/tmp";
        {
            std::ofstream f(dir + "/pm_host_test.csv");
            f << "x,y,z,nx,ny,nz,intensity\n-3.5376 , 0.639553 , -1.37678 , -0.060181,0.173729,0.982953, 7\n1 , 2 , 3 , 0,0,1, 9\n";
        }
        DP c = DP::load(dir + "/pm_host_test.csv");
        CHECK(c.features.rows() == 4 && c.features.cols() == 2 && c.features(0, 0) == -3.5376f && c.features(2, 1) == 3.f && c.features(3, 0) == 1.f);
        CHECK(c.descriptorExists("normals") && c.getDescriptorDimension("normals") == 3 && c.descriptorExists("intensity"));
        CHECK(c.descriptors(c.getDescriptorStartingRow("normals") + 2, 0) == 0.982953f && c.descriptors(c.getDescriptorStartingRow("intensity"), 1) == 9.f);
        {
            std::ofstream f(dir + "/pm_host_test2.csv");
            f << "0.5 1.5\n2.5 3.5\n";
        }
        DP c2 = DP::load(dir + "/pm_host_test2.csv");  // no header, two columns: a 2-D cloud
        CHECK(c2.features.rows() == 3 && c2.features.cols() == 2 && c2.features(1, 1) == 3.5f && c2.features(2, 0) == 1.f);
        {
            std::ofstream f(dir + "/pm_host_test.vtk");
            f << "# vtk DataFile Version 3.0\ndata\nASCII\nDATASET POLYDATA\nPOINTS 3 float\n-3.79521 0.0131325 -0.95763 \n1 2 3 \n4 5 6 \n"
                 "VERTICES 3 6\n1 0\n1 1\n1 2\nPOINT_DATA 3\nNORMALS normals float\n0 0 1\n0 1 0\n1 0 0\nSCALARS densities float 1\nLOOKUP_TABLE default\n0.5\n1.5\n2.5\n";
        }
        DP v = DP::load(dir + "/pm_host_test.vtk");
        CHECK(v.features.rows() == 4 && v.features.cols() == 3 && v.features(0, 0) == -3.79521f && v.features(2, 2) == 6.f && v.features(3, 1) == 1.f);
        CHECK(v.descriptorExists("normals") && v.descriptors(v.getDescriptorStartingRow("normals") + 1, 1) == 1.f);
        CHECK(v.descriptorExists("densities") && v.descriptors(v.getDescriptorStartingRow("densities"), 2) == 2.5f);
        CHECK(throws<std::runtime_error>([&] { DP::load(dir + "/does_not_exist.vtk"); }));
        CHECK(throws<std::runtime_error>([&] { DP::load(dir + "/pm_host_test.xyz"); }));
    }
    // rigid transformation: non-orthogonal matrices are rejected, correctParameters is idempotent
    // (utest/ui/Transformations.cpp:40-131)
    {
        PM::RigidTransformation rigid;
        PM::TransformationParameters Tm = PM::Matrix::Identity(4, 4);
        Tm(0, 0) = 1.2f;
        CHECK(!rigid.checkParameters(Tm));
        DP cloud;
        cloud.features = PM::Matrix::Constant(4, 3, 1.f);
        CHECK(throws<PM::TransformationError>([&] { rigid.compute(cloud, Tm); }));
        const PM::TransformationParameters c1 = rigid.correctParameters(Tm), c2 = rigid.correctParameters(c1);
        CHECK(rigid.checkParameters(c1));
        for (size_t i = 0; i < c1.size(); ++i) CHECK(std::fabs(c1(i) - c2(i)) < 1e-6f);
    }
    // double / 2-D: explicit ConfigurationError, never silently wrong
    {
        PointMatcher<double>::DataPoints d;
        d.features = PointMatcher<double>::Matrix::Constant(4, 2, 1.0);
        PointMatcher<double>::KDTreeMatcher km;
        CHECK(throws<PointMatcher<double>::ConfigurationError>([&] { km.init(d); }));
    }
    {
        // VTKFileInspector (InspectorsImpl.cpp:158-450, ASCII): the files of one iteration, byte for byte as the reference's
        // stream operators lay them out (Eigen's aligned columns), on a hand-made case
        const char* tmp = std::getenv("PM_TEST_TMP");
        const std::string base = std::string(tmp ? tmp : "/tmp") + "/pm_vtk_test";
        auto insp = pm.InspectorRegistrar.create("VTKFileInspector", {{"baseFileName", base}, {"dumpIterationInfo", "1"}, {"dumpDataLinks", "1"},
                                                                      {"dumpReading", "1"}, {"dumpReference", "1"}});
        CHECK(!insp->isNull() && insp->needsIterationData());
        CHECK(throws<PM::ConfigurationError>([&] { pm.InspectorRegistrar.create("VTKFileInspector", {{"writeBinary", "1"}}); }));
        auto perf = pm.InspectorRegistrar.create("PerformanceInspector", {{"dumpStats", "1"}});
        CHECK(!perf->isNull() && !perf->needsIterationData());  // statistics only: the loop stays on the device
        DP ref, rd;
        ref.features = PM::Matrix(4, 3);
        const float rp[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 2.5f, 0}};
        for (int j = 0; j < 3; ++j) { for (int i = 0; i < 3; ++i) ref.features(i, j) = rp[j][i]; ref.features(3, j) = 1; }
        PM::Matrix nrm(3, 3);
        for (int j = 0; j < 3; ++j) { nrm(0, j) = 0; nrm(1, j) = 0; nrm(2, j) = 1; }
        ref.addDescriptor("normals", nrm);
        rd.features = PM::Matrix(4, 2);
        const float qp[2][3] = {{0.1f, 0, 0}, {1, 0.25f, 0}};
        for (int j = 0; j < 2; ++j) { for (int i = 0; i < 3; ++i) rd.features(i, j) = qp[j][i]; rd.features(3, j) = 1; }
        PM::Matches m(1, 2);
        m.ids(0, 0) = 0; m.ids(0, 1) = 1;
        m.dists(0, 0) = 0.01f; m.dists(0, 1) = 0.0625f;
        PM::OutlierWeights w(1, 2);
        w(0, 0) = 1.f; w(0, 1) = 0.5f;
        PM::TransformationCheckers checkers;
        checkers.push_back(pm.TransformationCheckerRegistrar.create("CounterTransformationChecker", {{"maxIterationCount", "40"}}));
        bool it = true;
        checkers.init(PM::Matrix::Identity(4, 4), it);
        insp->init();
        insp->dumpIteration(0, PM::Matrix::Identity(4, 4), ref, rd, m, w, checkers);
        insp->finish(1);
        auto slurp = [](const std::string& path) {
            std::ifstream f(path);
            std::stringstream ss;
            ss << f.rdbuf();
            return ss.str();
        };
        CHECK(slurp(base + "-link-0.vtk") ==
              "# vtk DataFile Version 3.0\ncomment\nASCII\nDATASET POLYDATA\nPOINTS 5 float\n"
              "  0   0   0\n  1   0   0\n  0 2.5   0\n"
              " 0.1    0    0\n   1 0.25    0\n"
              "LINES 2 6\n2 3 0\n2 4 1\nCELL_DATA 2\nSCALARS outlier float 1\nLOOKUP_TABLE default\n1\n0.5\n");
        CHECK(slurp(base + "-reading-0.vtk") ==
              "# vtk DataFile Version 3.0\nFile created by libpointmatcher\nASCII\nDATASET POLYDATA\nPOINTS 2 float\n"
              " 0.1    0    0\n   1 0.25    0\nVERTICES 2 4\n1 0\n1 1\nPOINT_DATA 2\n");
        CHECK(slurp(base + "-reference-0.vtk") ==
              "# vtk DataFile Version 3.0\nFile created by libpointmatcher\nASCII\nDATASET POLYDATA\nPOINTS 3 float\n"
              "  0   0   0\n  1   0   0\n  0 2.5   0\nVERTICES 3 6\n1 0\n1 1\n1 2\nPOINT_DATA 3\nNORMALS normals float\n0 0 1\n0 0 1\n0 0 1\n");
        const std::string info = slurp(base + "-iterationInfo.csv");
        CHECK(info.find(", ") != std::string::npos && std::count(info.begin(), info.end(), '\n') == 2 && info.find("40") != std::string::npos);
    }
    std::printf("host cpu tests ok\n");
    return 0;
}

static bool read_f32(const char* path, size_t count, float* dst) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.read(reinterpret_cast<char*>(dst), count * sizeof(float));
    return (size_t)f.gcount() == count * sizeof(float);
}

static int run_icp(int argc, char** argv) {
    if (argc < 7) return 2;
    const int nq = std::atoi(argv[4]), nr = std::atoi(argv[6]);
    DP reading, reference;
    reading.features = PM::Matrix(4, nq);
    reference.features = PM::Matrix(4, nr);
    if (!read_f32(argv[3], 4 * (size_t)nq, reading.features.data()) || !read_f32(argv[5], 4 * (size_t)nr, reference.features.data())) {
        std::fprintf(stderr, "cannot read the clouds\n");
        return 2;
    }
    if (argc > 7) {
        PM::Matrix normals(3, nr);
        if (!read_f32(argv[7], 3 * (size_t)nr, normals.data())) return 2;
        reference.addDescriptor("normals", normals);
    }
    PM::ICP icp;
    // PM_TEST_INSPECTOR: an inspector that is not the NullInspector sees every iteration (ICP.cpp:403-405) — matches, weights and
    // the clouds of that iteration materialised on the host for it, and only for it
    struct Recorder : public PM::Inspector {
        int calls = 0, bad = 0;
        size_t lastIteration = 0;
        double kept = 0;
        void dumpIteration(const size_t it, const PM::TransformationParameters&, const DP& ref, const DP& rd, const PM::Matches& m, const PM::OutlierWeights& w,
                           const PM::TransformationCheckers&) override {
            ++calls;
            lastIteration = it;
            kept = 0;
            // the match distances are the squared distances between the clouds the inspector is shown, bit for bit
            for (int j = 0; j < m.ids.cols(); j += 97) {
                const int id = m.ids(0, j);
                if (id < 0) continue;
                const float dx = rd.features(0, j) - ref.features(0, id), dy = rd.features(1, j) - ref.features(1, id), dz = rd.features(2, j) - ref.features(2, id);
                volatile float a = dx * dx, b = dy * dy, c = dz * dz;
                volatile float ab = a + b;
                const float d2 = ab + c;
                if (d2 != m.dists(0, j)) ++bad;
            }
            for (int j = 0; j < w.cols(); ++j) kept += w(0, j);
        }
    };
    auto recorder = std::make_shared<Recorder>();
    if (std::string(argv[2]) == "default") {
        icp.setDefault();
    } else {
        std::ifstream cfg(argv[2]);
        if (!cfg) return 2;
        icp.loadFromYaml(cfg);
    }
    try {
        if (std::getenv("PM_TEST_SEQUENCE")) {
            // ICPSequence: the map is indexed once, the reading is registered against it twice
            PM::ICPSequence seq;
            if (std::string(argv[2]) == "default") seq.setDefault();
            else { std::ifstream cfg2(argv[2]); seq.loadFromYaml(cfg2); }
            if (seq.hasMap()) return 4;
            const PM::TransformationParameters I = seq(reading);  // no map yet: identity
            if (!(I == PM::Matrix::Identity(4, 4))) return 4;
            if (!seq.setMap(reference) || !seq.hasMap()) return 4;
            const PM::TransformationParameters T1 = seq(reading), T2 = seq(reading);
            if (!(T1 == T2)) return 5;  // the resident map is not disturbed by a registration
            std::printf("iterations %zu fused %d maxreached %d overlap %.9g\n", seq.getIterationCount(), seq.usedFusedLoop() ? 1 : 0,
                        seq.getMaxNumIterationsReached() ? 1 : 0, (double)seq.errorMinimizer->getWeightedPointUsedRatio());
            for (int i = 0; i < 4; ++i) std::printf("%.9g %.9g %.9g %.9g\n", T1(i, 0), T1(i, 1), T1(i, 2), T1(i, 3));
            std::printf("cov 0 0 0 0 0 0\n");
            return 0;
        }
        if (std::getenv("PM_TEST_INSPECTOR")) icp.inspector = recorder;
        const PM::TransformationParameters Tm = icp(reading, reference);
        if (std::getenv("PM_TEST_INSPECTOR"))
            std::printf("inspector calls %d last %zu mismatches %d kept %.9g\n", recorder->calls, recorder->lastIteration, recorder->bad, recorder->kept);
        std::printf("iterations %zu fused %d maxreached %d overlap %.9g\n", icp.getIterationCount(), icp.usedFusedLoop() ? 1 : 0,
                    icp.getMaxNumIterationsReached() ? 1 : 0, (double)icp.errorMinimizer->getWeightedPointUsedRatio());
        for (int i = 0; i < 4; ++i) std::printf("%.9g %.9g %.9g %.9g\n", Tm(i, 0), Tm(i, 1), Tm(i, 2), Tm(i, 3));
        const PM::Matrix cov = icp.errorMinimizer->getCovariance();
        std::printf("cov");
        for (int i = 0; i < 6; ++i) std::printf(" %.9g", cov(i, i));
        std::printf("\n");
        if (std::getenv("PM_TEST_ELEMENTS")) {
            // ErrorElements of the last iteration on request, and what the reference derives from them
            const auto ee = icp.errorMinimizer->getErrorElements();
            auto* gm = dynamic_cast<PM::GpuErrorMinimizer*>(icp.errorMinimizer.get());
            std::printf("elements %d rejM %d rejP %d used %.9g weighted %.9g residual %.9g noiseOverlap %.9g\n", (int)ee.reading.features.cols(), ee.nbRejectedMatches,
                        ee.nbRejectedPoints, (double)ee.pointUsedRatio, (double)ee.weightedPointUsedRatio, (double)gm->getResidualError(),
                        (double)icp.errorMinimizer->getOverlap());
        }
    } catch (const std::exception& e) {
        std::printf("exception %s\n", e.what());
        return 3;
    }
    return 0;
}

// test_host filters <features.f32> <n> <normals.f32> <densities.f32> <out.bin>
// applies each of the per-cloud host filters of the golden chain files to a fresh copy of the cloud
// (srand(1) before each) and writes, per filter: int32 count, int32 descriptor rows, features (4 x count),
// descriptors (rows x count) — compared bit for bit with the Python mirror (tests/test_prefilters.py).
static int run_filters(int argc, char** argv) {
    if (argc < 7) return 2;
    const int n = std::atoi(argv[3]);
    DP cloud;
    cloud.features = PM::Matrix(4, n);
    PM::Matrix normals(3, n), densities(1, n);
    if (!read_f32(argv[2], 4 * (size_t)n, cloud.features.data()) || !read_f32(argv[4], 3 * (size_t)n, normals.data()) ||
        !read_f32(argv[5], (size_t)n, densities.data()))
        return 2;
    cloud.addDescriptor("normals", normals);
    cloud.addDescriptor("densities", densities);
    const std::vector<std::pair<std::string, PM::Parameters>> list = {
        {"BoundingBoxDataPointsFilter", {{"xMin", "0.2"}}},
        {"BoundingBoxDataPointsFilter", {{"xMin", "-3"}, {"xMax", "2"}, {"yMin", "-1"}, {"yMax", "4"}, {"zMin", "-2"}, {"zMax", "2"}, {"removeInside", "0"}}},
        {"DistanceLimitDataPointsFilter", {{"dist", "3"}, {"removeInside", "0"}}},
        {"DistanceLimitDataPointsFilter", {{"dim", "1"}, {"dist", "-0.5"}}},
        {"FixStepSamplingDataPointsFilter", {{"startStep", "7"}, {"endStep", "3"}, {"stepMult", "0.7"}}},
        {"MaxPointCountDataPointsFilter", {{"maxCount", "500"}}},
        {"MaxPointCountDataPointsFilter", {{"maxCount", "100000"}, {"seed", "5"}}},
        {"MaxQuantileOnAxisDataPointsFilter", {{"ratio", "0.72"}}},
        {"MaxQuantileOnAxisDataPointsFilter", {{"dim", "2"}, {"ratio", "0.333"}}},
        {"RemoveNaNDataPointsFilter", {}},
        {"MaxDensityDataPointsFilter", {{"maxDensity", "0.3"}}},
        {"ShadowDataPointsFilter", {{"eps", "0.3"}}},
        {"SimpleSensorNoiseDataPointsFilter", {{"gain", "2"}}},
        {"SimpleSensorNoiseDataPointsFilter", {{"sensorType", "3"}}},
        {"SimpleSensorNoiseDataPointsFilter", {{"sensorType", "4"}}},
    };
    std::ofstream out(argv[6], std::ios::binary);
    for (const auto& item : list) {
        pmgpu_host_srand(1);
        auto f = PM::get().DataPointsFilterRegistrar.create(item.first, item.second);
        DP c = f->filter(cloud);
        if (item.first == "FixStepSamplingDataPointsFilter") c = f->filter(c);  // second call: the step has moved on
        const int32_t head[2] = {(int32_t)c.features.cols(), (int32_t)c.descriptors.rows()};
        out.write(reinterpret_cast<const char*>(head), sizeof(head));
        out.write(reinterpret_cast<const char*>(c.features.data()), sizeof(float) * 4 * (size_t)head[0]);
        out.write(reinterpret_cast<const char*>(c.descriptors.data()), sizeof(float) * (size_t)head[1] * (size_t)head[0]);
    }
    return out.good() ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "cpu") return run_cpu();
    if (argc >= 2 && std::string(argv[1]) == "icp") return run_icp(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "filters") return run_filters(argc, argv);
    std::fprintf(stderr, "usage: test_host cpu | icp <config.yaml|default> <reading.f32> <nq> <reference.f32> <nr> [normals.f32]\n");
    return 2;
}
