// End-to-end registrations through the C++ host mirror (libpointmatcher_b200/host/PointMatcher.h), the host side the north
// star names: what bench.py reports as `e2e` when this binary is faster than the Python mirror.
//   host_e2e <reading.f32> <reference.f32> <n_reading> <n_reference> <config: c2plane|c2|c3|c4> <iterations> <reps>
// Clouds are raw float32, 4 floats per point (x y z 1) = PointMatcher's 4 x N column-major features.  Prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "PointMatcher.h"

typedef PointMatcher<float> PM;
typedef PM::DataPoints DP;

static DP load(const char* path, int n) {
    DP::Labels labels;
    labels.push_back(DP::Label("x", 1)); labels.push_back(DP::Label("y", 1)); labels.push_back(DP::Label("z", 1)); labels.push_back(DP::Label("pad", 1));
    PM::Matrix f(4, n);
    std::ifstream in(path, std::ios::binary);
    in.read(reinterpret_cast<char*>(f.data()), (std::streamsize)((size_t)n * 4 * sizeof(float)));
    if (!in) { fprintf(stderr, "short read: %s\n", path); exit(2); }
    return DP(f, labels);
}

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: host_e2e reading reference n_reading n_reference config iterations reps\n"); return 2; }
    const int nq = atoi(argv[3]), nr = atoi(argv[4]), iterations = atoi(argv[6]), reps = atoi(argv[7]);
    const std::string config = argv[5];
    DP reading = load(argv[1], nq), reference = load(argv[2], nr);
    pmgpu_host_pin(reading.features.data(), reading.features.size() * sizeof(float));
    pmgpu_host_pin(reference.features.data(), reference.features.size() * sizeof(float));
    const bool plane = config != "c2", c4 = config == "c4";
    std::string yaml;
    if (plane) yaml += "referenceDataPointsFilters:\n  - SurfaceNormalDataPointsFilter:\n      knn: 20\n\n";
    yaml += std::string("matcher:\n  KDTreeMatcher:\n    knn: ") + (c4 ? "10\n    maxDist: 2.0\n" : "1\n") + "\n";
    yaml += c4 ? "outlierFilters:\n  - MaxDistOutlierFilter:\n      maxDist: 1.0\n  - MedianDistOutlierFilter:\n      factor: 3.0\n\n"
               : "outlierFilters:\n  - TrimmedDistOutlierFilter:\n      ratio: 0.75\n\n";
    yaml += std::string("errorMinimizer:\n  ") + (c4 ? "PointToPlaneWithCovErrorMinimizer" : (plane ? "PointToPlaneErrorMinimizer" : "PointToPointErrorMinimizer")) + "\n\n";
    yaml += "transformationCheckers:\n  - CounterTransformationChecker:\n      maxIterationCount: " + std::to_string(iterations) + "\n\n";
    yaml += "inspector:\n  NullInspector\n\nlogger:\n  NullLogger\n";
    PM::ICP icp;
    std::istringstream is(yaml);
    icp.loadFromYaml(is);
    std::vector<double> secs;
    PM::TransformationParameters T;
    for (int i = 0; i <= reps; ++i) {  // the first call is the warm-up (allocations, first-use costs)
        const auto t0 = std::chrono::steady_clock::now();
        T = icp(reading, reference);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (i > 0) secs.push_back(dt);
    }
    std::sort(secs.begin(), secs.end());
    const double med = secs[secs.size() / 2];
    printf("{\"seconds_per_registration\": %.9f, \"registrations_timed\": %d, \"iterations\": %d, \"T\": [", med, (int)secs.size(), iterations);
    for (int i = 0; i < 16; ++i) printf("%s%.9g", i ? ", " : "", (double)T(i % 4, i / 4));
    printf("]}\n");
    return 0;
}
