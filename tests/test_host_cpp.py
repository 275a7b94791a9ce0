"""The C++ boundary (libpointmatcher_b200/host/PointMatcher.h): same class names, YAML layout and
exceptions as the reference.  The CPU part checks the plugin runtime; the GPU part runs PM::ICP
the way examples/icp_simple.cpp does and compares with the oracle and with the Python mirror."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_transform_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_bin():
    src = os.path.join(ROOT, "tests", "host", "test_host.cpp")
    out = os.path.join(ROOT, "tests", "host", "_build", "test_host")
    host = os.path.join(ROOT, "libpointmatcher_b200", "host")
    lib = os.path.join(ROOT, "libpointmatcher_b200", "libpmgpu.so")
    deps = [src, lib, os.path.join(ROOT, "include", "pmgpu.h")] + [os.path.join(host, f) for f in os.listdir(host)]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-I", host, "-o", out, src, "-L", os.path.dirname(lib), "-lpmgpu",
                               "-Wl,-rpath," + os.path.dirname(lib)])
    return out


def test_plugin_runtime(host_bin):
    r = subprocess.run([host_bin, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host cpu tests ok" in r.stdout


CONFIG = """
referenceDataPointsFilters:
  - IdentityDataPointsFilter

matcher:
  KDTreeMatcher:
    knn: 1
    epsilon: 0

outlierFilters:
  - TrimmedDistOutlierFilter:
      ratio: 0.75

errorMinimizer:
  {minimizer}

transformationCheckers:
  - CounterTransformationChecker:
      maxIterationCount: {iters}
{differential}
inspector:
  NullInspector

logger:
  NullLogger
"""
DIFF = """  - DifferentialTransformationChecker:
      minDiffRotErr: 0.001
      minDiffTransErr: 0.001
      smoothLength: 3
"""
BOUND = """  - BoundTransformationChecker:
      maxRotationNorm: 0.8
      maxTranslationNorm: 5
"""


def _run_icp(host_bin, tmp_path, config, rd, rf, nrm, sequence=False, elements=False, inspector=False):
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(config)
    rd.astype(np.float32).tofile(tmp_path / "rd.f32")
    rf.astype(np.float32).tofile(tmp_path / "rf.f32")
    args = [host_bin, "icp", str(cfg), str(tmp_path / "rd.f32"), str(len(rd)), str(tmp_path / "rf.f32"), str(len(rf))]
    if nrm is not None:
        np.ascontiguousarray(nrm, np.float32).tofile(tmp_path / "nrm.f32")
        args.append(str(tmp_path / "nrm.f32"))
    env = dict(os.environ, PM_TEST_SEQUENCE="1") if sequence else (dict(os.environ, PM_TEST_ELEMENTS="1") if elements else None)
    if inspector:
        env = dict(os.environ, PM_TEST_INSPECTOR="1")
    r = subprocess.run(args, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    seen = None
    if inspector:
        tok = lines.pop(0).split()
        seen = dict(calls=int(tok[2]), last=int(tok[4]), mismatches=int(tok[6]), kept=float(tok[8]))
    head = lines[0].split()
    T = np.array([[float(x) for x in l.split()] for l in lines[1:5]], np.float32)
    out = dict(iterations=int(head[1]), fused=int(head[3]), maxreached=int(head[5]), T=T, inspector=seen)
    if elements:
        tok = lines[6].split()
        out.update({tok[i]: float(tok[i + 1]) for i in range(0, len(tok), 2)})
    return out


@pytest.mark.gpu
def test_cpp_inspector_sees_every_iteration(host_bin, tmp_path, oracle, synth):
    """Inspector::dumpIteration (ICP.cpp:403-405): a non-Null inspector switches the loop to one stage at a time and is shown, every
    iteration, the reading as that iteration sees it, the centred reference, the exact matches and the outlier weights — host copies
    made for it alone; the registration itself is the fused loop's (same transform, same iteration count)"""
    rd, rf, _ = synth.scan_pair(40000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer", iters=9, differential="")
    plain = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm)
    seen = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm, inspector=True)
    assert plain["fused"] == 1 and seen["fused"] == 0
    ins = seen["inspector"]
    assert ins["calls"] == 9 == seen["iterations"] == plain["iterations"] and ins["last"] == 8
    assert ins["mismatches"] == 0                              # dists == |stepReading - reference[id]|^2 on what it was shown
    assert abs(ins["kept"] / len(rd) - 0.75) < 0.01            # TrimmedDist 0.75 weights
    assert_transform_close(seen["T"], plain["T"], 1e-5, 1e-5)


@pytest.mark.gpu
def test_cpp_vtk_file_inspector_dumps_every_iteration(host_bin, tmp_path, oracle, synth):
    """a chain with the reference's VTKFileInspector (examples/data/default.yaml has one): one link / reading file per iteration, the
    match count of the link file = the finite matches, the same transform as with the NullInspector"""
    rd, rf, _ = synth.scan_pair(6000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    base = str(tmp_path / "dump")
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer", iters=4, differential="")
    plain = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm)
    vtk = cfg.replace("inspector:\n  NullInspector", "inspector:\n  VTKFileInspector:\n    baseFileName: %s\n    dumpDataLinks: 1\n    dumpReading: 1\n    dumpIterationInfo: 1" % base)
    assert vtk != cfg
    seen = _run_icp(host_bin, tmp_path, vtk, rd, rf, nrm)
    assert seen["fused"] == 0 and seen["iterations"] == plain["iterations"] == 4
    assert_transform_close(seen["T"], plain["T"], 1e-5, 1e-5)
    for it in range(4):
        link = open("%s-link-%d.vtk" % (base, it)).read().split("\n")
        assert link[4] == "POINTS %d float" % (len(rd) + len(rf))
        lines_at = link.index("LINES %d %d" % (len(rd), 3 * len(rd)))          # knn 1, maxDist inf: every reading point has a match
        assert link[lines_at + 1].startswith("2 %d " % len(rf))
        reading = open("%s-reading-%d.vtk" % (base, it)).read().split("\n")
        assert reading[4] == "POINTS %d float" % len(rd) and "VERTICES %d %d" % (len(rd), 2 * len(rd)) in reading
    info = open(base + "-iterationInfo.csv").read().strip().split("\n")
    assert len(info) == 5                                                       # header + one line per iteration


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["counter", "differential", "bound"])
def test_cpp_icp_matches_oracle_and_python(host_bin, tmp_path, oracle, synth, variant):
    rd, rf, T_gt = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    iters = 12 if variant == "counter" else 40
    extra = {"counter": "", "differential": DIFF, "bound": BOUND + DIFF}[variant]
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer", iters=iters, differential=extra)
    res = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm)
    diff = (0.001, 0.001, 3) if variant != "counter" else None
    ref = oracle.icp(rd, rf, ref_normals=nrm, filters=[(2, 0.75)], minimizer=1, max_iterations=iters, differential=diff, nthreads=8, acc_double=True)
    assert res["iterations"] == ref["iterations"]
    assert_transform_close(res["T"], ref["T"], 1e-5, 1e-5)
    # Counter / Differential chains run as one fused device loop; Bound needs the host checkers
    assert res["fused"] == (0 if variant == "bound" else 1)
    assert res["maxreached"] == (1 if variant == "counter" else 0)


@pytest.mark.gpu
def test_cpp_icp_default_chain(host_bin, tmp_path, synth):
    """icp.setDefault() = the reference's default chain (ICP.cpp:100-113): RandomSampling on the reading and
    SamplingSurfaceNormal on the reference run on the host (same std::rand stream in both mirrors), the loop on
    the GPU.  The C++ and the Python mirror must agree, and the chain must find the pose of the re-scan."""
    from libpointmatcher_b200 import capi, pm
    rd, rf, T_gt = synth.scan_pair(60000)
    rd.astype(np.float32).tofile(tmp_path / "rd.f32")
    rf.astype(np.float32).tofile(tmp_path / "rf.f32")
    r = subprocess.run([host_bin, "icp", "default", str(tmp_path / "rd.f32"), str(len(rd)), str(tmp_path / "rf.f32"), str(len(rf))],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    T = np.array([[float(x) for x in l.split()] for l in lines[1:5]], np.float64)
    assert np.linalg.norm(T[:3, 3] - T_gt[:3, 3]) < 0.03
    capi.lib.pmgpu_host_srand(1)   # a fresh C++ process starts from the same seed
    icp = pm.ICP()
    icp.setDefault()
    Tp = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    icp.ctx.close()
    assert int(lines[0].split()[1]) == icp.iterationCount
    assert_transform_close(T, Tp, 1e-6, 1e-6)


@pytest.mark.gpu
def test_cpp_icp_sequence_matches_plain_icp(host_bin, tmp_path, oracle, synth):
    """ICPSequence (setMap once, register twice) gives the transform ICP::operator() gives"""
    rd, rf, _ = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer", iters=15, differential="")
    plain = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm)
    seq = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm, sequence=True)
    assert seq["iterations"] == plain["iterations"] == 15
    assert (seq["T"] == plain["T"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["defaultPointToPlaneMinDistDataPointsFilter", "defaultRobustOutlierFilter", "force4DOFForPointToPlaneMinimizer",
                                  "defaultSimilarityPointToPointMinDistDataPointsFilter", "defaultMaxDensityDataPointsFilter",
                                  "defaultShadowDataPointsFilter", "defaultMaxPointCountDataPointsFilter", "defaultMaxQuantileOnAxisDataPointsFilter"])
def test_cpp_runs_reference_yaml_chain_to_golden(host_bin, tmp_path, name):
    """the reference's own chain files through the C++ mirror (`icp.loadFromYaml(ifs); icp(data, ref)`, utest/utest.cpp:81-160):
    host pre-filters, GPU loop (fused, or staged with the host Bound checker for the force4DOF chain), 3 % criterion"""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fixture.npz"))
    ref = np.ascontiguousarray(np.c_[fx["cloud0"], np.ones(len(fx["cloud0"]))].astype(np.float32))
    data = np.ascontiguousarray(np.c_[fx["cloud1"], np.ones(len(fx["cloud1"]))].astype(np.float32))
    cfg = str(fx["yaml_" + name])
    res = _run_icp(host_bin, tmp_path, cfg, data, ref, None)
    cur = res["T"].astype(np.float64) @ data.T.astype(np.float64)
    gold = fx["golden_" + name].astype(np.float64) @ data.T.astype(np.float64)
    assert np.median(np.abs(cur - gold)) / np.median(np.abs(cur)) < 0.03
    assert res["fused"] == (0 if name.startswith("force4DOF") else 1)


@pytest.mark.gpu
def test_cpp_trailing_normals_filter_matches_python(host_bin, tmp_path, synth):
    """reference filters ending in SurfaceNormalDataPointsFilter: both mirrors compute the normals on the matcher's structure
    (pmgpu_ref_set -> pmgpu_ref_compute_normals -> pmgpu_ref_center) and agree"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(60000)
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer", iters=10, differential="").replace(
        "  - IdentityDataPointsFilter\n", "  - IdentityDataPointsFilter\n  - SurfaceNormalDataPointsFilter:\n      knn: 12\n")
    res = _run_icp(host_bin, tmp_path, cfg, rd, rf, None)
    icp = pm.ICP()
    icp.loadFromYaml(cfg)
    Tp = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    icp.ctx.close()
    assert res["iterations"] == icp.iterationCount == 10 and res["fused"] == 1
    assert_transform_close(res["T"], Tp, 1e-6, 1e-6)


def test_cpp_host_filters_match_python_mirror(host_bin, tmp_path):
    """the nine per-cloud host filters of the golden chain files: C++ mirror == Python mirror, bit for bit"""
    from libpointmatcher_b200 import capi, pm
    rng = np.random.default_rng(3)
    n = 3000
    f = np.c_[rng.normal(0, 2.5, (n, 3)), np.ones(n)].astype(np.float32)
    f[3, 1] = f[77, 1] = np.nan      # not on an axis MaxQuantileOnAxis selects on: NaN has no rank
    normals = rng.normal(0, 1, (n, 3)).astype(np.float32)
    normals[::97] = 0
    dens = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    dens[::50] = dens.max()
    for name, arr in (("f", f), ("n", normals), ("d", dens)):
        arr.tofile(tmp_path / (name + ".f32"))
    out = tmp_path / "out.bin"
    subprocess.check_call([host_bin, "filters", str(tmp_path / "f.f32"), str(n), str(tmp_path / "n.f32"), str(tmp_path / "d.f32"), str(out)])
    raw = np.fromfile(out, np.uint8)
    configs = [("BoundingBoxDataPointsFilter", {"xMin": "0.2"}),
               ("BoundingBoxDataPointsFilter", {"xMin": "-3", "xMax": "2", "yMin": "-1", "yMax": "4", "zMin": "-2", "zMax": "2", "removeInside": "0"}),
               ("DistanceLimitDataPointsFilter", {"dist": "3", "removeInside": "0"}),
               ("DistanceLimitDataPointsFilter", {"dim": "1", "dist": "-0.5"}),
               ("FixStepSamplingDataPointsFilter", {"startStep": "7", "endStep": "3", "stepMult": "0.7"}),
               ("MaxPointCountDataPointsFilter", {"maxCount": "500"}),
               ("MaxPointCountDataPointsFilter", {"maxCount": "100000", "seed": "5"}),
               ("MaxQuantileOnAxisDataPointsFilter", {"ratio": "0.72"}),
               ("MaxQuantileOnAxisDataPointsFilter", {"dim": "2", "ratio": "0.333"}),
               ("RemoveNaNDataPointsFilter", {}),
               ("MaxDensityDataPointsFilter", {"maxDensity": "0.3"}),
               ("ShadowDataPointsFilter", {"eps": "0.3"}),
               ("SimpleSensorNoiseDataPointsFilter", {"gain": "2"}),
               ("SimpleSensorNoiseDataPointsFilter", {"sensorType": "3"}),
               ("SimpleSensorNoiseDataPointsFilter", {"sensorType": "4"})]
    pos = 0
    for name, params in configs:
        capi.lib.pmgpu_host_srand(1)
        flt = pm.DataPointsFilterRegistrar.create(name, params)
        c = flt.filter(pm.DataPoints(f, {"normals": normals, "densities": dens}))
        if name.startswith("FixStep"):
            c = flt.filter(c)
        count, rows = raw[pos:pos + 8].view(np.int32)
        pos += 8
        feat = raw[pos:pos + 16 * count].view(np.float32).reshape(count, 4)
        pos += 16 * count
        desc = raw[pos:pos + 4 * rows * count].view(np.float32).reshape(count, rows)
        pos += 4 * rows * count
        mine = np.concatenate(list(c.descriptors.values()), axis=1)
        assert count == len(c.features) and rows == mine.shape[1], name
        assert feat.tobytes() == c.features.tobytes(), name
        assert desc.tobytes() == np.ascontiguousarray(mine).tobytes(), name
    assert pos == len(raw)


@pytest.mark.gpu
def test_cpp_var_trimmed_and_force2d_chain_matches_oracle_and_python(host_bin, tmp_path, oracle, synth):
    """VarTrimmedDistOutlierFilter + PointToPlaneErrorMinimizer force2D from a YAML file through the C++ mirror: same answer as
    the oracle (1e-5) and as the Python mirror"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    cfg = CONFIG.format(minimizer="PointToPlaneErrorMinimizer:\n    force2D: 1", iters=10, differential="")
    cfg = cfg.replace("  - TrimmedDistOutlierFilter:\n      ratio: 0.75", "  - VarTrimmedDistOutlierFilter:\n      minRatio: 0.2\n      maxRatio: 0.95\n      lambda: 2.0")
    assert "VarTrimmedDistOutlierFilter" in cfg
    res = _run_icp(host_bin, tmp_path, cfg, rd, rf, nrm)
    oracle.set_var_trimmed_ratios(0.2, 0.95)
    try:
        ref = oracle.icp(rd, rf, ref_normals=nrm, filters=[(oracle.FILTER_VARTRIMMEDDIST, 2.0)], minimizer=oracle.MIN_P2PLANE | oracle.MIN_FORCE2D,
                         max_iterations=10, nthreads=8, acc_double=True)
    finally:
        oracle.set_var_trimmed_ratios()
    assert res["iterations"] == ref["iterations"] == 10 and res["fused"] == 1
    assert_transform_close(res["T"], ref["T"], 1e-5, 1e-5)
    icp = pm.ICP()
    icp.loadFromYaml(cfg)
    Tp = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    icp.ctx.close()
    assert_transform_close(res["T"], Tp, 1e-6, 1e-6)
    assert Tp[2, 3] == 0 and Tp[2, 2] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("minimizer", ["PointToPointErrorMinimizer", "PointToPlaneErrorMinimizer"])
def test_cpp_error_elements_residual_and_overlap_match_python(host_bin, tmp_path, oracle, synth, minimizer):
    """errorMinimizer->getErrorElements() / getOverlap() / the residual error after icp(reading, reference) through the C++ mirror:
    built on request from the resident matches, equal to the Python mirror's (whose numbers are checked against the oracle)"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(30000)
    cfg = CONFIG.format(minimizer=minimizer, iters=8, differential="")
    cfg = cfg.replace("referenceDataPointsFilters:\n  - IdentityDataPointsFilter", "readingDataPointsFilters:\n  - SimpleSensorNoiseDataPointsFilter\n\n"
                      "referenceDataPointsFilters:\n  - SimpleSensorNoiseDataPointsFilter\n  - SurfaceNormalDataPointsFilter:\n      knn: 10\n      keepDensities: 1")
    cfg = cfg.replace("    knn: 1\n", "    knn: 2\n")
    assert "keepDensities" in cfg and "knn: 2" in cfg
    res = _run_icp(host_bin, tmp_path, cfg, rd, rf, None, elements=True)
    icp = pm.ICP()
    icp.loadFromYaml(cfg)
    Tp = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    ee = icp.errorMinimizer.getErrorElements()
    assert_transform_close(res["T"], Tp, 1e-6, 1e-6)
    assert (res["elements"], res["rejM"], res["rejP"]) == (len(ee.reading.features), ee.nbRejectedMatches, ee.nbRejectedPoints)
    assert abs(res["used"] - ee.pointUsedRatio) < 1e-6 and abs(res["weighted"] - ee.weightedPointUsedRatio) < 1e-6
    residual, overlap = icp.errorMinimizer.getResidualError(), icp.errorMinimizer.getOverlap()
    icp.ctx.close()
    assert abs(res["residual"] - residual) < 1e-5 * residual
    assert abs(res["noiseOverlap"] - overlap) < 1e-6 and 0 < overlap <= 1
