"""The C-ABI library loads on a CPU-only box and exports every symbol include/pmgpu.h declares.
No compute call is made here (there is no CPU path to call)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pmgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from libpointmatcher_b200 import capi
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(capi.lib, n), "libpmgpu.so does not export %s" % n
    assert sorted(capi.SIGNATURES) == names, "capi.py and pmgpu.h disagree on the entry points"


def test_no_device_is_a_loud_error():
    import torch
    from libpointmatcher_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PmGpuError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_CUDA


def test_status_strings():
    from libpointmatcher_b200 import capi
    assert capi.lib.pmgpu_status_string(capi.ERR_NO_OUTLIER_TO_FILTER) == b"no outlier to filter"
    assert capi.lib.pmgpu_status_string(capi.ERR_NO_POINT_TO_MINIMIZE) == b"ErrorMnimizer: no point to minimize"
    assert capi.lib.pmgpu_status_string(capi.ERR_NO_NORMALS) == b"Field normals not found"


def test_product_never_touches_the_oracle():
    """the product package must not import, link or call anything under oracle/"""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "libpointmatcher_b200")):
        if "_build" in root:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, fn), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|orc_", text):
                    bad.append(os.path.join(root, fn))
    assert not bad, bad


def test_module_parameters_match_reference_tables():
    """names / defaults / bounds of SURVEY Appendix A"""
    from libpointmatcher_b200 import pm
    m = pm.KDTreeMatcher()
    assert (m.knn, m.epsilon, m.searchType, m.maxDist) == (1, 0.0, 1, float("inf"))
    assert pm.TrimmedDistOutlierFilter().value == 0.85
    assert pm.MedianDistOutlierFilter().value == 3.0
    assert pm.MaxDistOutlierFilter().value == 1.0
    assert pm.SurfaceNormalDataPointsFilter().knn == 5
    assert pm.CounterTransformationChecker().maxIterationCount == 40
    d = pm.DifferentialTransformationChecker()
    assert (d.minDiffRotErr, d.minDiffTransErr, d.smoothLength) == (0.001, 0.001, 3)
    with pytest.raises(pm.InvalidParameter):
        pm.KDTreeMatcher({"knn": "0"})
    with pytest.raises(pm.InvalidParameter):
        pm.KDTreeMatcher({"notAParam": "1"})
    with pytest.raises(pm.InvalidParameter):
        pm.SurfaceNormalDataPointsFilter({"knn": "2"})
    with pytest.raises(pm.ConfigurationError):
        pm.PointToPlaneErrorMinimizer({"force2D": "1", "force4DOF": "1"})
    with pytest.raises(pm.InvalidElement):
        pm.MatcherRegistrar.create("NoSuchMatcher")
