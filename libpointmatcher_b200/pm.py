"""Python mirror of the libpointmatcher plugin surface for the GPU hot path.

Same module names, parameter names, defaults, bounds and error behaviour as the reference
(`availableParameters()` of each class), so tests and configs read like
the reference's.  Every module marshals to the C ABI (capi.py -> libpmgpu.so); nothing is
computed in Python except the 4x4 bookkeeping of ICP::compute (ICP.cpp:264-313, 345-347, 448),
which the reference also does on the host.  The C++ twin of this file is
libpointmatcher_b200/host/PointMatcher.h.
"""
import ctypes as C
import os

import numpy as np

from . import capi


# ---- exceptions (names of the reference) ----------------------------------------------------
class InvalidParameter(RuntimeError):      # Parametrizable.h:101-104
    pass


class InvalidElement(RuntimeError):        # Registrar.h:69-72
    pass


class InvalidField(RuntimeError):          # PointMatcher.h:250-253
    pass


class ConvergenceError(RuntimeError):      # PointMatcher.h:148-151
    pass


class TransformationError(RuntimeError):   # PointMatcherSupport, TransformationsImpl.cpp:62
    pass


class ConfigurationError(RuntimeError):    # PointMatcher.h:83-100
    pass


_STATUS_TO_EXC = {
    capi.ERR_UNSUPPORTED: ConfigurationError,
    capi.ERR_NO_OUTLIER_TO_FILTER: ConvergenceError,
    capi.ERR_BAD_QUANTILE: ConvergenceError,
    capi.ERR_NO_POINT_TO_MINIMIZE: ConvergenceError,
    capi.ERR_NAN: ConvergenceError,
    capi.ERR_NO_NORMALS: InvalidField,
    capi.ERR_NOT_ORTHOGONAL: TransformationError,
    capi.ERR_BAD_ARG: InvalidParameter,
}


def _translate(fn, *a, **kw):
    try:
        return fn(*a, **kw)
    except capi.PmGpuError as e:
        exc = _STATUS_TO_EXC.get(e.code, RuntimeError)
        raise exc(str(e)) from None


# ---- Parametrizable (Parametrizable.cpp:170-207) ---------------------------------------------
def _cast(kind, s):
    s = str(s)
    if kind is float:
        return float(s)  # accepts "inf", "-inf", "nan" like lexical_cast_scalar_to_string (Parametrizable.h:53-64)
    if kind is bool:
        return bool(int(float(s)))
    if kind is int:
        return int(s)
    return s


class Parametrizable:
    className = ""
    #: rows of (name, doc, default, min, max, type) — availableParameters()
    PARAMS = ()

    def __init__(self, params=None):
        params = dict(params or {})
        self.parameters = {}
        self.parametersUsed = set()
        docs = {p[0]: p for p in self.PARAMS}
        for name in params:
            if name not in docs:
                raise InvalidParameter("Parameter %s for module %s was set but is not used" % (name, self.className))
        for name, _doc, default, lo, hi, kind in self.PARAMS:
            val = str(params.get(name, default))
            if kind in (int, float) and lo is not None:
                try:
                    v = _cast(kind, val)
                except ValueError:
                    raise InvalidParameter("Value %s of parameter %s in class %s cannot be parsed" % (val, name, self.className))
                if v < _cast(kind, lo):
                    raise InvalidParameter("Value %s of parameter %s in class %s is smaller than minimum admissible value %s"
                                           % (val, name, self.className, lo))
                if v > _cast(kind, hi):
                    raise InvalidParameter("Value %s of parameter %s in class %s is larger than maximum admissible value %s"
                                           % (val, name, self.className, hi))
            self.parameters[name] = val

    def get(self, name, kind=None):
        if name not in self.parameters:
            raise InvalidParameter("Parameter %s does not exist in class %s" % (name, self.className))
        self.parametersUsed.add(name)
        kind = kind or {p[0]: p[5] for p in self.PARAMS}[name]
        return _cast(kind, self.parameters[name])

    @classmethod
    def availableParameters(cls):
        return list(cls.PARAMS)


# ---- DataPoints / Matches (PointMatcher.h:207-391) -------------------------------------------
class DataPoints:
    """features: (N, 4) float32 (x, y, z, 1), or (N, 3) (x, y, 1) for a 2-D cloud; descriptors: dict name -> (N, span) float32."""

    def __init__(self, features, descriptors=None):
        self.features = np.ascontiguousarray(features, np.float32)
        self.descriptors = dict(descriptors or {})

    def descriptorExists(self, name):
        return name in self.descriptors

    def getDescriptorViewByName(self, name):
        if name not in self.descriptors:
            raise InvalidField("Field %s not found" % name)
        return self.descriptors[name]

    def copy(self):
        return DataPoints(self.features.copy(), {k: v.copy() for k, v in self.descriptors.items()})

    def concatenate(self, dp):
        """DataPoints::concatenate (DataPoints.cpp:225-330): append dp's points; only the descriptors both clouds carry
        (same name, same dimension) survive, in this cloud's order — what align_sequence / build_map grow their map with"""
        if self.features.shape[1] != dp.features.shape[1]:
            raise InvalidField("Cannot concatenate DataPoints because the dimension of the features are not the same. Actual dimension: %d New dimension: %d"
                               % (self.features.shape[1], dp.features.shape[1]))
        merged = {}
        for name, mine in self.descriptors.items():
            if name in dp.descriptors:
                theirs = dp.descriptors[name]
                if mine.shape[1] != theirs.shape[1]:
                    raise InvalidField("The field %s has dimension %d in this, different than dimension %d in that" % (name, mine.shape[1], theirs.shape[1]))
                merged[name] = np.ascontiguousarray(np.concatenate([mine, theirs], axis=0))
        self.features = np.ascontiguousarray(np.concatenate([self.features, dp.features], axis=0))
        self.descriptors = merged

    # external column name -> (internal name, kind): IO.h:117-157
    _CSV_LABELS = dict(
        [(n, (n, "feature")) for n in ("x", "y", "z", "pad")]
        + [(n, ("normals", "descriptor")) for n in ("nx", "ny", "nz", "normal_x", "normal_y", "normal_z")]
        + [("observationDirections%d" % i, ("observationDirections", "descriptor")) for i in range(3)]
        + [(n, ("color", "descriptor")) for n in ("red", "green", "blue", "alpha")]
        + [("eigValues%d" % i, ("eigValues", "descriptor")) for i in range(3)]
        + [("eigVectors%d%s" % (i, a), ("eigVectors", "descriptor")) for i in range(3) for a in "XYZ"]
        + [("intensity", ("intensity", "descriptor"))])

    @staticmethod
    def load(fileName):
        """DataPoints::load (IO.cpp:376-392): format from the extension; the CSV and the ASCII legacy-VTK
        readers (IO.cpp:535-760, 949-1250) — the formats of the reference's example data."""
        ext = os.path.splitext(fileName)[1].lower()
        if not os.path.isfile(fileName):
            raise RuntimeError("Cannot open file " + os.path.abspath(fileName))
        if ext == ".csv":
            return DataPoints._load_csv(fileName)
        if ext == ".vtk":
            return DataPoints._load_vtk(fileName)
        raise RuntimeError('loadAnyFormat(): Unknown extension "%s" for file "%s", extension must be either ".vtk" or ".csv"' % (ext, fileName))

    @staticmethod
    def _load_csv(fileName):
        with open(fileName) as f:
            lines = [ln.strip() for ln in f.read().splitlines()]
        lines = lines[:lines.index("")] if "" in lines else lines  # the reader stops at the first empty line
        if not lines:
            raise RuntimeError("CSV parse error: empty file")

        def split(ln):
            return [t for t in ln.replace(",", " ").replace(";", " ").replace("\t", " ").split(" ") if t]

        has_header = any(c not in " ,+-.1234567890Ee" for c in lines[0])
        if has_header:
            header, rows = split(lines[0]), lines[1:]
        else:
            rows = lines
            dim = len(split(lines[0]))
            if dim not in (2, 3):
                raise RuntimeError("CSV parse error: %d columns and no header: not obvious which columns to load for x, y or z" % dim)
            header = ["x", "y", "z"][:dim]
        data = np.array([[float(t) for t in split(ln)] for ln in rows], np.float64).astype(np.float32).reshape(len(rows), len(header))
        feats, descs = [], {}
        for name in ("x", "y", "z", "pad"):  # features in the table's order, descriptors grouped by internal name
            if name in header:
                feats.append(data[:, header.index(name)])
        if "x" not in header or "y" not in header:
            raise RuntimeError("CSV parse error: no x / y column")
        for j, name in enumerate(header):
            internal, kind = DataPoints._CSV_LABELS.get(name, (name, "descriptor"))
            if name == "time":
                raise RuntimeError("CSV parse error: time columns are not supported by this reader")
            if kind == "descriptor":
                descs.setdefault(internal, []).append(data[:, j])
        if "pad" not in header:
            feats.append(np.ones(len(data), np.float32))  # homogeneous row (IO.cpp:752-757)
        return DataPoints(np.stack(feats, axis=1), {k: np.stack(v, axis=1) for k, v in descs.items()})

    @staticmethod
    def _load_vtk(fileName):
        with open(fileName, "rb") as f:
            raw = f.read()
        head = raw[:256].decode("ascii", "replace").splitlines()
        if len(head) < 4 or not head[0].startswith("# vtk DataFile Version"):
            raise RuntimeError("Header in VTK file " + fileName + " is not valid")
        if head[2].strip() != "ASCII":
            raise RuntimeError("VTK reader: only ASCII legacy files are supported here, got " + head[2].strip())
        tok = raw.decode("ascii", "replace").split("\n", 3)[3].split()
        if tok[0] != "DATASET" or tok[1] not in ("POLYDATA", "UNSTRUCTURED_GRID"):
            raise RuntimeError("Invalid data type %s in VTK file: only POLYDATA and UNSTRUCTURED_GRID are supported" % " ".join(tok[:2]))
        i, n, pts, descs = 2, None, None, {}
        while i < len(tok):
            key = tok[i]
            if key == "POINTS":
                n = int(tok[i + 1])
                pts = np.array(tok[i + 3:i + 3 + 3 * n], np.float64).astype(np.float32).reshape(n, 3)
                i += 3 + 3 * n
            elif key in ("VERTICES", "LINES", "POLYGONS", "TRIANGLE_STRIPS", "CELLS"):
                i += 3 + int(tok[i + 2])
            elif key == "CELL_TYPES":
                i += 2 + int(tok[i + 1])
            elif key == "POINT_DATA":
                if int(tok[i + 1]) != n:
                    raise RuntimeError("The number of points is greater than the amount of point data.")
                i += 2
            elif key in ("NORMALS", "VECTORS"):
                name = "normals" if key == "NORMALS" else tok[i + 1]
                descs[name] = np.array(tok[i + 3:i + 3 + 3 * n], np.float64).astype(np.float32).reshape(n, 3)
                i += 3 + 3 * n
            elif key == "SCALARS":  # SCALARS name type [numComp] / LOOKUP_TABLE default / values
                name = tok[i + 1]
                comps, j = (int(tok[i + 3]), i + 4) if tok[i + 3].isdigit() else (1, i + 3)
                if tok[j] == "LOOKUP_TABLE":
                    j += 2
                descs[name] = np.array(tok[j:j + comps * n], np.float64).astype(np.float32).reshape(n, comps)
                i = j + comps * n
            elif key == "COLOR_SCALARS":
                comps = int(tok[i + 2])
                descs["color"] = np.array(tok[i + 3:i + 3 + comps * n], np.float64).astype(np.float32).reshape(n, comps)
                i += 3 + comps * n
            else:
                raise RuntimeError("VTK reader: unsupported section " + key)
        if pts is None:
            raise RuntimeError("VTK reader: no POINTS section")
        return DataPoints(np.c_[pts, np.ones(n, np.float32)].astype(np.float32), descs)


class Matches:
    InvalidId = -1
    InvalidDist = np.float32(np.inf)

    def __init__(self, dists, ids):
        self.dists, self.ids = dists, ids

    def getDistsQuantile(self, quantile):
        """Matches.cpp:60-87 on the GPU select kernel (through a throw-away TrimmedDist filter)."""
        raise NotImplementedError("use TrimmedDistOutlierFilter / the limits returned by OutlierFilters.compute")


# ---- pipeline: the device context shared by the modules of one ICP chain ----------------------
class _Bound:
    """Modules talk to one capi.Context; a stand-alone module creates its own lazily."""
    _ctx = None

    def bind(self, ctx):
        self._ctx = ctx

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = capi.Context(0)
        return self._ctx


# ---- Matcher (MatchersImpl.h:74-103) -----------------------------------------------------------
class KDTreeMatcher(Parametrizable, _Bound):
    className = "KDTreeMatcher"
    PARAMS = (
        ("knn", "number of nearest neighbors to consider it the reference", "1", "1", "2147483647", int),
        ("epsilon", "approximation to use for the nearest-neighbor search", "0", "0", "inf", float),
        ("searchType", "Nabo search type. 0: brute force, 1: kd-tree linear heap, 2: kd-tree tree heap", "1", "0", "2", int),
        ("maxDist", "maximum distance to consider for neighbors", "inf", "0", "inf", float),
    )

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.knn = self.get("knn")
        self.epsilon = self.get("epsilon")
        self.searchType = self.get("searchType")
        self.maxDist = self.get("maxDist")
        self.visitCounter = 0

    def init(self, filteredReference):
        nrm = filteredReference.descriptors.get("normals")
        _translate(self.ctx.set_reference, filteredReference.features, nrm)
        self._ref = filteredReference

    def initCentered(self, filteredReference):
        """init() on the reference centred on its mean (the preamble of ICP::compute, ICP.cpp:291-302),
        without a host copy of the cloud.  Returns the mean."""
        nrm = filteredReference.descriptors.get("normals")
        self._ref = filteredReference
        return _translate(self.ctx.set_reference_centered, filteredReference.features, nrm)

    def findClosests(self, filteredReading, T=None):
        """Matches of T * filteredReading against the reference passed to init()."""
        ctx = self.ctx
        if getattr(ctx, "_reading_obj", None) is not filteredReading:
            _translate(ctx.set_reading, filteredReading.features)
            ctx._reading_obj = filteredReading
        ids, dists, visits = _translate(ctx.knn, T, self.knn, self.epsilon, self.maxDist)
        self.visitCounter += visits
        return Matches(dists, ids)

    def getVisitCount(self):
        return self.visitCounter

    def resetVisitCount(self):
        self.visitCounter = 0


class KDTreeVarDistMatcher(KDTreeMatcher):
    """MatchersImpl.h:105-127, MatchersImpl.cpp:105-150: like KDTreeMatcher, with one maximum search distance per reading
    point, read from the reading's `maxDistField` descriptor."""
    className = "KDTreeVarDistMatcher"
    PARAMS = KDTreeMatcher.PARAMS[:3] + (
        ("maxDistField", "descriptor field name used to set a maximum distance to consider for neighbors per point", "maxSearchDist", None, None, str),)

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.knn, self.epsilon, self.searchType = self.get("knn"), self.get("epsilon"), self.get("searchType")
        self.maxDistField = self.get("maxDistField")
        self.maxDist = -1.0   # "per-point distances" for pmgpu_knn / pmgpu_icp_params
        self.visitCounter = 0

    def uploadMaxDists(self, reading):
        _translate(self.ctx.set_reading_max_dists, reading.getDescriptorViewByName(self.maxDistField))

    def findClosests(self, filteredReading, T=None):
        ctx = self.ctx
        if getattr(ctx, "_reading_obj", None) is not filteredReading:
            _translate(ctx.set_reading, filteredReading.features)
            ctx._reading_obj = filteredReading
            self.uploadMaxDists(filteredReading)
        ids, dists, visits = _translate(ctx.knn, T, self.knn, self.epsilon, -1.0)
        self.visitCounter += visits
        return Matches(dists, ids)


# ---- OutlierFilters (OutlierFiltersImpl.h:76-139) ---------------------------------------------
class _DistFilter(Parametrizable, _Bound):
    TYPE = None
    PARAM = None

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.value = self.get(self.PARAM)

    def spec(self):
        return (self.TYPE, self.value)

    def prepare(self, ctx):
        """parameters that do not fit the (type, value) pair go to the context before the chain is evaluated"""

    def compute(self, filteredReading, filteredReference, matches):
        self.prepare(self.ctx)
        w, _ = _translate(self.ctx.weights, [self.spec()])
        return w


class NullOutlierFilter(Parametrizable, _Bound):
    """Does nothing (OutlierFiltersImpl.cpp:45-58: weights of 1): in a chain it is a factor of one, so it never reaches the device"""
    className = "NullOutlierFilter"

    def spec(self):
        return None

    def prepare(self, ctx):
        pass

    def compute(self, filteredReading, filteredReference, matches):
        return np.ones(np.asarray(matches.ids).shape, np.float32)


class MaxDistOutlierFilter(_DistFilter):
    className = "MaxDistOutlierFilter"
    TYPE, PARAM = capi.FILTER_MAXDIST, "maxDist"
    PARAMS = (("maxDist", "threshold distance (Euclidean norm)", "1", "0.0000001", "inf", float),)


class MinDistOutlierFilter(_DistFilter):
    """OutlierFiltersImpl.cpp:87-101: links shorter than the threshold are outliers"""
    className = "MinDistOutlierFilter"
    TYPE, PARAM = capi.FILTER_MINDIST, "minDist"
    PARAMS = (("minDist", "threshold distance (Euclidean norm)", "1", "0.0000001", "inf", float),)


class MedianDistOutlierFilter(_DistFilter):
    className = "MedianDistOutlierFilter"
    TYPE, PARAM = capi.FILTER_MEDIANDIST, "factor"
    PARAMS = (("factor", "points farther away factor * median will be considered outliers.", "3", "0.0000001", "inf", float),)


class TrimmedDistOutlierFilter(_DistFilter):
    className = "TrimmedDistOutlierFilter"
    TYPE, PARAM = capi.FILTER_TRIMMEDDIST, "ratio"
    PARAMS = (("ratio", "percentage to keep", "0.85", "0.0000001", "1.0", float),)


class VarTrimmedDistOutlierFilter(_DistFilter):
    """OutlierFiltersImpl.h:147-172, OutlierFiltersImpl.cpp:152-218: TrimmedDist with the ratio that minimises the
    fractional RMS distance; sorted, summed (serially, in float) and minimised on the device."""
    className = "VarTrimmedDistOutlierFilter"
    TYPE, PARAM = capi.FILTER_VARTRIMMEDDIST, "lambda"
    PARAMS = (("minRatio", "min ratio", "0.05", "0.0000001", "1", float), ("maxRatio", "max ratio", "0.99", "0.0000001", "1", float),
              ("lambda", "lambda (part of the term that balance the rmsd: 1/ratio^lambda", "2.35", None, None, float))

    def __init__(self, params=None):
        _DistFilter.__init__(self, params)
        self.minRatio, self.maxRatio = self.get("minRatio"), self.get("maxRatio")
        if self.minRatio >= self.maxRatio:
            raise InvalidParameter("VarTrimmedDistOutlierFilter: minRatio (%g) should be smaller than maxRatio (%g)" % (self.minRatio, self.maxRatio))

    def prepare(self, ctx):
        _translate(ctx.set_var_trimmed_ratios, self.minRatio, self.maxRatio)


class SurfaceNormalOutlierFilter(_DistFilter):
    """OutlierFiltersImpl.h:180-197, OutlierFiltersImpl.cpp:222-285: weight 0 where the angle between the reading's and the
    matched reference point's normal exceeds maxAngle (|n_r . n_q| < cos maxAngle); all ones when a cloud has no normals."""
    className = "SurfaceNormalOutlierFilter"
    TYPE, PARAM = capi.FILTER_SURFACENORMAL, "maxAngle"
    PARAMS = (("maxAngle", "Maximum authorised angle between the 2 surface normals (in radian)", "1.57", "0.0", "3.1416", float),)


class RobustOutlierFilter(_DistFilter):
    """OutlierFiltersImpl.h:199-262, OutlierFiltersImpl.cpp:420-598: M-estimator weights (cauchy, welsch, sc, gm, tukey,
    huber, L1, student) of e^2 = dist / scale^2, scale = sqrt(MAD), Bergstrom's decreasing estimator, sqrt(std) or 1, with the
    `approximation` cut — evaluated on the device."""
    className = "RobustOutlierFilter"
    TYPE, PARAM = capi.FILTER_ROBUST, "tuning"
    FCTS = dict(cauchy=0, welsch=1, sc=2, gm=3, tukey=4, huber=5, L1=6, student=7)
    ESTIMATORS = dict(none=0, mad=1, berg=2, std=3)
    PARAMS = (
        ("robustFct", "Type of robust function used. Available fct: 'cauchy', 'welsch', 'sc'(aka Switchable-Constraint), 'gm' (aka Geman-McClure), "
                      "'tukey', 'huber' and 'L1'. (Default: cauchy)", "cauchy", None, None, str),
        ("tuning", "Tuning parameter used to limit the influence of outliers.", "1.0", "0.0000001", "inf", float),
        ("scaleEstimator", "The scale estimator is used to convert the error distance into a Mahalanobis distance: 'none', 'mad', 'berg'", "mad", None, None, str),
        ("nbIterationForScale", "For how many iteration the 'scaleEstimator' is recalculated. 0 means at each iteration.", "0", "0", "100", int),
        ("distanceType", "Type of error distance used, either point to point ('point2point') or point to plane('point2plane').", "point2point", None, None, str),
        ("approximation", "If the matched distance is larger than this threshold, its weight will be forced to zero.", "inf", "0.0", "inf", float),
    )

    def __init__(self, params=None):
        _DistFilter.__init__(self, params)
        fct, est = self.get("robustFct"), self.get("scaleEstimator")
        if fct not in self.FCTS:
            raise InvalidParameter("Invalid robust function name.")
        if est not in self.ESTIMATORS:
            raise InvalidParameter("Invalid scale estimator name.")
        if self.get("distanceType") not in ("point2point", "point2plane"):
            raise InvalidParameter("Invalid distance type name.")
        self.approximation = self.get("approximation")
        self.word = capi.FILTER_ROBUST | (self.FCTS[fct] << 8) | (self.ESTIMATORS[est] << 16) | (self.get("nbIterationForScale") << 20)
        if self.get("distanceType") == "point2plane":
            self.word |= capi.ROBUST_P2PLANE

    def prepare(self, ctx):
        _translate(ctx.set_robust_approximation, self.approximation)

    def spec(self):
        return (self.word, self.value)


class OutlierFilters(list, _Bound):
    """OutlierFilter.cpp:63-103: product of all filters' weights; empty chain -> dist != inf."""

    def compute(self, filteredReading, filteredReference, matches):
        for f in self:
            f.prepare(self.ctx)
        w, limits = _translate(self.ctx.weights, self.specs())
        self.limits = limits
        return w

    def specs(self):
        """the (type, value) words of the chain; a NullOutlierFilter is a factor of one and contributes none"""
        return [f.spec() for f in self if f.spec() is not None]


# ---- ErrorElements, materialised on request (ErrorMinimizer.cpp:58-193) -------------------------------
def rigid_apply(T, cloud):
    """RigidTransformation::compute (TransformationsImpl.cpp:49-87) on a host cloud, float, the GEMM's left-to-right sums:
    features' = T * features; "normals" / "observationDirections" turn with the rotation block"""
    T = np.asarray(T, np.float32)
    f = cloud.features
    out = np.empty_like(f)
    dim = f.shape[1]  # 4, or 3 for a 2-D cloud
    for r in range(dim):
        acc = T[r, 0] * f[:, 0]
        for c in range(1, dim):
            acc = (acc + T[r, c] * f[:, c]).astype(np.float32)
        out[:, r] = acc
    desc = dict(cloud.descriptors)
    for name in ("normals", "observationDirections"):
        if name in desc:
            d = desc[name]
            rot = np.empty_like(d)
            for r in range(d.shape[1]):
                acc = T[r, 0] * d[:, 0]
                for c in range(1, d.shape[1]):
                    acc = (acc + T[r, c] * d[:, c]).astype(np.float32)
                rot[:, r] = acc
            desc[name] = rot
    return DataPoints(out, desc)


class ErrorElements:
    """ErrorMinimizer::ErrorElements (PointMatcher.h:507-525, ErrorMinimizer.cpp:58-193): the kept (reading point, match) pairs in
    reading order, k innermost; `reading` repeats a point once per kept match, `reference` gathers the matched columns."""

    def __init__(self, requestedPts, sourcePts, outlierWeights, matches):
        ids, dists, w = np.asarray(matches.ids), np.asarray(matches.dists), np.asarray(outlierWeights, np.float32)
        n, knn = ids.shape
        keep = (dists != Matches.InvalidDist) & (w != 0)
        if not (w != 0).any():
            raise ConvergenceError("ErrorMnimizer: no point to minimize")
        i_idx, k_idx = np.nonzero(keep)                       # row-major over (point, k): i outer, k inner
        self.reading = DataPoints(requestedPts.features[i_idx], {k: v[i_idx] for k, v in requestedPts.descriptors.items()})
        kept_ids = ids[i_idx, k_idx]
        self.reference = DataPoints(sourcePts.features[kept_ids], {k: v[kept_ids] for k, v in sourcePts.descriptors.items()})
        self.weights = w[i_idx, k_idx]
        self.matches = Matches(dists[i_idx, k_idx][:, None], kept_ids[:, None])
        self.nbRejectedMatches = int(((dists != Matches.InvalidDist) & (w == 0)).sum())
        self.nbRejectedPoints = int((~keep.any(axis=1)).sum())
        self.pointUsedRatio = float(np.float32(len(i_idx)) / np.float32(knn * n))
        self.weightedPointUsedRatio = float(np.float32(self.weights.sum(dtype=np.float64)) / np.float32(knn * n))


def _delta_norms(ee):
    d = (ee.reading.features[:, :-1] - ee.reference.features[:, :-1]).astype(np.float32)
    return _float_norm(d)


def point_to_point_residual(ee):
    """PointToPointErrorMinimizer::computeResidualError (PointToPoint.cpp:153-163): sum of |reading - reference| over the kept pairs"""
    return float(_delta_norms(ee).sum(dtype=np.float64))


def point_to_plane_residual(ee, force2D=False):
    """PointToPlaneErrorMinimizer::computeResidualError (PointToPlane.cpp:314-352): sum w (n . (reading - reference))^2"""
    dims = 2 if force2D else ee.reading.features.shape[1] - 1
    n = ee.reference.getDescriptorViewByName("normals")
    dot = np.zeros(len(n), np.float32)
    for a in range(dims):
        dot = (dot + (ee.reading.features[:, a] - ee.reference.features[:, a]).astype(np.float32) * n[:, a]).astype(np.float32)
    return float((ee.weights * (dot * dot).astype(np.float32)).astype(np.float32).sum(dtype=np.float64))


def point_to_point_overlap(ee):
    """PointToPointErrorMinimizer::getOverlap (PointToPoint.cpp:116-151); None: no sensor noise, use weightedPointUsedRatio"""
    if not ee.reading.descriptorExists("simpleSensorNoise"):
        return None
    dists = _delta_norms(ee)
    mean = np.float32(np.float32(dists.sum(dtype=np.float64)) / np.float32(len(dists)))
    return float(np.float32(int((dists < mean + ee.reading.descriptors["simpleSensorNoise"][:, 0]).sum())) / np.float32(len(dists)))


def point_to_plane_overlap(ee):
    """PointToPlaneErrorMinimizer::getOverlap (PointToPlane.cpp:369-466); None: neither cloud has sensor noise"""
    rn = ee.reading.descriptors.get("simpleSensorNoise")
    fn = ee.reference.descriptors.get("simpleSensorNoise")
    dens = ee.reference.descriptors.get("densities")
    if rn is not None and fn is not None and dens is not None:
        values = dens.reshape(-1)
        median = np.partition(values, int(len(values) * 0.5))[int(len(values) * 0.5)]
        radius = np.float32(1.0 / np.power(np.float64(median), 1 / 3.0))
        unc = ((radius + rn[:, 0]).astype(np.float32) + fn[:, 0]).astype(np.float32)
    elif rn is not None and fn is not None:
        unc = (rn[:, 0] + fn[:, 0]).astype(np.float32)
    elif rn is not None:
        unc = rn[:, 0]
    elif fn is not None:
        unc = fn[:, 0]
    else:
        return None
    dists = _delta_norms(ee)
    f = ee.reading.features
    count, unique = 0, 1
    last = f[0] * 2
    for i in range(len(f)):            # sequential by construction: "last valid point" carries from one pair to the next
        if (last != f[i]).any() and abs(dists[i]) < unc[i]:
            last = f[i]
            count += 1
        if i > 0 and (f[i] != f[i - 1]).any():
            unique += 1
    return float(np.float32(count) / np.float32(unique + ee.nbRejectedPoints))


# ---- ErrorMinimizers ---------------------------------------------------------------------------
class _Minimizer(Parametrizable, _Bound):
    KIND = None

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.sensorStdDev = self.get("sensorStdDev") if any(p[0] == "sensorStdDev" for p in self.PARAMS) else 0.01
        if any(p[0] == "force2D" for p in self.PARAMS):
            force2D, force4DOF = self.get("force2D"), self.get("force4DOF")
            if force2D and force4DOF:
                raise ConfigurationError("Force 2D cannot be used together with force4DOF.")
            if force2D and self.KIND != capi.MIN_P2PLANE:
                raise ConfigurationError("GPU module: force2D is not supported together with the covariance")
            self.force2D = bool(force2D)
            self.force4DOF = bool(force4DOF)
        self._cov = np.zeros((6, 6), np.float32)
        self._stats = dict(pointUsedRatio=-1.0, weightedPointUsedRatio=-1.0)

    def compute(self, filteredReading, filteredReference, outlierWeights, matches):
        """Uses the matches / weights resident on the device (the arguments are the host copies
        the reference interface passes around)."""
        T, cov, stats = _translate(self.ctx.minimize, self.kind_word(), self.sensorStdDev)
        if cov is not None:
            self._cov = cov
        self._stats = stats
        return T

    def kind_word(self):
        """minimiser id with the force4DOF / force2D bit (pmgpu.h)"""
        return self.KIND | (capi.MIN_FORCE4DOF if getattr(self, "force4DOF", False) else 0) | (capi.MIN_FORCE2D if getattr(self, "force2D", False) else 0)

    def getCovariance(self):
        return self._cov

    def getPointUsedRatio(self):
        return self._stats["pointUsedRatio"]

    def getWeightedPointUsedRatio(self):
        return self._stats["weightedPointUsedRatio"]

    #: set by ICP: builds the ErrorElements of the last iteration from the resident matches (downloaded on request only)
    _error_elements_provider = None

    def getErrorElements(self):
        """ErrorMinimizer::getErrorElements (ErrorMinimizer.cpp:232-236) — lazily: nothing is copied back unless this is called"""
        if self._error_elements_provider is None:
            raise RuntimeError("Error, last error element empty. Error minimizer needs to be called at least once before using this method.")
        return self._error_elements_provider()

    def getResidualError(self, *_):
        """PointToPoint.cpp:101-114 / PointToPlane.cpp:354-367; the reference's four arguments are the host copies of what is resident"""
        ee = self.getErrorElements()
        if self.KIND in (capi.MIN_P2PLANE, capi.MIN_P2PLANE_COV):
            return point_to_plane_residual(ee, getattr(self, "force2D", False))
        return point_to_point_residual(ee)

    def getOverlap(self):
        """ErrorMinimizer.cpp:227-230, PointToPoint.cpp:116-151, PointToPlane.cpp:369-466: the noise-based estimate when the clouds
        carry simpleSensorNoise, else the weighted ratio of the outlier filters"""
        if self._error_elements_provider is not None and self.KIND != capi.MIN_P2POINT_SIM:
            ee = self.getErrorElements()
            plane = self.KIND in (capi.MIN_P2PLANE, capi.MIN_P2PLANE_COV)
            ov = point_to_plane_overlap(ee) if plane else point_to_point_overlap(ee)
            if ov is not None:
                return ov
        return self._stats["weightedPointUsedRatio"]


_P2PLANE_PARAMS = (
    ("force2D", "If set to true(1), the minimization will be forced to give a solution in 2D.", "0", "0", "1", bool),
    ("force4DOF", "If set to true(1), the minimization will optimize only yaw and translation.", "0", "0", "1", bool),
)
_COV_PARAM = (("sensorStdDev", "sensor standard deviation", "0.01", "0.", "inf", float),)


class PointToPointErrorMinimizer(_Minimizer):
    className = "PointToPointErrorMinimizer"
    KIND = capi.MIN_P2POINT
    PARAMS = ()


class PointToPointWithCovErrorMinimizer(_Minimizer):
    className = "PointToPointWithCovErrorMinimizer"
    KIND = capi.MIN_P2POINT_COV
    PARAMS = _COV_PARAM


class PointToPointSimilarityErrorMinimizer(_Minimizer):
    """PointToPointSimilarity.cpp:49-101: rotation, translation and one scale factor"""
    className = "PointToPointSimilarityErrorMinimizer"
    KIND = capi.MIN_P2POINT_SIM
    PARAMS = ()


class PointToPlaneErrorMinimizer(_Minimizer):
    className = "PointToPlaneErrorMinimizer"
    KIND = capi.MIN_P2PLANE
    PARAMS = _P2PLANE_PARAMS


class PointToPlaneWithCovErrorMinimizer(_Minimizer):
    className = "PointToPlaneWithCovErrorMinimizer"
    KIND = capi.MIN_P2PLANE_COV
    PARAMS = _P2PLANE_PARAMS + _COV_PARAM


# ---- SurfaceNormalDataPointsFilter (SurfaceNormal.h:65-80) --------------------------------------
class SurfaceNormalDataPointsFilter(Parametrizable, _Bound):
    className = "SurfaceNormalDataPointsFilter"
    PARAMS = (
        ("knn", "number of nearest neighbors to consider, including the point itself", "5", "3", "2147483647", int),
        ("maxDist", "maximum distance to consider for neighbors", "inf", "0", "inf", float),
        ("epsilon", "approximation to use for the nearest-neighbor search", "0", "0", "inf", float),
        ("keepNormals", "whether the normals should be added as descriptors to the resulting cloud", "1", None, None, bool),
        ("keepDensities", "whether the point densities should be added as descriptors", "0", None, None, bool),
        ("keepEigenValues", "whether the eigen values should be added as descriptors", "0", None, None, bool),
        ("keepEigenVectors", "whether the eigen vectors should be added as descriptors", "0", None, None, bool),
        ("keepMatchedIds", "whether the identifiers of matches points should be added as descriptors", "0", None, None, bool),
        ("keepMeanDist", "whether the distance to the nearest neighbor mean should be added as descriptors", "0", None, None, bool),
        ("sortEigen", "whether the eigenvalues and eigenvectors should be sorted (ascending)", "0", None, None, bool),
        ("smoothNormals", "whether the normal vector should be average with the nearest neighbors", "0", None, None, bool),
    )

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        for name, *_ in self.PARAMS:
            setattr(self, name, self.get(name))

    def init(self):
        pass

    def filter(self, cloud):
        out = cloud.copy()
        self.inPlaceFilter(out)
        return out

    def inPlaceFilter(self, cloud):
        keep = [name for flag, name in ((self.keepNormals, "normals"), (self.keepDensities, "densities"),
                                        (self.keepEigenValues, "eigValues"), (self.keepEigenVectors, "eigVectors"),
                                        (self.keepMatchedIds, "matchedIds"), (self.keepMeanDist, "meanDists")) if flag]
        res = _translate(self.ctx.normals, cloud.features, self.knn, self.epsilon, self.maxDist, self.sortEigen, keep, self.smoothNormals)
        self.degenerateCount = res.pop("degenerate")
        cloud.descriptors.update(res)


# ---- host-side pre-filters of the default chain -------------------------------
# Run once per cloud on the CPU, as in the reference: RandomSampling / SamplingSurfaceNormal go
# through the C entry points of csrc/host_filters.cu (same std::rand stream and std::nth_element
# order as the reference's C++), the threshold filters are plain column selections.
def _select_columns(cloud, keep):
    cloud.features = np.ascontiguousarray(cloud.features[keep])
    cloud.descriptors = {k: np.ascontiguousarray(v[keep]) for k, v in cloud.descriptors.items()}


class _HostFilter(Parametrizable):
    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        for name, *_ in self.PARAMS:
            setattr(self, name, self.get(name))

    def bind(self, ctx):
        pass

    def init(self):
        pass

    def filter(self, cloud):
        out = cloud.copy()
        self.inPlaceFilter(out)
        return out


class IdentityDataPointsFilter(_HostFilter):
    className = "IdentityDataPointsFilter"
    PARAMS = ()

    def inPlaceFilter(self, cloud):
        pass


class RandomSamplingDataPointsFilter(_HostFilter):
    """RandomSampling.h:58-65, RandomSampling.cpp:58-75"""
    className = "RandomSamplingDataPointsFilter"
    PARAMS = (("prob", "probability to keep a point, one over decimation factor ", "0.75", "0", "1", float),)

    def inPlaceFilter(self, cloud):
        n = cloud.features.shape[0]
        keep = np.empty(max(n, 1), np.int32)
        m = capi.lib.pmgpu_host_random_sampling(n, float(np.float32(self.prob)), keep.ctypes.data)
        _select_columns(cloud, keep[:m])


class _AxisThresholdFilter(_HostFilter):
    def _values(self, cloud):
        if self.dim >= cloud.features.shape[1] - 1:
            raise InvalidParameter("%s: Error, filtering on dimension number %d, larger than feature dimensionality %d"
                                   % (self.className, self.dim, cloud.features.shape[1] - 2))
        if self.dim == -1:  # Euclidean norm of the point, in float like Eigen's norm()
            f = cloud.features[:, :-1]
            acc = np.zeros(len(f), np.float32)
            for a in range(f.shape[1]):
                acc = (acc + f[:, a] * f[:, a]).astype(np.float32)
            return np.sqrt(acc)
        return cloud.features[:, self.dim]


class MinDistDataPointsFilter(_AxisThresholdFilter):
    """MinDist.h:56-62, MinDist.cpp:60-100"""
    className = "MinDistDataPointsFilter"
    PARAMS = (("dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", int),
              ("minDist", "minimum value authorized. If dim is set to -1 (radius), the absolute value of minDist will be used. "
                          "All points before that will be filtered.", "1", "-inf", "inf", float))

    def inPlaceFilter(self, cloud):
        lim = np.float32(abs(self.minDist) if self.dim == -1 else self.minDist)
        _select_columns(cloud, np.nonzero(self._values(cloud) > lim)[0])


class MaxDistDataPointsFilter(_AxisThresholdFilter):
    """MaxDist.h:56-62, MaxDist.cpp:60-100"""
    className = "MaxDistDataPointsFilter"
    PARAMS = (("dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", int),
              ("maxDist", "maximum distance authorized. If dim is set to -1 (radius), the absolute value of minDist will be used. "
                          "All points beyond that will be filtered.", "1", "-inf", "inf", float))

    def inPlaceFilter(self, cloud):
        lim = np.float32(abs(self.maxDist) if self.dim == -1 else self.maxDist)
        _select_columns(cloud, np.nonzero(self._values(cloud) < lim)[0])


class ObservationDirectionDataPointsFilter(_HostFilter):
    """ObservationDirection.h:63-70, ObservationDirection.cpp:61-88: descriptor observationDirections = sensor centre - point"""
    className = "ObservationDirectionDataPointsFilter"
    PARAMS = (("x", "x-coordinate of sensor", "0", None, None, float), ("y", "y-coordinate of sensor", "0", None, None, float),
              ("z", "z-coordinate of sensor", "0", None, None, float))

    def inPlaceFilter(self, cloud):
        dim = cloud.features.shape[1] - 1
        if dim not in (2, 3):
            raise InvalidField("ObservationDirectionDataPointsFilter: Error, works only in 2 or 3 dimensions, cloud has %d dimensions." % dim)
        centre = np.array([self.x, self.y, self.z][:dim], np.float32)
        cloud.descriptors["observationDirections"] = (centre[None, :] - cloud.features[:, :dim]).astype(np.float32)


class OrientNormalsDataPointsFilter(_HostFilter):
    """OrientNormals.h:59-64, OrientNormals.cpp:60-92: flip the normals toward (or away from) the observation point"""
    className = "OrientNormalsDataPointsFilter"
    PARAMS = (("towardCenter", "If set to true(1), all the normals will point inside the surface (i.e. toward the observation points).",
               "1", "0", "1", bool),)

    def inPlaceFilter(self, cloud):
        if not cloud.descriptorExists("normals"):
            raise InvalidField("OrientNormalsDataPointsFilter: Error, cannot find normals in descriptors.")
        if not cloud.descriptorExists("observationDirections"):
            raise InvalidField("OrientNormalsDataPointsFilter: Error, cannot find observation directions in descriptors.")
        n, o = cloud.descriptors["normals"], cloud.descriptors["observationDirections"]
        scalar = np.zeros(len(n), np.float64)          # `const double scalar = vecP.dot(vecN)`: a float dot product
        acc = np.zeros(len(n), np.float32)
        for a in range(n.shape[1]):
            acc = (acc + o[:, a] * n[:, a]).astype(np.float32)
        scalar[:] = acc
        flip = scalar < 0 if self.towardCenter else scalar > 0
        out = n.copy()
        out[flip] = -out[flip]
        cloud.descriptors["normals"] = out


def _float_norm(f):
    """Eigen's norm() of each row's first columns in float: sqrt of the squares summed left to right"""
    acc = np.zeros(len(f), np.float32)
    for a in range(f.shape[1]):
        acc = (acc + f[:, a] * f[:, a]).astype(np.float32)
    return np.sqrt(acc)


class BoundingBoxDataPointsFilter(_HostFilter):
    """BoundingBox.h:56-66, BoundingBox.cpp:71-103: keep (or remove) the points strictly inside an axis-aligned box"""
    className = "BoundingBoxDataPointsFilter"
    PARAMS = tuple((a + m, "%simum value on %s-axis defining one side of the bounding box" % (m.lower(), a), "-1" if m == "Min" else "1",
                    "-inf", "inf", float) for a in "xyz" for m in ("Min", "Max")) + (
        ("removeInside", "If set to true (1), remove points inside the bounding box; else (0), remove points outside the bounding box",
         "1", "0", "1", bool),)

    def inPlaceFilter(self, cloud):
        f = cloud.features
        lo = np.array([self.xMin, self.yMin, self.zMin], np.float32)
        hi = np.array([self.xMax, self.yMax, self.zMax], np.float32)
        dims = f.shape[1] - 1            # a 2-D cloud (3 rows) has no z test
        inside = np.all((f[:, :dims] > lo[:dims]) & (f[:, :dims] < hi[:dims]), axis=1)
        _select_columns(cloud, np.nonzero(~inside if self.removeInside else inside)[0])


class DistanceLimitDataPointsFilter(_AxisThresholdFilter):
    """DistanceLimit.h:56-62, DistanceLimit.cpp:66-127: MinDist (removeInside 1) or MaxDist (0) in one filter"""
    className = "DistanceLimitDataPointsFilter"
    PARAMS = (("dim", "dimension on which the filter will be applied. x=0, y=1, z=2, radius=-1", "-1", "-1", "2", int),
              ("dist", "distance limit of the filter. If dim is set to -1 (radius), the absolute value of dist will be used", "1", "-inf", "inf", float),
              ("removeInside", "If set to true (1), remove points before the distance limit; else (0), remove points beyond the distance limit",
               "1", "0", "1", bool))

    def _values(self, cloud):
        if self.dim >= cloud.features.shape[1] - 1:
            raise InvalidParameter("DistanceLimitDataPointsFilter: Error, filtering on dimension number %d, larger than authorized axis id %d"
                                   % (self.dim, cloud.features.shape[1] - 2))
        return _AxisThresholdFilter._values(self, cloud)

    def inPlaceFilter(self, cloud):
        lim = np.float32(abs(self.dist) if self.dim == -1 else self.dist)
        v = self._values(cloud)
        _select_columns(cloud, np.nonzero(v > lim if self.removeInside else v < lim)[0])


class FixStepSamplingDataPointsFilter(_HostFilter):
    """FixStepSampling.h:56-80, FixStepSampling.cpp:62-110: every step-th point from a random phase; the step
    is multiplied by stepMult after every call until it reaches endStep, init() rewinds it"""
    className = "FixStepSamplingDataPointsFilter"
    PARAMS = (("startStep", "initial number of point to skip (initial decimation factor)", "10", "1", "2147483647", int),
              ("endStep", "maximal or minimal number of points to skip (final decimation factor)", "10", "1", "2147483647", int),
              ("stepMult", "multiplication factor to compute the new decimation factor for each iteration", "1", "0.0000001", "inf", float))

    def __init__(self, params=None):
        _HostFilter.__init__(self, params)
        self.step = float(self.startStep)

    def init(self):
        self.step = float(self.startStep)

    def inPlaceFilter(self, cloud):
        i_step = int(self.step)
        phase = capi.lib.pmgpu_host_rand() % i_step
        _select_columns(cloud, np.arange(phase, cloud.features.shape[0], i_step))
        delta = self.startStep * self.stepMult - self.startStep
        self.step *= self.stepMult
        if (delta < 0 and self.step < self.endStep) or (delta > 0 and self.step > self.endStep):
            self.step = float(self.endStep)


class MaxPointCountDataPointsFilter(_HostFilter):
    """MaxPointCount.h:56-66, MaxPointCount.cpp:71-110: maxCount columns drawn with srand(seed); the draw itself
    (and the reference's swap through an Eigen view, which copies rather than swaps) is pmgpu_host_max_point_count"""
    className = "MaxPointCountDataPointsFilter"
    PARAMS = (("seed", "srand seed", "1", "0", "2147483647", int), ("maxCount", "maximum number of points", "1000", "0", "2147483647", int))

    def inPlaceFilter(self, cloud):
        n = cloud.features.shape[0]
        if n > 0 and self.maxCount <= n - 1:
            order = np.empty(n, np.int32)
            m = capi.lib.pmgpu_host_max_point_count(n, self.seed, self.maxCount, order.ctypes.data)
            _select_columns(cloud, order[:m])


class MaxQuantileOnAxisDataPointsFilter(_HostFilter):
    """MaxQuantileOnAxis.h:56-62, MaxQuantileOnAxis.cpp:65-103: keep the points below the ratio-quantile of one coordinate"""
    className = "MaxQuantileOnAxisDataPointsFilter"
    PARAMS = (("dim", "dimension on which the filter will be applied. x=0, y=1, z=2", "0", "0", "2", int),
              ("ratio", "maximum quantile authorized. All points beyond that will be filtered.", "0.5", "0.0000001", "0.9999999", float))

    def inPlaceFilter(self, cloud):
        rows = cloud.features.shape[1]
        if self.dim >= rows:
            raise InvalidParameter("MaxQuantileOnAxisDataPointsFilter: Error, filtering on dimension number %d, larger than feature dimensionality %d"
                                   % (self.dim, rows))
        values = cloud.features[:, self.dim]
        if len(values) == 0:
            return
        rank = int(np.float32(len(values)) * np.float32(self.ratio))   # `nbPointsIn * ratio` in T, truncated
        limit = np.partition(values, rank)[rank]
        _select_columns(cloud, np.nonzero(values < limit)[0])


class RemoveNaNDataPointsFilter(_HostFilter):
    """RemoveNaN.cpp:52-72: drop the points with a NaN among their features"""
    className = "RemoveNaNDataPointsFilter"
    PARAMS = ()

    def inPlaceFilter(self, cloud):
        _select_columns(cloud, np.nonzero(~np.isnan(cloud.features).any(axis=1))[0])


class MaxDensityDataPointsFilter(_HostFilter):
    """MaxDensity.h:56-61, MaxDensity.cpp:60-105: thin out the points whose "densities" descriptor exceeds maxDensity"""
    className = "MaxDensityDataPointsFilter"
    PARAMS = (("maxDensity", "Maximum density of points to target. Unit: number of points per m^3.", "10", "0.0000001", "inf", float),)

    def inPlaceFilter(self, cloud):
        if not cloud.descriptorExists("densities"):
            raise InvalidField("MaxDensityDataPointsFilter: Error, no densities found in descriptors.")
        dens = np.ascontiguousarray(cloud.descriptors["densities"][:, 0], np.float32)
        keep = np.empty(max(len(dens), 1), np.int32)
        m = capi.lib.pmgpu_host_max_density(dens.ctypes.data, 1, len(dens), float(np.float32(self.maxDensity)), keep.ctypes.data)
        _select_columns(cloud, keep[:m])


class ShadowDataPointsFilter(_HostFilter):
    """Shadow.h:58-63, Shadow.cpp:42-90: drop the points whose normal is (within eps) perpendicular to the line of sight"""
    className = "ShadowDataPointsFilter"
    PARAMS = (("eps", "Small angle (in rad) around which a normal shoudn't be observable", "0.1", "0.0", "3.1416", float),)

    def inPlaceFilter(self, cloud):
        if not cloud.descriptorExists("normals"):
            raise InvalidField("ShadowDataPointsFilter, Error: cannot find normals in descriptors")
        n = cloud.descriptors["normals"]
        p = cloud.features[:, :-1]
        with np.errstate(invalid="ignore", divide="ignore"):
            nn = _float_norm(n)
            pn = _float_norm(p)
            nu = np.where(nn[:, None] > 0, n / nn[:, None], n).astype(np.float32)   # Eigen's normalized() leaves a zero vector alone
            pu = np.where(pn[:, None] > 0, p / pn[:, None], p).astype(np.float32)
        dot = np.zeros(len(n), np.float32)
        for a in range(n.shape[1]):
            dot = (dot + nu[:, a] * pu[:, a]).astype(np.float32)
        _select_columns(cloud, np.nonzero(np.abs(dot) > np.sin(np.float32(self.eps)))[0])


class SimpleSensorNoiseDataPointsFilter(_HostFilter):
    """SimpleSensorNoise.h:58-82, SimpleSensorNoise.cpp:44-140: descriptor simpleSensorNoise from a per-sensor range model
    (the reference reads `gain` and never uses it; so does this)"""
    className = "SimpleSensorNoiseDataPointsFilter"
    PARAMS = (("sensorType", "Type of the sensor used. Choices: 0=Sick LMS-1xx, 1=Hokuyo URG-04LX, 2=Hokuyo UTM-30LX, 3=Kinect/Xtion",
               "0", "0", "2147483647", int),
              ("gain", "If the point cloud is coming from an untrusty source, you can use the gain to augment the uncertainty", "1", "1", "inf", float))
    LASERS = {0: (0.012, 0.0068, 0.0008), 1: (0.028, 0.0013, 0.0001), 2: (0.018, 0.0006, 0.0015), 4: (0.004, 0.0053, -0.0092)}

    def __init__(self, params=None):
        _HostFilter.__init__(self, params)
        if self.sensorType >= 5:
            raise InvalidParameter("SimpleSensorNoiseDataPointsFilter: Error, sensorType id %d does not exist." % self.sensorType)

    def inPlaceFilter(self, cloud):
        norm = _float_norm(cloud.features[:, :-1])
        if self.sensorType == 3:     # Kinect / Xtion
            noise = ((norm * norm).astype(np.float32) * np.float32(0.5 * 0.00285)).astype(np.float32)
        else:
            min_radius, beam_angle, beam_const = (np.float32(v) for v in self.LASERS[self.sensorType])
            noise = np.maximum(((beam_angle * norm).astype(np.float32) + beam_const).astype(np.float32), min_radius)
        cloud.descriptors["simpleSensorNoise"] = noise[:, None]


class SamplingSurfaceNormalDataPointsFilter(_HostFilter):
    """SamplingSurfaceNormal.h:60-74, SamplingSurfaceNormal.cpp:80-342: kd-split bins of <= knn points,
    one normal per bin, random (0) or one-per-bin (1) subsampling."""
    className = "SamplingSurfaceNormalDataPointsFilter"
    PARAMS = (
        ("ratio", "ratio of points to keep with random subsampling. Matrix (normal, density, etc.) will be associated to all points in the same bin.",
         "0.5", "0.0000001", "1.0", float),
        ("knn", "determined how many points are used to compute the normals. Direct link with the rapidity of the computation (large = fast). "
                "Technically, limit over which a box is splitted in two", "7", "3", "2147483647", int),
        ("samplingMethod", "if set to 0, random subsampling using the parameter ratio. If set to 1, bin subsampling with the resulting number of "
                           "points being 1/knn.", "0", "0", "1", int),
        ("maxBoxDim", "maximum length of a box above which the box is discarded", "inf", None, None, float),
        ("averageExistingDescriptors", "whether the filter keep the existing point descriptors and average them or should it drop them", "1", None, None, bool),
        ("keepNormals", "whether the normals should be added as descriptors to the resulting cloud", "1", None, None, bool),
        ("keepDensities", "whether the point densities should be added as descriptors to the resulting cloud", "0", None, None, bool),
        ("keepEigenValues", "whether the eigen values should be added as descriptors to the resulting cloud", "0", None, None, bool),
        ("keepEigenVectors", "whether the eigen vectors should be added as descriptors to the resulting cloud", "0", None, None, bool),
    )

    def inPlaceFilter(self, cloud):
        n = cloud.features.shape[0]
        names = list(cloud.descriptors)
        spans = [cloud.descriptors[k].shape[1] for k in names]
        desc = np.ascontiguousarray(np.concatenate([cloud.descriptors[k] for k in names], axis=1), np.float32) if names else None
        feat = np.ascontiguousarray(cloud.features, np.float32).copy()
        flags = (1 if self.keepNormals else 0) | (2 if self.keepDensities else 0) | (4 if self.keepEigenValues else 0) | (8 if self.keepEigenVectors else 0)
        keep = np.empty(max(n, 1), np.int32)
        rows = feat.shape[1]
        dn = rows - 1          # spans follow the cloud's dimension (2-D clouds: normals 2, eigVectors 4)
        normals = np.zeros((n, dn), np.float32)
        dens = np.zeros(n, np.float32)
        eva = np.zeros((n, dn), np.float32)
        eve = np.zeros((n, dn * dn), np.float32)
        unfit = C.c_int(0)
        m = capi.lib.pmgpu_host_sampling_surface_normal(
            feat.ctypes.data, rows, n, desc.ctypes.data if desc is not None else None, desc.shape[1] if desc is not None else 0,
            float(np.float32(self.ratio)), int(self.knn), int(self.samplingMethod), float(np.float32(self.maxBoxDim)),
            1 if self.averageExistingDescriptors else 0, flags, keep.ctypes.data, normals.ctypes.data, dens.ctypes.data, eva.ctypes.data,
            eve.ctypes.data, C.byref(unfit))
        if m < 0:
            raise RuntimeError("SamplingSurfaceNormalDataPointsFilter: bad argument")
        k = keep[:m]
        self.unfitPointsCount = unfit.value
        cloud.features = np.ascontiguousarray(feat[k])
        out, col = {}, 0
        for name, span in zip(names, spans):
            out[name] = np.ascontiguousarray(desc[k, col:col + span])
            col += span
        if self.keepNormals:
            out["normals"] = normals[k]
        if self.keepDensities:
            out["densities"] = dens[k, None]
        if self.keepEigenValues:
            out["eigValues"] = eva[k]
        if self.keepEigenVectors:
            out["eigVectors"] = eve[k]
        cloud.descriptors = out


# ---- TransformationCheckers (host objects carrying the parameters; evaluated on the device
#      inside the fused loop, TransformationCheckersImpl.cpp:45-158) ------------------------------
class CounterTransformationChecker(Parametrizable):
    className = "CounterTransformationChecker"
    PARAMS = (("maxIterationCount", "maximum number of iterations ", "40", "0", "2147483647", int),)

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.maxIterationCount = self.get("maxIterationCount")


class DifferentialTransformationChecker(Parametrizable):
    className = "DifferentialTransformationChecker"
    PARAMS = (
        ("minDiffRotErr", "threshold for rotation error (radian)", "0.001", "0.", "6.2831854", float),
        ("minDiffTransErr", "threshold for translation error", "0.001", "0.", "inf", float),
        ("smoothLength", "number of iterations over which to average the differencial error", "3", "0", "2147483647", int),
    )

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.minDiffRotErr = self.get("minDiffRotErr")
        self.minDiffTransErr = self.get("minDiffTransErr")
        self.smoothLength = self.get("smoothLength")


def _quat_from_matrix(m):
    """Eigen's rotation matrix -> quaternion (w, x, y, z), in float"""
    m = np.asarray(m, np.float32)
    t = np.float32(m[0, 0] + m[1, 1] + m[2, 2])
    q = np.zeros(4, np.float32)
    if t > 0:
        t = np.sqrt(np.float32(t + np.float32(1)))
        q[0] = np.float32(0.5) * t
        t = np.float32(0.5) / t
        q[1], q[2], q[3] = (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(np.float32(m[i, i] - m[j, j] - m[k, k] + np.float32(1)))
        q[1 + i] = np.float32(0.5) * t
        t = np.float32(0.5) / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


class BoundTransformationChecker(Parametrizable):
    """TransformationCheckersImpl.h:117-123, TransformationCheckersImpl.cpp:165-225: ConvergenceError when the transform leaves a
    bound around its initial value.  A host checker: with it in the chain the loop runs one device iteration per step."""
    className = "BoundTransformationChecker"
    PARAMS = (("maxRotationNorm", "rotation bound", "1", "0", "inf", float), ("maxTranslationNorm", "translation bound", "1", "0", "inf", float))

    def __init__(self, params=None):
        Parametrizable.__init__(self, params)
        self.maxRotationNorm, self.maxTranslationNorm = self.get("maxRotationNorm"), self.get("maxTranslationNorm")
        self.conditionVariables = [0.0, 0.0]

    def init(self, T):
        T = np.asarray(T, np.float32)
        self._dim = T.shape[0]
        if self._dim == 3:   # 2-D: the rotation is acos(T(0, 0)) (TransformationCheckersImpl.cpp:190-191)
            self._a0 = np.float32(np.arccos(T[0, 0]))
            self._t0 = T[:2, 2].copy()
            return
        self._q0 = _quat_from_matrix(T[:3, :3])
        self._t0 = T[:3, 3].copy()

    def check(self, T):
        if self._dim == 3:
            T = np.asarray(T, np.float32)
            a = np.float32(np.arccos(T[0, 0])) - self._a0
            while a > np.pi:     # normalizeAngle (TransformationCheckersImpl.cpp:229-236)
                a -= 2 * np.pi
            while a < -np.pi:
                a += 2 * np.pi
            rot, tr = float(a), float(np.linalg.norm(T[:2, 2] - self._t0))
            self.conditionVariables = [rot, tr]
            if rot > self.maxRotationNorm or tr > self.maxTranslationNorm:
                raise ConvergenceError("limit out of bounds: rot: %g/%g tr: %g/%g" % (rot, self.maxRotationNorm, tr, self.maxTranslationNorm))
            return
        q, q0 = _quat_from_matrix(np.asarray(T)[:3, :3]), self._q0
        # angularDistance: d = q * conj(q0); 2 * atan2(|d.vec|, |d.w|)
        w = q[0] * q0[0] + q[1] * q0[1] + q[2] * q0[2] + q[3] * q0[3]
        v = np.array([-q[0] * q0[1] + q[1] * q0[0] - q[2] * q0[3] + q[3] * q0[2],
                      -q[0] * q0[2] + q[1] * q0[3] + q[2] * q0[0] - q[3] * q0[1],
                      -q[0] * q0[3] - q[1] * q0[2] + q[2] * q0[1] + q[3] * q0[0]], np.float32)
        rot = float(2 * np.arctan2(np.float32(np.linalg.norm(v)), np.float32(abs(w))))
        tr = float(np.linalg.norm(np.asarray(T, np.float32)[:3, 3] - self._t0))
        self.conditionVariables = [rot, tr]
        if rot > self.maxRotationNorm or tr > self.maxTranslationNorm:
            raise ConvergenceError("limit out of bounds: rot: %g/%g tr: %g/%g" % (rot, self.maxRotationNorm, tr, self.maxTranslationNorm))


# ---- Inspector (PointMatcher.h:640-664, InspectorsImpl.h:47-98) -------------------------------------
class Inspector(Parametrizable):
    """What ICP reports to while it runs.  `dumpIteration` is called once per iteration with host copies of what that iteration
    matched (ICP.cpp:403-405); they are only made — and the loop only leaves the fused device path — for an inspector whose
    `needsIterationData()` is true, which is the case for every subclass that overrides `dumpIteration`."""
    className = "Inspector"

    def init(self):
        pass

    def addStat(self, name, data):
        pass

    def dumpStats(self, stream):
        pass

    def dumpStatsHeader(self, stream):
        pass

    def dumpIteration(self, iterationNumber, parameters, filteredReference, reading, matches, outlierWeights, transformationCheckers):
        pass

    def finish(self, iterationCount):
        pass

    def needsIterationData(self):
        return type(self).dumpIteration is not Inspector.dumpIteration


class NullInspector(Inspector):
    """Does nothing."""
    className = "NullInspector"


class PerformanceInspector(Inspector):
    """Keep statistics on performance."""
    className = "PerformanceInspector"
    PARAMS = (("baseFileName", "base file name for the statistics files (if empty, disabled)", "", None, None, str),
              ("dumpPerfOnExit", "dump performance statistics to stderr on exit", "0", None, None, int),
              ("dumpStats", "dump the statistics on first and last step", "0", None, None, int))

    def __init__(self, params=None):
        Inspector.__init__(self, params)
        self.baseFileName = self.get("baseFileName")
        self.bDumpPerfOnExit, self.bDumpStats = bool(self.get("dumpPerfOnExit")), bool(self.get("dumpStats"))
        self.stats = {}

    def addStat(self, name, data):          # InspectorsImpl.cpp:74-84
        if self.bDumpStats:
            self.stats.setdefault(name, []).append(float(data))

    def dumpStats(self, stream):
        """one "name: count mean min max" group per statistic, like the C++ host layer (the reference prints its Histogram
        class's bins in a format it documents as bound to change)"""
        stream.write(", ".join("%s: %d %g %g %g" % (k, len(v), sum(v) / len(v) if v else 0.0, min(v) if v else 0.0, max(v) if v else 0.0)
                               for k, v in sorted(self.stats.items())))

    def dumpStatsHeader(self, stream):
        stream.write(", ".join(sorted(self.stats)))


# ---- Registrar (Registrar.h:75-218) --------------------------------------------------------------
class Registrar(dict):
    def create(self, name, params=None):
        if name not in self:
            raise InvalidElement("Trying to instanciate unknown element %s from registrar" % name)
        return self[name](params)

    def getDescription(self, name):
        return self[name].__doc__ or ""


MatcherRegistrar = Registrar(KDTreeMatcher=KDTreeMatcher, KDTreeVarDistMatcher=KDTreeVarDistMatcher)
OutlierFilterRegistrar = Registrar(NullOutlierFilter=NullOutlierFilter, MaxDistOutlierFilter=MaxDistOutlierFilter,
                                   MinDistOutlierFilter=MinDistOutlierFilter, MedianDistOutlierFilter=MedianDistOutlierFilter,
                                   TrimmedDistOutlierFilter=TrimmedDistOutlierFilter, RobustOutlierFilter=RobustOutlierFilter,
                                   VarTrimmedDistOutlierFilter=VarTrimmedDistOutlierFilter,
                                   SurfaceNormalOutlierFilter=SurfaceNormalOutlierFilter)
ErrorMinimizerRegistrar = Registrar(PointToPointErrorMinimizer=PointToPointErrorMinimizer,
                                    PointToPointWithCovErrorMinimizer=PointToPointWithCovErrorMinimizer,
                                    PointToPointSimilarityErrorMinimizer=PointToPointSimilarityErrorMinimizer,
                                    PointToPlaneErrorMinimizer=PointToPlaneErrorMinimizer,
                                    PointToPlaneWithCovErrorMinimizer=PointToPlaneWithCovErrorMinimizer)
DataPointsFilterRegistrar = Registrar(SurfaceNormalDataPointsFilter=SurfaceNormalDataPointsFilter, IdentityDataPointsFilter=IdentityDataPointsFilter,
                                      RandomSamplingDataPointsFilter=RandomSamplingDataPointsFilter,
                                      SamplingSurfaceNormalDataPointsFilter=SamplingSurfaceNormalDataPointsFilter,
                                      MinDistDataPointsFilter=MinDistDataPointsFilter, MaxDistDataPointsFilter=MaxDistDataPointsFilter,
                                      ObservationDirectionDataPointsFilter=ObservationDirectionDataPointsFilter,
                                      OrientNormalsDataPointsFilter=OrientNormalsDataPointsFilter,
                                      BoundingBoxDataPointsFilter=BoundingBoxDataPointsFilter,
                                      DistanceLimitDataPointsFilter=DistanceLimitDataPointsFilter,
                                      FixStepSamplingDataPointsFilter=FixStepSamplingDataPointsFilter,
                                      MaxPointCountDataPointsFilter=MaxPointCountDataPointsFilter,
                                      MaxQuantileOnAxisDataPointsFilter=MaxQuantileOnAxisDataPointsFilter,
                                      RemoveNaNDataPointsFilter=RemoveNaNDataPointsFilter,
                                      MaxDensityDataPointsFilter=MaxDensityDataPointsFilter,
                                      ShadowDataPointsFilter=ShadowDataPointsFilter,
                                      SimpleSensorNoiseDataPointsFilter=SimpleSensorNoiseDataPointsFilter)
TransformationCheckerRegistrar = Registrar(CounterTransformationChecker=CounterTransformationChecker,
                                           DifferentialTransformationChecker=DifferentialTransformationChecker,
                                           BoundTransformationChecker=BoundTransformationChecker)
InspectorRegistrar = Registrar(NullInspector=NullInspector, PerformanceInspector=PerformanceInspector)


# ---- float32 4x4 helpers with the reference's GEMM accumulation order -----------------------------
def mat4_mul(A, B):
    """dim x dim product (4 x 4, or 3 x 3 for 2-D clouds) in float32, left-to-right sums"""
    A = np.asarray(A, np.float32)
    B = np.asarray(B, np.float32)
    d = A.shape[0]
    out = np.zeros((d, d), np.float32)
    for i in range(d):
        for j in range(d):
            acc = np.float32(A[i, 0] * B[0, j])
            for k in range(1, d):
                acc = np.float32(acc + np.float32(A[i, k] * B[k, j]))
            out[i, j] = acc
    return out


def sequential_mean(features):
    """`features.rowwise().sum() / N` in float, column after column (ICP.cpp:292)."""
    n = features.shape[0]
    s = np.cumsum(features, axis=0, dtype=np.float32)[-1]
    return (s / np.float32(n)).astype(np.float32)


# ---- ICP (ICP.cpp:99-113, 243-449) -------------------------------------------------------------------
class ICP:
    """ICP chain over the GPU modules.  `icp(reading, reference[, T_init])` returns the 4x4
    transform like `PointMatcher<T>::ICP::operator()`."""

    def __init__(self, device=0):
        self._device, self._ctx = device, None  # the device context is created on first use
        self.readingDataPointsFilters = []
        self.referenceDataPointsFilters = []
        self.matcher = None
        self.outlierFilters = OutlierFilters()
        self.errorMinimizer = None
        self.transformationCheckers = []
        self.inspector = NullInspector()
        self.maxNumIterationsReached = False
        self.iterationCount = 0
        self._shard = None

    def setSharded(self, rank, world):
        """this process registers columns [rank n / world, (rank + 1) n / world) of every (filtered)
        reading against the replicated reference; `self.ctx` must have been given the communicator of the `world`
        ranks (libpointmatcher_b200.dist.init_comm).  Every rank returns the same transform."""
        self._shard = (int(rank), int(world)) if world > 1 else None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = capi.Context(self._device)
        return self._ctx

    def setDefault(self):
        """ICPChainBase::setDefault (ICP.cpp:100-113): RandomSampling on the reading, SamplingSurfaceNormal
        on the reference (both on the host, once per cloud), KDTreeMatcher, TrimmedDist(0.85), PointToPlane,
        Counter(40) + Differential."""
        self.readingDataPointsFilters = [RandomSamplingDataPointsFilter()]
        self.referenceDataPointsFilters = [SamplingSurfaceNormalDataPointsFilter()]
        self.matcher = KDTreeMatcher()
        self.outlierFilters = OutlierFilters([TrimmedDistOutlierFilter()])
        self.errorMinimizer = PointToPlaneErrorMinimizer()
        self.transformationCheckers = [CounterTransformationChecker(), DifferentialTransformationChecker()]
        self.inspector = NullInspector()

    def loadFromYaml(self, text):
        """ICPChainBase::loadFromYaml (ICP.cpp:116-167): the reference's YAML layout — module lists for the
        filters / outlier filters / checkers, a single module for matcher and errorMinimizer; a module is
        `Name`, `Name:` or `Name: {param: value}`.  Unknown names raise InvalidElement (Registrar.h:150-162).
        inspector: NullInspector or PerformanceInspector (VTKFileInspector lives in the C++ host layer); logger: NullLogger only."""
        import yaml
        doc = yaml.safe_load(text) or {}

        def one(node):
            if isinstance(node, str):
                return node, {}
            if isinstance(node, dict) and len(node) == 1:
                (name, params), = node.items()
                return name, {k: str(v) for k, v in (params or {}).items()}
            raise ConfigurationError("ICP YAML: a module must be `Name` or `Name: {parameters}`")

        def many(key, registrar):
            return [registrar.create(*one(n)) for n in (doc.get(key) or [])]

        self.readingDataPointsFilters = many("readingDataPointsFilters", DataPointsFilterRegistrar)
        self.referenceDataPointsFilters = many("referenceDataPointsFilters", DataPointsFilterRegistrar)
        self.outlierFilters = OutlierFilters(many("outlierFilters", OutlierFilterRegistrar))
        self.transformationCheckers = many("transformationCheckers", TransformationCheckerRegistrar)
        self.matcher = MatcherRegistrar.create(*one(doc["matcher"])) if "matcher" in doc else KDTreeMatcher()
        self.errorMinimizer = ErrorMinimizerRegistrar.create(*one(doc["errorMinimizer"])) if "errorMinimizer" in doc else PointToPlaneErrorMinimizer()
        self.inspector = InspectorRegistrar.create(*one(doc["inspector"])) if "inspector" in doc else NullInspector()
        if "logger" in doc and one(doc["logger"])[0] != "NullLogger":
            raise InvalidElement("Trying to instanciate unknown element %s from registrar" % one(doc["logger"])[0])

    def _params(self):
        counter = [c for c in self.transformationCheckers if isinstance(c, CounterTransformationChecker)]
        diff = [c for c in self.transformationCheckers if isinstance(c, DifferentialTransformationChecker)]
        m = self.matcher
        for f in self.outlierFilters:
            f.prepare(self.ctx)
        return capi.make_params(
            knn=m.knn, epsilon=m.epsilon, max_dist=m.maxDist, filters=self.outlierFilters.specs(),
            minimizer=self.errorMinimizer.kind_word(), sensor_std_dev=self.errorMinimizer.sensorStdDev,
            max_iterations=counter[0].maxIterationCount if counter else 0x7FFFFFFF,
            differential=(diff[0].minDiffRotErr, diff[0].minDiffTransErr, diff[0].smoothLength) if diff else None)

    def __call__(self, readingIn, referenceIn, T_refIn_dataIn=None):
        return self.compute(readingIn, referenceIn, T_refIn_dataIn)

    def _bind(self):
        if self.matcher is None:
            raise RuntimeError("You must setup a matcher before running ICP")
        if self.errorMinimizer is None:
            raise RuntimeError("You must setup an error minimizer before running ICP")
        for mod in ([self.matcher, self.outlierFilters, self.errorMinimizer] + list(self.outlierFilters) + list(self.referenceDataPointsFilters)
                    + list(self.readingDataPointsFilters)):
            mod.bind(self.ctx)

    def _set_reference(self, referenceIn):
        """reference filters, centring on the mean and matcher init (ICP.cpp:285-302).  Returns T_refIn_refMean."""
        reference = referenceIn if isinstance(referenceIn, DataPoints) else DataPoints(referenceIn)
        filters = list(self.referenceDataPointsFilters)
        # A trailing SurfaceNormalDataPointsFilter that only adds normals is run on the structure the matcher needs anyway
        # (one upload, one build): set -> normals on the resident cloud -> centre.  Same kernel on the same coordinates as the
        # filter's own private run, and normals do not change under the centring translation.
        last = filters[-1] if filters else None
        fuse_normals = (type(last) is SurfaceNormalDataPointsFilter and type(self.matcher) is KDTreeMatcher and last.keepNormals
                        and not (last.keepDensities or last.keepEigenValues or last.keepEigenVectors or last.keepMatchedIds or last.keepMeanDist
                                 or last.sortEigen or last.smoothNormals))
        if fuse_normals:
            filters = filters[:-1]
        if filters:
            reference = reference.copy()  # inputs are never mutated (ICP.cpp:285)
            for f in filters:
                f.inPlaceFilter(reference)
        if fuse_normals:
            _translate(self.ctx.set_reference, reference.features, None)
            _translate(self.ctx.ref_compute_normals, last.knn, last.epsilon, last.maxDist)
            mean = _translate(self.ctx.ref_center, reference.features)
            self.matcher._ref = reference
        else:
            mean = self.matcher.initCentered(reference)
        self._reference_filtered, self._normals_on_device = reference, fuse_normals
        d = reference.features.shape[1]  # 4, or 3 for 2-D clouds
        T_refIn_refMean = np.eye(d, dtype=np.float32)
        T_refIn_refMean[:d - 1, d - 1] = mean[:d - 1]
        return T_refIn_refMean

    def _register(self, readingIn, T_refIn_refMean, T_refIn_dataIn):
        """computeWithTransformedReference (ICP.cpp:316-449) against the resident reference"""
        d = T_refIn_refMean.shape[0]
        T_init = np.eye(d, dtype=np.float32) if T_refIn_dataIn is None else np.asarray(T_refIn_dataIn, np.float32)
        if T_init.ndim != 2 or T_init.shape[0] != T_init.shape[1]:
            raise RuntimeError("The initial transformation matrix must be squared.")
        if T_init.shape[0] != d:
            raise RuntimeError("The shape of initial transformation matrix must be NxN. Where N is the number of rows in the read/reference scans.")
        reading = readingIn if isinstance(readingIn, DataPoints) else DataPoints(readingIn)
        if self.readingDataPointsFilters:
            reading = reading.copy()  # ICP.cpp:324-326
            for f in self.readingDataPointsFilters:
                f.inPlaceFilter(reading)
        # reading into the refMean frame (ICP.cpp:345-347); T_refIn_refMean is a pure translation
        T_refMean_refIn = np.eye(d, dtype=np.float32)
        T_refMean_refIn[:d - 1, d - 1] = -T_refIn_refMean[:d - 1, d - 1]
        T_refMean_dataIn = mat4_mul(T_refMean_refIn, T_init)
        if self._shard is not None and not reading.descriptors and reading.features.shape[1] == 4:
            # chunks of columns dealt out round-robin (dist.shard_columns): balanced ranks, full local density; uploaded straight
            # from the caller's matrix (one strided copy on the device side, no gather on the host)
            from . import dist as _pmdist
            rank, world = self._shard
            _translate(self.ctx.set_reading_sharded, reading.features, rank, world, _pmdist.SHARD_CHUNK)
            full = reading
            reading = _LazyShard(full, rank, world)   # the host copy of the slice is only made if ErrorElements are asked for
        else:
            if self._shard is not None:
                from . import dist as _pmdist
                rank, world = self._shard
                reading = DataPoints(_pmdist.shard_take(reading.features, rank, world), {k: _pmdist.shard_take(v, rank, world) for k, v in reading.descriptors.items()})
            _translate(self.ctx.set_reading, reading.features)
        self.ctx._reading_obj = None
        self._reading_filtered, self._T_refMean_dataIn = reading, T_refMean_dataIn
        self._T_refMean_refIn = T_refMean_refIn
        self.errorMinimizer._error_elements_provider = self.getErrorElements
        if isinstance(self.matcher, KDTreeVarDistMatcher):
            self.matcher.uploadMaxDists(reading)
        if reading.descriptorExists("normals"):  # they turn with the reading (TransformationsImpl.cpp:71-84)
            _translate(self.ctx.set_reading_normals, reading.descriptors["normals"])
        _translate(self.ctx.reading_apply_transform, T_refMean_dataIn)
        bounds = [c for c in self.transformationCheckers if isinstance(c, BoundTransformationChecker)]
        if self.inspector is None:
            raise RuntimeError("You must setup an inspector before running ICP")
        self.inspector.init()
        if bounds or self.inspector.needsIterationData():
            res = self._run_stepwise(bounds)
        else:
            res = _translate(self.ctx.icp_run, self._params())
        self.iterationCount = res["iterations"]
        counter = [c for c in self.transformationCheckers if isinstance(c, CounterTransformationChecker)]
        self.maxNumIterationsReached = bool(counter) and res["iterations"] >= max(1, counter[0].maxIterationCount)
        self.errorMinimizer._stats = res["stats"]
        self.errorMinimizer._cov = res["cov"]
        self.T_iter = res["T_iter"]
        self.inspector.addStat("IterationsCount", float(res["iterations"]))     # ICP.cpp:432-437
        self.inspector.addStat("OverlapRatio", float(res["stats"]["weightedPointUsedRatio"]))
        self.inspector.finish(res["iterations"])
        return mat4_mul(mat4_mul(T_refIn_refMean, res["T_iter"]), T_refMean_dataIn)

    def getErrorElements(self):
        """ErrorElements of the last executed iteration (ErrorMinimizer.cpp:58-193), from the resident matches: the reading as that
        iteration saw it (T_match * T_refMean_dataIn * filtered reading, ICP.cpp:345-347,381), the centred reference, the kept pairs"""
        matches, w, T_match = self.getMatches()
        filtered = self._reading_filtered.get() if isinstance(self._reading_filtered, _LazyShard) else self._reading_filtered
        step = rigid_apply(T_match, rigid_apply(self._T_refMean_dataIn, filtered))
        return ErrorElements(step, self._centred_reference(), w, matches)

    def getMatches(self):
        """What the reference's inspectors and ErrorMinimizer::getErrorElements read after the fact (ErrorMinimizer.cpp:58-193):
        the matches, outlier weights and reading transform of the last executed iteration, downloaded on request only.
        Returns (Matches, weights (nq, k), T_match); after a capped fused loop the rejected far matches read id -2 / FLT_MAX / 0."""
        ids, dists, w, T = _translate(self.ctx.matches)
        return Matches(dists, ids), w, T

    def _run_stepwise(self, bounds):
        """The loop one iteration at a time (pmgpu_icp_step: exact matching, the covariance of every iteration), for a host that
        looks at every iteration: host-side checkers see T_iter after it (TransformationCheckers::check, ICP.cpp:414-427) and an
        inspector that wants them is shown the iteration's reading, matches and weights (ICP.cpp:403-405).  Counter / Differential
        still decide on the device."""
        params = self._params()
        _translate(self.ctx.icp_reset, None)
        for b in bounds:
            b.init(np.eye(self.ctx.dimh, dtype=np.float32))
        want = self.inspector.needsIterationData()
        reference = reading0 = None
        if want:
            reference = self._centred_reference()
            filtered = self._reading_filtered.get() if isinstance(self._reading_filtered, _LazyShard) else self._reading_filtered
            reading0 = rigid_apply(self._T_refMean_dataIn, filtered)
        done, res = 0, None
        while True:
            res = _translate(self.ctx.icp_step, params)
            if res["iterations"] == done:
                break                          # the device checkers had already stopped the loop
            if want:
                matches, w, T_match = self.getMatches()
                self.inspector.dumpIteration(done, T_match, reference, rigid_apply(T_match, reading0), matches, w, self.transformationCheckers)
            done = res["iterations"]
            for b in bounds:
                b.check(res["T_iter"])
        return res

    def _centred_reference(self):
        """the filtered reference in the frame of its mean, with the normals a fused SurfaceNormal filter made on the device"""
        ref = self._reference_filtered
        desc = dict(ref.descriptors)
        if self._normals_on_device:
            desc["normals"] = _translate(self.ctx.ref_normals)
        centred = ref.features.copy()
        d = centred.shape[1]
        centred[:, :d - 1] = (centred[:, :d - 1] + self._T_refMean_refIn[:d - 1, d - 1][None, :]).astype(np.float32)   # minus the mean (ICP.cpp:291-299)
        return DataPoints(centred, desc)

    def compute(self, readingIn, referenceIn, T_refIn_dataIn=None):
        self._bind()
        T_refIn_refMean = self._set_reference(referenceIn)
        return self._register(readingIn, T_refIn_refMean, T_refIn_dataIn)


class _LazyShard:
    """this rank's slice of a sharded reading, gathered on the host only when somebody asks (ErrorElements)"""

    def __init__(self, full, rank, world):
        self.full, self.rank, self.world, self._slice = full, rank, world, None

    def descriptorExists(self, name):
        return False

    def get(self):
        if self._slice is None:
            from . import dist as _pmdist
            self._slice = DataPoints(_pmdist.shard_take(self.full.features, self.rank, self.world))
        return self._slice


class ICPSequence(ICP):
    """ICP against a map that stays resident on the device (ICP.cpp:455-609): `setMap` filters,
    centres and indexes the map once; every `icp(cloud[, T_init])` registers a new reading against
    it without touching the reference structure (BASELINE config 3: a 10 M-point map, readings
    streaming in)."""

    def __init__(self, device=0):
        super().__init__(device)
        self._T_refIn_refMean = None

    def hasMap(self):
        return self._T_refIn_refMean is not None

    def setMap(self, mapPointCloud):
        self._bind()
        self._T_refIn_refMean = self._set_reference(mapPointCloud)
        return True

    def clearMap(self):
        self._T_refIn_refMean = None

    def __call__(self, cloudIn, T_dataInOld_dataInNew=None):
        # without a map the reference warns and answers the identity (ICP.cpp:598-604)
        if not self.hasMap():
            return np.eye(4, dtype=np.float32)
        self._bind()
        return self._register(cloudIn, self._T_refIn_refMean, T_dataInOld_dataInNew)
