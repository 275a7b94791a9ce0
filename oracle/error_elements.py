"""Oracle restatement of ErrorMinimizer::ErrorElements and what the reference derives from it after an iteration.
TEST INFRASTRUCTURE ONLY — nothing in the product imports this file.  Plain loops, one pair at a time.

  ErrorElements ctor                          pointmatcher/ErrorMinimizer.cpp:58-193
  PointToPoint computeResidualError           pointmatcher/ErrorMinimizers/PointToPoint.cpp:153-163
  PointToPoint getOverlap                     pointmatcher/ErrorMinimizers/PointToPoint.cpp:116-151
  PointToPlane computeResidualError           pointmatcher/ErrorMinimizers/PointToPlane.cpp:314-352
  PointToPlane getOverlap                     pointmatcher/ErrorMinimizers/PointToPlane.cpp:369-466

Parity status: no reference test pins these numbers (utest/ui/ErrorMinimizers.cpp only checks that getOverlap / getResidualError
run): parity unpinned; Eigen's reduction order in `sum()` is unpinned too, so sums are compared with a relative tolerance.
"""
import numpy as np

INF = np.float32(np.inf)


def error_elements(reading, reading_desc, reference, reference_desc, weights, ids, dists):
    """reading (N, 4), reference (M, 4), *_desc: dict name -> (n, span); weights / ids / dists (N, k).
    Returns dict(reading, reading_desc, reference, reference_desc, weights, ids, dists, pointUsedRatio, weightedPointUsedRatio,
    nbRejectedMatches, nbRejectedPoints)."""
    n, knn = ids.shape
    if int((weights != 0).sum()) == 0:
        raise RuntimeError("ErrorMnimizer: no point to minimize")
    kept_i, kept_id, kept_d, kept_w = [], [], [], []
    rejected_matches = rejected_points = 0
    wsum = np.float32(0)
    for i in range(n):
        exist = False
        for k in range(knn):
            if dists[i, k] == INF:
                continue
            if weights[i, k] != 0:
                kept_i.append(i)
                kept_id.append(int(ids[i, k]))
                kept_d.append(dists[i, k])
                kept_w.append(weights[i, k])
                wsum = np.float32(wsum + weights[i, k])
                exist = True
            else:
                rejected_matches += 1
        if not exist:
            rejected_points += 1
    kept_i, kept_id = np.array(kept_i, np.int64), np.array(kept_id, np.int64)
    return dict(reading=reading[kept_i], reading_desc={k: v[kept_i] for k, v in reading_desc.items()},
                reference=reference[kept_id], reference_desc={k: v[kept_id] for k, v in reference_desc.items()},
                weights=np.array(kept_w, np.float32), ids=kept_id, dists=np.array(kept_d, np.float32),
                pointUsedRatio=float(np.float32(len(kept_i)) / np.float32(knn * n)),
                weightedPointUsedRatio=float(wsum / np.float32(knn * n)),
                nbRejectedMatches=rejected_matches, nbRejectedPoints=rejected_points)


def _norm(v):
    acc = np.float32(0)
    for a in v:
        acc = np.float32(acc + np.float32(a) * np.float32(a))
    return np.float32(np.sqrt(acc))


def _delta_norms(e):
    return [_norm((e["reading"][i, :3] - e["reference"][i, :3]).astype(np.float32)) for i in range(len(e["reading"]))]


def point_to_point_residual(e):
    return float(np.sum(np.array(_delta_norms(e), np.float64)))


def point_to_plane_residual(e, force2d=False):
    n = e["reference_desc"]["normals"]
    total = 0.0
    for i in range(len(n)):
        dot = np.float32(0)
        for a in range(2 if force2d else 3):
            dot = np.float32(dot + np.float32(e["reading"][i, a] - e["reference"][i, a]) * n[i, a])
        total += float(np.float32(e["weights"][i] * np.float32(dot * dot)))
    return total


def point_to_point_overlap(e):
    if "simpleSensorNoise" not in e["reading_desc"]:
        return None
    d = _delta_norms(e)
    mean = np.float32(np.float32(np.sum(np.array(d, np.float64))) / np.float32(len(d)))
    count = sum(1 for i in range(len(d)) if d[i] < np.float32(mean + e["reading_desc"]["simpleSensorNoise"][i, 0]))
    return float(np.float32(count) / np.float32(len(d)))


def point_to_plane_overlap(e):
    rn, fn = e["reading_desc"].get("simpleSensorNoise"), e["reference_desc"].get("simpleSensorNoise")
    dens = e["reference_desc"].get("densities")
    m = len(e["reading"])
    if rn is not None and fn is not None and dens is not None:
        values = sorted(float(x) for x in dens.reshape(-1))
        median = np.float32(values[int(len(values) * 0.5)])
        radius = np.float32(1.0 / (float(median) ** (1 / 3.0)))
        unc = [np.float32(np.float32(radius + rn[i, 0]) + fn[i, 0]) for i in range(m)]
    elif rn is not None and fn is not None:
        unc = [np.float32(rn[i, 0] + fn[i, 0]) for i in range(m)]
    elif rn is not None:
        unc = [rn[i, 0] for i in range(m)]
    elif fn is not None:
        unc = [fn[i, 0] for i in range(m)]
    else:
        return None
    d = _delta_norms(e)
    count, unique = 0, 1
    last = e["reading"][0] * 2
    for i in range(m):
        point = e["reading"][i]
        if (last != point).any():
            if abs(d[i]) < unc[i]:
                last = point
                count += 1
        if i > 0 and (point != e["reading"][i - 1]).any():
            unique += 1
    return float(np.float32(count) / np.float32(unique + e["nbRejectedPoints"]))
