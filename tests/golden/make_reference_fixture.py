"""Packs the reference's own example clouds and golden transforms into tests/golden/reference_fixture.npz.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_reference_fixture.py

Contents (data only, no reference source code):
  cloud0, cloud1   examples/data/cloud.00000.vtk / cloud.00001.vtk   (N, 3) float32  — utest icpTest (utest/utest.cpp:81-160)
  car400, car401   examples/data/car_cloud400.csv / car_cloud401.csv (N, 6) float32 x y z nx ny nz — utest validate3dTransformation
  golden_<name>    examples/data/icp_data/<name>.ref_trans           (4, 4) float64
  yaml_<name>      examples/data/icp_data/<name>.yaml, yaml_default = examples/data/default.yaml   (text)
  validT3d         utest/utest.cpp:352-356
  box2d_one, box2d_two   examples/data/2D_oneBox.csv / 2D_twoBoxes.csv   (N, 2) float32 — utest validate2dTransformation
  validT2d         utest/utest.cpp:347-350
"""
import os

import numpy as np

REF = "/root/reference"
DATA = os.path.join(REF, "examples", "data")


def load_vtk_points(path):
    with open(path) as f:
        lines = f.read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.startswith("POINTS"))
    n = int(lines[i].split()[1])
    vals = []
    k = i + 1
    while len(vals) < 3 * n:
        vals.extend(float(x) for x in lines[k].split())
        k += 1
    return np.array(vals[: 3 * n], np.float32).reshape(n, 3)


def load_csv(path):
    """x, y, z [, nx, ny, nz]: comma- or blank-separated, optional header line"""
    rows = []
    for line in open(path):
        tok = line.replace(",", " ").split()
        try:
            rows.append([float(t) for t in tok])
        except ValueError:
            continue  # header
    return np.array(rows, np.float64).astype(np.float32)


def load_trans(name):
    return np.loadtxt(os.path.join(DATA, "icp_data", name + ".ref_trans")).reshape(4, 4)


out = dict(
    cloud0=load_vtk_points(os.path.join(DATA, "cloud.00000.vtk")),
    cloud1=load_vtk_points(os.path.join(DATA, "cloud.00001.vtk")),
    car400=load_csv(os.path.join(DATA, "car_cloud400.csv")),
    car401=load_csv(os.path.join(DATA, "car_cloud401.csv")),
    validT3d=np.array([[0.982304, 0.166685, -0.0854066, 0.0446816], [-0.150189, 0.973488, 0.172524, 0.191998],
                       [0.111899, -0.156644, 0.981296, -0.0356313], [0, 0, 0, 1]]),
    box2d_one=load_csv(os.path.join(DATA, "2D_oneBox.csv")),
    box2d_two=load_csv(os.path.join(DATA, "2D_twoBoxes.csv")),
    validT2d=np.array([[0.987498, 0.157629, 0.0859918], [-0.157629, 0.987498, 0.203247], [0, 0, 1]]),
)
ICP_DATA = os.path.join(DATA, "icp_data")
# every chain file of utest icpTest (utest/utest.cpp:81-160) with its golden transform
for name in sorted(f[:-5] for f in os.listdir(ICP_DATA) if f.endswith(".yaml")):
    out["golden_" + name] = load_trans(name)
    out["yaml_" + name] = np.array(open(os.path.join(ICP_DATA, name + ".yaml")).read())
for name in ("defaultPointToPlaneWithCovErrorMinimizer", "defaultPointToPointWithCovErrorMinimizer"):
    if os.path.exists(os.path.join(ICP_DATA, name + ".ref_trans")):
        out["golden_" + name] = load_trans(name)
out["yaml_default"] = np.array(open(os.path.join(DATA, "default.yaml")).read())  # BASELINE config 1
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixture.npz")
np.savez_compressed(path, **out)
print({k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path))
