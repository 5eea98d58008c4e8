"""Generates the input-pipeline fixtures from the REAL reference code. Run once in the build
container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pipeline.py

It imports /root/reference/src/datasets/modelnet.py unchanged and stores
* cloud_a.ply / cloud_b.ply    two small ASCII PLY files written here (b announces more than 2048 vertices and mixes
                               number formats), and pipeline_ply.npz = what the reference's ``ply_reader`` returns for them;
* pipeline_norm.npz            inputs (as parsed by the reference's reader) and outputs of the reference's
                               ``FewShotSubModelNet.__getitem__`` (centre on the mean point, divide by the largest norm)
                               for several clouds.
/root/reference does not exist on the GPU box; nothing at test time reads it -- only this script does.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")
import datasets.modelnet as M  # noqa: E402
from PIL import Image  # noqa: E402

HEADER = "ply\nformat ascii 1.0\ncomment fpsg_b200 test fixture\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n"


def write_ply(path, pts, fmt):
    with open(path, "w") as f:
        f.write(HEADER % len(pts))
        for i, p in enumerate(pts):
            f.write(fmt(i, p))


def reference_item(ply_path, n_pts):
    with tempfile.TemporaryDirectory() as d:
        img = os.path.join(d, "x.png")
        Image.new("RGB", (4, 4)).save(img)
        cfg = os.path.join(d, "cfg.txt")
        open(cfg, "w").write(f"{img}\t{ply_path}\n")
        ds = M.FewShotSubModelNet(cfg, n_pts=n_pts)
        _img, pc = ds[0]
        return pc.numpy()


def main():
    rng = np.random.default_rng(20261018)
    a = rng.standard_normal((300, 3)) * np.array([1.0, 0.4, 2.5]) + np.array([0.3, -1.2, 4.0])
    b = rng.uniform(-50, 50, size=(2100, 3))
    write_ply(os.path.join(HERE, "cloud_a.ply"), a, lambda i, p: "%.6f %.6f %.6f\n" % tuple(p))
    fmts = ["%.9g %.9g %.9g\n", "%e %e %e \n", "%+.4f %.3f %.7f\n", "%.12f %.1f %d\n"]
    write_ply(os.path.join(HERE, "cloud_b.ply"), b, lambda i, p: fmts[i % 4] % ((p[0], p[1], int(p[2])) if i % 4 == 3 else tuple(p)))
    ra = np.asarray(M.ply_reader(os.path.join(HERE, "cloud_a.ply")), dtype=np.float32)
    rb = np.asarray(M.ply_reader(os.path.join(HERE, "cloud_b.ply")), dtype=np.float32)
    assert ra.shape == (300, 3) and rb.shape == (2048, 3)
    np.savez_compressed(os.path.join(HERE, "pipeline_ply.npz"), a_xyz=ra, b_xyz=rb)

    xs, ys = {}, {}
    cases = {"a300": (os.path.join(HERE, "cloud_a.ply"), 300), "b2048": (os.path.join(HERE, "cloud_b.ply"), 2048)}
    with tempfile.TemporaryDirectory() as d:
        for name, (n, scale, shift) in {"n2": (2, 1.0, 0.0), "n513": (513, 3.0, 100.0), "n2048": (2048, 0.01, -7.0)}.items():
            pts = rng.standard_normal((n, 3)) * scale + shift
            path = os.path.join(d, name + ".ply")
            write_ply(path, pts, lambda i, p: "%.9g %.9g %.9g\n" % tuple(p))
            xs[name] = np.asarray(M.ply_reader(path), dtype=np.float32)
            ys[name] = reference_item(path, n)
    for name, (path, n) in cases.items():
        xs[name] = np.asarray(M.ply_reader(path), dtype=np.float32)
        ys[name] = reference_item(path, n)
    out = {}
    for k in xs:
        out["x_" + k] = xs[k]
        out["y_" + k] = ys[k]
        print(k, xs[k].shape, ys[k].shape, float(np.abs(ys[k]).max()))
    np.savez_compressed(os.path.join(HERE, "pipeline_norm.npz"), **out)


if __name__ == "__main__":
    main()
