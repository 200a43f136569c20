"""Generates tests/golden/ztools_normals.npz by running the REFERENCE's own prototype of the screen-space normal
reconstruction: absMin and calcNormals of demo/dumps/ztools.py:19-44 (plain numpy; the two function bodies are taken
from the file as they stand and executed, the rest of that Python-2 script is not needed).

    python tests/golden/make_ztools_golden.py           (needs /root/reference)
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ZTOOLS = "/root/reference/demo/dumps/ztools.py"


def load_calc_normals(path=ZTOOLS):
    """absMin / calcNormals exactly as written in the reference (fov = 70 degrees is the module's global, :16)."""
    src = open(path).read()
    ns = {}
    exec("from numpy import *\nfov = radians(70.0)\n", ns)
    for name in ("absMin", "calcNormals"):
        m = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, src, re.S | re.M)
        exec(m.group(0), ns)
    return ns["calcNormals"]


def z_buffers():
    """Seeded smooth depth images with steps (silhouettes), W x H as the oracle test uses them."""
    rng = np.random.RandomState(3)
    out = []
    for W, H in ((64, 48), (97, 33)):
        y, x = np.mgrid[:H, :W].astype(np.float64)
        z = 1.5 + 0.3 * np.sin(x / 9.0 + rng.rand()) * np.cos(y / 7.0 + rng.rand()) + 0.002 * x
        z[:, W // 2:] += 0.25                                  # a depth step
        z += rng.rand(H, W) * 1e-3
        out.append(z.astype(np.float32))
    return out


def main():
    calc = load_calc_normals()
    data = {}
    for i, z in enumerate(z_buffers()):
        data["z%d" % i] = z
        data["n%d" % i] = calc(z.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "ztools_normals.npz"), **data)
    print("wrote", sorted(data))


if __name__ == "__main__":
    main()
