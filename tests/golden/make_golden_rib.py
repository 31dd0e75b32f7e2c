"""Regenerates tests/golden/rib_neff.json: the eight rib-waveguide known answers of the reference's second
mode-solver benchmark (docs/benchmarks/mode_solver.py:52-68, after Hadley 2002; judged at 5e-6, :160) next to the
CPU oracle's values on the committed synthetic meshes (femwell_b200.mesh.rib_benchmark_mesh, h0 = 0.075,
ratio 1.2; lambda = 1.15, order 2, the 3-point rule a default Basis(mesh, ElementTriP0()) carries)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from femwell_b200.element import triangle_quadrature  # noqa: E402
from femwell_b200.mesh import rib_benchmark_mesh  # noqa: E402
from oracle import femwell_oracle as fo  # noqa: E402

SLABS = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7]
PAPER = [3.412022, 3.412126, 3.412279, 3.412492, 3.412774, 3.413132, 3.413571, 3.414100]
INDICES = {"core": 3.44, "box": 3.40, "clad": 1.0}


def rib_problem(slab, h0=0.075, ratio=1.2):
    m = rib_benchmark_mesh(slab, h0=h0, ratio=ratio)
    eps = np.zeros(m.nelements)
    for name, n in INDICES.items():
        eps[m.subdomains[name]] = n**2
    return m, eps


if __name__ == "__main__":
    X, W = triangle_quadrature(2)
    out = {"slab_thickness": SLABS, "paper": PAPER, "h0": 0.075, "ratio": 1.2, "wavelength": 1.15, "oracle": [],
           "n_triangles": []}
    for slab in SLABS:
        m, eps = rib_problem(slab)
        r = fo.compute_modes(m.p, m.t, 0, eps, 1.15, X, W, num_modes=1, order=2, with_h=False)
        out["oracle"].append(float(r.n_effs[0].real))
        out["n_triangles"].append(int(m.nelements))
        print(slab, m.nelements, r.n_effs[0].real, r.n_effs[0].real - PAPER[len(out["oracle"]) - 1], flush=True)
    json.dump(out, open(os.path.join(HERE, "rib_neff.json"), "w"), indent=1)
