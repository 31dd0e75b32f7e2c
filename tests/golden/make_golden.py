"""Regenerates tests/golden/*.json.  The reference itself (scikit-fem missing) cannot be imported in
this image, so the fixtures hold (a) the literature n_eff values the reference's own benchmarks use
(docs/benchmarks/mode_solver_rectangle.py:45, docs/benchmarks/mode_solver.py:58-68) and (b) outputs of
the pinned CPU oracle on committed synthetic inputs (regression pins for the GPU box)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from common import RECT_CASES, mode_element, rectangle_problem, strip_problem  # noqa: E402
from femwell_b200.basis import Basis  # noqa: E402
from femwell_b200.element import triangle_quadrature  # noqa: E402
from oracle import femwell_oracle as fo  # noqa: E402

out = {"paper": [c[2] for c in RECT_CASES], "paper_rib": [3.412022, 3.412126, 3.412279, 3.412492, 3.412774,
                                                          3.413132, 3.413571, 3.414100]}
X, W = triangle_quadrature(4)
vals = []
for case in range(4):
    m, eps, bnd, fsel, ref = rectangle_problem(20, case)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, X, W, num_modes=1, order=2, dirichlet_facets=fsel, with_h=False)
    vals.append(float(r.n_effs[0].real))
out["order2_n20_intorder4"] = vals
json.dump(out, open(os.path.join(HERE, "rectangle_neff.json"), "w"), indent=1)

# strip waveguide (C1-like, 2*16^2 triangles): n_eff, te_fraction, powers for the GPU parity tests
m, eps = strip_problem(16)
b = Basis(m, mode_element(2))
v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, X, W, num_modes=2, order=2, v0=v0)
ed = b.element_dofs.astype(np.int64)
strip = {
    "n": 16, "wavelength": 1.55,
    "n_eff": [[float(z.real), float(z.imag)] for z in r.n_effs],
    "te_fraction": [float(fo.te_fraction(m.p, m.t, ed, 2, X, W, r.E[:, i])) for i in range(2)],
    "nnzA": int(r.A.nnz), "nnzB": int(r.B.nnz),
    "sumA": float(r.A.sum()), "sumabsA": float(abs(r.A).sum()), "sumabsB": float(abs(r.B).sum()),
}
json.dump(strip, open(os.path.join(HERE, "strip_n16.json"), "w"), indent=1)
print(out, strip)
