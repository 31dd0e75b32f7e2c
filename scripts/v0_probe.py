"""Operator applications of compute_modes with the default (random) start vector vs a smooth start vector that is
concentrated on the high-index region (Whitney dofs of a low-order polynomial field weighted by eps - eps_min)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250


def smart_v0(m, eps, N, seed_noise=0.0):
    nv, nf = m.nvertices, m.nfacets
    wv = np.zeros(nv)
    for i in range(3):
        np.maximum.at(wv, m.t[i], eps - eps.min())
    wv /= wv.max()
    fa, fb = m.facets
    mid = 0.5 * (m.p[:, fa] + m.p[:, fb]); tang = m.p[:, fb] - m.p[:, fa]
    wf = 0.5 * (wv[fa] + wv[fb])
    sel = wf > 0.5
    xc, yc = mid[0] - mid[0][sel].mean(), mid[1] - mid[1][sel].mean()
    xn, yn = xc / np.abs(xc[sel]).max(), yc / np.abs(yc[sel]).max()
    Fx = wf * (1 + 0.9 * xn + 0.8 * (xn**2 - 0.3) + 0.7 * xn**3) * (1 + 0.6 * yn)
    Fy = wf * (0.95 + 0.85 * xn + 0.75 * (xn**2 - 0.3) + 0.65 * xn**3) * (1 + 0.5 * yn)
    v0 = np.zeros(N)
    v0[nv + 3 * np.arange(nf)] = Fx * tang[0] + Fy * tang[1]
    if seed_noise:
        v0 += seed_noise * np.abs(v0).max() * np.random.default_rng(1).uniform(-1, 1, N)
    return v0


for slab in (0.11, 0.0):
    m, eps = strip_problem(n, slab_h=slab)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    for k in (1, 2, 4, 6):
        out = []
        for label in ("uniform", "smooth (library default)", "python smooth"):
            opts = {"start_vector": "uniform"} if label == "uniform" else {}
            if label.startswith("python"):
                N = m.nvertices + 3 * m.nfacets + 2 * m.nelements
                opts["v0"] = smart_v0(m, eps, N, 0.0)
            modes = compute_modes(b0, eps, 1.55, num_modes=k, order=2, solver_options=opts)
            st = modes.stats
            out.append(f"{label}: n_op {st['n_op']} restarts {st['restarts']} res {st['max_residual']:.1e}")
        print(f"slab {slab} k={k}: " + " | ".join(out), "n_eff", np.round(modes.n_effs.real, 6), flush=True)
