"""Operator applications / Arnoldi time of compute_modes vs Krylov dimension (config 2 mesh)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m, eps = strip_problem(n, slab_h=0.11)
b0 = Basis(m, ElementTriP0(), intorder=4)
ncvs = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 20, 24, 28, 32, 40]
for ncv in ncvs:
    opts = {} if ncv == 0 else {"ncv": ncv}
    t0 = time.time()
    modes = compute_modes(b0, eps, 1.55, num_modes=k, order=2, solver_options=opts)
    st = modes.stats
    print("ncv", ncv, "n_op", st["n_op"], "restarts", st["restarts"], "arnoldi_ms %.1f" % st["t_arnoldi_ms"],
          "solve_ms %.1f" % st["t_solve_ms_total"], "res %.2e" % st["max_residual"], "wall %.3f" % (time.time() - t0),
          "neff", np.round(modes.n_effs.real, 9))
