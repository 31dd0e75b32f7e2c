"""LU solve accuracy vs mesh size (residual of the real single-rhs path and of the 2-rhs path)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.context import Context
from femwell_b200.element import ElementTriP0
ctx = Context(0)
for n in [int(a) for a in sys.argv[1:]]:
    m, eps = strip_problem(n, slab_h=0.11)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
    k0 = 2 * np.pi / 1.55
    ctx.set_space(b); ctx.symbolic(); ctx.assemble(0, eps, k0)
    sigma = k0**2 * eps.max() * 1.1
    ctx.lu_factor(sigma, None, 0)
    rng = np.random.default_rng(0)
    rhs = rng.uniform(-1, 1, b.N)
    def resid(x, rhs):
        r = -ctx.spmv(0, x) + sigma * ctx.spmv(1, x) - rhs
        return np.linalg.norm(r) / np.linalg.norm(rhs)
    x1 = ctx.lu_solve(rhs, force_real=True)
    x2 = ctx.lu_solve(rhs + 1j * rng.uniform(-1, 1, b.N))
    # one step of refinement on x1
    r1 = rhs - (-ctx.spmv(0, x1) + sigma * ctx.spmv(1, x1))
    x1r = x1 + ctx.lu_solve(r1, force_real=True)
    print(f"n={n} N={b.N} |x|={np.linalg.norm(x1):.3e} resid nrhs1 {resid(x1, rhs):.3e} refined {resid(x1r, rhs):.3e} nrhs2(re) {resid(x2.real, rhs):.3e} x1-x2 {np.linalg.norm(x1-x2.real)/np.linalg.norm(x1):.3e}", flush=True)
