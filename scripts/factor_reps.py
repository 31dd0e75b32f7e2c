"""Three LU factorisations at config-2 size: the event-timed factor stage of each (knob experiments of the numeric phase)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.context import Context
from femwell_b200.element import ElementTriP0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ctx = Context(0)
m, eps = strip_problem(n, slab_h=0.11)
b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
k0 = 2 * np.pi / 1.55
ctx.set_space(b); ctx.symbolic(); ctx.assemble(0, eps, k0)
sigma = k0**2 * eps.max() * 1.1
out = []
for rep in range(4):
    ctx.lu_factor(sigma)
    out.append(ctx.timings()["lu_factor_ms"])
rhs = np.random.default_rng(0).uniform(-1, 1, b.N)
x = ctx.lu_solve(rhs, refine=0, force_real=True)
r = -ctx.spmv(0, x) + sigma * ctx.spmv(1, x) - rhs
print("factor ms", " ".join(f"{v:.1f}" for v in out), " residual %.2e" % (np.linalg.norm(r) / np.linalg.norm(rhs)))
