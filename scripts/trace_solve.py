"""One LU factorisation + two solves at config-2 size with per-level timing (FW_TRACE=2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.context import Context
from femwell_b200.element import ElementTriP0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = Context(0)
m, eps = strip_problem(n, slab_h=0.11)
b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
k0 = 2 * np.pi / 1.55
ctx.set_space(b); ctx.symbolic(); ctx.assemble(0, eps, k0)
sigma = k0**2 * eps.max() * 1.1
ctx.lu_factor(sigma, None, leaf)  # warm (allocations)
os.environ["FW_TRACE"] = os.environ.get("FW_FACTOR_TRACE", "1")
ctx.lu_factor(sigma, None, leaf)
rhs = np.random.default_rng(0).uniform(-1, 1, b.N)
os.environ["FW_TRACE"] = "0"
x = ctx.lu_solve(rhs)  # warm
os.environ["FW_TRACE"] = "2"
t0 = time.time(); x = ctx.lu_solve(rhs); print("solve wall", time.time() - t0)
os.environ["FW_TRACE"] = "0"
y = ctx.spmv(0, x.real); z = ctx.spmv(1, x.real)
r = -y + sigma * z - rhs
print("residual", np.linalg.norm(r) / np.linalg.norm(rhs))
