"""Cost of a triangular-solve pair with one and with two right-hand sides (config-2 mesh): CUDA-graph replays timed on the host."""
import os, sys, time
os.environ["FW_SOLVE_TIME"] = "1"
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
ctx.lu_factor(k0**2 * eps.max() * 1.1)
rng = np.random.default_rng(0)
for name, rhs in (("nrhs=1 (real rhs)", rng.uniform(-1, 1, b.N)), ("nrhs=2 (complex rhs, real factor)", rng.uniform(-1, 1, b.N) + 1j * rng.uniform(-1, 1, b.N))):
    for rep in range(5):
        t0 = time.perf_counter(); x = ctx.lu_solve(rhs, refine=0, force_real=not np.iscomplexobj(rhs)); dt = time.perf_counter() - t0
        print(f"{name}: rep {rep} wall {dt*1e3:.2f} ms (includes the H2D/D2H of the vectors)")
print("timings", {k: v for k, v in ctx.timings().items() if 'solve' in k})
