"""Which triangle pays the accuracy price of the DINV chains?  One config-2 factorisation, then the substitution pair with the
inverse diagonal blocks in both passes (default), in the forward pass only, in the backward pass only and in neither:
device time of the pair (no CUDA graph) and residual of the solve."""
import os, sys, time
os.environ["FW_SOLVE_TIME"] = "1"
os.environ["FW_NO_SOLVE_GRAPH"] = "1"
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
ctx.lu_factor(sigma)
rhs = np.random.default_rng(0).uniform(-1, 1, b.N)
for name, env in (("both passes (default)", {}), ("forward only", {"FW_DINV_FWD_ONLY": "1"}),
                  ("backward only", {"FW_DINV_BWD_ONLY": "1"}), ("neither", {"FW_NO_DINV_CHAIN": "1"})):
    for k in ("FW_DINV_FWD_ONLY", "FW_DINV_BWD_ONLY", "FW_NO_DINV_CHAIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print("==", name, flush=True)
    for rep in range(3):
        x = ctx.lu_solve(rhs, refine=0, force_real=True)
    r = -ctx.spmv(0, x) + sigma * ctx.spmv(1, x) - rhs
    print("   residual %.3e" % (np.linalg.norm(r) / np.linalg.norm(rhs)), flush=True)
