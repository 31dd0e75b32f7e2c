"""Symbolic + assembly + SpMV only (config 2 mesh) -- the short command profiled under ncu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.context import Context
from femwell_b200.element import ElementTriP0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = Context(0)
m, eps = strip_problem(n, slab_h=0.11)
b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
k0 = 2 * np.pi / 1.55
ctx.set_space(b)
nnz = ctx.symbolic()
for r in range(reps):
    ctx.assemble(0, eps, k0)
    t = ctx.timings()
    print(f"rep {r}: element {t['element_kernel_ms']:.4f} ms reduce {t['reduce_ms']:.4f} ms nnz {nnz}")
x = np.random.default_rng(0).standard_normal(b.N)
y = ctx.spmv(0, x)
print("spmv checksum", float(np.abs(y).sum()))
