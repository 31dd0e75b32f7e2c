"""Debug helper: compare the GPU multifrontal factors front by front with the numpy emulator."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.sparse.linalg as spla

from common import mode_element, strip_problem
from femwell_b200 import _lib as L
from femwell_b200.basis import Basis
from femwell_b200.context import Context
from femwell_b200.element import ElementTriP0
from mf_emulator import Emulator, lu_symbolic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
leaf = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ctx = Context(0)
m, eps = strip_problem(n)
b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
k0 = 2 * np.pi / 1.55
ctx.set_space(b)
ctx.symbolic()
ctx.assemble(0, eps, k0)
A, B = ctx.matrix(0, eliminate_zeros=False), ctx.matrix(1, eliminate_zeros=False)
sigma = k0**2 * eps.max() * 1.1
OP = (-A + sigma * B).tocsr()
ctx.lu_factor(sigma, None, leaf)
c = m.centroids()
S = lu_symbolic(b.N, b.element_dofs, c[0], c[1], np.ones(b.N, dtype=np.uint8), leaf)
print({k: S[k] for k in ("n_active", "n_nodes", "n_levels", "factor_entries", "max_m", "max_ns")})
F = np.zeros(S["factor_entries"])
piv = np.zeros(S["n_active"], dtype=np.int32)
L.check(ctx.lib.fw_debug_lu_get(ctx._h, 0, L.ptr(F)))
L.check(ctx.lib.fw_debug_lu_get(ctx._h, 1, L.ptr(piv)))
E = Emulator(S, OP)
worst = []
for lev in range(S["n_levels"] - 1, -1, -1):
    for v in S["level_nodes"][S["level_start"][lev]:S["level_start"][lev + 1]]:
        mm, ns, os_ = S["m"][v], S["ns"][v], S["own_start"][v]
        G = F[S["front_off"][v]:S["front_off"][v] + mm * mm].reshape(mm, mm).T  # column-major
        R = E.F[v]
        pg = piv[os_:os_ + ns]
        pe = E.piv[v]
        # compare the factor parts only (S22 is scratch)
        mask = np.ones((mm, mm), dtype=bool)
        mask[ns:, ns:] = False
        d = np.abs(G - R)[mask].max() if mask.any() else 0.0
        sc = np.abs(R)[mask].max() if mask.any() else 1.0
        d22 = np.abs(G - R)[~mask].max() if (~mask).any() else 0.0
        worst.append((d / sc, lev, v, mm, ns, int((pg != pe).sum()), d22))
worst.sort(reverse=True)
print("worst fronts (rel diff, level, node, m, ns, piv mismatches, S22 diff):")
for w in worst[:8]:
    print("  ", w)
byl = {}
for w in worst:
    byl.setdefault(w[1], []).append(w[0])
for lev in sorted(byl, reverse=True):
    print("level", lev, "nfronts", len(byl[lev]), "max rel diff", max(byl[lev]))
rhs = np.random.default_rng(0).uniform(-1, 1, b.N)
x = ctx.lu_solve(rhs)
xe = E.solve(rhs)
xs = spla.splu(OP.tocsc()).solve(rhs)
print("gpu resid", np.linalg.norm(OP @ x - rhs) / np.linalg.norm(rhs), "emu resid",
      np.linalg.norm(OP @ xe - rhs) / np.linalg.norm(rhs), "gpu-vs-splu", np.linalg.norm(x - xs) / np.linalg.norm(xs))
