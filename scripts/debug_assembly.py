import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.context import Context
from oracle import femwell_oracle as fo
ctx = Context(0)
order = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m, eps = strip_problem(10)
b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
k0 = 2*np.pi/1.55
A_ref, B_ref = fo.assemble_waveguide(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), order, b.X, b.W, fo.interpolate_eps(0, eps, m.nelements, b.X), k0, 1.3, np.inf, n=b.N)
ctx.set_space(b); print("nnz struct", ctx.symbolic())
ctx.assemble(0, eps, k0, 1.3, np.inf)
for name, M, R in (("A", ctx.matrix(0), A_ref), ("B", ctx.matrix(1), B_ref)):
    print(name, "nnz gpu", M.nnz, "ref", R.nnz)
    D = (M - R).tocoo()
    big = np.abs(D.data) > 1e-12*np.abs(R).max()
    print(" entries differing:", big.sum(), "max diff", np.abs(D.data).max() if D.nnz else 0, "scale", np.abs(R).max())
    Pm = (M != 0).astype(int) ; Pr = (R != 0).astype(int)
    Mp = M.copy(); Mp.data[:] = 1; Rp = R.copy(); Rp.data[:] = 1
    X = (Mp - Rp).tocoo()
    only_gpu = X.data > 0; only_ref = X.data < 0
    print(" pattern: only gpu", only_gpu.sum(), "only ref", only_ref.sum())
    kinds = np.zeros(b.N, dtype=int); kinds[b.split_indices()[1]] = 1
    for lab, sel in (("gpu", only_gpu), ("ref", only_ref)):
        r, c = X.row[sel][:8], X.col[sel][:8]
        for i, j in zip(r, c):
            print("  only", lab, (i, j), "kinds", kinds[i], kinds[j], "gpu val", M[i, j], "ref val", R[i, j])
    idx = np.argsort(-np.abs(D.data))[:6]
    for q in idx:
        i, j = D.row[q], D.col[q]
        print("  diff", (i, j), "kinds", kinds[i], kinds[j], "gpu", M[i, j], "ref", R[i, j])
