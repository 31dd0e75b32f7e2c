import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, json
from common import mode_element, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes, get_context
from oracle import femwell_oracle as fo
order, n, k = 1, 24, 3
m, eps = strip_problem(n)
b0 = Basis(m, ElementTriP0(), intorder=4)
modes = compute_modes(b0, eps, 1.55, num_modes=k, order=order)
print(modes.stats)
b = modes[0].basis
v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, b.X, b.W, num_modes=k + 3, order=order, v0=v0)
print("oracle n_eff", r.n_effs)
print("gpu    n_eff", modes.n_effs)
ed = b.element_dofs.astype(np.int64)
k0 = 2*np.pi/1.55
zd = b.split_indices()[1]
for i, md in enumerate(modes):
    x = md.E.copy()
    lam = md.k**2
    x[zd] *= 1j*np.sqrt(lam/k0**4)   # redo the scaling
    res = np.linalg.norm(-r.A@x + lam*(r.B@x))/np.linalg.norm(r.A@x)
    ovs = [fo.calculate_overlap(m.p, m.t, ed, order, b.X, b.W, md.E, md.H, r.E[:, j], r.H[:, j]) for j in range(k+3)]
    print(i, "eig residual", res, "overlaps with oracle modes", np.round(np.abs(ovs), 4))
    Href = fo.calculate_hfield(m.p, m.t.astype(np.int64), ed, order, b.X, b.W, md.E, md.k, md.k0*299792458.0, b.N)
    print("   H vs oracle H", np.linalg.norm(md.H-Href)/np.linalg.norm(Href))
for i in range(k+3):
    print("oracle self overlaps", i, [np.round(abs(fo.calculate_overlap(m.p, m.t, ed, order, b.X, b.W, r.E[:, i], r.H[:, i], r.E[:, j], r.H[:, j])),4) for j in range(k+3)])
# assembly on structured mesh
ctx = get_context()
for jit, nn in ((0.0, 8), (0.15, 16)):
    m, eps = strip_problem(nn, jitter=jit)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
    ctx.set_space(b); ctx.symbolic(); ctx.assemble(0, eps, k0)
    A = ctx.matrix(0); As = ctx.matrix(0, eliminate_zeros=False)
    Ar, Br = fo.assemble_waveguide(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), 2, b.X, b.W, fo.interpolate_eps(0, eps, m.nelements, b.X), k0, n=b.N)
    print("jitter", jit, "nnz gpu", A.nnz, "struct", As.nnz, "oracle", Ar.nnz, "B gpu", ctx.matrix(1).nnz, "oracle", Br.nnz)
    Mp, Rp = A.copy(), Ar.copy(); Mp.data[:] = 1; Rp.data[:] = 1
    X = (Mp-Rp).tocoo(); sel = X.data != 0
    vals = np.abs(np.asarray(A[X.row[sel], X.col[sel]]).ravel()); valsr = np.abs(np.asarray(Ar[X.row[sel], X.col[sel]]).ravel())
    print("   pattern diffs", sel.sum(), "max |gpu| there", vals.max() if sel.any() else 0, "max |ref| there", valsr.max() if sel.any() else 0, "scale", abs(Ar).max(), "sumabs diff", abs(abs(A).sum()-abs(Ar).sum())/abs(Ar).sum())
