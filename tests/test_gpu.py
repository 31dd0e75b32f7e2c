"""GPU parity suite: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): matrix entries 1e-12 relative, n_eff 1e-9 relative, normalised
mode overlaps within 1e-6.  Integer/index work (CSR patterns) is compared bit-exactly.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from common import RECT_CASES, csr_rel_diff, mode_element, rectangle_problem, strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementDG, ElementTriP0, ElementTriP1, ElementVector, triangle_quadrature
from femwell_b200.mesh import MeshTri
from oracle import femwell_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from femwell_b200.context import Context

    c = Context(0)
    yield c
    c.close()


def _oracle_AB(m, b, eps_kind, eps, k0, mu_r=1.0, radius=np.inf, eliminate_zeros=True):
    order = b.elem.order
    return fo.assemble_waveguide(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), order, b.X, b.W,
                                 fo.interpolate_eps(eps_kind, eps, m.nelements, b.X), k0, mu_r, radius, n=b.N,
                                 eliminate_zeros=eliminate_zeros)


def _pattern_diff_is_roundoff(M, R):
    """eliminate_zeros makes the pattern depend on whether a mathematically-zero local integral rounds to
    exactly 0.0: entries present in only one of the two matrices must be round-off zeros."""
    scale = abs(R).max()
    Mp, Rp = M.copy(), R.copy()
    Mp.data = np.ones(Mp.nnz)
    Rp.data = np.ones(Rp.nnz)
    X = (Mp - Rp).tocoo()
    sel = X.data != 0
    vals = np.concatenate([np.abs(np.asarray(M[X.row[sel], X.col[sel]]).ravel()),
                           np.abs(np.asarray(R[X.row[sel], X.col[sel]]).ravel())]) if sel.any() else np.zeros(0)
    return bool((vals <= 1e-14 * scale).all()), int(sel.sum())


# ----------------------------------------------------------------------------- K2: COO -> CSR
@pytest.mark.parametrize("n_rows,n_entries,cplx,seed", [(1, 1, False, 0), (7, 0, False, 1), (50, 400, False, 2),
                                                        (1000, 20000, True, 3), (100000, 3000000, False, 4),
                                                        (257, 70000, False, 5)])
def test_coo_to_csr_matches_scipy(ctx, n_rows, n_entries, cplx, seed):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n_rows, n_entries).astype(np.int32)
    cols = rng.integers(0, n_rows, n_entries).astype(np.int32)
    data = rng.standard_normal(n_entries)
    if cplx:
        data = data + 1j * rng.standard_normal(n_entries)
    if n_entries:
        data[rng.integers(0, n_entries, n_entries // 10)] = 0  # explicit zeros must be eliminated first
    ref = sp.coo_matrix((data, (rows, cols)), shape=(n_rows, n_rows))
    ref.eliminate_zeros()
    ref = ref.tocsr()
    got = ctx.coo_to_csr(n_rows, rows, cols, data)
    assert np.array_equal(got.indptr, ref.indptr)
    assert np.array_equal(got.indices, ref.indices)
    np.testing.assert_allclose(got.data, ref.data, rtol=1e-13, atol=1e-13)


def test_coo_to_csr_cancelling_duplicates_stay_explicit(ctx):
    rows = np.array([0, 0, 1], dtype=np.int32)
    cols = np.array([1, 1, 0], dtype=np.int32)
    data = np.array([2.0, -2.0, 3.0])
    got = ctx.coo_to_csr(2, rows, cols, data)
    assert got.nnz == 2 and got[0, 1] == 0.0 and got[1, 0] == 3.0


# ----------------------------------------------------------------------------- K1+K2: assembly
CASES = [
    # order, n, jitter, eps_kind, complex, radius, intorder
    (1, 10, 0.15, 0, False, np.inf, 2),
    (2, 10, 0.15, 0, False, np.inf, 4),
    (2, 8, 0.0, 0, False, np.inf, 4),      # structured: accidental exact zeros are eliminated (R8)
    (2, 9, 0.15, 0, True, np.inf, 4),
    (2, 9, 0.15, 1, True, 5.0, 4),        # DG-P1 complex eps + bend (config 3 style)
    (2, 9, 0.15, 2, False, np.inf, 4),    # diagonal anisotropy
    (2, 9, 0.15, 2, True, 7.5, 6),
    (1, 9, 0.15, 1, False, 3.0, 5),
    (2, 7, 0.15, 0, False, np.inf, 2),     # P0-default 3-point rule (R3)
]


def _make_eps(m, eps_kind, cplx, seed=0):
    rng = np.random.default_rng(seed)
    ne = m.nelements
    base = np.full(ne, 1.444**2)
    base[m.subdomains["core"]] = 3.4777**2
    if eps_kind == 0:
        eps = base.copy()
    elif eps_kind == 1:
        eps = np.repeat(base, 3) * (1 + 0.05 * rng.uniform(-1, 1, 3 * ne))
    else:
        eps = np.repeat(base, 3) * np.tile([1.0, 1.1, 0.9], ne)
    if cplx:
        eps = eps.astype(complex) - 1j * rng.uniform(0, 0.5, eps.size)
    return eps


@pytest.mark.parametrize("order,n,jitter,eps_kind,cplx,radius,intorder", CASES)
def test_assembly_matches_oracle(ctx, order, n, jitter, eps_kind, cplx, radius, intorder):
    m, _ = strip_problem(n, jitter=jitter)
    b = Basis(m, ElementTriP0(), intorder=intorder).with_element(mode_element(order))
    eps = _make_eps(m, eps_kind, cplx)
    k0 = 2 * np.pi / 1.55
    A_ref, B_ref = _oracle_AB(m, b, eps_kind, eps, k0, 1.3, radius)
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(eps_kind, eps, k0, 1.3, radius)
    # (1) structural pattern (before eliminate_zeros): bit-exact integer parity
    As, Bs = ctx.matrix(0, eliminate_zeros=False), ctx.matrix(1, eliminate_zeros=False)
    As_ref, Bs_ref = _oracle_AB(m, b, eps_kind, eps, k0, 1.3, radius, eliminate_zeros=False)
    for got, ref in ((As, As_ref), (Bs, Bs_ref)):
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert abs(got - ref).max() < 1e-12 * abs(ref).max()
    # (2) after eliminate_zeros: values to 1e-12; pattern equal up to round-off zeros
    A, B = ctx.matrix(0), ctx.matrix(1)
    assert A.dtype == A_ref.dtype and B.dtype == B_ref.dtype
    dA, sameA = csr_rel_diff(A, A_ref)
    dB, sameB = csr_rel_diff(B, B_ref)
    assert dA < 1e-12 and dB < 1e-12, (dA, dB)
    okA, nA = _pattern_diff_is_roundoff(A, A_ref)
    okB, nB = _pattern_diff_is_roundoff(B, B_ref)
    assert okA and okB, "CSR sparsity pattern differs from the reference path beyond round-off zeros"
    assert nA <= 0.02 * A_ref.nnz and nB <= 0.02 * B_ref.nnz
    if jitter == 0.0:
        # structured mesh: the oracle eliminates exact zeros; the GPU keeps them only as round-off
        assert A_ref.nnz <= A.nnz <= As.nnz


@pytest.mark.parametrize("order,eps_kind,cplx", [(2, 0, False), (2, 2, True), (1, 0, True), (1, 2, False)])
def test_constant_coefficient_kernel_matches_quadrature_kernel(ctx, order, eps_kind, cplx):
    """radius = inf and element-wise constant eps take the pre-integrated K1 (k_element_blocks_const);
    radius = 1e300 forces the quadrature-loop K1 on the same problem (bend factor 1 + x/R == 1 exactly)."""
    m, _ = strip_problem(11, jitter=0.15)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    eps = _make_eps(m, eps_kind, cplx, seed=3)
    k0 = 2 * np.pi / 1.31
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(eps_kind, eps, k0, 1.0, np.inf)
    assert ctx.timings()["element_kernel_launches"] == 1
    A1, B1 = ctx.matrix(0, eliminate_zeros=False), ctx.matrix(1, eliminate_zeros=False)
    ctx.assemble(eps_kind, eps, k0, 1.0, 1e300)
    assert ctx.timings()["element_kernel_launches"] == 3
    A2, B2 = ctx.matrix(0, eliminate_zeros=False), ctx.matrix(1, eliminate_zeros=False)
    for x, y in ((A1, A2), (B1, B2)):
        assert np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices)
        assert abs(x - y).max() < 1e-13 * abs(y).max()


def test_assembly_golden_checksums(ctx):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "strip_n16.json")))
    m, eps = strip_problem(g["n"])
    b = Basis(m, mode_element(2), intorder=4)
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(0, eps, 2 * np.pi / g["wavelength"])
    A, B = ctx.matrix(0), ctx.matrix(1)
    # nnz after eliminate_zeros differs from the oracle's only by round-off zeros
    assert 0 <= A.nnz - g["nnzA"] <= 0.01 * g["nnzA"] and 0 <= B.nnz - g["nnzB"] <= 0.01 * g["nnzB"]
    assert abs(abs(A).sum() - g["sumabsA"]) < 1e-11 * g["sumabsA"]
    assert abs(abs(B).sum() - g["sumabsB"]) < 1e-11 * g["sumabsB"]


def test_hfield_operators_match_oracle(ctx):
    m, eps = strip_problem(8)
    b = Basis(m, mode_element(2), intorder=4)
    beta = 7.3 + 0.2j
    C_ref, M_ref = fo.assemble_hfield_operators(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), 2, b.X,
                                                b.W, beta, b.N)
    ctx.set_space(b)
    ctx.symbolic()
    Mass, C0, C1 = ctx.matrix(2), ctx.matrix(3), ctx.matrix(4)
    dM, _ = csr_rel_diff(Mass.astype(complex), M_ref)
    dC, _ = csr_rel_diff((C0 + 1j * beta * C1).tocsr(), C_ref)
    assert dM < 1e-12 and dC < 1e-12


# ----------------------------------------------------------------------------- K4: SpMV
@pytest.mark.parametrize("cplx_eps,cplx_x", [(False, False), (False, True), (True, True), (True, False)])
def test_spmv_matches_scipy(ctx, cplx_eps, cplx_x):
    m, _ = strip_problem(12)
    b = Basis(m, mode_element(2), intorder=4)
    eps = _make_eps(m, 0, cplx_eps)
    k0 = 2 * np.pi / 1.55
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(0, eps, k0)
    A = ctx.matrix(0)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(b.N) + (1j * rng.standard_normal(b.N) if cplx_x else 0)
    y = ctx.spmv(0, x)
    ref = A @ x
    assert np.linalg.norm(y - ref) < 1e-13 * np.linalg.norm(ref)


# ----------------------------------------------------------------------------- sparse LU
@pytest.mark.parametrize("order,n,cplx,dirichlet,leaf", [(2, 12, False, False, 0), (2, 24, False, True, 16),
                                                       (1, 30, False, False, 8), (2, 16, True, False, 0),
                                                       (2, 40, False, False, 0),
                                                       # top separators > 384 dofs: cluster chains (register-resident
                                                       # f64 kernels with inverted diagonal blocks; c128 generic)
                                                       (2, 110, False, False, 0), (2, 100, True, True, 0)])
def test_lu_solve_matches_scipy(ctx, order, n, cplx, dirichlet, leaf):
    m, _ = strip_problem(n)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    eps = _make_eps(m, 0, cplx)
    k0 = 2 * np.pi / 1.55
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(0, eps, k0)
    A, B = ctx.matrix(0), ctx.matrix(1)
    sigma = k0**2 * np.max(eps.real) * 1.1
    D = b.get_dofs(None) if dirichlet else None
    ctx.lu_factor(sigma, D, leaf)
    act = np.ones(b.N, dtype=bool)
    if dirichlet:
        act[D] = False
    rng = np.random.default_rng(3)
    for rhs_c in (False, True):
        rhs = rng.standard_normal(b.N) + (1j * rng.standard_normal(b.N) if rhs_c else 0)
        rhs = rhs * act
        x = ctx.lu_solve(rhs)
        I = np.nonzero(act)[0]
        OP = (-A + sigma * B).tocsr()[I][:, I]
        xs = spla.splu(OP.tocsc().astype(complex)).solve(rhs[I].astype(complex))
        assert np.all(x[~act] == 0)
        assert np.linalg.norm(OP @ x[I] - rhs[I]) < 1e-11 * np.linalg.norm(rhs)
        assert np.linalg.norm(x[I] - xs) < 1e-9 * np.linalg.norm(xs)


@pytest.mark.parametrize("order,n,cplx,leaf", [(2, 40, False, 0), (2, 30, True, 8), (1, 60, False, 4)])
def test_lu_solve_schedules_agree(ctx, order, n, cplx, leaf, monkeypatch):
    """The triangular solves run either level by level or as dataflow bands (persistent kernels that draw fronts from
    a ticket counter and wait on per-front completion flags).  Same factors, same arithmetic per front: the schedules
    must agree to round-off, for one and two right-hand sides, also when the whole tree (root included) is one band
    and when the solve is repeated (flag epochs, CUDA-graph replay of the third call)."""
    m, _ = strip_problem(n)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    eps = _make_eps(m, 0, cplx)
    k0 = 2 * np.pi / 1.55
    ctx.set_space(b)
    ctx.symbolic()
    ctx.assemble(0, eps, k0)
    A, B = ctx.matrix(0), ctx.matrix(1)
    sigma = k0**2 * np.max(eps.real) * 1.1
    ctx.lu_factor(sigma, None, leaf)
    OP = (-A + sigma * B).tocsr()
    rng = np.random.default_rng(11)
    rhs1 = rng.standard_normal(b.N)
    rhs2 = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    out = {}
    for name, env in (("levels", {}), ("bands", {"FW_TREE": "1"}), ("one band", {"FW_TREE": "1", "FW_TREE_MIN_FRONTS": "1"}),
                      ("levels + L2 prefetch", {"FW_SOLVE_PF_KB": "48"}),
                      ("levels, substitution chains instead of inverse pivot blocks", {"FW_NO_TINV_SOLVE": "1"}),
                      ("levels, serial 32-step diagonal solves instead of the pre-inverted diagonal blocks",
                       {"FW_NO_DINV_CHAIN": "1"}),
                      ("one band, 64 threads", {"FW_TREE": "1", "FW_TREE_MIN_FRONTS": "1", "FW_TREE_SMALL_M": "100000",
                                                "FW_TREE_THREADS_A": "64"})):
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            mp.setenv("FW_NO_SOLVE_GRAPH", "1")  # a captured graph would keep the schedule of its first capture
            xs = [ctx.lu_solve(r) for r in (rhs1, rhs2, rhs1, rhs1)]
        assert np.array_equal(xs[0], xs[2]) and np.array_equal(xs[0], xs[3]), name  # deterministic, epoch-safe
        for r, x in zip((rhs1, rhs2), xs):
            assert np.linalg.norm(OP @ x - r) < 1e-11 * np.linalg.norm(r), name
        out[name] = xs
    for name in out:
        for a, c in zip(out[name][:2], out["levels"][:2]):
            assert np.linalg.norm(a - c) < 1e-11 * np.linalg.norm(c), name


# ----------------------------------------------------------------------------- eigen-solve
def _match(l_gpu, l_ref):
    """pair eigenvalues by value"""
    idx = [int(np.argmin(abs(l_ref - z))) for z in l_gpu]
    assert sorted(idx) == list(range(len(l_ref))), (l_gpu, l_ref)
    return idx


@pytest.mark.parametrize("case", range(4))
def test_eigs_rectangle_known_answers_and_oracle(ctx, case):
    """metallic boundaries; reference literature values (mode_solver_rectangle.py:45) and oracle parity."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps, bnd, fsel, ref = rectangle_problem(20, case)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.5, num_modes=1, order=2, metallic_boundaries=bnd)
    X, W = triangle_quadrature(4)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, X, W, num_modes=1, order=2, dirichlet_facets=fsel, with_h=False)
    assert abs(modes[0].n_eff - r.n_effs[0]) < 1e-9 * abs(r.n_effs[0])
    assert abs(modes[0].n_eff.real - ref) < 5e-5


@pytest.mark.parametrize("order,n,k,cplx", [(2, 16, 2, False), (1, 24, 3, False), (2, 14, 2, True), (2, 20, 6, False)])
def test_compute_modes_matches_oracle(ctx, order, n, k, cplx):
    from femwell_b200.maxwell.waveguide import calculate_overlap, compute_modes

    m, eps = strip_problem(n)
    if cplx:
        eps = eps.astype(complex)
        eps[m.subdomains["clad"]] -= 0.05j
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=k, order=order)
    b = modes[0].basis
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, b.X, b.W, num_modes=k, order=order, v0=v0)
    idx = _match(np.array([md.k**2 for md in modes]), r.lams)
    ed = b.element_dofs.astype(np.int64)
    k0 = 2 * np.pi / 1.55
    zd = b.split_indices()[1]

    def eig_residual(E, lam):
        x = np.array(E, dtype=complex)
        x[zd] *= 1j * np.sqrt(lam / k0**4)  # back to the solved variable E_3 = i beta E_z
        return np.linalg.norm(-(r.A @ x) + lam * (r.B @ x)) / np.linalg.norm(r.A @ x)

    for i, md in enumerate(modes):
        j = idx[i]
        assert abs(md.n_eff - r.n_effs[j]) < 1e-9 * abs(r.n_effs[j]), (md.n_eff, r.n_effs[j])
        # the GPU eigenvector satisfies the oracle's pencil
        assert eig_residual(md.E, md.k**2) < 1e-10
        # normalised overlap with the oracle's mode (phase-free): |<gpu, ref>| = 1 within 1e-6.
        # scipy/ARPACK returns vectors polluted by null(M) (the E_z block) for the last Ritz pairs --
        # eigenvalue exact, eigen-residual O(1) -- so the field comparison is only meaningful where the
        # oracle's own vector is an eigenvector.
        ov = fo.calculate_overlap(m.p, m.t, ed, order, b.X, b.W, md.E, md.H, r.E[:, j], r.H[:, j])
        res_ref = eig_residual(r.E[:, j], r.lams[j])
        print("mode", i, "->", j, "n_eff", md.n_eff, r.n_effs[j], "overlap", ov, "oracle eig residual", res_ref)
        if res_ref < 1e-9:
            assert abs(abs(ov) - 1) < 1e-6, ov
        # power normalisation through the GPU functional
        assert abs(md.calculate_power() - 1) < 1e-9
        assert abs(calculate_overlap(b, md.E, md.H, b, md.E, md.H) - 1) < 1e-9
    # modes come back nearest-to-sigma first
    lam = np.array([md.k**2 for md in modes])
    sig = modes.stats and (2 * np.pi / 1.55) ** 2 * np.max(eps) * 1.1
    assert np.all(np.diff(abs(lam - sig)) >= -1e-9 * abs(sig))


def test_golden_strip_modes(ctx):
    from femwell_b200.maxwell.waveguide import compute_modes

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "strip_n16.json")))
    m, eps = strip_problem(g["n"])
    modes = compute_modes(Basis(m, ElementTriP0(), intorder=4), eps, g["wavelength"], num_modes=2, order=2)
    for i in range(2):
        assert abs(modes[i].n_eff.real - g["n_eff"][i][0]) < 1e-9 * g["n_eff"][i][0]
        assert abs(modes[i].te_fraction - g["te_fraction"][i]) < 1e-6


def test_bend_with_pml_complex_dgp1(ctx):
    """config-3 style: DG-P1 complex eps with PML ramp, finite radius, n_guess (bent_waveguide.py:73-119)."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m = __import__("femwell_b200.mesh", fromlist=["waveguide_mesh"]).waveguide_mesh(
        14, slab_h=0.11, bbox=(-1.25, 4.25, -1.0, 1.22), extra_x=[2.25])
    b0 = Basis(m, ElementDG(ElementTriP1()), intorder=4)
    ne = m.nelements
    base = np.full(ne, 1.0)
    for name, nn in {"core": 3.48, "slab": 3.48, "box": 1.48, "clad": 1.0}.items():
        base[m.subdomains[name]] = nn**2
    eps = np.repeat(base, 3).astype(complex)
    xv = m.p[0, m.t].T.reshape(-1)  # DG-P1 dof (e, a) sits on vertex a of element e
    eps += -10j * np.maximum(0, xv - 2.25) ** 2
    straight = compute_modes(b0, eps, 1.55, num_modes=1, order=2)
    bent = compute_modes(b0, eps, 1.55, num_modes=1, order=2, radius=5.0, n_guess=straight[0].n_eff)
    X, W = b0.X, b0.W
    r = fo.compute_modes(m.p, m.t, 1, eps, 1.55, X, W, num_modes=1, order=2, radius=5.0,
                         n_guess=straight[0].n_eff, with_h=False)
    assert abs(bent[0].n_eff - r.n_effs[0]) < 1e-9 * abs(r.n_effs[0])
    assert bent[0].n_eff.imag != 0


# ----------------------------------------------------------------------------- post-processing
def test_hfield_and_functionals_match_oracle(ctx):
    from femwell_b200.maxwell.waveguide import calculate_hfield, compute_modes, speed_of_light

    m, eps = strip_problem(14)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=2, order=2)
    md, md2 = modes[0], modes[1]
    b = md.basis
    ed = b.element_dofs.astype(np.int64)
    args = (m.p, m.t.astype(np.int64), ed, 2, b.X, b.W)
    H_ref = fo.calculate_hfield(*args, md.E, md.k, md.k0 * speed_of_light, b.N)
    H = calculate_hfield(b, md.E, md.k, md.k0 * speed_of_light)
    assert np.linalg.norm(H - H_ref) < 1e-9 * np.linalg.norm(H_ref)
    assert np.linalg.norm(md.H - H_ref) < 1e-9 * np.linalg.norm(H_ref)
    assert abs(md.te_fraction - fo.te_fraction(*args, md.E)) < 1e-12
    assert abs(md.transversality - fo.transversality(*args, md.E)) < 1e-12
    for f in ("xy", "x", "y"):
        ref = fo.effective_area(*args, md.E, f)
        assert abs(md.calculate_effective_area(f) - ref) < 1e-10 * abs(ref)
    ov_ref = fo.calculate_overlap(*args, md.E, md.H, md2.E, md2.H)
    assert abs(md.calculate_overlap(md2) - ov_ref) < 1e-10
    core = m.subdomains["core"]
    p_ref = fo.calculate_overlap(*args, md.E, md.H, md.E, md.H, subset=core)
    assert abs(md.calculate_power(elements="core") - p_ref) < 1e-10
    assert 0.5 < md.calculate_power(elements="core").real < 1.0
    deps = np.zeros(m.nelements)
    deps[core] = 0.01
    cc_ref = fo.coupling_coefficient(*args, md.E, md2.E, fo.interpolate_eps(0, deps, m.nelements, b.X))
    cc = md.calculate_coupling_coefficient(md2, deps)
    assert abs(cc - cc_ref) < 1e-10 * max(1.0, abs(cc_ref))
    cf_ref = fo.confinement_factor(*args, md.E, fo.interpolate_eps(0, eps[core], core.size, b.X), core)
    cf = md.calculate_confinement_factor("core")
    assert abs(cf - cf_ref) < 1e-10 * abs(cf_ref)
    assert abs(md.calculate_pertubated_neff(deps) - md.n_eff) > 0
    # a16: intensity projected onto DG-P1
    basis2, inten = md.calculate_intensity()
    inten_ref = fo.calculate_intensity(*args, md.E, md.H)
    assert inten.shape == (3 * m.nelements,) and basis2.N == 3 * m.nelements
    assert np.linalg.norm(inten - inten_ref) < 1e-10 * np.linalg.norm(inten_ref)


@pytest.mark.parametrize("order", [1, 2])
def test_poynting_matches_oracle(ctx, order):
    """Mode.poynting / Sx, Sy, Sz (waveguide.py:74-110): DG projection of E x H."""
    from femwell_b200.maxwell.waveguide import Mode

    m, eps = strip_problem(9, jitter=0.15)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    rng = np.random.default_rng(11)
    E = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    H = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    ref = fo.poynting(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), order, b.X, b.W, E, H)
    mode = Mode(frequency=1.0, k=1.0, basis_epsilon_r=b.with_element(ElementTriP0()), epsilon_r=eps, basis=b, E=E, H=H)
    pb, P = mode.poynting
    nl = 3 if order == 1 else 6
    assert P.shape == (m.nelements * 3 * nl,) and pb.N == P.size
    assert np.abs(P - ref).max() < 1e-11 * np.abs(ref).max()
    for c, (bc, S) in enumerate((mode.Sx, mode.Sy, mode.Sz)):
        assert bc.N == S.size == m.nelements * nl
        assert np.array_equal(S, P[c::3])


@pytest.mark.parametrize("order", [1, 2])
def test_energy_current_density_matches_oracle(ctx, order):
    from femwell_b200.maxwell.waveguide import calculate_energy_current_density

    m, _ = strip_problem(9, jitter=0.15)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    rng = np.random.default_rng(12)
    xs = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    ref = fo.energy_current_density(m.p, m.t.astype(np.int64), b.element_dofs.astype(np.int64), order, b.X, b.W, xs)
    be, got = calculate_energy_current_density(b, xs)
    assert be.N == got.size == m.nelements
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("order_i,order_j", [(2, 2), (1, 2), (2, 1)])
def test_cross_mesh_overlap_matches_oracle(ctx, order_i, order_j):
    """calculate_overlap / calculate_scalar_product between modes on DIFFERENT meshes (waveguide.py:737-782)."""
    from femwell_b200.maxwell.waveguide import calculate_overlap, calculate_scalar_product

    mi, _ = strip_problem(9, jitter=0.15)
    mj, _ = strip_problem(12, jitter=0.1)
    bi = Basis(mi, ElementTriP0(), intorder=4).with_element(mode_element(order_i))
    bj = Basis(mj, ElementTriP0(), intorder=4).with_element(mode_element(order_j))
    rng = np.random.default_rng(21)

    def smooth(b):
        # smooth-ish fields: low-order random combination so that the P1 projection is well resolved
        return rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)

    E_i, H_i, E_j, H_j = smooth(bi), smooth(bi), smooth(bj), smooth(bj)
    args = (mi.p, mi.t.astype(np.int64), bi.element_dofs.astype(np.int64), order_i, bi.X, bi.W, E_i, H_i,
            mj.p, mj.t.astype(np.int64), bj.element_dofs.astype(np.int64), order_j, bj.X, bj.W, E_j, H_j)
    ref = fo.overlap_cross(*args, mode=0)
    got = calculate_overlap(bi, E_i, H_i, bj, E_j, H_j)
    assert abs(got - ref) < 1e-10 * abs(ref), (got, ref)
    ref1 = fo.overlap_cross(*args, mode=1)
    got1 = calculate_scalar_product(bi, E_i, bj, H_j)
    assert abs(got1 - ref1) < 1e-10 * abs(ref1), (got1, ref1)
    # same basis: the scalar product reduces to the Functional on one mesh
    same = calculate_scalar_product(bi, E_i, bi, H_i)
    full = calculate_overlap(bi, E_i, H_i, bi, E_i, H_i)
    assert abs(0.5 * (same + np.conj(same)) - full) < 1e-12 * abs(full)
    # a foreign mesh that does not cover basis_i is an error, like scikit-fem's element finder
    mk = MeshTri(mj.p * 0.5, mj.t)
    bk = Basis(mk, ElementTriP0(), intorder=4).with_element(mode_element(order_j))
    with pytest.raises(Exception, match="outside of the mesh"):
        calculate_overlap(bi, E_i, H_i, bk, E_j, H_j)


def test_wavelength_sweep_reuses_symbolic_work_and_matches_cold_solves(ctx):
    """Fixed-mesh sweep (vary_wavelength.py:83-95): CSR pattern and LU analysis are reused between the points;
    every point must equal a cold solve of the same problem."""
    from femwell_b200.maxwell.waveguide import compute_modes, get_context
    from femwell_b200.sweep import group_index, wavelength_sweep

    m, eps = strip_problem(14)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    wls = np.linspace(1.5, 1.6, 4)
    eps_of = lambda wl: eps * (1.0 + 0.02 * (wl - 1.55))
    table = wavelength_sweep(b0, eps_of, wls, num_modes=2, order=2)
    assert table.shape == (4, 2)
    assert get_context().timings()["lu_symbolic_ms"] < 5.0  # last point: analysis reused
    for i, wl in enumerate(wls):
        get_context()._space_key = None  # cold: new space, new pattern, new analysis
        cold = compute_modes(b0, eps_of(wl), wl, num_modes=2, order=2).n_effs
        assert np.abs(table[i] - cold).max() < 1e-11
    assert np.all(np.isfinite(group_index(wls, table[:, 0].real)))


def test_sweep_two_gpus_nccl():
    """One process per GPU, block partition of the sweep, one NCCL all_gather (skipped on a single-GPU box)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "sweep_nccl_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SWEEP_NCCL_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_argument_validation(ctx):
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(6)
    b0 = Basis(m, ElementTriP0())
    with pytest.raises(ValueError):
        compute_modes(b0, eps, 1.55, solver="scipy")
    with pytest.raises(AssertionError):
        compute_modes(b0, eps, 1.55, order=3)


# ----------------------------------------------------------------------------- size-independent properties at larger size
def test_large_assembly_properties(ctx):
    """~180k triangles: linearity in eps, row-sum identities, spmv consistency (no oracle run)."""
    m, eps = strip_problem(300)
    b = Basis(m, mode_element(2), intorder=4)
    k0 = 2 * np.pi / 1.55
    ctx.set_space(b)
    nnz = ctx.symbolic()
    ctx.assemble(0, eps, k0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(b.N)
    y1 = ctx.spmv(1, x)
    ctx.assemble(0, 2 * eps, k0)
    y2 = ctx.spmv(1, x)
    # B does not depend on eps
    assert np.linalg.norm(y1 - y2) == 0
    # mass matrix integrates the constant: sum over P-block of M = area
    Mass = ctx.matrix(2)
    zd = b.split_indices()[1]
    one_z = np.zeros(b.N)
    one_z[zd] = 1.0  # sum of all P2 shape functions == 1
    area = (m.p[0].max() - m.p[0].min()) * (m.p[1].max() - m.p[1].min())
    assert abs(one_z @ (Mass @ one_z) - area) < 1e-10 * area
    assert abs(one_z @ ctx.spmv(2, one_z) - area) < 1e-10 * area
    assert nnz >= Mass.nnz


# ----------------------------------------------------------------------------- EigenSolver protocol on generic CSR
@pytest.mark.parametrize("order,n,cplx,with_coords,condensed", [(2, 12, False, False, False), (1, 20, False, True, True),
                                                              (2, 10, True, False, False), (2, 24, False, False, True)])
def test_solver_eigen_b200_protocol_matches_scipy(ctx, order, n, cplx, with_coords, condensed):
    """solver(K, M) -> (lams, xs) on matrices assembled elsewhere (here: by the oracle), as skfem.solve would
    call it (femwell/solver.py:87-142, waveguide.py:611-625)."""
    from femwell_b200.solver import solver_eigen_b200

    m, eps = strip_problem(n)
    if cplx:
        eps = eps.astype(complex)
        eps[m.subdomains["clad"]] -= 0.03j
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    k0 = 2 * np.pi / 1.55
    A, B = _oracle_AB(m, b, 0, eps, k0)
    K, M = (-A).tocsr(), (-B).tocsr()
    coords = None
    if condensed:  # what skfem's condense hands to the solver: the interior sub-matrices
        I = np.setdiff1d(np.arange(b.N), b.get_dofs(None))
        K, M = K[I][:, I], M[I][:, I]
    sigma = k0**2 * np.max(eps.real) * 1.1
    if with_coords:
        # dof coordinates: vertices, facet midpoints (order 1: Nv + Nf unknowns)
        xy = np.concatenate([m.p, m.facet_midpoints()], axis=1)
        coords = xy[:, I] if condensed else xy
    solver = solver_eigen_b200(k=3, sigma=sigma, coords=coords)
    lams, xs = solver(K, M)
    ref, _ = spla.eigs(K, M=M, k=3, sigma=sigma, v0=np.random.default_rng(0).uniform(-1, 1, K.shape[0]))
    assert xs.shape == (K.shape[0], 3)
    for i, lam in enumerate(lams):
        assert np.min(abs(ref - lam)) < 1e-9 * abs(lam)
        x = xs[:, i]
        assert np.linalg.norm(K @ x - lam * (M @ x)) < 1e-9 * np.linalg.norm(K @ x)
        assert abs(np.linalg.norm(x) - 1) < 1e-12
    assert solver.stats["nconv"] == 3


def test_solver_eigen_b200_generic_stencil(ctx):
    """not an FE matrix at all: 2-D 5-point Laplacian pencil with a diagonal mass matrix."""
    from femwell_b200.solver import solver_eigen_b200

    nx = 40
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nx, nx))
    K = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))).tocsr()
    M = sp.diags(np.linspace(1.0, 2.0, nx * nx)).tocsr()
    lams, xs = solver_eigen_b200(k=4, sigma=0.05)(K, M)
    ref = spla.eigs(K, M=M, k=4, sigma=0.05, return_eigenvectors=False)
    for lam in lams:
        assert np.min(abs(ref - lam)) < 1e-10 * abs(lam)


# ============================================================================= round 2: BASELINE configs and branches
def _z_rescale(b, E, lam, k0):
    """E (un-scaled E_z) back to the solved variable E_3 = i beta E_z (waveguide.py:627 inverted)"""
    x = np.array(E, dtype=complex)
    x[b.split_indices()[1]] *= 1j * np.sqrt(lam / k0**4)
    return x


def _compare_modes_with_oracle(modes, r, m, b, wavelength, order, check_h=True, dirichlet=None):
    """n_eff 1e-9, eigen-residual against the oracle's pencil, normalised overlap with the oracle's mode 1e-6,
    H against the oracle's calculate_hfield 1e-9, unit power."""
    from femwell_b200.maxwell.waveguide import speed_of_light

    k0 = 2 * np.pi / wavelength
    ed = b.element_dofs.astype(np.int64)
    args = (m.p, m.t.astype(np.int64), ed, order, b.X, b.W)
    idx = _match(np.array([md.k**2 for md in modes]), r.lams)
    keep = np.ones(b.N, dtype=bool)
    if dirichlet is not None:
        keep[dirichlet] = False
    for i, md in enumerate(modes):
        j = idx[i]
        assert abs(md.n_eff - r.n_effs[j]) < 1e-9 * abs(r.n_effs[j]), (md.n_eff, r.n_effs[j])
        x = _z_rescale(b, md.E, md.k**2, k0)
        res = np.linalg.norm((-(r.A @ x) + md.k**2 * (r.B @ x))[keep]) / np.linalg.norm((r.A @ x)[keep])
        assert res < 1e-10, res
        if dirichlet is not None:
            assert np.all(md.E[dirichlet] == 0)
        if check_h:
            H_ref = fo.calculate_hfield(*args, md.E, md.k, md.k0 * speed_of_light, b.N)
            assert np.linalg.norm(md.H - H_ref) < 1e-9 * np.linalg.norm(H_ref)
            assert abs(md.calculate_power() - 1) < 1e-9
            xr = _z_rescale(b, r.E[:, j], r.lams[j], k0)
            res_ref = np.linalg.norm((-(r.A @ xr) + r.lams[j] * (r.B @ xr))[keep]) / np.linalg.norm((r.A @ xr)[keep])
            if res_ref < 1e-9:  # ARPACK's last vectors may be polluted by null(M), DESIGN.md section 2
                ov = fo.calculate_overlap(*args, md.E, md.H, r.E[:, j], r.H[:, j])
                assert abs(abs(ov) - 1) < 1e-6, ov


def test_config1_strip_5k_triangles_matches_oracle(ctx):
    """BASELINE config 1 exactly: Si strip 500x220 nm in SiO2, lambda 1.55, 2*50^2 = 5000 triangles, order 2,
    num_modes 2, intorder 4 (the reference's CPU-runnable case)."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(50)
    assert m.nelements == 5000
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=2, order=2)
    b = modes[0].basis
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, b.X, b.W, num_modes=2, order=2, v0=v0)
    _compare_modes_with_oracle(modes, r, m, b, 1.55, 2)
    assert modes[0].n_eff.real > modes[1].n_eff.real > 1.444  # guided, fundamental first
    assert modes[0].te_fraction > 0.9


def test_bend_with_pml_fields_match_oracle(ctx):
    """config-3 style (bent_waveguide.py:73-119): DG-P1 complex eps with a PML ramp, radius 5, n_guess -- H field,
    power normalisation and overlap with the oracle's mode, not only n_eff."""
    from femwell_b200.maxwell.waveguide import compute_modes
    from femwell_b200.mesh import waveguide_mesh

    m = waveguide_mesh(14, slab_h=0.11, bbox=(-1.25, 4.25, -1.0, 1.22), extra_x=[2.25])
    b0 = Basis(m, ElementDG(ElementTriP1()), intorder=4)
    base = np.full(m.nelements, 1.0)
    for name, nn in {"core": 3.48, "slab": 3.48, "box": 1.48, "clad": 1.0}.items():
        base[m.subdomains[name]] = nn**2
    eps = np.repeat(base, 3).astype(complex)
    xv = m.p[0, m.t].T.reshape(-1)
    eps += -10j * np.maximum(0, xv - 2.25) ** 2
    straight = compute_modes(b0, eps, 1.55, num_modes=1, order=2)
    bent = compute_modes(b0, eps, 1.55, num_modes=2, order=2, radius=5.0, n_guess=straight[0].n_eff)
    b = bent[0].basis
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    r = fo.compute_modes(m.p, m.t, 1, eps, 1.55, b.X, b.W, num_modes=2, order=2, radius=5.0,
                         n_guess=straight[0].n_eff, v0=v0)
    _compare_modes_with_oracle(bent, r, m, b, 1.55, 2)
    assert bent[0].n_eff.imag != 0 and bent[0].H.imag.any()
    # lossy problem: the power is complex before normalisation, exactly 1 after (waveguide.py:636-638)
    assert abs(bent[0].calculate_overlap(bent[0]) - 1) < 1e-9


@pytest.mark.parametrize("cplx", [False, True])
def test_anisotropic_epsilon_through_compute_modes(ctx, cplx):
    """ElementVector(ElementTriP0(), 3): eps_t = diag(eps_0, eps_1), eps_z = eps_2 (waveguide.py:579-586,
    waveguide_anisotropic.py:81-84) through the whole solve, not only the assembly."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m, _ = strip_problem(16)
    b0 = Basis(m, ElementVector(ElementTriP0(), 3), intorder=4)
    eps = _make_eps(m, 2, cplx, seed=5)
    assert eps.size == 3 * m.nelements
    modes = compute_modes(b0, eps, 1.55, num_modes=2, order=2)
    b = modes[0].basis
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    r = fo.compute_modes(m.p, m.t, 2, eps, 1.55, b.X, b.W, num_modes=2, order=2, v0=v0)
    _compare_modes_with_oracle(modes, r, m, b, 1.55, 2)
    # the isotropic solve of the same structure differs: the anisotropy really entered
    iso = compute_modes(Basis(m, ElementTriP0(), intorder=4), eps.reshape(-1, 3)[:, 0].copy(), 1.55, num_modes=1,
                        order=2)
    assert abs(iso[0].n_eff - modes[0].n_eff) > 1e-4


@pytest.mark.parametrize("case,order", [(0, 1), (2, 1), (1, 2)])
def test_metallic_boundaries_order_1_and_2(ctx, case, order):
    """condense on named boundaries (waveguide.py:610-619) at both orders; fields and H, not only n_eff."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps, bnd, fsel, ref = rectangle_problem(24, case)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.5, num_modes=2, order=order, metallic_boundaries=bnd)
    b = modes[0].basis
    D = b.get_dofs(bnd)
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, b.X, b.W, num_modes=2, order=order, dirichlet_facets=fsel, v0=v0)
    assert np.array_equal(np.sort(D), fo.boundary_dofs(order, m.nvertices, m.facets.astype(np.int64), fsel))
    _compare_modes_with_oracle(modes, r, m, b, 1.5, order, dirichlet=D)
    tol = 5e-5 if order == 2 else 5e-4  # n = 24: discretisation error vs the literature value (O(h^2) at order 1)
    assert abs(modes[0].n_eff.real - ref) < tol


def test_config2_at_scale_residual_and_leaf_size_independence(ctx):
    """BASELINE config 2 at full size (500k triangles, order 2, 4 modes): the assembled pencil is exported and the
    eigen-residual of every returned mode is evaluated ON THE HOST (scipy SpMV), the power normalisation is
    re-checked, and n_eff must not depend on the nested-dissection leaf size (a different elimination order)."""
    import sys

    from femwell_b200.maxwell.waveguide import compute_modes, get_context

    sys.path.insert(0, ROOT)
    from bench import make_problem

    m, eps = make_problem(500)
    assert m.nelements == 500000
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=4, order=2)
    b = modes[0].basis
    assert b.N == 3504001
    c = get_context()
    A, B = c.matrix(0), c.matrix(1)
    assert A.shape == (b.N, b.N) and A.nnz > 79_000_000 and B.nnz > 28_000_000
    k0 = 2 * np.pi / 1.55
    for md in modes:
        x = _z_rescale(b, md.E, md.k**2, k0)
        Ax = A @ x
        res = np.linalg.norm(-Ax + md.k**2 * (B @ x)) / np.linalg.norm(Ax)
        assert res < 1e-10, res
        assert abs(md.calculate_power() - 1) < 1e-9
    assert modes.stats["max_residual"] < 1e-10 and modes.stats["nconv"] == 4
    assert modes.stats["timings"]["pcg_rel_residual"] <= 1e-10
    n1 = modes.n_effs
    del A, B
    other = compute_modes(b0, eps, 1.55, num_modes=4, order=2, solver_options={"leaf_elements": 32})
    assert other.stats["lu_fronts"] != modes.stats["lu_fronts"]
    assert np.abs(other.n_effs - n1).max() < 1e-10 * abs(n1).max()
    # mode-field overlaps between the two solves (normalised, phase free): the same modes
    for a, bb in zip(modes, other):
        assert abs(abs(a.calculate_overlap(bb)) - 1) < 1e-6


@pytest.mark.parametrize("case", range(8))
def test_rib_known_answers(ctx, case):
    """The eight rib values of docs/benchmarks/mode_solver.py:52-68 at the reference's 5e-6 (:160), and the oracle's
    value on the same mesh (tests/golden/rib_neff.json, pinned by the CPU suite) at 1e-9."""
    import sys

    from femwell_b200.maxwell.waveguide import compute_modes

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_rib import PAPER, SLABS, rib_problem

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "rib_neff.json")))
    m, eps = rib_problem(SLABS[case])
    modes = compute_modes(Basis(m, ElementTriP0()), eps, 1.15, num_modes=1, order=2)  # default 3-point rule
    assert abs(modes[0].n_eff.real - PAPER[case]) < 5e-6
    assert abs(modes[0].n_eff.real - g["oracle"][case]) < 1e-9 * g["oracle"][case]


def test_rib_known_answer_on_a_fine_mesh(ctx):
    """the same benchmark on a ~100k-triangle mesh the CPU oracle would not finish in test time"""
    import sys

    from femwell_b200.maxwell.waveguide import compute_modes

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_rib import PAPER, rib_problem

    m, eps = rib_problem(0.4, h0=0.02, ratio=1.1)
    assert m.nelements > 50000
    modes = compute_modes(Basis(m, ElementTriP0()), eps, 1.15, num_modes=1, order=2)
    assert abs(modes[0].n_eff.real - PAPER[4]) < 5e-6


# ----------------------------------------------------------------------------- boundary: foreign basis, subsets, errors
@pytest.mark.parametrize("order", [1, 2])
def test_bridge_with_foreign_local_basis(ctx, order):
    """compute_modes_from_skfem(tables="skfem") with duck-typed scikit-fem objects whose ``lbasis`` is a deliberately
    different local basis of the same space: matrix entries equal the oracle driven by the same tables, n_eff is
    invariant, and the dof vectors are those of the foreign basis (x_builtin = T^T x_foreign entity by entity)."""
    from types import SimpleNamespace

    from test_cpu import FakeSkfemElement, foreign_tables_of

    from femwell_b200 import bridge
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(12)
    X, W = triangle_quadrature(4)
    kind, vec, sc, T = foreign_tables_of(order, X, W)
    own = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))

    class ElementTriP0_:  # scikit-fem's class name decides the epsilon kind
        pass

    ElementTriP0_.__name__ = "ElementTriP0"
    sk_mesh = SimpleNamespace(p=m.p, t=m.t, subdomains=dict(m.subdomains), boundaries=dict(m.boundaries))
    sk_eps_basis = SimpleNamespace(mesh=sk_mesh, elem=ElementTriP0_(), X=X, W=W, tind=None)
    sk_mode_basis = SimpleNamespace(elem=FakeSkfemElement(kind, vec, sc), element_dofs=own.element_dofs, N=own.N)
    RefMode = lambda **kw: SimpleNamespace(**kw)
    RefModes = lambda modes: list(modes)
    got = bridge.compute_modes_from_skfem(sk_eps_basis, eps, 1.55, num_modes=2, order=order, Mode=RefMode, Modes=RefModes,
                                          mode_basis=sk_mode_basis)
    assert got[0].basis is sk_mode_basis and got[0].basis_epsilon_r is sk_eps_basis
    ref = compute_modes(Basis(m, ElementTriP0(), intorder=4), eps, 1.55, num_modes=2, order=order)
    # (a) matrices of the foreign basis vs the oracle driven by the same tables
    from femwell_b200.maxwell.waveguide import get_context

    bt = own.with_tables(kind, vec, sc, element_dofs=own.element_dofs, n_dofs=own.N)
    c = get_context()
    c.set_space(bt)
    c.symbolic()
    k0 = 2 * np.pi / 1.55
    c.assemble(0, eps, k0)
    A, B = c.matrix(0), c.matrix(1)
    with fo.foreign_tables((kind, vec, sc)):
        A_ref, B_ref = _oracle_AB(m, own, 0, eps, k0)
    assert csr_rel_diff(A, A_ref)[0] < 1e-12 and csr_rel_diff(B, B_ref)[0] < 1e-12
    # (b) n_eff invariant, (c) dof vectors related by the change of basis
    ed = own.element_dofs
    for md_f, md_b in zip(got, ref):
        assert abs(md_f.k - md_b.k) < 1e-9 * abs(md_b.k)
        # element-wise: local coefficients x_builtin = T^T x_foreign (up to the common phase of the eigenvector)
        xf, xb = md_f.E[ed], md_b.E[ed]  # (nb, Ne)
        xb_from_f = T.T @ xf
        ph = np.vdot(xb_from_f.ravel(), xb.ravel())
        ph /= abs(ph)
        assert np.abs(xb_from_f * ph - xb).max() < 1e-7 * np.abs(xb).max()
        hb_from_f = T.T @ md_f.H[ed]
        assert np.abs(hb_from_f * ph - md_b.H[ed]).max() < 1e-7 * np.abs(md_b.H[ed]).max()


def test_element_subset_of_the_basis_is_honoured(ctx):
    """basis_epsilon_r.with_element keeps ``tind`` (waveguide.py:574): the forms are assembled over the subset only.
    The reference's pencil then has empty rows for every dof outside the subset and scipy's splu raises
    RuntimeError("Factor is exactly singular"); so does the GPU factorisation."""
    from femwell_b200.maxwell.waveguide import compute_modes, get_context

    m, eps = strip_problem(10)
    sel = np.nonzero(m.centroids()[0] < 0.4)[0].astype(np.int32)
    b0 = Basis(m, ElementTriP0(), intorder=4, elements=sel)
    b = b0.with_element(mode_element(2))
    assert np.array_equal(b.tind, sel)
    c = get_context()
    c.set_space(b)
    c.symbolic()
    k0 = 2 * np.pi / 1.55
    c.assemble(0, eps, k0)
    A, B = c.matrix(0), c.matrix(1)
    ed = b.element_dofs.astype(np.int64)
    eq = fo.interpolate_eps(0, eps[sel], sel.size, b.X)
    A_ref, B_ref = fo.assemble_waveguide(m.p, m.t.astype(np.int64), ed, 2, b.X, b.W, eq, k0, n=b.N, subset=sel)
    assert csr_rel_diff(A, A_ref)[0] < 1e-12 and csr_rel_diff(B, B_ref)[0] < 1e-12
    assert _pattern_diff_is_roundoff(A, A_ref)[0]
    untouched = np.setdiff1d(np.arange(b.N), np.unique(ed[:, sel]))
    assert untouched.size > 0 and A[untouched].nnz == 0
    with pytest.raises(RuntimeError, match="Factor is exactly singular"):
        compute_modes(b0, eps, 1.55, num_modes=1, order=2)
    # the whole mesh given as an explicit element list is the ordinary solve
    full = compute_modes(Basis(m, ElementTriP0(), intorder=4, elements=np.arange(m.nelements)), eps, 1.55, order=2)
    plain = compute_modes(Basis(m, ElementTriP0(), intorder=4), eps, 1.55, order=2)
    assert abs(full[0].n_eff - plain[0].n_eff) < 1e-12


def test_no_convergence_raises_arpack_compatible_error_with_partial_pairs(ctx):
    """scipy raises ArpackNoConvergence carrying the converged pairs (arpack.py:303-329)."""
    from scipy.sparse.linalg import ArpackNoConvergence

    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(16)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    with pytest.raises(ArpackNoConvergence) as ei:
        compute_modes(b0, eps, 1.55, num_modes=6, order=2, solver_options={"ncv": 8, "maxiter": 1, "tol": 1e-15})
    e = ei.value
    assert isinstance(e, RuntimeError) and e.code == -3
    n = Basis(m, mode_element(2)).N
    nconv = e.eigenvalues.shape[0]
    assert 0 <= nconv < 6 and e.eigenvectors.shape == (n, nconv)
    if nconv:
        good = compute_modes(b0, eps, 1.55, num_modes=6, order=2)
        lam = np.array([md.k**2 for md in good])
        for z in e.eigenvalues:
            assert np.min(abs(lam - z)) < 1e-8 * abs(z)
    # the handle stays usable
    assert compute_modes(b0, eps, 1.55, num_modes=1, order=2)[0].n_eff.real > 1.444


def test_hfield_rejects_non_finite_input_and_reports_pcg_stats(ctx):
    """The mass PCG must not return an unconverged H silently."""
    from femwell_b200._lib import FemwellB200Error
    from femwell_b200.maxwell.waveguide import calculate_hfield, compute_modes, get_context, speed_of_light

    m, eps = strip_problem(12)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=2, order=2)
    t = modes.stats["timings"]
    assert 0 < t["pcg_iterations"] < 600 and t["pcg_rel_residual"] <= 1e-10
    md = modes[0]
    bad = md.E.copy()
    bad[7] = np.nan
    with pytest.raises(FemwellB200Error):
        calculate_hfield(md.basis, bad, md.k, md.k0 * speed_of_light)
    # tighter and looser tolerances both land on the oracle's H within their own accuracy
    H13 = compute_modes(b0, eps, 1.55, num_modes=2, order=2, solver_options={"pcg_tol_exp": 13})[0].H
    H6 = compute_modes(b0, eps, 1.55, num_modes=2, order=2, solver_options={"pcg_tol_exp": 6})[0].H
    ph = np.vdot(H13, md.H)
    ph /= abs(ph)
    assert np.linalg.norm(H13 * ph - md.H) < 1e-8 * np.linalg.norm(md.H)
    ph6 = np.vdot(H13, H6)
    ph6 /= abs(ph6)
    assert 1e-13 < np.linalg.norm(H13 * ph6 - H6) / np.linalg.norm(H6) < 1e-4


def test_real_problem_returns_real_transverse_field(ctx):
    """real eps: ARPACK's real driver returns real eigenvectors for real eigenvalues; after the E_z un-scaling E_t is
    real and E_z imaginary up to one global phase, and the packed real PCG (two real systems per complex slot) must
    reproduce the general complex PCG."""
    import os as _os

    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(14)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=3, order=2)
    b = modes[0].basis
    td, zd = b.split_indices()
    for md in modes:
        assert md.k.imag == 0
        assert np.abs(md.E[td].imag).max() == 0 and np.abs(md.E[zd].real).max() == 0
        assert np.abs(md.H[td].imag).max() == 0 and np.abs(md.H[zd].real).max() == 0


@pytest.mark.parametrize("order,eps_kind,cplx", [(1, 0, False), (2, 0, False), (2, 1, True), (1, 1, False)])
def test_error_estimator_matches_oracle(ctx, order, eps_kind, cplx):
    """Mode.eval_error_estimator (waveguide.py:473-502) on a solved mode and on a random field."""
    from femwell_b200.maxwell.waveguide import Mode, compute_modes

    m, _ = strip_problem(12, jitter=0.15)
    elem = ElementTriP0() if eps_kind == 0 else ElementDG(ElementTriP1())
    b0 = Basis(m, elem, intorder=4)
    eps = _make_eps(m, eps_kind, cplx, seed=7)
    modes = compute_modes(b0, eps, 1.55, num_modes=1, order=order)
    md = modes[0]
    b = md.basis
    ed = b.element_dofs.astype(np.int64)
    eta = md.eval_error_estimator()
    ref = fo.eval_error_estimator(m.p, m.t, ed, order, md.E, eps_kind, eps)
    assert eta.shape == (m.nelements,) and np.all(eta >= 0)
    assert np.abs(eta - ref).max() < 1e-10 * ref.max()
    # the indicator concentrates where the field is: the element with the largest value touches the core region
    c = m.centroids()[:, int(np.argmax(eta))]
    assert abs(c[0]) < 0.6 and -0.3 < c[1] < 0.6
    rng = np.random.default_rng(5)
    E = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    rnd = Mode(frequency=1.0, k=1.0, basis_epsilon_r=md.basis_epsilon_r, epsilon_r=eps, basis=b, E=E, H=E)
    ref2 = fo.eval_error_estimator(m.p, m.t, ed, order, E, eps_kind, eps)
    assert np.abs(rnd.eval_error_estimator() - ref2).max() < 1e-11 * ref2.max()


def test_mass_solve_direct_fallback_matches_pcg_and_rescues_elongated_meshes(ctx):
    """calculate_hfield (waveguide.py:657-677) solves the mass system with spsolve.  Here: block PCG, and when it
    stagnates (elongated triangles make the diagonally scaled Nedelec mass matrix ill conditioned) the sparse LU of
    the mass block.  (a) both solvers agree on a well-shaped mesh, (b) the graded rib-benchmark mesh, on which the PCG
    does not converge, still yields an H field that satisfies the oracle's projection."""
    import sys

    from femwell_b200._lib import B200NoConvergence
    from femwell_b200.maxwell.waveguide import compute_modes, get_context, speed_of_light

    m, eps = strip_problem(14)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    a = compute_modes(b0, eps, 1.55, num_modes=2, order=2, solver_options={"mass_solver": "pcg"})
    assert a.stats["timings"]["mass_direct_blocks"] == 0
    d = compute_modes(b0, eps, 1.55, num_modes=2, order=2, solver_options={"mass_solver": "direct"})
    assert d.stats["timings"]["mass_direct_blocks"] == 2
    for x, y in zip(a, d):
        assert np.linalg.norm(x.H - y.H) < 1e-8 * np.linalg.norm(x.H)
        assert abs(y.calculate_power() - 1) < 1e-9
    # (b) elongated, jittered triangles (aspect ratio ~200): Jacobi-type PCG needs thousands of iterations
    from femwell_b200.maxwell.waveguide import calculate_hfield, mu_0

    ms = MeshTri.structured(np.linspace(0, 200, 21), np.linspace(0, 1, 21), jitter=0.2)
    bs = Basis(ms, ElementTriP0(), intorder=4).with_element(mode_element(2))
    rng = np.random.default_rng(2)
    E = rng.standard_normal(bs.N) + 1j * rng.standard_normal(bs.N)
    beta, omega = 7.3, 2 * np.pi * speed_of_light / 1.55
    c = get_context()
    c.set_space(bs)
    c.symbolic()
    c.set_option(2, 1)
    with pytest.raises(B200NoConvergence, match="PCG"):
        calculate_hfield(bs, E, beta, omega)
    c.set_option(2, 0)
    H = calculate_hfield(bs, E, beta, omega)
    assert c.timings()["mass_direct_blocks"] >= 1
    C_ref, M_ref = fo.assemble_hfield_operators(ms.p, ms.t.astype(np.int64), bs.element_dofs.astype(np.int64), 2, bs.X,
                                                bs.W, beta, bs.N)
    rhs = C_ref @ E
    Hs = H / (-1j / mu_0 / omega)
    assert np.linalg.norm(M_ref @ Hs - rhs) < 1e-8 * np.linalg.norm(rhs)  # backward error: the projection is satisfied
    H_ref = fo.calculate_hfield(ms.p, ms.t.astype(np.int64), bs.element_dofs.astype(np.int64), 2, bs.X, bs.W, E, beta,
                                omega, bs.N)
    assert np.linalg.norm(H - H_ref) < 1e-5 * np.linalg.norm(H_ref)  # forward error: limited by cond(M) of this mesh


@pytest.mark.parametrize("order,n,leaf,dirichlet", [(2, 12, 16, False), (1, 17, 8, True), (2, 21, 4, True),
                                                    (2, 9, 32, False), (2, 150, 0, False)])
def test_device_lu_analysis_equals_host_analysis(ctx, order, n, leaf, dirichlet):
    """The nested-dissection analysis runs on the GPU (lu_symbolic_gpu.cu); the threaded host version
    (lu_symbolic.cpp, exercised by the CPU suite) must produce the same tables bit by bit: integer work."""
    from mf_emulator import lu_symbolic

    m, _ = strip_problem(n)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    act = np.ones(b.N, dtype=np.uint8)
    D = b.get_dofs(None) if dirichlet else None
    if dirichlet:
        act[D] = 0
    ctx.set_space(b)
    G = ctx.debug_lu_analysis(D, leaf)
    c = (m.p[:, m.t[0]] + m.p[:, m.t[1]] + m.p[:, m.t[2]]) / 3.0
    S = lu_symbolic(b.N, b.element_dofs, c[0], c[1], act, leaf if leaf else 16)
    for key in ("n_active", "n_nodes", "n_levels", "factor_entries", "max_m", "max_ns"):
        assert G[key] == S[key], key
    for key in ("perm", "iperm", "node_of", "m", "ns", "own_start", "blist", "rel", "blist_off"):
        assert np.array_equal(np.asarray(G[key], dtype=np.int64), np.asarray(S[key], dtype=np.int64)), key


# ----------------------------------------------------------------------------- adaptive refinement loop (refinement.py:58-105)
def test_adaptive_refinement_loop_matches_oracle_and_converges(ctx):
    """solve -> eval_error_estimator -> adaptive_theta -> mesh.refined, four rounds on rectangle case 0 (order 1, PEC
    sides): on every (unstructured) refined mesh n_eff equals the oracle's to 1e-9 and the indicator the oracle's to
    1e-9; the error against the literature value falls."""
    from femwell_b200.maxwell.waveguide import compute_modes
    from femwell_b200.mesh import adaptive_theta

    (ec, ecl), bnd, ref = RECT_CASES[0]
    m = MeshTri.rectangle(6, frozen_x=[0.5], frozen_y=[0.5]).with_subdomains(
        {"core": lambda c: (c[0] < 0.5) & (c[1] < 0.5)})
    errs = []
    for it in range(4):
        b0 = Basis(m, ElementTriP0(), intorder=2)
        eps = np.full(m.nelements, ecl)
        eps[m.subdomains["core"]] = ec
        modes = compute_modes(b0, eps, 1.5, num_modes=1, order=1, metallic_boundaries=bnd)
        md = modes[0]
        fsel = np.unique(np.concatenate([m.boundaries[b] for b in bnd]))
        r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, md.basis.X, md.basis.W, num_modes=1, order=1,
                             dirichlet_facets=fsel, with_h=False)
        assert abs(md.n_eff - r.n_effs[0]) < 1e-9 * abs(r.n_effs[0]), (it, md.n_eff, r.n_effs[0])
        errs.append(abs(md.n_eff.real - ref))
        eta = md.eval_error_estimator()
        eta_ref = fo.eval_error_estimator(m.p, m.t, md.basis.element_dofs.astype(np.int64), 1, md.E, 0, eps)
        assert np.abs(eta - eta_ref).max() < 1e-9 * eta_ref.max()
        sel = adaptive_theta(eta, theta=0.5)
        assert 0 < sel.size < m.nelements
        m = m.refined(sel)
    assert errs[-1] < 0.5 * errs[0], errs


# ----------------------------------------------------------------------------- config 4: coupled-waveguide geometry sweep
def test_coupler_geometry_sweep_matches_oracle(ctx):
    """BASELINE config 4 (crosstalk.py:147-243, coupled_mode_theory.py:41-147) at test size: per geometry its own mesh,
    supermodes + isolated mode, same-mesh overlaps and two cross-mesh overlaps; n_eff 1e-9 and |overlap|^2 1e-6 vs the
    oracle on every geometry."""
    from femwell_b200.mesh import coupled_waveguide_mesh
    from femwell_b200.sweep import COUPLER_COLUMNS, coupler_geometry_sweep

    gaps, widths, n = [0.2, 0.35], [0.44, 0.5], 22
    table = coupler_geometry_sweep(gaps, widths, n=n, order=2)
    assert table.shape == (4, len(COUPLER_COLUMNS) + 1) and (table[:, -1] == 0).all()
    col = {c: i for i, c in enumerate(COUPLER_COLUMNS)}
    X, W = triangle_quadrature(4)
    pts = [(g, w) for g in gaps for w in widths]
    for row, (g, w) in zip(table, pts):
        m = coupled_waveguide_mesh(n, gap=g, w_core_1=w, w_core_2=w + 0.01)
        eps1 = np.full(m.nelements, 1.444**2)
        eps1[m.subdomains["core_1"]] = 3.4777**2
        eps2 = eps1.copy()
        eps2[m.subdomains["core_2"]] = 3.4777**2
        N = m.nvertices + 3 * m.nfacets + 2 * m.nelements
        v0 = np.random.default_rng(0).uniform(-1, 1, N)
        rb = fo.compute_modes(m.p, m.t, 0, eps2, 1.55, X, W, num_modes=2, order=2, v0=v0)
        r1 = fo.compute_modes(m.p, m.t, 0, eps1, 1.55, X, W, num_modes=1, order=2, v0=v0)
        nb = np.sort(rb.n_effs.real)[::-1]
        assert abs(row[col["n_eff_sym"]] - nb[0]) < 1e-9 * nb[0]
        assert abs(row[col["n_eff_asym"]] - nb[1]) < 1e-9 * nb[1]
        assert abs(row[col["n_eff_1"]] - r1.n_effs[0].real) < 1e-9 * nb[0]
        b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
        ed = b.element_dofs.astype(np.int64)
        order_b = np.argsort(-rb.n_effs.real)
        ov = [abs(fo.calculate_overlap(m.p, m.t.astype(np.int64), ed, 2, X, W, r1.E[:, 0], r1.H[:, 0],
                                       rb.E[:, j], rb.H[:, j])) ** 2 for j in order_b]
        # compare fields only where the oracle's own vector is an eigenvector (ARPACK null(M) pollution, see above)
        k0, zd = 2 * np.pi / 1.55, b.split_indices()[1]

        def eig_residual(rr, j):
            x = np.array(rr.E[:, j], dtype=complex)
            x[zd] *= 1j * np.sqrt(rr.lams[j] / k0**4)
            return np.linalg.norm(-(rr.A @ x) + rr.lams[j] * (rr.B @ x)) / np.linalg.norm(rr.A @ x)

        for c, j in zip(("A_into_1", "A_into_2"), order_b):
            if max(eig_residual(rb, j), eig_residual(r1, 0)) < 1e-9:
                assert abs(row[col[c]] - ov[list(order_b).index(j)]) < 1e-6, (c, row[col[c]], ov)
        assert abs(row[col["A_into_1"]] + row[col["A_into_2"]] - 1) < 0.05  # two supermodes carry the isolated mode
        assert row[col["coupling_length_um"]] > 0 and 0 <= row[col["crosstalk_coeff"]] <= 1
    # the first geometry is the cross-mesh reference of the others: overlap of near-identical isolated modes ~ 1
    assert np.isnan(table[0, col["ref_overlap_iso"]])
    assert np.all(table[1:, col["ref_overlap_iso"]] > 0.8) and np.all(table[1:, col["ref_overlap_iso"]] < 1.05)
    # coupling gets weaker (coupling length longer) with the gap at fixed width
    assert table[2, col["coupling_length_um"]] > table[0, col["coupling_length_um"]]


def test_smooth_start_vector_gives_the_same_modes_in_fewer_steps(ctx):
    """Default start vector of compute_modes: a smooth field weighted by the index contrast (csrc/start_vector.cu)
    instead of scipy's uniform(-1, 1) noise.  Same eigenvalues (1e-11), not more operator applications; an element
    count that is not a multiple of the warp size exercises the ragged reductions; repeated calls are bit-identical."""
    from femwell_b200.maxwell.waveguide import compute_modes

    m, eps = strip_problem(25, slab_h=0.11)
    assert m.nelements % 32 != 0
    b0 = Basis(m, ElementTriP0(), intorder=4)
    res = {}
    for sv in ("uniform", "smooth", "smooth"):
        modes = compute_modes(b0, eps, 1.55, num_modes=3, order=2, solver_options={"start_vector": sv})
        res.setdefault(sv, []).append((modes.n_effs.copy(), modes.stats["n_op"], modes[0].E.copy()))
    (nu, opu, _), (ns, ops, Es), (ns2, ops2, Es2) = res["uniform"][0], res["smooth"][0], res["smooth"][1]
    assert np.abs(nu - ns).max() < 1e-11 * np.abs(nu).max()
    assert ops <= opu, (ops, opu)
    assert ops == ops2 and np.array_equal(ns, ns2) and np.array_equal(Es, Es2)
    # complex DG-P1 epsilon goes through the same kernels
    eps_c = _make_eps(m, 1, True, seed=3)
    b1 = Basis(m, ElementDG(ElementTriP1()), intorder=4)
    a = compute_modes(b1, eps_c, 1.55, num_modes=1, order=1, solver_options={"start_vector": "uniform"})
    c = compute_modes(b1, eps_c, 1.55, num_modes=1, order=1)
    assert abs(a.n_effs[0] - c.n_effs[0]) < 1e-11 * abs(a.n_effs[0])


def test_pattern_from_incidence_equals_full_sort_and_falls_back_on_high_valence(ctx, monkeypatch):
    """The structural pattern of the FE space is built from the dof -> element incidence (one warp-level sort per row);
    the full radix sort of all COO keys is the fallback for rows with more than 512 candidates.  Both must give the
    same CSR pattern AND the same values bit for bit (same summation order), and a vertex of valence 40 (560
    candidates at order 2) must take the fallback transparently."""
    k0 = 2 * np.pi / 1.55
    m, eps = strip_problem(18, jitter=0.15)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(2))
    mats = []
    for full in (False, True):
        with monkeypatch.context() as mp:
            if full:
                mp.setenv("FW_SYMBOLIC_FULL_SORT", "1")
            ctx.set_space(b)
            ctx.symbolic()
            ctx.assemble(0, eps, k0)
            mats.append((ctx.matrix(0, eliminate_zeros=False), ctx.matrix(1, eliminate_zeros=False)))
    for (a, c) in zip(mats[0], mats[1]):
        assert np.array_equal(a.indptr, c.indptr) and np.array_equal(a.indices, c.indices)
        assert np.array_equal(a.data, c.data)
    # a fan of 40 triangles around one vertex
    nfan = 40
    ang = np.linspace(0, 2 * np.pi, nfan, endpoint=False)
    p = np.concatenate([[[0.0], [0.0]], np.stack([np.cos(ang), np.sin(ang)]) * (1 + 0.1 * np.cos(3 * ang))], axis=1)
    t = np.stack([np.zeros(nfan, dtype=int), 1 + np.arange(nfan), 1 + (np.arange(nfan) + 1) % nfan])
    fan = MeshTri(p, t)
    bf = Basis(fan, ElementTriP0(), intorder=4).with_element(mode_element(2))
    epsf = np.linspace(1.0, 2.0, nfan)
    ctx.set_space(bf)
    ctx.symbolic()
    ctx.assemble(0, epsf, k0)
    A, B = ctx.matrix(0), ctx.matrix(1)
    Ar, Br = _oracle_AB(fan, bf, 0, epsf, k0)
    for M, R in ((A, Ar), (B, Br)):
        d, _ = csr_rel_diff(M, R)
        assert d < 1e-12
        ok, _n = _pattern_diff_is_roundoff(M, R)
        assert ok


def test_inverse_pivot_blocks_fall_back_to_substitution_when_the_residual_check_fails(ctx, monkeypatch):
    """Selective inversion (lu_tinv.cu) keeps the leaf levels out because their pivot blocks are ill conditioned.  With
    every level forced in, the eigen-solver's residual check must still deliver modes to 1e-9: either the inverses were
    accurate enough on this mesh, or the first retry switches the substitutions back to chains."""
    from femwell_b200.maxwell.waveguide import compute_modes, get_context

    m, eps = strip_problem(60, slab_h=0.11)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    ref = compute_modes(b0, eps, 1.55, num_modes=3, order=2)
    with monkeypatch.context() as mp:
        mp.setenv("FW_TINV_ALL_LEVELS", "1")
        mp.setenv("FW_TINV_MIN_NS", "1")
        get_context()._space_key = None  # new analysis: the level selection is made there
        forced = compute_modes(b0, eps, 1.55, num_modes=3, order=2)
    get_context()._space_key = None
    assert np.abs(forced.n_effs - ref.n_effs).max() < 1e-9 * np.abs(ref.n_effs).max()
    assert forced.stats["max_residual"] <= 1e-10 and ref.stats["max_residual"] <= 1e-10
    print("operator applications: default", ref.stats["n_op"], "forced", forced.stats["n_op"])
