"""CPU suite (no GPU): oracle vs the reference's known answers, host logic, C-ABI surface."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from common import RECT_CASES, mode_element, rectangle_problem, strip_problem
from femwell_b200 import _lib as L
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0, element_tables, triangle_quadrature
from femwell_b200.mesh import MeshTri
from oracle import femwell_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- ABI
def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "femwell_b200.h")).read()
    declared = set(re.findall(r"\b(fw_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.EXPORTS)
    assert lib.fw_version() >= 100


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from femwell_b200.context import Context

    with pytest.raises(L.FemwellB200Error):
        Context(0)


# ----------------------------------------------------------------------------- quadrature / tables
@pytest.mark.parametrize("intorder", [2, 4, 5, 6])
def test_quadrature_exact_for_monomials(intorder):
    import math

    X, W = triangle_quadrature(intorder)
    for a in range(intorder + 1):
        for b in range(intorder + 1 - a):
            exact = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
            assert abs((W * X[0] ** a * X[1] ** b).sum() - exact) < 1e-14


@pytest.mark.parametrize("order", [1, 2])
def test_reference_tables_match_oracle_physical_basis(order):
    """Piola-mapped reference tables == the oracle's physical-gradient basis on a skewed triangle."""
    p = np.array([[0.1, 1.3, 0.4], [0.2, 0.5, 1.7]])
    t = np.array([[0], [1], [2]])
    X, W = triangle_quadrature(4)
    tb = element_tables(order, X, W)
    fns, xq, dx = fo.physical_basis(order, p, t, X)
    e1, e2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    A = np.stack([e1, e2], axis=1)
    AinvT = np.linalg.inv(A).T
    det = np.linalg.det(A)
    for l, fn in enumerate(fns):
        assert fn.kind == tb.kind[l]
        if fn.kind == 0:
            np.testing.assert_allclose(AinvT @ tb.vec[l], fn.t[:, 0, :], atol=1e-13)
            np.testing.assert_allclose(tb.sc[l] / det, fn.curl[0], atol=1e-13)
        else:
            np.testing.assert_allclose(AinvT @ tb.vec[l], fn.grad[:, 0, :], atol=1e-13)
            np.testing.assert_allclose(tb.sc[l], fn.z[0], atol=1e-13)


@pytest.mark.parametrize("order", [1, 2])
def test_dof_numbering_matches_oracle(order):
    m = MeshTri.rectangle(5, jitter=0.1)
    b = Basis(m, mode_element(order))
    facets, t2f = fo.topology(m.p, m.t)
    assert np.array_equal(facets, m.facets) and np.array_equal(t2f, m.t2f)
    ed = fo.element_dofs(order, m.nvertices, m.nfacets, m.t.astype(np.int64), t2f)
    assert np.array_equal(ed, b.element_dofs)
    assert b.N == ed.max() + 1
    tdofs, zdofs = b.split_indices()
    assert len(tdofs) + len(zdofs) == b.N
    D = b.get_dofs(["left"])
    assert np.array_equal(D, fo.boundary_dofs(order, m.nvertices, m.facets, m.boundaries["left"].astype(np.int64)))


# ----------------------------------------------------------------------------- oracle pinned to the reference's known answers
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "rectangle_neff.json")))


@pytest.mark.parametrize("case", range(4))
def test_oracle_reproduces_rectangle_known_answers(case):
    """docs/benchmarks/mode_solver_rectangle.py:45 values; O(h^1.5) convergence on uniform meshes
    (dielectric corner singularity), so the tolerance is mesh-dependent, and the stored oracle value
    at n=20 is a regression pin."""
    m, eps, bnd, fsel, ref = rectangle_problem(20, case)
    X, W = triangle_quadrature(4)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, X, W, num_modes=1, order=2, dirichlet_facets=fsel, with_h=False)
    neff = r.n_effs[0].real
    assert abs(neff - ref) < 5e-5
    assert abs(neff - GOLDEN["order2_n20_intorder4"][case]) < 1e-9
    assert abs(GOLDEN["paper"][case] - ref) == 0


def test_oracle_converges_towards_known_answer():
    errs = []
    X, W = triangle_quadrature(4)
    for n in (10, 20, 40):
        m, eps, bnd, fsel, ref = rectangle_problem(n, 0)
        r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, X, W, num_modes=1, order=2, dirichlet_facets=fsel, with_h=False)
        errs.append(abs(r.n_effs[0].real - ref))
    assert errs[2] < errs[1] < errs[0] and errs[2] < 5e-6


def test_oracle_power_normalisation_and_orthogonality():
    m, eps = strip_problem(10)
    X, W = triangle_quadrature(4)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, X, W, num_modes=2, order=2,
                         v0=np.random.default_rng(0).uniform(-1, 1, Basis(m, mode_element(2)).N))
    b = Basis(m, mode_element(2))
    ed = b.element_dofs.astype(np.int64)
    p00 = fo.calculate_overlap(m.p, m.t, ed, 2, X, W, r.E[:, 0], r.H[:, 0], r.E[:, 0], r.H[:, 0])
    p01 = fo.calculate_overlap(m.p, m.t, ed, 2, X, W, r.E[:, 0], r.H[:, 0], r.E[:, 1], r.H[:, 1])
    # discrete modes are only approximately orthogonal under the unconjugated-curl overlap on a coarse mesh
    assert abs(p00 - 1) < 1e-10 and abs(p01) < 1e-2
    assert 1.444 < r.n_effs[0].real < 3.4777


# ----------------------------------------------------------------------------- host numerics of the library
@pytest.mark.parametrize("n,seed,real", [(1, 0, False), (2, 1, False), (7, 2, True), (21, 3, False), (40, 4, True)])
def test_host_dense_eig_matches_numpy(n, seed, real):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + (0 if real else 1j * rng.standard_normal((n, n)))
    A = np.asfortranarray(A.astype(np.complex128))
    w = np.zeros(n, dtype=np.complex128)
    V = np.zeros((n, n), dtype=np.complex128, order="F")
    L.check(L.load().fw_host_dense_eig(n, L.ptr(A), L.ptr(w), L.ptr(V)))
    wn = np.linalg.eigvals(A)
    for z in w:
        assert np.min(abs(wn - z)) < 1e-10
    for z in wn:
        assert np.min(abs(w - z)) < 1e-10
    for i in range(n):
        v = V[:, i]
        assert np.linalg.norm(A @ v - w[i] * v) < 1e-10 * max(1, np.abs(A).max() * n)
        assert abs(np.linalg.norm(v) - 1) < 1e-12


def test_host_dense_eig_hessenberg_like_arnoldi_matrix():
    # upper Hessenberg with widely spread eigenvalues, as produced by shift-invert Arnoldi
    rng = np.random.default_rng(5)
    n = 20
    H = np.triu(rng.standard_normal((n, n)), -1) * 1e-3
    H[0, 0], H[1, 1] = 50.0, 20.0
    A = np.asfortranarray(H.astype(np.complex128))
    w = np.zeros(n, dtype=np.complex128)
    V = np.zeros((n, n), dtype=np.complex128, order="F")
    L.check(L.load().fw_host_dense_eig(n, L.ptr(A), L.ptr(w), L.ptr(V)))
    wn = np.linalg.eigvals(H)
    big = sorted(w, key=abs)[-2:]
    bign = sorted(wn, key=abs)[-2:]
    assert np.allclose(big, bign, rtol=1e-13)


@pytest.mark.parametrize("order,n,leaf,dirichlet", [(1, 16, 8, False), (2, 12, 16, False), (2, 14, 8, True)])
def test_lu_analysis_drives_a_correct_multifrontal_solve(order, n, leaf, dirichlet):
    """The host analysis (nested dissection, fronts, extend-add maps) + a numpy emulation of the GPU
    numeric phase (windowed pivoting) solves the shifted pencil to machine precision."""
    from mf_emulator import Emulator, lu_symbolic

    m, eps = strip_problem(n)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    k0 = 2 * np.pi / 1.55
    ed = b.element_dofs.astype(np.int64)
    A, B = fo.assemble_waveguide(m.p, m.t.astype(np.int64), ed, order, b.X, b.W,
                                 fo.interpolate_eps(0, eps, m.nelements, b.X), k0, n=b.N)
    sigma = k0**2 * eps.max() * 1.1
    OP = (-A + sigma * B).tocsr()
    act = np.ones(b.N, dtype=np.uint8)
    if dirichlet:
        act[b.get_dofs(None)] = 0
    c = m.centroids()
    S = lu_symbolic(b.N, b.element_dofs, c[0], c[1], act, leaf)
    assert S["n_active"] == int(act.sum())
    assert sorted(S["perm"].tolist()) == np.nonzero(act)[0].tolist()
    assert (S["ns"].sum() == S["n_active"]) and (S["m"] >= S["ns"]).all()
    E = Emulator(S, OP, window=64)
    rhs = np.random.default_rng(0).uniform(-1, 1, b.N) * act
    x = E.solve(rhs)
    I = np.nonzero(act)[0]
    OPc = OP[I][:, I]
    assert np.linalg.norm(OPc @ x[I] - rhs[I]) < 1e-12 * np.linalg.norm(rhs)
    assert np.all(x[act == 0] == 0)
    xs = spla.splu(OPc.tocsc()).solve(rhs[I])
    assert np.linalg.norm(x[I] - xs) < 1e-10 * np.linalg.norm(xs)


# ----------------------------------------------------------------------------- scikit-fem facing bridge (duck-typed)
def test_bridge_extracts_arrays_from_duck_typed_basis():
    from types import SimpleNamespace

    from femwell_b200 import bridge
    from femwell_b200.element import EPS_DG_P1, eps_kind_of

    m, eps = strip_problem(6)

    class ElementDG:  # same class names as scikit-fem's
        def __init__(self, elem):
            self.elem = elem

    class ElementTriP1:
        pass

    sk_mesh = SimpleNamespace(p=m.p, t=m.t, subdomains={"core": m.subdomains["core"]}, boundaries={"left": m.boundaries["left"]})
    mm = bridge.mesh_from_skfem(sk_mesh)
    assert np.array_equal(mm.t, m.t) and np.array_equal(mm.facets, m.facets)
    assert np.array_equal(mm.subdomains["core"], m.subdomains["core"])
    assert eps_kind_of(bridge._eps_element(ElementDG(ElementTriP1()))) == EPS_DG_P1
    # lbasis-style table extraction reproduces the built-in tables when fed the built-in basis
    X, W = triangle_quadrature(4)
    tb = element_tables(2, X, W)

    class FakeElem:
        def lbasis(self, Xq, i):
            f = SimpleNamespace(value=tb.vec[i], curl=tb.sc[i]) if tb.kind[i] == 0 else None
            g = SimpleNamespace(value=tb.sc[i], grad=tb.vec[i]) if tb.kind[i] == 1 else None
            return (f, g)

    kind, vec, sc = bridge.tables_from_lbasis(FakeElem(), X, 2)
    assert np.array_equal(kind, tb.kind) and np.allclose(vec, tb.vec) and np.allclose(sc, tb.sc)


# ----------------------------------------------------------------------------- oracle: analytic pins of the widened rows
def _constant_field_dofs(m, b, c):
    """order-1 dofs of the constant transverse field c (Whitney dofs are tangential line integrals)."""
    x = np.zeros(b.N, dtype=complex)
    lo, hi = m.facets
    x[m.nvertices:] = c[0] * (m.p[0, hi] - m.p[0, lo]) + c[1] * (m.p[1, hi] - m.p[1, lo])
    return x


def test_oracle_cross_mesh_overlap_and_poynting_of_constant_fields():
    """Constant transverse fields are in the N1 space and in P1: the cross-mesh overlap (P1 projection on mesh j,
    point location, quadrature on mesh i) must return the analytic value c x c' * area; E x H is constant too."""
    mi, _ = strip_problem(5, jitter=0.15)
    mj, _ = strip_problem(7, jitter=0.1)
    bi = Basis(mi, ElementTriP0(), intorder=4).with_element(mode_element(1))
    bj = Basis(mj, ElementTriP0(), intorder=4).with_element(mode_element(1))
    cEi, cHi = np.array([1.0 + 0.5j, -2.0j]), np.array([0.3, 1.0 - 1.0j])
    cEj, cHj = np.array([0.7j, 1.5]), np.array([-1.0 + 0.2j, 0.4])
    E_i, H_i = _constant_field_dofs(mi, bi, cEi), _constant_field_dofs(mi, bi, cHi)
    E_j, H_j = _constant_field_dofs(mj, bj, cEj), _constant_field_dofs(mj, bj, cHj)
    ti, tj = mi.t.astype(np.int64), mj.t.astype(np.int64)
    di, dj = bi.element_dofs.astype(np.int64), bj.element_dofs.astype(np.int64)
    U = fo.project_transverse_p1(mj.p, tj, dj, 1, bj.X, bj.W, H_j)
    assert np.abs(U - cHj[:, None]).max() < 1e-12
    e1 = mi.p[:, ti[1]] - mi.p[:, ti[0]]
    e2 = mi.p[:, ti[2]] - mi.p[:, ti[0]]
    area = 0.5 * np.abs(e1[0] * e2[1] - e1[1] * e2[0]).sum()
    cross = lambda a, b: a[0] * b[1] - a[1] * b[0]
    got = fo.overlap_cross(mi.p, ti, di, 1, bi.X, bi.W, E_i, H_i, mj.p, tj, dj, 1, bj.X, bj.W, E_j, H_j, mode=0)
    want = 0.5 * (cross(np.conj(cEi), cHj) + cross(cEj, np.conj(cHi))) * area
    assert abs(got - want) < 1e-12 * abs(want)
    got1 = fo.overlap_cross(mi.p, ti, di, 1, bi.X, bi.W, E_i, None, mj.p, tj, dj, 1, bj.X, bj.W, None, H_j, mode=1)
    assert abs(got1 - cross(np.conj(cEi), cHj) * area) < 1e-12 * abs(want)
    # same mesh: the cross-mesh formula with identical meshes equals the one-mesh Functional for P1-exact fields
    same = fo.calculate_overlap(mi.p, ti, di, 1, bi.X, bi.W, E_i, H_i, E_i, H_i)
    assert abs(same - 0.5 * (cross(np.conj(cEi), cHi) + cross(cEi, np.conj(cHi))) * area) < 1e-12 * abs(same)
    # Poynting vector of constant transverse fields: only P_z = Ex Hy - Ey Hx, constant in every DG dof
    P = fo.poynting(mi.p, ti, di, 1, bi.X, bi.W, E_i, H_i).reshape(-1, 3, 3)
    assert np.abs(P[:, :, 2] - cross(cEi, cHi)).max() < 1e-12 and np.abs(P[:, :, :2]).max() < 1e-12
    # energy density |ex|^2 + |ey|^2 (+ |ez| = 0)
    ed = fo.energy_current_density(mi.p, ti, di, 1, bi.X, bi.W, E_i)
    assert np.abs(ed - (abs(cEi[0]) ** 2 + abs(cEi[1]) ** 2)).max() < 1e-12


# ----------------------------------------------------------------------------- .msh ingest
@pytest.mark.parametrize("name", ["square_v41.msh", "square_v22.msh"])
def test_read_msh_fixture(name):
    from femwell_b200.msh import read_msh

    m = read_msh(os.path.join(ROOT, "tests", "golden", name))
    assert m.nvertices == 4 and m.nelements == 2  # the unused fifth node is pruned
    assert np.array_equal(m.p, np.array([[0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0]]))
    assert np.array_equal(m.t, np.array([[0, 0], [1, 2], [2, 3]]))
    assert {k: v.tolist() for k, v in m.subdomains.items()} == {"core": [0], "clad": [1]}
    fac = {k: [tuple(m.facets[:, f]) for f in v] for k, v in m.boundaries.items()}
    assert fac == {"bottom": [(0, 1)], "left": [(0, 3)]}


def test_msh_round_trip_keeps_mesh_and_groups(tmp_path):
    from femwell_b200.msh import read_msh, write_msh

    m, _ = strip_problem(6, jitter=0.1)
    m = m.with_boundaries({"left": lambda x: x[0] < x[0].min() + 1e-9, "top": lambda x: x[1] > x[1].max() - 1e-9})
    path = str(tmp_path / "strip.msh")
    write_msh(m, path)
    r = read_msh(path)
    assert np.array_equal(r.p, m.p) and np.array_equal(r.t, m.t)
    assert np.array_equal(r.subdomains["core"], m.subdomains["core"])
    for k in ("left", "top"):
        assert np.array_equal(np.sort(r.boundaries[k]), np.sort(m.boundaries[k]))


def test_read_msh_entity_with_two_physical_groups(tmp_path):
    """gmsh 4.1: an entity may carry several physical tags; every one of them becomes a sub-domain."""
    from femwell_b200.msh import read_msh

    src = open(os.path.join(ROOT, "tests", "golden", "square_v41.msh")).read()
    src = src.replace('4\n1 20 "bottom"', '5\n2 12 "everything"\n1 20 "bottom"')
    src = src.replace("1 0 0 0 1 1 0 1 10 3 1 2 3", "1 0 0 0 1 1 0 2 10 12 3 1 2 3")
    src = src.replace("2 0 0 0 1 1 0 1 11 3 3 4 1", "2 0 0 0 1 1 0 2 11 12 3 3 4 1")
    path = tmp_path / "two_groups.msh"
    path.write_text(src)
    m = read_msh(str(path))
    assert m.subdomains["everything"].tolist() == [0, 1]
    assert m.subdomains["core"].tolist() == [0] and m.subdomains["clad"].tolist() == [1]


def test_basis_dof_counts_of_projection_spaces():
    from femwell_b200.element import ElementDG, ElementTriP1, ElementTriP2, ElementVector

    m, _ = strip_problem(4)
    b = Basis(m, ElementTriP0(), intorder=4)
    ne = m.nelements
    assert b.N == ne
    assert b.with_element(ElementDG(ElementTriP1())).N == 3 * ne
    assert b.with_element(ElementDG(ElementTriP2())).N == 6 * ne
    assert b.with_element(ElementVector(ElementDG(ElementTriP2()), 3)).N == 18 * ne
    assert b.with_element(ElementVector(ElementTriP0(), 3)).N == 3 * ne


def test_overlap_dispatch_same_basis_vs_cross_mesh():
    """calculate_overlap takes the one-mesh Functional only for the same mesh, element and quadrature
    (waveguide.py:729-731); anything else goes to the cross-mesh branch."""
    from femwell_b200.maxwell.waveguide import _same_quadrature_space

    m1, _ = strip_problem(4)
    m2, _ = strip_problem(4)
    b1 = Basis(m1, ElementTriP0(), intorder=4).with_element(mode_element(2))
    assert _same_quadrature_space(b1, b1.with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, Basis(m2, ElementTriP0(), intorder=4).with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, Basis(m1, ElementTriP0(), intorder=2).with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, b1.with_element(mode_element(1)))
