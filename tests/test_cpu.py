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


@pytest.mark.parametrize("version,binary", [("2.2", True), ("4.1", False), ("4.1", True)])
def test_msh_binary_and_v41_round_trip(tmp_path, version, binary):
    """gmsh writes binary files by default when Mesh.Binary=1 (2.2: int32 records; 4.1: size_t tags, entity blocks)."""
    from femwell_b200.msh import read_msh, write_msh

    m, _ = strip_problem(7, slab_h=0.11, jitter=0.1)
    m = m.with_boundaries({"left": lambda x: x[0] < x[0].min() + 1e-9, "top": lambda x: x[1] > x[1].max() - 1e-9})
    path = str(tmp_path / "strip.msh")
    write_msh(m, path, version=version, binary=binary)
    raw = open(path, "rb").read()
    assert raw.startswith(b"$MeshFormat\n" + version.encode() + (b" 1 8" if binary else b" 0 8"))
    r = read_msh(path)
    assert np.array_equal(r.p, m.p)
    if version == "2.2":
        assert np.array_equal(r.t, m.t)
        for k in ("core", "slab", "box", "clad"):
            assert np.array_equal(r.subdomains[k], m.subdomains[k])
    else:  # 4.1 groups the elements by entity: same triangles and groups, permuted
        key = lambda mm, idx: sorted(map(tuple, mm.t[:, idx].T.tolist()))  # noqa: E731
        assert key(r, np.arange(r.nelements)) == key(m, np.arange(m.nelements))
        for k in ("core", "slab", "box", "clad"):
            assert key(r, r.subdomains[k]) == key(m, m.subdomains[k])
    for k in ("left", "top"):
        fk = lambda mm: sorted(map(tuple, mm.facets[:, mm.boundaries[k]].T.tolist()))  # noqa: E731
        assert fk(r) == fk(m)


def test_read_msh_big_endian_binary(tmp_path):
    from femwell_b200.msh import read_msh

    nodes = np.empty(3, dtype=np.dtype([("id", ">i4"), ("x", ">f8", 3)]))
    nodes["id"], nodes["x"] = [1, 2, 3], [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    body = (b"$MeshFormat\n2.2 1 8\n" + np.array([1], dtype=">i4").tobytes() + b"\n$EndMeshFormat\n$Nodes\n3\n"
            + nodes.tobytes() + b"\n$EndNodes\n$Elements\n1\n" + np.array([2, 1, 0, 1, 1, 2, 3], dtype=">i4").tobytes()
            + b"\n$EndElements\n")
    path = tmp_path / "be.msh"
    path.write_bytes(body)
    m = read_msh(str(path))
    assert m.nelements == 1 and np.array_equal(m.p, [[0, 1, 0], [0, 0, 1]])


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
    m2, _ = strip_problem(5)
    m3, _ = strip_problem(4)  # an equal copy of m1: same geometry, another object
    b1 = Basis(m1, ElementTriP0(), intorder=4).with_element(mode_element(2))
    assert _same_quadrature_space(b1, b1.with_element(mode_element(2)))
    assert _same_quadrature_space(b1, Basis(m3, ElementTriP0(), intorder=4).with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, Basis(m2, ElementTriP0(), intorder=4).with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, Basis(m1, ElementTriP0(), intorder=2).with_element(mode_element(2)))
    assert not _same_quadrature_space(b1, b1.with_element(mode_element(1)))


# ----------------------------------------------------------------------------- oracle pin: the eight rib known answers
@pytest.mark.parametrize("case", range(8))
def test_oracle_reproduces_rib_known_answers(case):
    """docs/benchmarks/mode_solver.py:52-68 (after Hadley 2002, stated error +-3e-6), judged at 5e-6 by the reference
    (:160).  Synthetic graded tensor mesh instead of gmsh (~6000 triangles), the 3-point rule a default
    Basis(mesh, ElementTriP0()) carries, order 2, lambda = 1.15."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_rib import PAPER, SLABS, rib_problem

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "rib_neff.json")))
    assert g["paper"] == PAPER and g["slab_thickness"] == SLABS
    m, eps = rib_problem(SLABS[case])
    X, W = triangle_quadrature(2)
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.15, X, W, num_modes=1, order=2, with_h=False)
    assert abs(r.n_effs[0].real - PAPER[case]) < 5e-6
    assert abs(r.n_effs[0].real - g["oracle"][case]) < 1e-10  # the committed fixture the GPU test compares with


# ----------------------------------------------------------------------------- bridge: a foreign local basis
def foreign_tables_of(order, X, W):
    """A deliberately different local basis of the SAME finite-element space, as (kind, vec, sc) tables and the
    local change-of-basis matrix T (phi' = T phi): Whitney functions sign-flipped, the second edge function mixed
    with the first, Lagrange functions rescaled per entity type, interior functions mixed.  T is the same for all
    elements and couples only functions of one entity with the same role on every edge, so global conformity holds
    and dof vectors transform as x = T^T x' (entity by entity)."""
    tb = element_tables(order, X, W)
    nb = tb.nbfun
    T = np.zeros((nb, nb))
    if order == 1:
        for l in range(3):
            T[l, l] = 2.0  # P1 vertex functions
            T[3 + l, 3 + l] = -1.0  # Whitney
    else:
        for l in range(3):
            T[l, l] = 0.5
        for f in range(3):
            w, g, pz = 3 + 3 * f, 4 + 3 * f, 5 + 3 * f
            T[w, w] = -1.0
            T[g, g], T[g, w] = 2.0, 0.25
            T[pz, pz] = -3.0
        T[12, 12], T[12, 13] = 1.0, 0.5
        T[13, 13], T[13, 12] = 1.5, -0.3
    vec = np.einsum("ij,jcq->icq", T, tb.vec)
    sc = np.einsum("ij,jq->iq", T, tb.sc)
    return tb.kind, vec, sc, T


class FakeSkfemElement:
    """duck-typed ``ElementComposite``: ``lbasis(X, i)`` -> (Nedelec DiscreteField, Lagrange DiscreteField)"""

    def __init__(self, kind, vec, sc):
        self.kind, self.vec, self.sc = kind, vec, sc

    def lbasis(self, Xq, i):
        from types import SimpleNamespace

        nq = Xq.shape[1]
        if self.kind[i] == 0:
            return (SimpleNamespace(value=self.vec[i], curl=self.sc[i], grad=None),
                    SimpleNamespace(value=np.zeros(nq), grad=np.zeros((2, nq)), curl=None))
        return (SimpleNamespace(value=np.zeros((2, nq)), curl=np.zeros(nq), grad=None),
                SimpleNamespace(value=self.sc[i], grad=self.vec[i], curl=None))


@pytest.mark.parametrize("order", [1, 2])
def test_bridge_reads_a_foreign_local_basis(order):
    from femwell_b200 import bridge

    X, W = triangle_quadrature(4)
    kind, vec, sc, T = foreign_tables_of(order, X, W)
    k2, v2, s2 = bridge.tables_from_lbasis(FakeSkfemElement(kind, vec, sc), X, order)
    assert np.array_equal(k2, kind) and np.array_equal(v2, vec) and np.array_equal(s2, sc)
    # the foreign basis spans the same space: the oracle driven by the tables gives the same eigenvalues
    m, eps = strip_problem(8)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(order))
    ed = b.element_dofs.astype(np.int64)
    k0 = 2 * np.pi / 1.55
    eq = fo.interpolate_eps(0, eps, m.nelements, X)
    A0, B0 = fo.assemble_waveguide(m.p, m.t.astype(np.int64), ed, order, X, W, eq, k0, n=b.N)
    with fo.foreign_tables((kind, vec, sc)):
        A1, B1 = fo.assemble_waveguide(m.p, m.t.astype(np.int64), ed, order, X, W, eq, k0, n=b.N)
    assert abs(A1 - A0).max() > 1e-3 * abs(A0).max()  # genuinely different matrices
    sigma = k0**2 * eps.max() * 1.1
    v0 = np.random.default_rng(0).uniform(-1, 1, b.N)
    l0 = spla.eigs(-A0, M=-B0, k=2, sigma=sigma, v0=v0, return_eigenvectors=False)
    l1 = spla.eigs(-A1, M=-B1, k=2, sigma=sigma, v0=v0, return_eigenvectors=False)
    assert np.abs(np.sort(l0.real) - np.sort(l1.real)).max() < 1e-9 * abs(l0).max()
    # Basis.with_tables carries the tables to the C ABI struct
    bt = b.with_tables(kind, vec, sc, element_dofs=b.element_dofs, n_dofs=b.N)
    assert bt.tables_id is not None and b.tables_id is None and bt.N == b.N
    assert np.array_equal(bt.tables().vec, vec)
    with pytest.raises(ValueError):
        b.with_tables(kind[::-1].copy(), vec, sc)


def test_bridge_refuses_builtin_tables_next_to_a_foreign_basis():
    from types import SimpleNamespace

    from femwell_b200 import bridge

    m, eps = strip_problem(4)

    class ElementTriP0:  # scikit-fem's class name
        pass

    X, W = triangle_quadrature(4)
    b0 = SimpleNamespace(mesh=SimpleNamespace(p=m.p, t=m.t, subdomains={}, boundaries={}), elem=ElementTriP0(), X=X, W=W)
    with pytest.raises(ValueError, match="builtin"):
        bridge.compute_modes_from_skfem(b0, eps, 1.55, order=2, tables="builtin", Mode=object, Modes=object)
    with pytest.raises(ValueError):
        bridge.compute_modes_from_skfem(b0, eps, 1.55, order=2, tables="nope")


# ----------------------------------------------------------------------------- error conventions
def test_error_classes_mirror_scipy():
    from scipy.sparse.linalg import ArpackNoConvergence

    e = L.B200NoConvergence(-3, "ARPACK-style error: no convergence (1/2 eigenvalues converged after 3 restarts)",
                            np.array([1.0 + 0j]), np.ones((5, 1), dtype=complex))
    assert isinstance(e, ArpackNoConvergence) and isinstance(e, L.FemwellB200Error) and isinstance(e, RuntimeError)
    assert e.eigenvalues.shape == (1,) and e.eigenvectors.shape == (5, 1) and e.code == -3
    s = L.B200SingularError(-4, "whatever the library said")
    assert isinstance(s, RuntimeError) and str(s) == "Factor is exactly singular"


def test_rib_benchmark_mesh_hits_every_interface():
    from femwell_b200.mesh import rib_benchmark_mesh

    for slab in (0.0, 0.3):
        m = rib_benchmark_mesh(slab, h0=0.2, ratio=1.3)
        xs, ys = np.unique(m.p[0]), np.unique(m.p[1])
        for x in (-9.0, -1.5, 1.5, 9.0):
            assert np.isclose(xs, x).any()
        for y in [-10.0, 0.0, 1.0, 3.0] + ([slab] if slab else []):
            assert np.isclose(ys, y).any()
        c = m.centroids()
        core = m.subdomains["core"]
        assert np.all(c[1, core] > 0) and np.all(c[1, core] < 1.0)
        assert core.size + m.subdomains["box"].size + m.subdomains["clad"].size == m.nelements


# ----------------------------------------------------------------------------- oracle pin: facet-jump error estimator
def test_oracle_error_estimator_of_a_constant_field_is_analytic():
    """Mode.eval_error_estimator (waveguide.py:473-502): for a constant transverse field c, a globally linear E_z and
    a piecewise-constant epsilon the only jump is [eps] c.n, so tmp[f] = |f|^2 (eps_1 - eps_2)^2 |c.n|^2 on the material
    interfaces and eta_E = sum(0.5 tmp[t2f])."""
    m, eps = strip_problem(6, jitter=0.1)
    b = Basis(m, ElementTriP0(), intorder=4).with_element(mode_element(1))
    c = np.array([1.0 + 0.5j, -2.0j])
    x = _constant_field_dofs(m, b, c)
    x[: m.nvertices] = 0.7 * m.p[0] - 0.2 * m.p[1]
    eta = fo.eval_error_estimator(m.p, m.t, b.element_dofs.astype(np.int64), 1, x, 0, eps)
    lo, hi = m.facets
    d = m.p[:, hi] - m.p[:, lo]
    length = np.hypot(*d)
    n = np.stack([d[1], -d[0]]) / length
    f2t = m.f2t
    inter = f2t[1] >= 0
    assert (~inter).sum() == m.boundary_facets().size
    tmp = np.zeros(m.nfacets)
    tmp[inter] = length[inter] ** 2 * (eps[f2t[0, inter]] - eps[f2t[1, inter]]) ** 2 \
        * np.abs(c[0] * n[0, inter] + c[1] * n[1, inter]) ** 2
    want = np.sum(0.5 * tmp[m.t2f], axis=0)
    assert want.max() > 1 and np.abs(eta - want).max() < 1e-12 * want.max()
    # facet tables handed to the kernel: the Whitney function of an edge has unit tangential moment on it
    s, w, vec = b.facet_tables()
    assert s.size == 2 and abs(w.sum() - 1) < 1e-15 and vec.shape == (6, 3, 2, 2)
    tang = [(1.0, 0.0), (-1.0, 1.0), (0.0, 1.0)]
    for k in range(3):
        for kk in range(3):
            mom = (vec[3 + kk, k, 0] * tang[k][0] + vec[3 + kk, k, 1] * tang[k][1]) @ w
            assert abs(mom - (1.0 if k == kk else 0.0)) < 1e-14


# ----------------------------------------------------------------------------- adaptive refinement (refinement.py:58-105)
def _areas(m):
    a, b, c = m.p[:, m.t[0]], m.p[:, m.t[1]], m.p[:, m.t[2]]
    return 0.5 * np.abs((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]))


def test_adaptive_refinement_is_conforming_and_keeps_groups():
    from femwell_b200.mesh import adaptive_theta

    m = MeshTri.rectangle(4, jitter=0.1).with_subdomains({"low": lambda c: c[1] < 0.5})
    q0 = None
    for _ in range(5):
        sel = np.nonzero(np.hypot(*(m.centroids() - 0.3)) < 0.2)[0]
        r = m.refined(sel)
        assert abs(_areas(r).sum() - 1) < 1e-12
        assert set(np.unique(r._facet_count)) <= {1, 2}  # no hanging node: every facet has one or two elements
        bl = r.boundary_facets()
        assert abs(np.hypot(*(r.p[:, r.facets[0, bl]] - r.p[:, r.facets[1, bl]])).sum() - 4) < 1e-12
        assert abs(_areas(r)[r.subdomains["low"]].sum() - _areas(m)[m.subdomains["low"]].sum()) < 1e-12
        for name in ("left", "right", "top", "bottom"):
            fl = r.boundaries[name]
            assert abs(np.hypot(*(r.p[:, r.facets[0, fl]] - r.p[:, r.facets[1, fl]])).sum() - 1) < 1e-12
        # every selected triangle was split into four; shape regularity does not degrade
        assert r.nelements >= m.nelements + 3 * sel.size
        assert np.array_equal(np.bincount(r.parent_elements)[sel], np.full(sel.size, 4))
        edge = np.max([np.hypot(*(r.p[:, r.t[i]] - r.p[:, r.t[j]])) for i, j in ((0, 1), (1, 2), (0, 2))], axis=0)
        q = (_areas(r) / edge**2).min()
        q0 = q if q0 is None else q0
        assert q > 0.99 * q0
        m = r
    u = MeshTri.rectangle(3).refined()
    assert u.nelements == 4 * 18 and u.nvertices == 16 + 33
    assert MeshTri.rectangle(3).refined(2).nelements == 16 * 18
    est = np.array([0.1, 1.0, 0.6, 0.49])
    assert adaptive_theta(est, theta=0.5).tolist() == [1, 2]
    assert adaptive_theta(est, theta=0.5, max=1.1).tolist() == [1, 2]


def test_oracle_adaptive_loop_converges_to_the_known_answer():
    """refinement.py:58-80 with the oracle in place of compute_modes: solve, estimate, adaptive_theta, refine.  The
    error against the literature value must fall while only part of the mesh is refined."""
    from femwell_b200.mesh import adaptive_theta

    (ec, ecl), bnd, ref = RECT_CASES[0]
    m = MeshTri.rectangle(6, frozen_x=[0.5], frozen_y=[0.5]).with_subdomains(
        {"core": lambda c: (c[0] < 0.5) & (c[1] < 0.5)})
    X, W = triangle_quadrature(2)
    errs, nes = [], []
    for _ in range(4):
        eps = np.full(m.nelements, ecl)
        eps[m.subdomains["core"]] = ec
        fsel = np.unique(np.concatenate([m.boundaries[b] for b in bnd]))
        r = fo.compute_modes(m.p, m.t, 0, eps, 1.5, X, W, num_modes=1, order=1, dirichlet_facets=fsel, with_h=False)
        errs.append(abs(r.n_effs[0].real - ref))
        nes.append(m.nelements)
        b = Basis(m, ElementTriP0(), intorder=2).with_element(mode_element(1))
        eta = fo.eval_error_estimator(m.p, m.t, b.element_dofs.astype(np.int64), 1, r.E[:, 0], 0, eps)
        sel = adaptive_theta(eta, theta=0.5)
        assert 0 < sel.size < m.nelements
        m = m.refined(sel)
    assert errs[-1] < 0.5 * errs[0] and nes[-1] < 16 * nes[0]


def test_coupled_waveguide_mesh_groups():
    from femwell_b200.mesh import coupled_waveguide_mesh

    m = coupled_waveguide_mesh(40, gap=0.3, w_core_1=0.45, w_core_2=0.46)
    a = _areas(m)
    assert abs(a[m.subdomains["core_1"]].sum() - 0.45 * 0.22) < 1e-12
    assert abs(a[m.subdomains["core_2"]].sum() - 0.46 * 0.22) < 1e-12
    assert abs(a[m.subdomains["box"]].sum() - 4.0) < 1e-12
    assert sum(v.size for v in m.subdomains.values()) == m.nelements == 3200
    c = m.centroids()
    assert c[0, m.subdomains["core_1"]].max() < -0.15 and c[0, m.subdomains["core_2"]].min() > 0.15


# ----------------------------------------------------------------------------- bench.py host logic (CPU ladder)
def test_bench_ladder_extrapolation_is_anchored_and_never_sublinear():
    """The CPU reference arm measures a ladder of small meshes and extrapolates the largest sample by the fitted power
    law: exact on a pure power law, never below linear scaling, anchored at the largest sample."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ne = [2592, 5000, 10082, 20000]
    t = [3e-4 * n**1.4 for n in ne]
    p, c = bench.fit_power_law(ne, t)
    assert abs(p - 1.4) < 1e-9 and abs(c - 3e-4) < 1e-9
    t_full, p2, t_lin = bench.extrapolate_ladder(list(zip(ne, t)), 500000)
    assert abs(p2 - 1.4) < 1e-9 and abs(t_full - 3e-4 * 500000**1.4) < 1e-6 * t_full
    assert abs(t_lin - t[-1] * 500000 / ne[-1]) < 1e-9 * t_lin and t_full > t_lin
    # a sub-linear fit (noise on tiny samples) is clamped to linear scaling of the largest sample
    t_full, p3, t_lin = bench.extrapolate_ladder([(1000, 1.0), (2000, 1.5)], 100000)
    assert p3 < 1 and abs(t_full - t_lin) < 1e-12 * t_lin
    assert "500000 triangles" in bench.workload_text(500, 4)
    assert bench.algorithmic_bytes_element_kernel(10, 8, 14, 8, 8, 8) == 10 * 20 + 8 * 16 + 10 * (196 + 64) * 8


def test_ncu_summary_converter_and_roofline_traffic_reader(tmp_path):
    """bench.py takes `roofline.traffic` from a committed per-launch ncu summary; scripts/ncu_raw_to_summary.py writes
    that format from an `ncu --page raw --csv` export.  Round trip on a synthetic export, then the committed capture of
    one substitution pair: its DRAM traffic cannot be below the algorithmic bytes the bench divides by."""
    import importlib.util
    import subprocess
    import sys

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    raw = tmp_path / "raw.csv"
    raw.write_text(
        '==PROF== Connected to process 1\n'
        '"ID","Process ID","Kernel Name","Block Size","Grid Size","dram__bytes_read.sum","dram__bytes_write.sum",'
        '"gpu__time_duration.sum","launch__registers_per_thread"\n'
        '"","","","","","byte","byte","ns","register/thread"\n'
        '"0","7","void fw::k_fwd_update<double, 1, 1>(fw::FrontTab, const int *)","(128, 1, 1)","(16, 2, 1)","1,000","24","5,000","62"\n'
        '"1","7","void fw::k_bwd_update<double, 1, 4>(fw::FrontTab, const int *)","(512, 1, 1)","(8, 1, 2)","3000","0","7000","62"\n')
    out = tmp_path / "summary.csv"
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_raw_to_summary.py"), str(raw), str(out), "synthetic"],
                   check=True, capture_output=True)
    assert out.read_text().startswith("# synthetic\n")
    assert bench.ncu_traffic("summary.csv", "k_", total=True, profiles_dir=str(tmp_path)) == 4024.0
    assert bench.ncu_traffic("summary.csv", "k_fwd_update", profiles_dir=str(tmp_path)) == 1024.0
    assert bench.ncu_traffic("summary.csv", "no_such_kernel", profiles_dir=str(tmp_path)) is None
    assert bench.ncu_traffic("missing.csv", "k_", profiles_dir=str(tmp_path)) is None
    pair = bench.ncu_traffic("ncu_r02_final_solve_pair.csv", "k_", total=True)
    assert pair is not None and 9.95e9 < pair < 2.0e10  # 9.95 GB of L and U entries per pair at config 2
