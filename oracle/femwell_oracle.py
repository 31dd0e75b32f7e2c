"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

A numpy/scipy restatement of the reference hot path
``femwell.maxwell.waveguide.compute_modes`` + ``Mode`` post-processing
(/root/reference/femwell/maxwell/waveguide.py:528-677, 699-736, 112-258) including the
scikit-fem assembly semantics it relies on (scikit-fem >= 8.1.0 is
the pinned third-party dependency, pyproject.toml:26, NOT vendored and NOT installable
in this image).  The solve half uses the reference's real code:
``scipy.sparse.linalg.eigs`` (ARPACK + SuperLU) and ``scipy.sparse.linalg.spsolve``.

PARITY PINNING.  The reference holds no golden matrices / fields for this path; the only
known answers are the literature n_eff values of docs/benchmarks/mode_solver_rectangle.py:45
and docs/benchmarks/mode_solver.py:58-68.  ``tests/test_oracle.py`` pins this oracle to the
four analytic rectangle values.  Matrix entries, eigenvectors, H, overlaps are
"parity unpinned" against scikit-fem itself (it cannot be imported here) -- they are
GPU-vs-this-restatement on identical meshes and identical basis tables.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.

Deliberately written differently from the CUDA path: basis functions are evaluated from
*physical* barycentric gradients per element (no reference tables, no Piola/metric
factorisation), forms are evaluated pair by pair exactly like scikit-fem's
``BilinearForm._assemble`` (python loop over (trial j, test i), vectorised over (Ne, nqp)).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SPEED_OF_LIGHT = 299792458.0
MU_0 = 1.25663706212e-06  # scipy.constants.mu_0 (CODATA 2018), waveguide.py:675
EPSILON_0 = 8.8541878128e-12

_EDGES = [(0, 1), (1, 2), (0, 2)]


# ----------------------------------------------------------------------------- topology
def topology(p, t):
    """facets (2,Nf), t2f (3,Ne) -- SURVEY.md A.1 (unique sorted vertex pairs, lexicographic)."""
    t = np.sort(np.asarray(t, dtype=np.int64), axis=0)
    pairs = np.concatenate([t[[0, 1]], t[[1, 2]], t[[0, 2]]], axis=1).T  # (3Ne, 2)
    uniq, inv = np.unique(pairs, axis=0, return_inverse=True)
    return uniq.T.copy(), inv.reshape(3, -1)


def element_dofs(order, nv, nf, t, t2f):
    """nodal, facet, interior numbering of N1xP1 / N2xP2."""
    ne = t.shape[1]
    if order == 1:
        return np.concatenate([t, nv + t2f], axis=0)
    rows = [t]
    for k in range(3):
        for c in range(3):
            rows.append((nv + 3 * t2f[k] + c)[None, :])
    rows.append((nv + 3 * nf + 2 * np.arange(ne))[None, :])
    rows.append((nv + 3 * nf + 2 * np.arange(ne) + 1)[None, :])
    return np.concatenate(rows, axis=0)


def local_kinds(order):
    """0 = Nedelec (transverse), 1 = Lagrange (z); entity-major local ordering."""
    if order == 1:
        return np.array([1, 1, 1, 0, 0, 0])
    return np.array([1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0])


# ----------------------------------------------------------------------------- basis
@dataclass
class LocalBasis:
    """Global (physical) basis function values at the quadrature points of every element."""

    kind: int
    t: Optional[np.ndarray] = None  # (2, Ne, nq) transverse vector   (Nedelec)
    curl: Optional[np.ndarray] = None  # (Ne, nq)
    z: Optional[np.ndarray] = None  # (Ne, nq)                        (Lagrange)
    grad: Optional[np.ndarray] = None  # (2, Ne, nq)


def physical_basis_from_tables(tables, p, t, X):
    """The same interface for a FOREIGN local basis given by reference tables ``(kind, vec, sc)`` at the points X
    (scikit-fem ``elem.lbasis``): covariant Piola map for the H(curl) functions
    (phi = A^-T phi_hat, curl phi = curl_hat / det A) and the gradient map for the H1 functions."""
    kind, vec, sc = tables
    p = np.asarray(p, dtype=np.float64)
    t = np.asarray(t)
    ne, nq = t.shape[1], X.shape[1]
    v = [p[:, t[a]] for a in range(3)]
    e1, e2 = v[1] - v[0], v[2] - v[0]
    det = e1[0] * e2[1] - e1[1] * e2[0]
    # A = [e1 e2]; A^-T = 1/det [[e2y, -e1y], [-e2x, e1x]]
    def piola(vh):  # (2, nq) -> (2, Ne, nq)
        return np.stack([(e2[1][:, None] * vh[0][None, :] - e1[1][:, None] * vh[1][None, :]) / det[:, None],
                         (-e2[0][:, None] * vh[0][None, :] + e1[0][:, None] * vh[1][None, :]) / det[:, None]])

    out = []
    for l in range(len(kind)):
        if kind[l] == 0:
            out.append(LocalBasis(0, t=piola(vec[l]), curl=sc[l][None, :] / det[:, None]))
        else:
            out.append(LocalBasis(1, z=np.broadcast_to(sc[l], (ne, nq)).copy(), grad=piola(vec[l])))
    ones = np.ones(nq)
    xq = v[0][:, :, None] + e1[:, :, None] * X[0][None, None, :] + e2[:, :, None] * X[1][None, None, :]
    dx = np.abs(det)[:, None] * ones[None, :]
    return out, xq, dx


# a foreign local basis for all assembly / interpolation routines below (tests only; None: Appendix-D basis)
_TABLES = None


class foreign_tables:
    """``with foreign_tables((kind, vec, sc)): ...`` -- every routine of this module then evaluates the local
    basis from the given reference tables instead of the built-in Appendix-D formulas."""

    def __init__(self, tables):
        self.tables = tables

    def __enter__(self):
        global _TABLES
        self.prev, _TABLES = _TABLES, self.tables

    def __exit__(self, *exc):
        global _TABLES
        _TABLES = self.prev


def physical_basis(order, p, t, X):
    """All local basis functions (composite entity-major order) at all points.
    Basis of SURVEY.md Appendix D written with physical gradients of the barycentrics."""
    if _TABLES is not None:
        return physical_basis_from_tables(_TABLES, p, t, X)
    p = np.asarray(p, dtype=np.float64)
    t = np.asarray(t)
    ne, nq = t.shape[1], X.shape[1]
    v = [p[:, t[a]] for a in range(3)]  # each (2, Ne)
    e1, e2 = v[1] - v[0], v[2] - v[0]
    det = e1[0] * e2[1] - e1[1] * e2[0]
    # physical gradients of barycentric coordinates: grad lam_1 = (e2_y, -e2_x)/det etc.
    g1 = np.stack([e2[1], -e2[0]]) / det
    g2 = np.stack([-e1[1], e1[0]]) / det
    g0 = -(g1 + g2)
    G = [g0, g1, g2]  # (2, Ne) each
    lam = np.stack([1 - X[0] - X[1], X[0], X[1]])  # (3, nq)
    ones = np.ones(nq)

    def cross(a, b):
        return a[0] * b[1] - a[1] * b[0]

    def whitney(a, b):
        val = G[b][:, :, None] * lam[a][None, None, :] - G[a][:, :, None] * lam[b][None, None, :]
        curl = (2 * cross(G[a], G[b]))[:, None] * ones[None, :]
        return val, curl

    ned = []
    for a, b in _EDGES:
        w, c = whitney(a, b)
        ned.append(LocalBasis(0, t=w, curl=c))
        if order == 2:
            g = G[b][:, :, None] * lam[a][None, None, :] + G[a][:, :, None] * lam[b][None, None, :]
            ned.append(LocalBasis(0, t=g, curl=np.zeros((ne, nq))))
    if order == 2:
        for c_, (a, b) in ((2, (0, 1)), (1, (0, 2))):
            w, cw = whitney(a, b)
            val = w * lam[c_][None, None, :]
            curl = cross(G[c_][:, :, None], w) + lam[c_][None, :] * cw
            ned.append(LocalBasis(0, t=val, curl=curl))
    lag = []
    if order == 1:
        for a in range(3):
            lag.append(LocalBasis(1, z=np.broadcast_to(lam[a], (ne, nq)).copy(), grad=G[a][:, :, None] * ones))
    else:
        for a in range(3):
            lag.append(
                LocalBasis(
                    1,
                    z=np.broadcast_to(lam[a] * (2 * lam[a] - 1), (ne, nq)).copy(),
                    grad=G[a][:, :, None] * (4 * lam[a] - 1)[None, None, :],
                )
            )
        for a, b in _EDGES:
            lag.append(
                LocalBasis(
                    1,
                    z=np.broadcast_to(4 * lam[a] * lam[b], (ne, nq)).copy(),
                    grad=4 * (G[b][:, :, None] * lam[a][None, None, :] + G[a][:, :, None] * lam[b][None, None, :]),
                )
            )
    # composite entity-major order
    if order == 1:
        out = lag + ned
    else:
        out = lag[:3]
        for f in range(3):
            out += [ned[2 * f], ned[2 * f + 1], lag[3 + f]]
        out += [ned[6], ned[7]]
    xq = (
        v[0][:, :, None]
        + e1[:, :, None] * X[0][None, None, :]
        + e2[:, :, None] * X[1][None, None, :]
    )  # (2, Ne, nq)
    dx = np.abs(det)[:, None] * ones[None, :]  # multiplied by W by the caller
    return out, xq, dx


def interpolate_eps(eps_kind, epsilon_r, ne, X):
    """``basis_epsilon_r.interpolate(epsilon_r)`` (waveguide.py:602): (Ne,nq) or (3,Ne,nq)."""
    nq = X.shape[1]
    if eps_kind == 0:
        return np.repeat(np.asarray(epsilon_r)[:, None], nq, axis=1)
    e3 = np.asarray(epsilon_r).reshape(ne, 3)
    if eps_kind == 1:
        lam = np.stack([1 - X[0] - X[1], X[0], X[1]])
        return e3 @ lam
    return np.repeat(e3.T[:, :, None], nq, axis=2)


# ----------------------------------------------------------------------------- assembly
def _assemble(form, basis_fns, dofs, dxw, n, dtype, eliminate_zeros=True):
    """scikit-fem BilinearForm._assemble restated: data[j,i,:] =
    sum_q form(u_j, v_i) dx ; rows = test dofs, cols = trial dofs ; eliminate_zeros ; tocsr."""
    rows, cols, data = [], [], []
    for j, u in enumerate(basis_fns):
        for i, v in enumerate(basis_fns):
            val = form(u, v)
            if val is None:
                continue
            rows.append(dofs[i])
            cols.append(dofs[j])
            data.append(np.sum(val * dxw, axis=1).astype(dtype))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    data = np.concatenate(data)
    m = sp.coo_matrix((data, (rows, cols)), shape=(n, n))
    if eliminate_zeros:
        m.eliminate_zeros()
    return m.tocsr()


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1]


def assemble_waveguide(p, t, dofs, order, X, W, eps_q, k0, mu_r=1.0, radius=np.inf, n=None, subset=None,
                       eliminate_zeros=True):
    """A, B of waveguide.py:577-603 as canonical CSR."""
    if subset is not None:
        t = t[:, subset]
        dofs = dofs[:, subset]
    fns, xq, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    f = 1 + xq[0] / radius
    if eps_q.ndim == 3:
        eps_t = eps_q[:2] * f
        eps_z = eps_q[2] / f
    else:
        eps_t = eps_q * f
        eps_z = eps_q / f
    mu_t = mu_r * f
    mu_z = mu_r / f
    dtype = eps_q.dtype

    def aform(u, v):
        out = None

        def add(o, term):
            return term if o is None else o + term

        if u.kind == 0 and v.kind == 0:
            out = add(out, 1 / mu_z * u.curl * v.curl / k0**2)
            out = add(out, -_dot(eps_t * u.t, v.t))
        if u.kind == 1 and v.kind == 0:
            out = add(out, _dot(1 / mu_t * u.grad, v.t))
        if u.kind == 0 and v.kind == 1:
            out = add(out, _dot(eps_t * u.t, v.grad))
        if u.kind == 1 and v.kind == 1:
            out = add(out, -eps_z * u.z * v.z * k0**2)
        return out

    def bform(u, v):
        if u.kind == 0 and v.kind == 0:
            return -_dot(1 / mu_t * u.t, v.t) / k0**2
        return None

    if n is None:
        n = int(dofs.max()) + 1
    A = _assemble(aform, fns, dofs, dxw, n, dtype, eliminate_zeros)
    B = _assemble(bform, fns, dofs, dxw, n, dtype, eliminate_zeros)
    return A, B


def assemble_hfield_operators(p, t, dofs, order, X, W, beta, n):
    """C(beta) and M_full of calculate_hfield (waveguide.py:657-668), in complex128
    (the reference declares complex64; SURVEY.md R7: compute in FP64)."""
    fns, xq, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]

    def cform(u, v):
        if u.kind == 0 and v.kind == 0:
            return (-1j * beta * u.t[1]) * v.t[0] + (1j * beta * u.t[0]) * v.t[1]
        if u.kind == 1 and v.kind == 0:
            return u.grad[1] * v.t[0] - u.grad[0] * v.t[1] + 0j
        if u.kind == 0 and v.kind == 1:
            return u.curl * v.z + 0j
        return None

    def mform(u, v):
        if u.kind == 0 and v.kind == 0:
            return _dot(u.t, v.t) + 0j
        if u.kind == 1 and v.kind == 1:
            return u.z * v.z + 0j
        return None

    C = _assemble(cform, fns, dofs, dxw, n, np.complex128)
    M = _assemble(mform, fns, dofs, dxw, n, np.complex128)
    return C, M


def calculate_hfield(p, t, dofs, order, X, W, xs, beta, omega, n):
    """waveguide.py:657-677."""
    C, M = assemble_hfield_operators(p, t, dofs, order, X, W, beta, n)
    return spla.spsolve(M.tocsc(), C @ xs.astype(complex)) * -1j / MU_0 / omega


# ----------------------------------------------------------------------------- functionals
def interpolate_mode(fns, dofs, x):
    """((Ex,Ey), Ez) at quadrature points, (Ne,nq) each."""
    ne, nq = fns[0].z.shape if fns[0].kind == 1 else fns[0].curl.shape
    et = np.zeros((2, ne, nq), dtype=complex)
    ez = np.zeros((ne, nq), dtype=complex)
    for l, fn in enumerate(fns):
        c = x[dofs[l]][:, None]
        if fn.kind == 0:
            et += fn.t * c[None]
        else:
            ez += fn.z * c
    return et, ez


def calculate_overlap(p, t, dofs, order, X, W, E_i, H_i, E_j, H_j, subset=None):
    """Same-basis branch of waveguide.py:699-736."""
    if subset is not None:
        t, dofs = t[:, subset], dofs[:, subset]
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    ei, _ = interpolate_mode(fns, dofs, E_i)
    hi, _ = interpolate_mode(fns, dofs, H_i)
    ej, _ = interpolate_mode(fns, dofs, E_j)
    hj, _ = interpolate_mode(fns, dofs, H_j)

    def cross(a, b):
        return a[0] * b[1] - a[1] * b[0]

    return 0.5 * np.sum((cross(np.conj(ei), hj) + cross(ej, np.conj(hi))) * dxw)


def te_fraction(p, t, dofs, order, X, W, E):
    """waveguide.py:112-127."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, _ = interpolate_mode(fns, dofs, E)
    ex = np.sum(np.abs(et[0]) ** 2 * dxw)
    ey = np.sum(np.abs(et[1]) ** 2 * dxw)
    return ex / (ex + ey)


def transversality(p, t, dofs, order, X, W, E):
    """waveguide.py:135-159."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, ez = interpolate_mode(fns, dofs, E)
    ex = np.sum(np.abs(et[0]) ** 2 * dxw)
    ey = np.sum(np.abs(et[1]) ** 2 * dxw)
    ea = np.sum(np.abs(et[0] * np.conj(et[0]) + et[1] * np.conj(et[1]) + ez * np.conj(ez)) * dxw)
    return (ex + ey) / ea


def effective_area(p, t, dofs, order, X, W, E, field="xy"):
    """waveguide.py:181-213."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, _ = interpolate_mode(fns, dofs, E)
    if field == "xy":
        e2 = np.conj(et[0]) * et[0] + np.conj(et[1]) * et[1]
    else:
        num = 0 if field == "x" else 1
        e2 = np.conj(et[num]) * et[num]
    return np.real(np.sum(e2 * dxw) ** 2 / np.sum(e2**2 * dxw))


def coupling_coefficient(p, t, dofs, order, X, W, E_i, E_j, delta_eps_q):
    """waveguide.py:167-179."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    eit, eiz = interpolate_mode(fns, dofs, E_i)
    ejt, ejz = interpolate_mode(fns, dofs, E_j)
    return np.sum(delta_eps_q * (np.conj(eit[0]) * ejt[0] + np.conj(eit[1]) * ejt[1] + np.conj(eiz) * ejz) * dxw)


def confinement_factor(p, t, dofs, order, X, W, E, eps_q, subset):
    """waveguide.py:225-249 (eps_q already restricted to ``subset``)."""
    t, dofs = t[:, subset], dofs[:, subset]
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, ez = interpolate_mode(fns, dofs, E)
    if eps_q.ndim == 3:
        eps_t, eps_z = eps_q[:2], eps_q[2]
        val = (
            np.sqrt(eps_t[0]) * np.conj(et[0]) * et[0]
            + np.sqrt(eps_t[1]) * np.conj(et[1]) * et[1]
            + np.sqrt(eps_z) * np.conj(ez) * ez
        )
    else:
        val = np.sqrt(eps_q) * (np.conj(et[0]) * et[0] + np.conj(et[1]) * et[1]) + np.sqrt(eps_q) * np.conj(ez) * ez
    return SPEED_OF_LIGHT * EPSILON_0 * np.sum(val * dxw)


def calculate_intensity(p, t, dofs, order, X, W, E, H):
    """waveguide.py:260-280: 0.5 Re(Ex Hy* - Ey Hx*) at the quadrature points projected onto DG-P1
    (basis2.project: per-element mass solve with the basis quadrature)."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, _ = interpolate_mode(fns, dofs, E)
    ht, _ = interpolate_mode(fns, dofs, np.conj(H))
    inten = 0.5 * np.real(et[0] * ht[1] - et[1] * ht[0])  # (Ne, nq)
    lam = np.stack([1 - X[0] - X[1], X[0], X[1]])  # (3, nq)
    M = np.einsum("eq,aq,bq->eab", dxw, lam, lam)
    b = np.einsum("eq,aq,eq->ea", dxw, lam, inten)
    return np.linalg.solve(M, b[:, :, None])[:, :, 0].reshape(-1)


def poynting(p, t, dofs, order, X, W, E, H):
    """waveguide.py:74-95: P = E x H (no conjugation) at the quadrature points, L2-projected component-wise onto
    ElementVector(ElementDG(Lagrange_order), 3) with the basis quadrature; dof = e*(3*nl) + a*3 + c
    (scikit-fem ElementVector interleaves the components, ElementDG keeps the local Lagrange order)."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    (ex, ey), ez = interpolate_mode(fns, dofs, E)
    (hx, hy), hz = interpolate_mode(fns, dofs, H)
    P = np.stack([ey * hz - ez * hy, ez * hx - ex * hz, ex * hy - ey * hx])  # (3, Ne, nq)
    lag = np.stack([f.z for f in fns if f.kind == 1])  # (nl, Ne, nq)
    M = np.einsum("eq,aeq,beq->eab", dxw, lag, lag)
    b = np.einsum("eq,aeq,ceq->eac", dxw, lag, P)
    u = np.linalg.solve(M.astype(complex), b)  # (Ne, nl, 3)
    return u.reshape(-1)


def project_transverse_p1(p, t, dofs, order, X, W, x):
    """basis.with_element(ElementVector(ElementTriP1())).project(E_t) (waveguide.py:739-746): global L2
    projection of the transverse field onto continuous P1, component-wise.  Returns (2, Nv) complex."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    et, _ = interpolate_mode(fns, dofs, x)
    lam = np.stack([1 - X[0] - X[1], X[0], X[1]])
    nv = p.shape[1]
    Ml = np.einsum("eq,aq,bq->abe", dxw, lam, lam)
    rows = np.repeat(t[:, None, :], 3, axis=1).ravel()
    cols = np.repeat(t[None, :, :], 3, axis=0).ravel()
    M = sp.coo_matrix((Ml.ravel(), (rows, cols)), shape=(nv, nv)).tocsc()
    lu = spla.splu(M)
    out = np.zeros((2, nv), dtype=complex)
    for c in range(2):
        b = np.zeros(nv, dtype=complex)
        bl = np.einsum("eq,aq,eq->ae", dxw, lam, et[c])
        np.add.at(b, t.ravel(), bl.ravel())
        out[c] = lu.solve(b.real) + 1j * lu.solve(b.imag)
    return out


def locate_points(p, t, pts):
    """index of a triangle containing every point (smallest index), barycentric coordinates (brute force)."""
    v0 = p[:, t[0]]
    e1, e2 = p[:, t[1]] - v0, p[:, t[2]] - v0
    det = e1[0] * e2[1] - e1[1] * e2[0]
    tri = np.full(pts.shape[1], -1)
    l1o, l2o = np.zeros(pts.shape[1]), np.zeros(pts.shape[1])
    for i in range(pts.shape[1]):
        dxp, dyp = pts[0, i] - v0[0], pts[1, i] - v0[1]
        l1 = (dxp * e2[1] - dyp * e2[0]) / det
        l2 = (dyp * e1[0] - dxp * e1[1]) / det
        ok = np.nonzero((l1 >= -1e-10) & (l2 >= -1e-10) & (l1 + l2 <= 1 + 1e-10))[0]
        if ok.size == 0:
            raise ValueError("Point is outside of the mesh.")
        tri[i], l1o[i], l2o[i] = ok[0], l1[ok[0]], l2[ok[0]]
    return tri, l1o, l2o


def overlap_cross(pi, ti, dofs_i, order_i, Xi, Wi, E_i, H_i, pj, tj, dofs_j, order_j, Xj, Wj, E_j, H_j, mode=0):
    """Different-mesh branch of calculate_overlap (waveguide.py:737-759; mode 0) and of
    calculate_scalar_product (:762-782; mode 1)."""
    fns, xq, dx = physical_basis(order_i, pi, ti, Xi)
    dxw = dx * Wi[None, :]
    ei, _ = interpolate_mode(fns, dofs_i, E_i)
    Hp = project_transverse_p1(pj, tj, dofs_j, order_j, Xj, Wj, H_j)
    pts = xq.reshape(2, -1)
    tri, l1, l2 = locate_points(pj, tj, pts)
    l0 = 1 - l1 - l2

    def at_points(U):
        v = U[:, tj[0, tri]] * l0 + U[:, tj[1, tri]] * l1 + U[:, tj[2, tri]] * l2
        return v.reshape(2, *xq.shape[1:])

    hj = at_points(Hp)
    val = np.conj(ei[0]) * hj[1] - np.conj(ei[1]) * hj[0]
    if mode == 0:
        hi, _ = interpolate_mode(fns, dofs_i, H_i)
        ej = at_points(project_transverse_p1(pj, tj, dofs_j, order_j, Xj, Wj, E_j))
        val = 0.5 * (val + ej[0] * np.conj(hi[1]) - ej[1] * np.conj(hi[0]))
    return np.sum(val * dxw)


def energy_current_density(p, t, dofs, order, X, W, xs):
    """waveguide.py:680-696: P0 mass solve of the linear form |e_x|^2 + |e_y|^2 + |e_z| (e_z not squared)."""
    fns, _, dx = physical_basis(order, p, t, X)
    dxw = dx * W[None, :]
    (ex, ey), ez = interpolate_mode(fns, dofs, xs)
    a = np.sum(dxw * (np.abs(ex) ** 2 + np.abs(ey) ** 2 + np.abs(ez)), axis=1).astype(complex)
    b = np.sum(dxw, axis=1)
    return a / b


def eval_error_estimator(p, t, dofs, order, E, eps_kind, epsilon_r):
    """Mode.eval_error_estimator (waveguide.py:473-502) restated: two InteriorFacetBasis objects (sides 0 / 1) on the
    Gauss-Legendre rule of degree 2 * maxdeg, ``w.h`` = facet length, ``edge_jump.elemental`` per interior facet,
    ``eta_E = sum(0.5 * tmp[t2f], axis=0)``.  Evaluated element by element with the physical basis at the facet
    points (no reference tables)."""
    t = np.asarray(t, dtype=np.int64)
    ne = t.shape[1]
    facets, t2f = topology(p, t)
    nf = facets.shape[1]
    npts = int(np.ceil((2 * order + 1) / 2))
    xg, wg = np.polynomial.legendre.leggauss(npts)
    sg, wg = 0.5 * (xg + 1), 0.5 * wg
    a, b = p[:, facets[0]], p[:, facets[1]]
    length = np.hypot(*(b - a))
    nrm = np.stack([(b - a)[1], -(b - a)[0]]) / length  # (2, nf)
    # traces of (grad E_z . n, eps E_t . n) from both neighbours: lists per facet
    gz = [[None] * nf, [None] * nf]
    dn = [[None] * nf, [None] * nf]
    seen = np.zeros(nf, dtype=int)
    ends = [((0.0, 0.0), (1.0, 0.0)), ((1.0, 0.0), (0.0, 1.0)), ((0.0, 0.0), (0.0, 1.0))]
    e3 = None if eps_kind == 0 else np.asarray(epsilon_r).reshape(ne, 3)
    for k in range(3):
        (xa, ya), (xb, yb) = ends[k]
        Xk = np.stack([xa + sg * (xb - xa), ya + sg * (yb - ya)])
        fns, _, _ = physical_basis(order, p, t, Xk)
        et, _ = interpolate_mode(fns, dofs, E)  # (2, Ne, nq)
        gzv = np.zeros((2, ne, npts), dtype=complex)
        for l, fn in enumerate(fns):
            if fn.kind == 1:
                gzv += fn.grad * E[dofs[l]][None, :, None]
        if eps_kind == 0:
            ev = np.asarray(epsilon_r)[:, None] * np.ones(npts)[None, :]
        else:
            lam = np.stack([1 - Xk[0] - Xk[1], Xk[0], Xk[1]])
            ev = e3 @ lam
        for e in range(ne):
            f = t2f[k, e]
            side = seen[f]
            seen[f] += 1
            n = nrm[:, f]
            gz[side][f] = gzv[0, e] * n[0] + gzv[1, e] * n[1]
            dn[side][f] = (et[0, e] * n[0] + et[1, e] * n[1]) * ev[e]
    tmp = np.zeros(nf)
    for f in range(nf):
        if seen[f] == 2:
            val = np.abs(gz[0][f] - gz[1][f]) ** 2 + np.abs(dn[0][f] - dn[1][f]) ** 2
            tmp[f] = np.sum(wg * length[f] * length[f] * val)
    return np.sum(0.5 * tmp[t2f], axis=0)


# ----------------------------------------------------------------------------- driver
@dataclass
class OracleModes:
    lams: np.ndarray
    n_effs: np.ndarray
    E: np.ndarray  # (N, k)
    H: np.ndarray  # (N, k)
    A: sp.csr_matrix
    B: sp.csr_matrix
    sigma: complex
    powers: np.ndarray
    timings: dict


def boundary_dofs(order, nv, facets, fsel):
    verts = np.unique(facets[:, fsel])
    if order == 1:
        return np.unique(np.concatenate([verts, nv + fsel]))
    return np.unique(np.concatenate([verts, nv + 3 * fsel, nv + 3 * fsel + 1, nv + 3 * fsel + 2]))


def compute_modes(
    p,
    t,
    eps_kind,
    epsilon_r,
    wavelength,
    X,
    W,
    *,
    mu_r=1,
    num_modes=1,
    order=1,
    dirichlet_facets=None,
    radius=np.inf,
    n_guess=None,
    v0=None,
    with_h=True,
) -> OracleModes:
    """waveguide.py:528-654 restated.  ``dirichlet_facets``: None (no metallic boundary) or
    an index array of facets whose dofs are condensed out (waveguide.py:610-619)."""
    import time

    tm = {}
    t0 = time.perf_counter()
    t = np.sort(np.asarray(t, dtype=np.int64), axis=0)
    nv, ne = p.shape[1], t.shape[1]
    facets, t2f = topology(p, t)
    nf = facets.shape[1]
    dofs = element_dofs(order, nv, nf, t, t2f)
    n = nv + nf if order == 1 else nv + 3 * nf + 2 * ne
    k0 = 2 * np.pi / wavelength
    tm["topology"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    eps_q = interpolate_eps(eps_kind, epsilon_r, ne, X)
    A, B = assemble_waveguide(p, t, dofs, order, X, W, eps_q, k0, mu_r, radius, n)
    tm["assembly"] = time.perf_counter() - t0
    if n_guess:
        sigma = k0**2 * n_guess**2
    else:
        sigma = k0**2 * np.max(epsilon_r) * 1.1
    t0 = time.perf_counter()
    if dirichlet_facets is not None:
        D = boundary_dofs(order, nv, facets, np.asarray(dirichlet_facets))
        I = np.setdiff1d(np.arange(n), D)
        K = (-A)[I][:, I]
        M = (-B)[I][:, I]
        v0c = None if v0 is None else v0[I]
        lams, xr = spla.eigs(K, M=M, k=num_modes, sigma=sigma, v0=v0c)
        xs = np.zeros((n, num_modes), dtype=complex)
        xs[I] = xr
    else:
        lams, xs = spla.eigs(-A, M=-B, k=num_modes, sigma=sigma, v0=v0)
        xs = xs.astype(complex)
    tm["eigs"] = time.perf_counter() - t0
    kinds = local_kinds(order)
    zdofs = np.unique(dofs[kinds == 1])
    xs[zdofs, :] /= 1j * np.sqrt(lams[np.newaxis, :] / k0**4)
    hs = np.zeros_like(xs)
    powers = np.zeros(num_modes, dtype=complex)
    t0 = time.perf_counter()
    if with_h:
        for i, lam in enumerate(lams):
            H = calculate_hfield(p, t, dofs, order, X, W, xs[:, i], np.sqrt(lam), k0 * SPEED_OF_LIGHT, n)
            power = calculate_overlap(p, t, dofs, order, X, W, xs[:, i], H, xs[:, i], H)
            xs[:, i] /= np.sqrt(power)
            H /= np.sqrt(power)
            hs[:, i] = H
            powers[i] = power
    tm["hfield_power"] = time.perf_counter() - t0
    return OracleModes(lams, np.sqrt(lams) / k0, xs, hs, A, B, sigma, powers, tm)
