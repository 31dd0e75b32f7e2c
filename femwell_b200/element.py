"""Reference-element tables for the mixed Nedelec x Lagrange elements of the
waveguide path (``ElementTriN1()*ElementTriP1()`` / ``ElementTriN2()*ElementTriP2()``,
reference femwell/maxwell/waveguide.py:567-572) and the triangle quadrature rules
the caller's epsilon-basis hands down (``waveguide.py:574-575``).

The CUDA kernels are *table driven* (SURVEY.md R2): they receive, per local basis
function ``l`` and quadrature point ``q``,

* ``kind[l]``      0 = H(curl) (Nedelec, transverse field), 1 = H1 (Lagrange, E_z),
* ``vec[l, :, q]`` reference vector: Nedelec value phi_hat, or Lagrange grad_hat psi,
* ``sc[l, q]``     reference scalar: Nedelec curl_hat, or Lagrange value psi,

and apply the affine / covariant-Piola map themselves
(``phi = A^-T phi_hat``, ``curl phi = curl_hat / det A``, ``grad psi = A^-T grad_hat psi``).
With scikit-fem present the same tables can be filled from ``elem.lbasis``; without it
the hierarchical basis of SURVEY.md Appendix D is used (same FE space).

Local ordering is entity-major as in scikit-fem's ``ElementComposite`` / ``Dofs``
(SURVEY.md A.3): vertices, then per facet, then interior.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# ----------------------------------------------------------------------------- quadrature


def triangle_quadrature(intorder: int):
    """Return ``X`` (2, nqp), ``W`` (nqp,) on the reference triangle (area 1/2)."""
    if intorder <= 2:
        X = np.array([[1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]])
        W = np.full(3, 1 / 6)
    elif intorder <= 4:
        a1, w1 = 0.445948490915965, 0.223381589678011 / 2
        a2, w2 = 0.091576213509771, 0.109951743655322 / 2
        pts = []
        for a in (a1, a2):
            b = 1 - 2 * a
            pts += [(a, a), (a, b), (b, a)]
        X = np.array(pts).T
        W = np.array([w1] * 3 + [w2] * 3)
    elif intorder <= 5:
        a1 = 0.470142064105115
        a2 = 0.101286507323456
        w0, w1, w2 = 0.225 / 2, 0.132394152788506 / 2, 0.125939180544827 / 2
        pts = [(1 / 3, 1 / 3)]
        for a in (a1, a2):
            b = 1 - 2 * a
            pts += [(a, a), (a, b), (b, a)]
        X = np.array(pts).T
        W = np.array([w0] + [w1] * 3 + [w2] * 3)
    elif intorder <= 6:
        a1, w1 = 0.249286745170910, 0.116786275726379 / 2
        a2, w2 = 0.063089014491502, 0.050844906370207 / 2
        a3, b3, w3 = 0.053145049844817, 0.310352451033784, 0.082851075618374 / 2
        pts, ws = [], []
        for a, w in ((a1, w1), (a2, w2)):
            b = 1 - 2 * a
            pts += [(a, a), (a, b), (b, a)]
            ws += [w] * 3
        c3 = 1 - a3 - b3
        for u, v in ((a3, b3), (b3, a3), (a3, c3), (c3, a3), (b3, c3), (c3, b3)):
            pts.append((u, v))
            ws.append(w3)
        X = np.array(pts).T
        W = np.array(ws)
    else:
        raise NotImplementedError("triangle quadrature implemented up to degree 6")
    return np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(W, dtype=np.float64)


# ----------------------------------------------------------------------------- elements


class _Elem:
    """Tiny stand-ins for the scikit-fem element classes named in the reference API."""

    maxdeg = 0

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return type(self).__name__ + "()"


class ElementTriP0(_Elem):
    maxdeg = 0


class ElementTriP1(_Elem):
    maxdeg = 1


class ElementTriP2(_Elem):
    maxdeg = 2


class ElementTriN1(_Elem):
    maxdeg = 1


class ElementTriN2(_Elem):
    maxdeg = 2

    def __mul__(self, other):
        return ElementComposite(self, other)


def _n1_mul(self, other):
    return ElementComposite(self, other)


ElementTriN1.__mul__ = _n1_mul


class ElementDG(_Elem):
    def __init__(self, elem):
        self.elem = elem
        self.maxdeg = elem.maxdeg

    def __repr__(self):
        return f"ElementDG({self.elem!r})"


class ElementVector(_Elem):
    def __init__(self, elem, dim=3):
        self.elem = elem
        self.dim = dim
        self.maxdeg = elem.maxdeg

    def __repr__(self):
        return f"ElementVector({self.elem!r}, {self.dim})"


class ElementComposite(_Elem):
    def __init__(self, *elems):
        self.elems = list(elems)
        self.maxdeg = max(e.maxdeg for e in elems)

    @property
    def order(self):
        if isinstance(self.elems[0], ElementTriN1) and isinstance(self.elems[1], ElementTriP1):
            return 1
        if isinstance(self.elems[0], ElementTriN2) and isinstance(self.elems[1], ElementTriP2):
            return 2
        raise AssertionError("Only order 1 and 2 implemented by now.")

    def __repr__(self):
        return " * ".join(repr(e) for e in self.elems)


# epsilon representation codes shared with the C-ABI (include/femwell_b200.h FW_EPS_*)
EPS_P0 = 0  # one value per element                       (ElementTriP0)
EPS_DG_P1 = 1  # three values per element, dof = 3*e + a       (ElementDG(ElementTriP1))
EPS_VEC_P0 = 2  # diagonal anisotropy, dof = 3*e + c            (ElementVector(ElementTriP0), 3)


def eps_kind_of(elem) -> int:
    if isinstance(elem, ElementTriP0):
        return EPS_P0
    if isinstance(elem, ElementDG) and isinstance(elem.elem, ElementTriP1):
        return EPS_DG_P1
    if isinstance(elem, ElementVector) and isinstance(elem.elem, ElementTriP0) and elem.dim == 3:
        return EPS_VEC_P0
    name = type(elem).__name__
    # scikit-fem objects, if a user hands them in
    if name == "ElementTriP0":
        return EPS_P0
    if name == "ElementDG" and type(getattr(elem, "elem", None)).__name__ == "ElementTriP1":
        return EPS_DG_P1
    if name == "ElementVector" and type(getattr(elem, "elem", None)).__name__ == "ElementTriP0":
        return EPS_VEC_P0
    raise NotImplementedError(f"epsilon basis element {elem!r} not supported on the b200 path")


# ----------------------------------------------------------------------------- tables

_GRAD_L = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # grad_hat lambda_a
_EDGES = [(0, 1), (1, 2), (0, 2)]  # local facets


def _lam(X):
    return np.stack([1 - X[0] - X[1], X[0], X[1]])  # (3, nqp)


def _cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _whitney(lam, a, b):
    """w_ab = lam_a grad lam_b - lam_b grad lam_a ; curl = 2 grad lam_a x grad lam_b."""
    val = lam[a][None, :] * _GRAD_L[b][:, None] - lam[b][None, :] * _GRAD_L[a][:, None]
    curl = np.full(lam.shape[1], 2 * _cross(_GRAD_L[a], _GRAD_L[b]))
    return val, curl


def _nedelec(order, X):
    lam = _lam(X)
    nq = X.shape[1]
    vals, curls = [], []
    for a, b in _EDGES:
        w, c = _whitney(lam, a, b)
        vals.append(w)
        curls.append(c)
        if order == 2:
            g = lam[a][None, :] * _GRAD_L[b][:, None] + lam[b][None, :] * _GRAD_L[a][:, None]
            vals.append(g)
            curls.append(np.zeros(nq))
    if order == 2:
        for c_, (a, b) in ((2, (0, 1)), (1, (0, 2))):
            w, cw = _whitney(lam, a, b)
            vals.append(lam[c_][None, :] * w)
            curls.append(_cross(_GRAD_L[c_][:, None], w) + lam[c_] * cw)
    return np.stack(vals), np.stack(curls)  # (nbt, 2, nq), (nbt, nq)


def _lagrange(order, X):
    lam = _lam(X)
    if order == 1:
        vals = [lam[a] for a in range(3)]
        grads = [np.repeat(_GRAD_L[a][:, None], X.shape[1], axis=1) for a in range(3)]
    else:
        vals = [lam[a] * (2 * lam[a] - 1) for a in range(3)]
        grads = [(4 * lam[a] - 1)[None, :] * _GRAD_L[a][:, None] for a in range(3)]
        for a, b in _EDGES:
            vals.append(4 * lam[a] * lam[b])
            grads.append(
                4 * (lam[a][None, :] * _GRAD_L[b][:, None] + lam[b][None, :] * _GRAD_L[a][:, None])
            )
    return np.stack(vals), np.stack(grads)  # (nbz, nq), (nbz, 2, nq)


@dataclass(frozen=True)
class ElementTables:
    order: int
    nbfun: int
    nqp: int
    kind: np.ndarray  # int32 (nbfun,)
    vec: np.ndarray  # float64 (nbfun, 2, nqp)
    sc: np.ndarray  # float64 (nbfun, nqp)
    X: np.ndarray  # (2, nqp)
    W: np.ndarray  # (nqp,)
    sub_index: np.ndarray  # int32 (nbfun,) index inside its own sub-element

    @property
    def nbt(self):
        return int((self.kind == 0).sum())

    @property
    def nbz(self):
        return int((self.kind == 1).sum())


def local_layout(order: int):
    """(kind, sub_index) per composite local dof, entity-major."""
    if order == 1:
        kind = [1, 1, 1, 0, 0, 0]
        sub = [0, 1, 2, 0, 1, 2]
    elif order == 2:
        kind, sub = [1, 1, 1], [0, 1, 2]
        for f in range(3):
            kind += [0, 0, 1]
            sub += [2 * f, 2 * f + 1, 3 + f]
        kind += [0, 0]
        sub += [6, 7]
    else:
        raise AssertionError("Only order 1 and 2 implemented by now.")
    return np.array(kind, dtype=np.int32), np.array(sub, dtype=np.int32)


def element_tables(order: int, X: np.ndarray, W: np.ndarray) -> ElementTables:
    X = np.ascontiguousarray(X, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    kind, sub = local_layout(order)
    nval, ncurl = _nedelec(order, X)
    pval, pgrad = _lagrange(order, X)
    nb, nq = kind.size, X.shape[1]
    vec = np.zeros((nb, 2, nq))
    sc = np.zeros((nb, nq))
    for l in range(nb):
        if kind[l] == 0:
            vec[l], sc[l] = nval[sub[l]], ncurl[sub[l]]
        else:
            vec[l], sc[l] = pgrad[sub[l]], pval[sub[l]]
    return ElementTables(order, nb, nq, kind, np.ascontiguousarray(vec), np.ascontiguousarray(sc), X, W, sub)
