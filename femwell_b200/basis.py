"""Light ``Basis`` stand-in exposing the attributes of scikit-fem's ``CellBasis``
that the waveguide path reads (reference femwell/maxwell/waveguide.py:574-575,
602-603, 615, 627, 727-729).

It is host-side bookkeeping only: mesh, element, quadrature ``(X, W)``, the
element subset ``tind`` and the DOF numbering.  All arithmetic happens in the
CUDA library; ``interpolate`` is provided for user convenience (plots, tests).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .element import (
    ElementComposite,
    ElementDG,
    ElementTriP0,
    ElementTriP1,
    ElementVector,
    element_tables,
    eps_kind_of,
    local_layout,
    triangle_quadrature,
    EPS_P0,
    EPS_DG_P1,
    EPS_VEC_P0,
)
from .mesh import MeshTri


def _dg_count(elem) -> int:
    """dofs per element of the element-local (P0 / discontinuous / vector-valued) spaces"""
    name = type(elem).__name__
    if name == "ElementTriP0":
        return 1
    if name == "ElementDG":
        return {"ElementTriP0": 1, "ElementTriP1": 3, "ElementTriP2": 6}[type(elem.elem).__name__]
    if name == "ElementVector":
        return elem.dim * _dg_count(elem.elem)
    raise NotImplementedError(f"no dof count for {elem!r}")


class Basis:
    def __init__(self, mesh: MeshTri, elem, intorder: Optional[int] = None, elements=None, quadrature=None):
        self.mesh = mesh
        self.elem = elem
        if quadrature is not None:
            self.X, self.W = (np.ascontiguousarray(a, dtype=np.float64) for a in quadrature)
        else:
            # scikit-fem: intorder = 2*maxdeg, orders <= 1 promoted to the 3-point rule
            self.X, self.W = triangle_quadrature(2 * elem.maxdeg if intorder is None else intorder)
        self.tind = None if elements is None else mesh.normalize_elements(elements)
        self._dofs = None
        # optional reference tables of a foreign local basis (scikit-fem's ``elem.lbasis``, femwell_b200.bridge);
        # None: the built-in hierarchical basis of femwell_b200.element
        self._tables = None
        self._n_dofs = None
        self._fw_device = None  # GPU the basis was solved on (set by compute_modes; follows derived bases)

    # -- scikit-fem compatible helpers ------------------------------------------------
    @property
    def quadrature(self):
        return self.X, self.W

    def with_element(self, elem) -> "Basis":
        """Same mesh, quadrature and element subset, other element."""
        b = Basis(self.mesh, elem, quadrature=(self.X, self.W))
        b.tind = self.tind
        b._fw_device = self._fw_device
        return b

    def with_elements(self, elements) -> "Basis":
        b = Basis(self.mesh, self.elem, quadrature=(self.X, self.W))
        b.tind = self.mesh.normalize_elements(elements)
        if self._dofs is not None:
            b._dofs = self._dofs
        b._tables, b._n_dofs, b._fw_device = self._tables, self._n_dofs, self._fw_device
        return b

    def with_tables(self, kind, vec, sc, element_dofs=None, n_dofs=None) -> "Basis":
        """The same mode basis with a foreign local basis: reference tables ``(kind, vec, sc)`` at this basis'
        quadrature points (layout of femwell_b200.element.ElementTables) and, optionally, the owner's own
        ``element_dofs`` / dof count.  All kernels are table driven, so dof vectors then refer to that basis."""
        from .element import ElementTables, local_layout

        if not self.is_mode_basis:
            raise ValueError("with_tables needs the mixed Nedelec x Lagrange basis")
        lk, sub = local_layout(self.elem.order)
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        if not np.array_equal(kind, lk):
            raise ValueError("local dof layout must be entity-major: vertex, facet, interior (SURVEY.md A.3)")
        b = Basis(self.mesh, self.elem, quadrature=(self.X, self.W))
        b.tind = self.tind
        b._fw_device = self._fw_device
        b._tables = ElementTables(self.elem.order, kind.size, self.X.shape[1], kind,
                                  np.ascontiguousarray(vec, dtype=np.float64), np.ascontiguousarray(sc, dtype=np.float64),
                                  self.X, self.W, sub)
        if element_dofs is not None:
            b._dofs = np.ascontiguousarray(element_dofs, dtype=np.int32)
            b._n_dofs = int(n_dofs) if n_dofs is not None else int(b._dofs.max()) + 1
        return b

    @property
    def tables_id(self):
        """hashable identity of the reference tables (None: built-in)"""
        if self._tables is None:
            return None
        return (self._tables.vec.tobytes(), self._tables.sc.tobytes())

    def __eq__(self, other):
        return (
            isinstance(other, Basis)
            and self.mesh is other.mesh
            and self.elem == other.elem
            and self.X.shape == other.X.shape
            and np.array_equal(self.X, other.X)
            and np.array_equal(self.W, other.W)
            and (
                (self.tind is None and other.tind is None)
                or (self.tind is not None and other.tind is not None and np.array_equal(self.tind, other.tind))
            )
        )

    __hash__ = object.__hash__

    # -- dof numbering ------------------------------------------------------------------
    @property
    def is_mode_basis(self):
        return isinstance(self.elem, ElementComposite)

    @property
    def N(self) -> int:
        m = self.mesh
        if self._n_dofs is not None:
            return self._n_dofs
        if self.is_mode_basis:
            if self.elem.order == 1:
                return m.nvertices + m.nfacets
            return m.nvertices + 3 * m.nfacets + 2 * m.nelements
        return _dg_count(self.elem) * m.nelements

    @property
    def element_dofs(self) -> np.ndarray:
        """int32 (Nbfun, Ne): vertex dofs, facet dofs, interior dofs."""
        if self._dofs is None and self.is_mode_basis:
            # the numbering depends on (mesh, order) only: share it between the Basis objects that
            # compute_modes re-creates on every call (sweeps on one mesh)
            cache = self.mesh.__dict__.setdefault("_dof_cache", {})
            if self.elem.order in cache:
                self._dofs = cache[self.elem.order]
        if self._dofs is None:
            m = self.mesh
            nv, nf, ne = m.nvertices, m.nfacets, m.nelements
            if self.is_mode_basis:
                if self.elem.order == 1:
                    ed = np.concatenate([m.t, nv + m.t2f], axis=0)
                else:
                    rows = [m.t]
                    for k in range(3):
                        base = nv + 3 * m.t2f[k]
                        rows.append(np.stack([base, base + 1, base + 2]))
                    e = np.arange(ne, dtype=np.int64)
                    rows.append(np.stack([nv + 3 * nf + 2 * e, nv + 3 * nf + 2 * e + 1]))
                    ed = np.concatenate(rows, axis=0)
            else:
                kind = eps_kind_of(self.elem)
                e = np.arange(ne, dtype=np.int64)
                ed = e[None, :] if kind == EPS_P0 else np.stack([3 * e, 3 * e + 1, 3 * e + 2])
            self._dofs = np.ascontiguousarray(ed, dtype=np.int32)
            if self.is_mode_basis:
                self.mesh.__dict__.setdefault("_dof_cache", {})[self.elem.order] = self._dofs
        return self._dofs

    def zeros(self, dtype=np.float64) -> np.ndarray:
        return np.zeros(self.N, dtype=dtype)

    def split_indices(self):
        """[transverse (Nedelec) dofs, longitudinal (Lagrange) dofs] (waveguide.py:627)."""
        kind, _ = local_layout(self.elem.order)
        ed = self.element_dofs
        return [np.unique(ed[kind == 0]), np.unique(ed[kind == 1])]

    def get_dofs(self, facets=None, elements=None) -> np.ndarray:
        """DOFs on boundary facets (``facets``: None = whole boundary, name, list of
        names, callable on midpoints, or index array) or inside ``elements``."""
        m = self.mesh
        if elements is not None:
            sel = m.normalize_elements(elements)
            return np.unique(self.element_dofs[:, sel])
        if facets is None:
            fsel = m.boundary_facets()
        elif isinstance(facets, str):
            fsel = m.boundaries[facets]
        elif callable(facets):
            fsel = m.facets_satisfying(facets)
        elif isinstance(facets, (list, tuple)) and len(facets) and isinstance(facets[0], str):
            fsel = np.unique(np.concatenate([m.boundaries[n] for n in facets]))
        else:
            fsel = np.asarray(facets, dtype=np.int32)
        nv = m.nvertices
        verts = np.unique(m.facets[:, fsel])
        if not self.is_mode_basis:
            raise NotImplementedError
        if self.elem.order == 1:
            return np.unique(np.concatenate([verts, nv + fsel])).astype(np.int32)
        f = fsel.astype(np.int64)
        return np.unique(np.concatenate([verts, nv + 3 * f, nv + 3 * f + 1, nv + 3 * f + 2])).astype(np.int32)

    # -- convenience (host, numpy) ---------------------------------------------------------
    def tables(self):
        if self._tables is not None:
            return self._tables
        return element_tables(self.elem.order, self.X, self.W)

    def facet_tables(self):
        """Gauss-Legendre rule on [0, 1] of degree 2 * maxdeg (scikit-fem's default for an InteriorFacetBasis of this
        element, waveguide.py:475) and the reference vectors of every local basis function at its points on the three
        local edges (0,1) (1,2) (0,2): ``(s [nqf], w [nqf], vec [nb, 3, 2, nqf])``."""
        if self._tables is not None and getattr(self, "_facet_vec", None) is None:
            raise NotImplementedError("a foreign local basis needs its own facet tables (Basis._facet_vec)")
        npts = int(np.ceil((2 * self.elem.maxdeg + 1) / 2))
        x, w = np.polynomial.legendre.leggauss(npts)
        s, w = 0.5 * (x + 1.0), 0.5 * w
        if self._tables is not None:
            return s, w, np.ascontiguousarray(self._facet_vec, dtype=np.float64)
        ends = [((0.0, 0.0), (1.0, 0.0)), ((1.0, 0.0), (0.0, 1.0)), ((0.0, 0.0), (0.0, 1.0))]
        nb = self.tables().nbfun
        vec = np.zeros((nb, 3, 2, npts))
        for k, (a, b) in enumerate(ends):
            Xk = np.stack([a[0] + s * (b[0] - a[0]), a[1] + s * (b[1] - a[1])])
            vec[:, k] = element_tables(self.elem.order, Xk, w).vec
        return np.ascontiguousarray(s), np.ascontiguousarray(w), np.ascontiguousarray(vec)

    def interpolate(self, x: np.ndarray):
        """Host-side evaluation at the quadrature points (not on the hot path).
        Mode basis: returns ``((Ex, Ey), Ez)`` with arrays (Ne_sel, nqp); epsilon basis:
        (Ne_sel, nqp) or (3, Ne_sel, nqp)."""
        m = self.mesh
        sel = np.arange(m.nelements) if self.tind is None else self.tind
        lam = np.stack([1 - self.X[0] - self.X[1], self.X[0], self.X[1]])
        if not self.is_mode_basis:
            kind = eps_kind_of(self.elem)
            if kind == EPS_P0:
                return np.repeat(x[sel][:, None], self.X.shape[1], axis=1)
            x3 = x.reshape(-1, 3)[sel]
            if kind == EPS_DG_P1:
                return x3 @ lam
            return np.repeat(x3.T[:, :, None], self.X.shape[1], axis=2)
        tb = self.tables()
        p, t = m.p, m.t[:, sel]
        a11 = p[0, t[1]] - p[0, t[0]]
        a12 = p[0, t[2]] - p[0, t[0]]
        a21 = p[1, t[1]] - p[1, t[0]]
        a22 = p[1, t[2]] - p[1, t[0]]
        det = a11 * a22 - a12 * a21
        # A^-T = 1/det [[a22, -a21], [-a12, a11]]
        xe = x[self.element_dofs[:, sel]]  # (nb, Ne)
        nk = tb.kind == 0
        vx = np.einsum("le,lq->eq", xe[nk], tb.vec[nk, 0])
        vy = np.einsum("le,lq->eq", xe[nk], tb.vec[nk, 1])
        ex = (a22[:, None] * vx - a21[:, None] * vy) / det[:, None]
        ey = (-a12[:, None] * vx + a11[:, None] * vy) / det[:, None]
        ez = np.einsum("le,lq->eq", xe[~nk], tb.sc[~nk])
        return (ex, ey), ez
