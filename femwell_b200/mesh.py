"""Triangle mesh container + synthetic structured mesh generators.

Mirrors the slice of scikit-fem's ``MeshTri`` that the waveguide path of femwell
touches (reference call sites: femwell/maxwell/waveguide.py:574-575, 615;
docs/benchmarks/mode_solver_rectangle.py:59-66).  Conventions follow SURVEY.md
Appendix A.1:

* ``p``  float64 (2, Nv) vertex coordinates,
* ``t``  int32   (3, Ne) vertex indices, **sorted ascending per column**,
* local edges ``(0,1), (1,2), (0,2)``; global ``facets`` (2, Nf) are the unique
  sorted vertex pairs in lexicographic order; ``t2f`` (3, Ne) is the inverse map,
* ``subdomains[name]`` -> element indices, ``boundaries[name]`` -> facet indices.

gmsh / shapely are not available in the target image, so the BASELINE configs
are generated synthetically here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Optional, Sequence

import numpy as np


@dataclass
class MeshTri:
    p: np.ndarray
    t: np.ndarray
    subdomains: Dict[str, np.ndarray] = field(default_factory=dict)
    boundaries: Dict[str, np.ndarray] = field(default_factory=dict)

    def __post_init__(self):
        self.p = np.ascontiguousarray(self.p, dtype=np.float64)
        t = np.sort(np.asarray(self.t, dtype=np.int64), axis=0)
        self.t = np.ascontiguousarray(t.astype(np.int32))
        self._build_facets()

    # ------------------------------------------------------------------ topology
    def _build_facets(self):
        nv = self.nvertices
        t = self.t.astype(np.int64)
        # local edges (0,1),(1,2),(0,2); t sorted => (low, high) already
        lo = np.concatenate([t[0], t[1], t[0]])
        hi = np.concatenate([t[1], t[2], t[2]])
        keys = lo * nv + hi
        uniq, inv = np.unique(keys, return_inverse=True)
        self.facets = np.ascontiguousarray(
            np.stack([uniq // nv, uniq % nv]).astype(np.int32)
        )
        self.t2f = np.ascontiguousarray(inv.reshape(3, -1).astype(np.int32))
        counts = np.bincount(inv, minlength=uniq.size)
        self._facet_count = counts

    @property
    def nvertices(self) -> int:
        return self.p.shape[1]

    @property
    def nelements(self) -> int:
        return self.t.shape[1]

    @property
    def nfacets(self) -> int:
        return self.facets.shape[1]

    @property
    def f2t(self) -> np.ndarray:
        """int32 (2, Nf): the two elements of every facet, ``-1`` on the boundary (scikit-fem ``mesh.f2t``); side 0
        is the element with the lower index."""
        if getattr(self, "_f2t", None) is None:
            ne = self.nelements
            f = self.t2f.ravel()
            e = np.tile(np.arange(ne, dtype=np.int64), 3)
            order = np.lexsort((e, f))
            fs, es = f[order], e[order]
            first = np.ones(fs.size, dtype=bool)
            first[1:] = fs[1:] != fs[:-1]
            out = np.full((2, self.nfacets), -1, dtype=np.int32)
            out[0, fs[first]] = es[first]
            out[1, fs[~first]] = es[~first]
            self._f2t = out
        return self._f2t

    def boundary_facets(self) -> np.ndarray:
        """Facets belonging to exactly one element."""
        return np.nonzero(self._facet_count == 1)[0].astype(np.int32)

    def facet_midpoints(self) -> np.ndarray:
        return 0.5 * (self.p[:, self.facets[0]] + self.p[:, self.facets[1]])

    def centroids(self) -> np.ndarray:
        return (self.p[:, self.t[0]] + self.p[:, self.t[1]] + self.p[:, self.t[2]]) / 3.0

    def facets_satisfying(self, test: Callable, boundaries_only: bool = True) -> np.ndarray:
        mid = self.facet_midpoints()
        ok = np.asarray(test(mid), dtype=bool)
        if boundaries_only:
            ok &= self._facet_count == 1
        return np.nonzero(ok)[0].astype(np.int32)

    def elements_satisfying(self, test: Callable) -> np.ndarray:
        return np.nonzero(np.asarray(test(self.centroids()), dtype=bool))[0].astype(np.int32)

    def with_boundaries(self, tests: Dict[str, Callable]) -> "MeshTri":
        for name, test in tests.items():
            self.boundaries[name] = self.facets_satisfying(test)
        return self

    def with_subdomains(self, tests: Dict[str, Callable]) -> "MeshTri":
        for name, test in tests.items():
            self.subdomains[name] = self.elements_satisfying(test)
        return self

    def normalize_elements(self, elements) -> np.ndarray:
        """Resolve an element selector (name, list of names, index array, callable)."""
        if elements is None:
            return np.arange(self.nelements, dtype=np.int32)
        if isinstance(elements, str):
            return np.asarray(self.subdomains[elements], dtype=np.int32)
        if callable(elements):
            return self.elements_satisfying(elements)
        if isinstance(elements, (list, tuple)) and len(elements) and isinstance(elements[0], str):
            return np.unique(np.concatenate([self.subdomains[n] for n in elements])).astype(np.int32)
        return np.asarray(elements, dtype=np.int32)

    # --------------------------------------------------------------- generators
    @classmethod
    def structured(
        cls,
        xs: Sequence[float],
        ys: Sequence[float],
        jitter: float = 0.0,
        seed: int = 1234,
        frozen_x: Iterable[float] = (),
        frozen_y: Iterable[float] = (),
    ) -> "MeshTri":
        """Tensor grid of ``len(xs)-1`` x ``len(ys)-1`` cells, each split along the
        same diagonal.  ``jitter`` (fraction of the local cell size) perturbs the
        interior vertices (avoids accidental exact zeros); vertices
        on the outer boundary or on the ``frozen`` grid lines (material
        interfaces) stay put.
        """
        xs = np.asarray(xs, dtype=np.float64)
        ys = np.asarray(ys, dtype=np.float64)
        nx, ny = xs.size, ys.size
        X, Y = np.meshgrid(xs, ys, indexing="xy")  # (ny, nx)
        if jitter > 0:
            rng = np.random.default_rng(seed)
            hx = np.minimum(np.diff(xs, prepend=xs[0] - 1e300), np.diff(xs, append=xs[-1] + 1e300))
            hy = np.minimum(np.diff(ys, prepend=ys[0] - 1e300), np.diff(ys, append=ys[-1] + 1e300))
            mx = np.ones(nx, dtype=bool)
            my = np.ones(ny, dtype=bool)
            mx[[0, -1]] = False
            my[[0, -1]] = False
            for f in frozen_x:
                mx[np.isclose(xs, f)] = False
            for f in frozen_y:
                my[np.isclose(ys, f)] = False
            dx = rng.uniform(-jitter, jitter, size=(ny, nx)) * hx[None, :]
            dy = rng.uniform(-jitter, jitter, size=(ny, nx)) * hy[None, :]
            # a vertex on a frozen x-line may still slide along it (dy) and vice versa,
            # so outer boundaries and material interfaces stay straight
            X = X + dx * mx[None, :]
            Y = Y + dy * my[:, None]
        p = np.stack([X.ravel(), Y.ravel()])
        idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
        v00 = idx[:-1, :-1].ravel()
        v10 = idx[:-1, 1:].ravel()
        v01 = idx[1:, :-1].ravel()
        v11 = idx[1:, 1:].ravel()
        t = np.concatenate(
            [np.stack([v00, v10, v11]), np.stack([v00, v11, v01])], axis=1
        )
        # interleave the two triangles of each cell so that element order is spatial
        ne = t.shape[1] // 2
        order = np.empty(2 * ne, dtype=np.int64)
        order[0::2] = np.arange(ne)
        order[1::2] = np.arange(ne) + ne
        t = t[:, order]
        m = cls(p, t)
        x0, x1, y0, y1 = xs[0], xs[-1], ys[0], ys[-1]
        tol = 1e-9 * max(x1 - x0, y1 - y0)
        m.with_boundaries(
            {
                "left": lambda q: np.abs(q[0] - x0) < tol,
                "right": lambda q: np.abs(q[0] - x1) < tol,
                "bottom": lambda q: np.abs(q[1] - y0) < tol,
                "top": lambda q: np.abs(q[1] - y1) < tol,
            }
        )
        return m

    @classmethod
    def rectangle(cls, n: int, bbox=(0.0, 1.0, 0.0, 1.0), **kw) -> "MeshTri":
        x0, x1, y0, y1 = bbox
        return cls.structured(np.linspace(x0, x1, n + 1), np.linspace(y0, y1, n + 1), **kw)


def graded_axis(breaks: Sequence[float], cells: Sequence[int]) -> np.ndarray:
    """Piecewise-uniform 1-D grid whose lines hit every break point exactly."""
    out = [np.array([breaks[0]], dtype=np.float64)]
    for a, b, n in zip(breaks[:-1], breaks[1:], cells):
        out.append(np.linspace(a, b, int(n) + 1)[1:])
    return np.concatenate(out)


def _split_cells(breaks, total):
    """Distribute ``total`` cells over the intervals proportionally to length (>=1 each)."""
    lens = np.diff(np.asarray(breaks, dtype=np.float64))
    raw = lens / lens.sum() * total
    cells = np.maximum(1, np.floor(raw).astype(int))
    while cells.sum() < total:
        cells[np.argmax(raw - cells)] += 1
    while cells.sum() > total:
        i = np.argmax(np.where(cells > 1, cells - raw, -np.inf))
        cells[i] -= 1
    return cells


def waveguide_mesh(
    n: int,
    core_w: float = 0.5,
    core_h: float = 0.22,
    slab_h: float = 0.0,
    bbox=(-1.5, 1.5, -1.0, 1.22),
    jitter: float = 0.15,
    seed: int = 1234,
    extra_x: Sequence[float] = (),
) -> MeshTri:
    """Strip (``slab_h=0``) or rib waveguide cross-section on an ``n x n`` grid
    (``2 n^2`` triangles) whose grid lines coincide with the material interfaces
    (geometry after
    docs/photonics/examples/bent_waveguide.py:43-49).
    Sub-domains: ``core``, ``slab`` (if any), ``box`` (y<0), ``clad``.
    """
    x0, x1, y0, y1 = bbox
    xb = sorted({x0, -core_w / 2, core_w / 2, x1, *extra_x})
    yb = sorted({y0, 0.0, core_h, y1} | ({slab_h} if slab_h > 0 else set()))
    xs = graded_axis(xb, _split_cells(xb, n))
    ys = graded_axis(yb, _split_cells(yb, n))
    m = MeshTri.structured(xs, ys, jitter=jitter, seed=seed, frozen_x=xb, frozen_y=yb)
    c = m.centroids()
    core = (np.abs(c[0]) < core_w / 2) & (c[1] > 0) & (c[1] < core_h)
    slab = (c[1] > 0) & (c[1] < slab_h) & ~core if slab_h > 0 else np.zeros_like(core)
    box = c[1] < 0
    clad = ~(core | slab | box)
    m.subdomains = {
        "core": np.nonzero(core)[0].astype(np.int32),
        "slab": np.nonzero(slab)[0].astype(np.int32),
        "box": np.nonzero(box)[0].astype(np.int32),
        "clad": np.nonzero(clad)[0].astype(np.int32),
    }
    return m


def geometric_axis(a: float, b: float, h0: float, ratio: float, fine_end: str) -> np.ndarray:
    """Grid from ``a`` to ``b`` whose spacing is ``h0`` at the ``fine_end`` ("a" or "b") and grows
    geometrically by ``ratio`` per cell towards the other end (rescaled to hit both ends exactly)."""
    length = b - a
    hs = []
    h = h0
    while sum(hs) + h < length:
        hs.append(h)
        h *= ratio
    hs = np.asarray(hs if hs else [length], dtype=np.float64)
    hs *= length / hs.sum()
    pts = np.concatenate([[0.0], np.cumsum(hs)])
    pts[-1] = length
    return a + pts if fine_end == "a" else b - pts[::-1]


def rib_benchmark_mesh(slab_thickness: float, h0: float = 0.075, ratio: float = 1.2, core_width: float = 3.0,
                       core_thickness: float = 1.0, slab_width: float = 18.0, box_thickness: float = 10.0,
                       clad_thickness: float = 3.0, jitter: float = 0.0) -> MeshTri:
    """Rib waveguide of the reference's second mode-solver benchmark (docs/benchmarks/mode_solver.py:52-104,
    after Hadley 2002): core ``core_width x core_thickness`` on a slab of ``slab_thickness`` spanning the whole
    width, box below, cladding above.  The reference meshes it with gmsh (resolution 0.1 in the core growing to
    0.3); here: a tensor grid with spacing ``h0`` over the core, growing geometrically by ``ratio`` away from it,
    grid lines on every material interface.  Sub-domains ``core`` (core + slab), ``box``, ``clad``."""
    hw, sw = core_width / 2, slab_width / 2
    xs = np.concatenate([
        geometric_axis(-sw, -hw, h0, ratio, "b")[:-1],
        np.linspace(-hw, hw, int(round(core_width / h0)) + 1)[:-1],
        geometric_axis(hw, sw, h0, ratio, "a"),
    ])
    ybr = sorted({0.0, core_thickness} | ({slab_thickness} if slab_thickness > 0 else set()))
    ys = [geometric_axis(-box_thickness, 0.0, h0, ratio, "b")[:-1]]
    for a, b in zip(ybr[:-1], ybr[1:]):
        ys.append(np.linspace(a, b, max(1, int(round((b - a) / h0))) + 1)[:-1])
    ys.append(geometric_axis(core_thickness, clad_thickness, h0, ratio, "a"))
    ys = np.concatenate(ys)
    m = MeshTri.structured(xs, ys, jitter=jitter, frozen_x=[-hw, hw], frozen_y=ybr)
    c = m.centroids()
    core = (np.abs(c[0]) < hw) & (c[1] > 0) & (c[1] < core_thickness)
    if slab_thickness > 0:
        core |= (c[1] > 0) & (c[1] < slab_thickness)
    box = c[1] < 0
    m.subdomains = {
        "core": np.nonzero(core)[0].astype(np.int32),
        "box": np.nonzero(box)[0].astype(np.int32),
        "clad": np.nonzero(~(core | box))[0].astype(np.int32),
    }
    return m


# ------------------------------------------------------------------ refinement (host bookkeeping)
def adaptive_theta(est: np.ndarray, theta: float = 0.5, max: Optional[float] = None) -> np.ndarray:
    """Elements whose indicator exceeds ``theta`` times the largest one (scikit-fem ``adaptive_theta`` as used by
    docs/photonics/examples/refinement.py:69)."""
    est = np.asarray(est)
    if max is None:
        return np.nonzero(theta * np.max(est) < est)[0]
    return np.nonzero(theta * max < est)[0]


def _refine_marked(mesh: MeshTri, marked: np.ndarray) -> MeshTri:
    """Split every triangle according to its marked facets (0: keep, 1: bisect, 2: bisect twice, 3: four similar
    children); a marked facet gets one new vertex at its midpoint, so the result is conforming as long as every
    element with a marked facet has its longest facet marked (guaranteed by the caller's closure)."""
    p, t, t2f, facets = mesh.p, mesh.t.astype(np.int64), mesh.t2f, mesh.facets
    nv = mesh.nvertices
    mid_of = np.full(mesh.nfacets, -1, dtype=np.int64)
    mf = np.nonzero(marked)[0]
    mid_of[mf] = nv + np.arange(mf.size)
    pm = 0.5 * (p[:, facets[0, mf]] + p[:, facets[1, mf]])
    p_new = np.concatenate([p, pm], axis=1)
    # facet lengths; the reference edge of an element is its longest facet (ties: lowest local index)
    d = p[:, facets[0]] - p[:, facets[1]]
    flen = np.hypot(d[0], d[1])
    el = flen[t2f]                       # (3, Ne): local edges (0,1), (1,2), (0,2)
    lo = np.argmax(el, axis=0)           # local index of the longest edge
    # rotate: a, b = ends of the longest edge, c = the opposite vertex
    ends = np.array([[0, 1], [1, 2], [0, 2]])
    opp = np.array([2, 0, 1])
    ne = mesh.nelements
    cols = np.arange(ne)
    ia, ib, ic = ends[lo, 0], ends[lo, 1], opp[lo]
    a, b, c = t[ia, cols], t[ib, cols], t[ic, cols]
    # local edge index joining two local vertices
    edge_of = np.full((3, 3), -1)
    edge_of[0, 1] = edge_of[1, 0] = 0
    edge_of[1, 2] = edge_of[2, 1] = 1
    edge_of[0, 2] = edge_of[2, 0] = 2
    f_ab = t2f[lo, cols]
    f_bc = t2f[edge_of[ib, ic], cols]
    f_ac = t2f[edge_of[ia, ic], cols]
    mab, mbc, mac = mid_of[f_ab], mid_of[f_bc], mid_of[f_ac]
    has_ab, has_bc, has_ac = mab >= 0, mbc >= 0, mac >= 0
    assert not np.any((has_bc | has_ac) & ~has_ab), "closure violated: longest facet unmarked"
    tri, parent = [], []

    def emit(sel, *verts):
        idx = np.nonzero(sel)[0]
        if idx.size:
            tri.append(np.stack([v[idx] for v in verts]))
            parent.append(idx)

    emit(~has_ab, a, b, c)
    green = has_ab & ~has_bc & ~has_ac
    emit(green, a, mab, c)
    emit(green, mab, b, c)
    blue_r = has_ab & has_bc & ~has_ac
    emit(blue_r, a, mab, c)
    emit(blue_r, mab, b, mbc)
    emit(blue_r, mab, mbc, c)
    blue_l = has_ab & ~has_bc & has_ac
    emit(blue_l, mab, b, c)
    emit(blue_l, a, mab, mac)
    emit(blue_l, mab, c, mac)
    red = has_ab & has_bc & has_ac
    emit(red, a, mab, mac)
    emit(red, mab, b, mbc)
    emit(red, mac, mbc, c)
    emit(red, mab, mbc, mac)
    t_new = np.concatenate(tri, axis=1)
    par = np.concatenate(parent)
    order = np.argsort(par, kind="stable")  # children stay next to each other in parent order (spatial locality)
    t_new, par = t_new[:, order], par[order]
    out = MeshTri(p_new, t_new)
    for name, els in mesh.subdomains.items():
        flag = np.zeros(ne, dtype=bool)
        flag[np.asarray(els, dtype=np.int64)] = True
        out.subdomains[name] = np.nonzero(flag[par])[0].astype(np.int32)
    if mesh.boundaries:
        # a child facet lies on a named boundary iff both its ends are (an end of / the midpoint of) a facet of it
        nvn = out.nvertices
        for name, fs in mesh.boundaries.items():
            fs = np.asarray(fs, dtype=np.int64)
            v0, v1 = facets[0, fs].astype(np.int64), facets[1, fs].astype(np.int64)
            mm = mid_of[fs]
            pairs = [np.stack([v0, v1])[:, mm < 0], np.stack([v0, mm])[:, mm >= 0], np.stack([v1, mm])[:, mm >= 0]]
            pr = np.sort(np.concatenate(pairs, axis=1), axis=0)
            keys = pr[0] * nvn + pr[1]
            okeys = out.facets[0].astype(np.int64) * nvn + out.facets[1].astype(np.int64)
            out.boundaries[name] = np.nonzero(np.isin(okeys, keys))[0].astype(np.int32)
    out.parent_elements = par.astype(np.int32)
    return out


def _refined(mesh: MeshTri, elements=None) -> MeshTri:
    """``mesh.refined(elements)`` of scikit-fem as used by docs/photonics/examples/refinement.py:80: red-green-blue
    refinement of the selected triangles (all three facets of a selected triangle are split; neighbours are closed by
    splitting their longest facet first so that no hanging node remains); ``elements=None`` or an integer: uniform
    red refinement (that many times).  Sub-domains and named boundaries are inherited by the children."""
    if elements is None or isinstance(elements, (int, np.integer)):
        times = 1 if elements is None else int(elements)
        m = mesh
        for _ in range(times):
            m = _refine_marked(m, np.ones(m.nfacets, dtype=bool))
        return m
    els = mesh.normalize_elements(elements).astype(np.int64)
    marked = np.zeros(mesh.nfacets, dtype=bool)
    marked[mesh.t2f[:, els].ravel()] = True
    d = mesh.p[:, mesh.facets[0]] - mesh.p[:, mesh.facets[1]]
    flen = np.hypot(d[0], d[1])
    longest = mesh.t2f[np.argmax(flen[mesh.t2f], axis=0), np.arange(mesh.nelements)]
    while True:  # closure: an element with any marked facet must have its longest facet marked
        touched = marked[mesh.t2f].any(axis=0)
        need = longest[touched]
        if marked[need].all():
            break
        marked[need] = True
    return _refine_marked(mesh, marked)


MeshTri.refined = _refined


def coupled_waveguide_mesh(n: int, gap: float = 0.4, w_core_1: float = 0.45, w_core_2: float = 0.46,
                           h_core: float = 0.22, w_sim: float = 4.0, h_clad: float = 1.0, h_box: float = 1.0,
                           jitter: float = 0.15, seed: int = 1234) -> MeshTri:
    """Two parallel strip waveguides (docs/photonics/examples/coupled_mode_theory.py:41-95, crosstalk.py:60-140) on
    an ``n x n`` tensor grid whose lines follow every material interface.  Sub-domains ``core_1`` (left),
    ``core_2`` (right), ``clad`` (y > 0 outside the cores), ``box`` (y < 0)."""
    xb = sorted({-w_sim / 2, -w_core_1 - gap / 2, -gap / 2, gap / 2, w_core_2 + gap / 2, w_sim / 2})
    yb = sorted({-h_box, 0.0, h_core, h_clad})
    xs = graded_axis(xb, _split_cells(xb, n))
    ys = graded_axis(yb, _split_cells(yb, n))
    m = MeshTri.structured(xs, ys, jitter=jitter, seed=seed, frozen_x=xb, frozen_y=yb)
    c = m.centroids()
    in_y = (c[1] > 0) & (c[1] < h_core)
    c1 = in_y & (c[0] > -w_core_1 - gap / 2) & (c[0] < -gap / 2)
    c2 = in_y & (c[0] > gap / 2) & (c[0] < w_core_2 + gap / 2)
    box = c[1] < 0
    m.subdomains = {
        "core_1": np.nonzero(c1)[0].astype(np.int32),
        "core_2": np.nonzero(c2)[0].astype(np.int32),
        "box": np.nonzero(box)[0].astype(np.int32),
        "clad": np.nonzero(~(c1 | c2 | box))[0].astype(np.int32),
    }
    return m
