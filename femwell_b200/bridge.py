"""scikit-fem facing bridge: what the `solver == "b200"` branch of the reference's ``compute_modes`` calls
(INTEGRATION.md section 1).  It only reads attributes of scikit-fem objects (``mesh.p``, ``mesh.t``,
``basis.element_dofs``, ``basis.X/W``, ``elem.lbasis``) and hands plain arrays to ``femwell_b200``.

scikit-fem is not installable in the build image, so this module is exercised with
duck-typed stand-ins: tests/test_cpu.py::test_bridge_* (array extraction, table extraction from a foreign ``lbasis``)
and tests/test_gpu.py::test_bridge_with_foreign_local_basis (a deliberately different local basis -- sign-flipped and
rescaled functions -- fed through ``lbasis``: n_eff invariant, matrix entries equal to the oracle driven by the same
tables, fields equal after the change of basis).

``tables="skfem"`` (default): the reference tables come from ``elem.lbasis`` and ``element_dofs`` / ``N`` from the
scikit-fem basis, so the returned dof vectors ARE scikit-fem's and may be wrapped in the reference's ``Mode``.
``tables="builtin"``: the hierarchical basis of ``femwell_b200.element`` (same FE space: identical eigenvalues,
overlaps, powers) -- its dof vectors must not be paired with a scikit-fem basis, so ``Mode=`` / ``Modes=`` are refused.
"""
from __future__ import annotations

import numpy as np

from .basis import Basis
from .element import (ElementComposite, ElementDG, ElementTriN1, ElementTriN2, ElementTriP0, ElementTriP1,
                      ElementTriP2, ElementVector, local_layout)
from .mesh import MeshTri


def _eps_element(skfem_elem):
    name = type(skfem_elem).__name__
    inner = type(getattr(skfem_elem, "elem", None)).__name__
    if name == "ElementTriP0":
        return ElementTriP0()
    if name == "ElementDG" and inner == "ElementTriP1":
        return ElementDG(ElementTriP1())
    if name == "ElementVector" and inner == "ElementTriP0":
        return ElementVector(ElementTriP0(), 3)
    raise NotImplementedError(f"epsilon basis element {skfem_elem!r} is not supported on the b200 path")


def mesh_from_skfem(mesh) -> MeshTri:
    """Same ``p``/``t`` (scikit-fem stores ``t`` column-sorted), sub-domains and boundaries by name."""
    m = MeshTri(np.asarray(mesh.p, dtype=np.float64), np.asarray(mesh.t))
    for name, idx in (getattr(mesh, "subdomains", None) or {}).items():
        m.subdomains[name] = np.asarray(idx, dtype=np.int32)
    for name, idx in (getattr(mesh, "boundaries", None) or {}).items():
        # our facet numbering (unique sorted vertex pairs, lexicographic) equals scikit-fem's
        m.boundaries[name] = np.asarray(idx, dtype=np.int32)
    return m


def tables_from_lbasis(elem, X, order):
    """Reference tables (kind, vec, sc) from scikit-fem's ``elem.lbasis(X, i)`` (composite element: a tuple of
    DiscreteFields, one per sub-element, the inactive one being zero -- SURVEY.md A.5).  ``kind`` is read off the
    active component and must be the entity-major layout (A.3)."""
    kind_expected, _ = local_layout(order)
    nb, nq = kind_expected.size, X.shape[1]
    kind = np.zeros(nb, dtype=np.int32)
    vec = np.zeros((nb, 2, nq))
    sc = np.zeros((nb, nq))

    def arr(a, shape):
        return None if a is None else np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(shape[:-1] + (-1,)), shape)

    for i in range(nb):
        fields = elem.lbasis(X, i)
        ft, fz = fields[0], fields[1]
        tv = arr(getattr(ft, "value", None), (2, nq))
        zv = arr(getattr(fz, "value", None), (nq,))
        t_active = tv is not None and np.any(tv != 0) or (getattr(ft, "curl", None) is not None
                                                          and np.any(np.asarray(ft.curl) != 0))
        if t_active:
            kind[i] = 0
            vec[i] = tv
            sc[i] = arr(ft.curl, (nq,))
        else:
            kind[i] = 1
            sc[i] = zv
            vec[i] = arr(fz.grad, (2, nq))
    if not np.array_equal(kind, kind_expected):
        raise NotImplementedError("the composite element's local dofs are not entity-major (vertex, facet, interior): "
                                  f"{kind.tolist()} instead of {kind_expected.tolist()}")
    return kind, vec, sc


def compute_modes_from_skfem(basis_epsilon_r, epsilon_r, wavelength, *, mu_r=1, num_modes=1, order=1,
                             metallic_boundaries=False, radius=np.inf, n_guess=None, Mode=None, Modes=None,
                             tables="skfem", mode_basis=None, solver_options=None):
    """Drop-in body for the reference's ``compute_modes`` (waveguide.py:565-654) with scikit-fem inputs.

    ``mode_basis``: the scikit-fem mode basis ``basis_epsilon_r.with_element(ElementTriN2() * ElementTriP2())``
    (built here when scikit-fem is importable; tests pass a duck-typed one)."""
    from .maxwell import waveguide as wg

    if tables not in ("skfem", "builtin"):
        raise ValueError("tables must be 'skfem' or 'builtin'")
    if tables == "builtin" and (Mode is not None or Modes is not None):
        raise ValueError("tables='builtin' returns dof vectors of femwell_b200's own local basis; they cannot be "
                         "wrapped in the reference's Mode next to a scikit-fem basis -- use tables='skfem'")
    mesh = mesh_from_skfem(basis_epsilon_r.mesh)
    b0 = Basis(mesh, _eps_element(basis_epsilon_r.elem),
               quadrature=(np.asarray(basis_epsilon_r.X), np.asarray(basis_epsilon_r.W)))
    tind = getattr(basis_epsilon_r, "tind", None)
    if tind is not None and len(tind) != mesh.nelements:
        b0.tind = np.asarray(tind, dtype=np.int32)  # element subset of the caller's basis (waveguide.py:574)
    opts = dict(solver_options or {})
    if tables == "skfem":
        if mode_basis is None:
            elem = _skfem_mode_element(order)
            if elem is None:
                raise ImportError("tables='skfem' needs scikit-fem (or a duck-typed `mode_basis`)")
            mode_basis = basis_epsilon_r.with_element(elem)
        kind, vec, sc = tables_from_lbasis(mode_basis.elem, b0.X, order)
        opts["mode_tables"] = dict(kind=kind, vec=vec, sc=sc, element_dofs=np.asarray(mode_basis.element_dofs),
                                   n_dofs=int(mode_basis.N))
    if metallic_boundaries and tables == "skfem" and hasattr(mode_basis, "get_dofs"):
        # scikit-fem's own boundary dofs (waveguide.py:615)
        d = mode_basis.get_dofs(None if metallic_boundaries is True else metallic_boundaries)
        opts["dirichlet_dofs"] = np.asarray(d.all() if hasattr(d, "all") and callable(d.all) and not isinstance(d, np.ndarray)
                                            else d, dtype=np.int32)
    modes = wg.compute_modes(b0, np.asarray(epsilon_r), wavelength, mu_r=mu_r, num_modes=num_modes, order=order,
                             metallic_boundaries=metallic_boundaries, radius=radius, n_guess=n_guess,
                             solver="b200", solver_options=opts)
    if Mode is None or Modes is None:
        return modes
    # fill the reference's own dataclasses (waveguide.py:641-654) next to the scikit-fem basis the tables came from
    return Modes(modes=[Mode(frequency=md.frequency, k=md.k, basis_epsilon_r=basis_epsilon_r, epsilon_r=epsilon_r,
                             basis=mode_basis, E=md.E, H=md.H) for md in modes])


def _skfem_mode_element(order):
    try:
        import skfem
    except ImportError:
        return None
    if order == 1:
        return skfem.ElementTriN1() * skfem.ElementTriP1()
    return skfem.ElementTriN2() * skfem.ElementTriP2()
