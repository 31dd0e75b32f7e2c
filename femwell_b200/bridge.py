"""scikit-fem facing bridge: what the `solver == "b200"` branch of the reference's ``compute_modes`` calls
(INTEGRATION.md section 1).  It only reads attributes of scikit-fem objects (``mesh.p``, ``mesh.t``,
``basis.element_dofs``, ``basis.X/W``, ``elem.lbasis``) and hands plain arrays to ``femwell_b200``.

scikit-fem is not installable in the build image, so this module is exercised only with
duck-typed stand-ins (tests/test_cpu.py::test_bridge_extracts_arrays_from_duck_typed_basis); the table
extraction from ``elem.lbasis`` follows SURVEY.md Appendix A.5/A.8 and must be re-checked the first time
scikit-fem is importable.  ``tables="builtin"`` uses the hierarchical basis of ``femwell_b200.element`` (same FE
space: identical eigenvalues, overlaps, powers; raw dof vectors then refer to that basis).
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
    """Reference tables (kind, vec, sc) from scikit-fem's ``elem.lbasis(X, i)`` (composite: a tuple of
    DiscreteFields, the inactive component being zero)."""
    kind, _ = local_layout(order)
    nb, nq = kind.size, X.shape[1]
    vec = np.zeros((nb, 2, nq))
    sc = np.zeros((nb, nq))
    for i in range(nb):
        fields = elem.lbasis(X, i)
        if kind[i] == 0:
            f = fields[0]
            vec[i] = np.asarray(f.value).reshape(2, -1)
            sc[i] = np.asarray(f.curl).reshape(-1)
        else:
            f = fields[1]
            sc[i] = np.asarray(f.value).reshape(-1)
            vec[i] = np.asarray(f.grad).reshape(2, -1)
    return kind, vec, sc


def compute_modes_from_skfem(basis_epsilon_r, epsilon_r, wavelength, *, mu_r=1, num_modes=1, order=1,
                             metallic_boundaries=False, radius=np.inf, n_guess=None, Mode=None, Modes=None,
                             tables="builtin", solver_options=None):
    """Drop-in body for the reference's ``compute_modes`` (waveguide.py:565-654) with scikit-fem inputs."""
    from .maxwell import waveguide as wg

    mesh = mesh_from_skfem(basis_epsilon_r.mesh)
    b0 = Basis(mesh, _eps_element(basis_epsilon_r.elem),
               quadrature=(np.asarray(basis_epsilon_r.X), np.asarray(basis_epsilon_r.W)))
    tind = getattr(basis_epsilon_r, "tind", None)
    if tind is not None and len(tind) != mesh.nelements:
        raise NotImplementedError("compute_modes on an element subset is not supported on the b200 path")
    if tables != "builtin":
        raise NotImplementedError("tables='skfem' needs scikit-fem at hand to be validated; see tables_from_lbasis")
    modes = wg.compute_modes(b0, np.asarray(epsilon_r), wavelength, mu_r=mu_r, num_modes=num_modes, order=order,
                             metallic_boundaries=metallic_boundaries, radius=radius, n_guess=n_guess,
                             solver="b200", solver_options=solver_options)
    if Mode is None or Modes is None:
        return modes
    # fill the reference's own dataclasses (waveguide.py:641-654); `basis` is rebuilt on the scikit-fem side
    elem = _skfem_mode_element(order)
    basis = basis_epsilon_r.with_element(elem) if elem is not None else modes[0].basis
    return Modes(modes=[Mode(frequency=md.frequency, k=md.k, basis_epsilon_r=basis_epsilon_r, epsilon_r=epsilon_r,
                             basis=basis, E=md.E, H=md.H) for md in modes])


def _skfem_mode_element(order):
    try:
        import skfem
    except ImportError:
        return None
    if order == 1:
        return skfem.ElementTriN1() * skfem.ElementTriP1()
    return skfem.ElementTriN2() * skfem.ElementTriP2()
