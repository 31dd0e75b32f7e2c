"""Waveguide eigenmode solver -- B200-native drop-in for ``femwell.maxwell.waveguide``.

Same public API as the reference (``/root/reference/femwell/maxwell/waveguide.py``):
``compute_modes(basis_epsilon_r, epsilon_r, wavelength, *, mu_r, num_modes, order,
metallic_boundaries, radius, n_guess, solver)`` returning ``Modes`` of ``Mode`` (:528-654), and the
``Mode`` scalar post-processing (:54-72, :112-258), ``calculate_hfield`` (:657-677),
``calculate_overlap`` same-basis branch (:699-736).

All arithmetic runs in libfemwell_b200 (hand-written sm_100a CUDA behind the C ABI of
``include/femwell_b200.h``); this module only validates arguments, moves numpy arrays across the
boundary and wraps the results in the unchanged dataclasses.  There is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import cached_property
from typing import List, Optional

import numpy as np

from ..basis import Basis
from ..context import (FUN_CONFINEMENT, FUN_COUPLING, FUN_E2_E4, FUN_EX2_EY2, FUN_EX2_EY2_EALL, FUN_OVERLAP,
                       Context)
from ..element import (ElementTriN1, ElementTriN2, ElementTriP1, ElementTriP2, eps_kind_of)

speed_of_light = 299792458.0
epsilon_0 = 8.8541878128e-12
mu_0 = 1.25663706212e-06

_contexts = {}


def get_context(device: Optional[int] = None) -> Context:
    """One library handle per (process, device)."""
    if device is None:
        import os

        device = int(os.environ.get("FEMWELL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        try:
            import torch

            if torch.cuda.is_available():
                device = device % torch.cuda.device_count()
        except Exception:
            pass
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def _bind(basis: Basis, device=None, subset=False) -> Context:
    """Handle that holds ``basis``: the device the basis was solved on (``solver_options={"device": ...}``) unless
    ``device`` is given.  ``subset``: also restrict the forms to ``basis.tind`` (only the bilinear forms need it;
    the functionals take their element selection per call)."""
    if device is None:
        device = getattr(basis, "_fw_device", None)
    ctx = get_context(device)
    if not ctx.same_space(basis) or ctx.nnz_structural is None:
        ctx.set_space(basis)
        ctx.symbolic()
    elif subset:
        ctx.set_element_subset(basis.tind)
    return ctx


@dataclass(frozen=True)
class Mode:
    frequency: float
    """Frequency of the light"""
    k: float
    """Propagation constant of the mode"""
    basis_epsilon_r: Basis
    """Basis used for epsilon_r"""
    epsilon_r: np.ndarray
    """Epsilon_r with which the mode was calculated"""
    basis: Basis
    """Basis on which the mode was calculated and E/H are defined"""
    E: np.ndarray
    """Electric field of the mode"""
    H: np.ndarray
    """Magnetic field of the mode"""

    @property
    def omega(self):
        return 2 * np.pi * self.frequency

    @property
    def k0(self):
        return self.omega / speed_of_light

    @property
    def wavelength(self):
        return speed_of_light / self.frequency

    @property
    def n_eff(self):
        return self.k / self.k0

    def _sel(self, elements):
        return None if elements is None else self.basis.mesh.normalize_elements(elements)

    @cached_property
    def poynting(self):
        """Poynting vector E x H projected onto ElementVector(ElementDG(Lagrange), 3) (waveguide.py:74-95).
        Returns ``(poynting_basis, P_proj)`` like the reference."""
        from ..element import ElementDG, ElementTriP1, ElementTriP2, ElementVector

        lag = ElementTriP1() if self.basis.elem.order == 1 else ElementTriP2()
        pbasis = self.basis.with_element(ElementVector(ElementDG(lag), 3))
        return pbasis, _bind(self.basis).poynting(self.E, self.H)

    def _poynting_component(self, c):
        from ..element import ElementDG, ElementTriP1, ElementTriP2

        _, P = self.poynting
        lag = ElementTriP1() if self.basis.elem.order == 1 else ElementTriP2()
        return self.basis.with_element(ElementDG(lag)), P[c::3]

    @property
    def Sx(self):
        """waveguide.py:97-100"""
        return self._poynting_component(0)

    @property
    def Sy(self):
        return self._poynting_component(1)

    @property
    def Sz(self):
        return self._poynting_component(2)

    @cached_property
    def te_fraction(self):
        """waveguide.py:112-127"""
        o = _bind(self.basis).functional(FUN_EX2_EY2, [self.E], elements=self.basis.tind)
        return (o[0] / (o[0] + o[1])).real

    @cached_property
    def tm_fraction(self):
        return 1 - self.te_fraction

    @cached_property
    def transversality(self):
        """waveguide.py:135-159"""
        o = _bind(self.basis).functional(FUN_EX2_EY2_EALL, [self.E], elements=self.basis.tind)
        return ((o[0] + o[1]) / o[2]).real

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(k: {self.k}, n_eff:{self.n_eff})"

    def calculate_overlap(self, mode, elements=None):
        return calculate_overlap(self.basis, self.E, self.H, mode.basis, mode.E, mode.H)

    def calculate_coupling_coefficient(self, mode, delta_epsilon):
        """waveguide.py:167-179"""
        o = _bind(self.basis).functional(FUN_COUPLING, [self.E, mode.E], aux=np.asarray(delta_epsilon),
                                         aux_kind=eps_kind_of(self.basis_epsilon_r.elem), elements=self.basis.tind)
        return o[0]

    def calculate_effective_area(self, field="xy"):
        """waveguide.py:181-213"""
        opt = {"xy": 0, "x": 1}.get(field, 2)
        o = _bind(self.basis).functional(FUN_E2_E4, [self.E], option=opt, elements=self.basis.tind)
        return np.real(o[0] ** 2 / o[1])

    def calculate_propagation_loss(self, distance):
        return -20 / np.log(10) * self.k0 * np.imag(self.n_eff) * distance

    def calculate_power(self, elements=None):
        """waveguide.py:218-223"""
        if elements is None or (hasattr(elements, "__len__") and len(elements) == 0):
            basis = self.basis
        else:
            basis = self.basis.with_elements(elements)
        return calculate_overlap(basis, self.E, self.H, basis, self.E, self.H)

    def calculate_confinement_factor(self, elements):
        """waveguide.py:225-249"""
        sel = self.basis.mesh.normalize_elements(elements)
        o = _bind(self.basis).functional(FUN_CONFINEMENT, [self.E], aux=np.asarray(self.epsilon_r),
                                         aux_kind=eps_kind_of(self.basis_epsilon_r.elem), elements=sel)
        return speed_of_light * epsilon_0 * o[0]

    def calculate_intensity(self):
        """Intensity 0.5 Re(E x H*)_z projected onto discontinuous P1 (waveguide.py:260-280).
        Returns ``(basis2, intensity2)`` like the reference."""
        from ..element import ElementDG

        basis2 = self.basis.with_element(ElementDG(ElementTriP1()))
        return basis2, _bind(self.basis).intensity(self.E, self.H)

    def eval_error_estimator(self):
        """Interior-facet jump estimator (waveguide.py:473-502): per element, half the sum over its facets of
        ``int_f h (|[grad E_z] . n|^2 + |[eps E_t] . n|^2) ds``; feeds ``adaptive_theta`` / ``mesh.refined``
        (docs/photonics/examples/refinement.py:58-80)."""
        return _bind(self.basis).error_estimator(self.E, eps_kind_of(self.basis_epsilon_r.elem),
                                                 np.asarray(self.epsilon_r))

    def calculate_pertubated_neff(self, delta_epsilon):
        return self.n_eff + self.calculate_coupling_coefficient(self, delta_epsilon) * epsilon_0 * speed_of_light * 0.5


@dataclass(frozen=True)
class Modes:
    modes: List
    stats: dict = field(default_factory=dict, compare=False, repr=False)

    def __getitem__(self, idx) -> Mode:
        return self.modes[idx]

    def __len__(self) -> int:
        return len(self.modes)

    def __repr__(self) -> str:
        modes = "\n\t" + "\n\t".join(repr(mode) for mode in self.modes) + "\n"
        return f"{self.__class__.__name__}(modes=({modes}))"

    def sorted(self, key):
        return Modes(modes=sorted(self.modes, key=key))

    @property
    def n_effs(self):
        return np.array([mode.n_eff for mode in self.modes])


def _mode_element(order):
    if order == 1:
        return ElementTriN1() * ElementTriP1()
    if order == 2:
        return ElementTriN2() * ElementTriP2()
    raise AssertionError("Only order 1 and 2 implemented by now.")


def compute_modes(
    basis_epsilon_r,
    epsilon_r,
    wavelength,
    *,
    mu_r=1,
    num_modes=1,
    order=1,
    metallic_boundaries=False,
    radius=np.inf,
    n_guess=None,
    solver="b200",
    solver_options=None,
) -> Modes:
    """Computes the modes of a waveguide (reference signature, waveguide.py:528-540).

    ``solver`` must be ``"b200"``; ``solver_options`` (dict, optional): ``v0``, ``ncv``, ``tol``,
    ``maxiter``, ``refine``, ``leaf_elements``, ``reuse_symbolic`` (default True: keep the CSR pattern and the LU
    analysis while the mesh stays), ``device``, ``pcg_tol_exp`` (H-field mass solves to 10^-value, default 10), ``start_vector`` (``"smooth"`` (default): contrast-weighted smooth start vector,
    ``"uniform"``: uniform(-1, 1) like scipy; an explicit ``v0`` wins), ``mass_solver``
    (``"auto"``: block PCG with a sparse-LU fallback when it stagnates, ``"pcg"``, ``"direct"``),
    ``mode_tables`` (``dict(kind=, vec=, sc=, element_dofs=, n_dofs=)``: a foreign local basis, see
    ``Basis.with_tables`` / ``femwell_b200.bridge``).

    Errors mirror the reference: unknown ``solver`` -> ``ValueError`` (waveguide.py:563), ``order`` not in (1, 2)
    -> ``AssertionError`` (:572), no convergence -> a ``scipy.sparse.linalg.ArpackNoConvergence`` subclass carrying
    the converged pairs, singular shifted pencil -> ``RuntimeError("Factor is exactly singular")`` (scipy ``splu``).
    """
    if solver != "b200":
        raise ValueError("`solver` must either be `scipy` or `slepc` (femwell's CPU paths, not carried by "
                         "femwell_b200) or `b200`")
    opts = dict(solver_options or {})
    k0 = 2 * np.pi / wavelength
    element = _mode_element(order)
    basis = basis_epsilon_r.with_element(element)
    mode_tables = opts.pop("mode_tables", None)
    if mode_tables is not None:
        basis = basis.with_tables(**mode_tables)
    basis_epsilon_r = basis.with_element(basis_epsilon_r.elem)  # adjust quadrature (waveguide.py:575)
    epsilon_r = np.asarray(epsilon_r)
    eps_kind = eps_kind_of(basis_epsilon_r.elem)

    if n_guess:
        sigma = k0**2 * n_guess**2
    else:
        sigma = k0**2 * np.max(epsilon_r) * 1.1

    dirichlet = opts.pop("dirichlet_dofs", None)  # precomputed by the scikit-fem bridge (its own get_dofs)
    if metallic_boundaries and dirichlet is None:
        dirichlet = basis.get_dofs(None if metallic_boundaries is True else metallic_boundaries)

    device = opts.pop("device", None)
    ctx = get_context(device)
    basis._fw_device = ctx.device  # the Mode post-processing binds to the same GPU
    reuse = bool(opts.pop("reuse_symbolic", True)) and ctx.same_space(basis) and ctx.nnz_structural is not None
    if not reuse:
        ctx.set_space(basis, solve_hint=(dirichlet, opts.get("leaf_elements", 0)))
    else:
        ctx.set_element_subset(basis.tind)
    # per-call options of the H-field mass solves (always set: a handle is shared by all calls on its device)
    ctx.set_option(1, int(opts.pop("pcg_tol_exp", 10)))
    ctx.set_option(2, {"auto": 0, "pcg": 1, "direct": 2}[opts.pop("mass_solver", "auto")])
    # default start vector: smooth and weighted by the index contrast ("smooth"), or scipy-like noise ("uniform")
    ctx.set_option(3, {"smooth": 1, "uniform": 0}[opts.pop("start_vector", "uniform" if mode_tables is not None else "smooth")])
    lams, E, H, powers, stats = ctx.compute_modes(
        eps_kind, epsilon_r, wavelength, num_modes, sigma, mu_r=float(mu_r), radius=float(radius),
        dirichlet=dirichlet, reuse_symbolic=reuse, **opts)
    stats["timings"] = ctx.timings()
    ctx.nnz_structural = stats["timings"]["nnz_structural"]
    return Modes(
        modes=[
            Mode(
                frequency=speed_of_light / wavelength,
                k=np.sqrt(lams[i]),
                basis_epsilon_r=basis_epsilon_r,
                epsilon_r=epsilon_r,
                basis=basis,
                E=E[i],  # row views of the (k, N) result arrays (contiguous, no extra host copy)
                H=H[i],
            )
            for i in range(num_modes)
        ],
        stats=stats,
    )


def calculate_hfield(basis, xs, beta, omega):
    """waveguide.py:657-677"""
    return _bind(basis, subset=True).hfield(xs, beta, omega, mu_0)


def calculate_energy_current_density(basis, xs):
    """waveguide.py:680-696; returns ``(basis_energy, values)`` on the P0 basis."""
    from ..element import ElementTriP0

    return basis.with_element(ElementTriP0()), _bind(basis).energy_current_density(xs)


def _same_mesh(mi, mj):
    return mi is mj or (mi.p.shape == mj.p.shape and mi.t.shape == mj.t.shape and np.array_equal(mi.t, mj.t)
                        and np.array_equal(mi.p, mj.p))


def _same_quadrature_space(basis_i, basis_j):
    """The reference's same-basis test (waveguide.py:727-729): ``basis_i == basis_j`` or equal quadrature rules.
    DEVIATION (documented in INTEGRATION.md): the reference takes the same-basis branch for ANY two bases with equal
    rules and then pairs the two meshes element by element -- a numpy broadcast error for different element counts,
    a meaningless number otherwise.  Here equal rules on different meshes go to the cross-mesh branch."""
    if basis_i is basis_j:
        return True
    if basis_i.elem != basis_j.elem or basis_i.tables_id != basis_j.tables_id:
        return False
    if basis_i.X.shape != basis_j.X.shape or not (np.isclose(basis_i.X, basis_j.X).all()
                                                   and np.isclose(basis_i.W, basis_j.W).all()):
        return False
    return _same_mesh(basis_i.mesh, basis_j.mesh)


def calculate_overlap(basis_i, E_i, H_i, basis_j, E_j, H_j):
    """waveguide.py:699-759: same-basis Functional, or (different meshes) P1 projection of mode j on its own mesh
    evaluated at the quadrature points of basis_i."""
    if _same_quadrature_space(basis_i, basis_j):
        return _bind(basis_i).functional(FUN_OVERLAP, [E_i, H_i, E_j, H_j], elements=basis_i.tind)[0]
    return _bind(basis_i).overlap_cross(basis_j, E_i, H_i, E_j, H_j, mode=0, elements=basis_i.tind)


def calculate_scalar_product(basis_i, E_i, basis_j, H_j):
    """waveguide.py:762-782: int conj(E_i) x H_j (transverse parts)."""
    if _same_quadrature_space(basis_i, basis_j):
        # FUN_OVERLAP = 0.5 (conj(E_i) x H_j + E_j x conj(H_i)) with H_i = 0
        zero = np.zeros(basis_i.N, dtype=np.complex128)
        return 2 * _bind(basis_i).functional(FUN_OVERLAP, [E_i, zero, zero, H_j], elements=basis_i.tind)[0]
    return _bind(basis_i).overlap_cross(basis_j, E_i, None, None, H_j, mode=1, elements=basis_i.tind)
