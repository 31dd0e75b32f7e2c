"""Multi-GPU sweeps: one independent solve per GPU, results gathered once at the end.

The reference runs sweeps as serial Python loops (docs/photonics/examples/vary_wavelength.py:83-95,
vary_width.py:46-65, bent_waveguide.py:104-119, crosstalk.py:147-243).  A single solve stays on
one GPU (BASELINE.json north_star item 5); sweep points are independent, so the index space is cut
into contiguous blocks (which keeps ``n_guess`` continuation chains on one rank) and the only
communication is one ``all_gather`` of a few scalars per point (NCCL on the GPU box, gloo in the
CPU tests).  No data-path collective exists or is invented.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import numpy as np


def block_partition(n_points: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_points, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def _dist():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def sweep(points: Sequence, solve_point: Callable[[int, object], Sequence[float]], width: int,
          device: Optional[str] = None) -> Optional[np.ndarray]:
    """Run ``solve_point(index, point) -> width float64 values`` over this rank's block and gather.

    Returns the full ``(len(points), width + 1)`` table on every rank (last column: status,
    0 = ok, 1 = failed; failed rows hold NaN -- a failed point does not abort the sweep,
    SURVEY.md section 5).  Works without an initialised process group (world size 1).
    """
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    n = len(points)
    lo, hi = block_partition(n, world, rank)
    cap = (n + world - 1) // world
    local = np.full((cap, width + 1), np.nan)
    for i in range(lo, hi):
        try:
            vals = np.asarray(solve_point(i, points[i]), dtype=np.float64).ravel()
            local[i - lo, :width] = vals
            local[i - lo, width] = 0.0
        except Exception as exc:  # noqa: BLE001 - mark the point, keep sweeping
            local[i - lo, width] = 1.0
            if os.environ.get("FEMWELL_B200_SWEEP_RAISE"):
                raise
            print(f"[sweep] point {i} failed on rank {rank}: {exc}")
    if not dist:
        return local[: hi - lo]
    import torch

    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(local).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    table = np.empty((n, width + 1))
    for r in range(world):
        a, b = block_partition(n, world, r)
        table[a:b] = out[r].cpu().numpy()[: b - a]
    return table


def pack_complex(z) -> List[float]:
    z = np.asarray(z, dtype=np.complex128).ravel()
    return np.stack([z.real, z.imag], axis=1).ravel().tolist()


def unpack_complex(row: np.ndarray, k: int) -> np.ndarray:
    return row[0:2 * k:2] + 1j * row[1:2 * k:2]


def wavelength_sweep(basis_epsilon_r, epsilon_of_wavelength: Callable[[float], np.ndarray], wavelengths, *,
                     num_modes=1, order=2, n_guess=None, **kw) -> np.ndarray:
    """Fixed-mesh dispersion sweep (vary_wavelength.py:83-95 / calculate_GVD.py:108-118): the mesh, FE
    space, structural CSR pattern and nested-dissection analysis are shared by all points of a rank.
    Returns complex n_eff [len(wavelengths), num_modes] (NaN where a point failed)."""
    from .maxwell.waveguide import compute_modes

    def solve_point(i, wl):
        modes = compute_modes(basis_epsilon_r, epsilon_of_wavelength(wl), wl, num_modes=num_modes, order=order,
                              n_guess=n_guess, **kw)
        return pack_complex(modes.n_effs)

    table = sweep(list(wavelengths), solve_point, 2 * num_modes)
    return np.stack([unpack_complex(row, num_modes) for row in table])


def group_index(wavelengths, n_eff):
    """n_g = n_eff - lambda d n_eff / d lambda by central differences (docs/math/dispersion.md:153-161)."""
    wavelengths = np.asarray(wavelengths, dtype=np.float64)
    return n_eff - wavelengths * np.gradient(n_eff, wavelengths)


# ------------------------------------------------------------------ BASELINE config 4: coupled-waveguide geometry sweep
COUPLER_COLUMNS = ("n_eff_sym", "n_eff_asym", "n_eff_1", "coupling_length_um", "A_into_1", "A_into_2",
                   "crosstalk_coeff", "ref_overlap_iso", "ref_overlap_sym")


def coupler_point(gap: float, width: float, *, n: int = 224, wavelength: float = 1.55, order: int = 2,
                  core_index: float = 3.4777, clad_index: float = 1.444, delta_width: float = 0.01,
                  reference_mode=None, solver_options=None):
    """One geometry of a coupled-waveguide sweep (docs/photonics/examples/crosstalk.py:147-243 and
    coupled_mode_theory.py:41-147): its own mesh, the two supermodes of the pair, the mode of waveguide 1 in isolation
    (same mesh: the CSR pattern and the LU analysis of the first solve are reused), the same-mesh overlaps of the
    isolated mode with both supermodes, the worst-case crosstalk coefficient and the coupling length.  With
    ``reference_mode`` (a ``Mode`` on ANOTHER mesh) two cross-mesh overlaps are added (waveguide.py:737-759).
    Returns ``(row, modes_both, modes_1)`` with ``row`` ordered as ``COUPLER_COLUMNS``."""
    from .basis import Basis
    from .element import ElementTriP0
    from .maxwell.waveguide import compute_modes
    from .mesh import coupled_waveguide_mesh

    mesh = coupled_waveguide_mesh(n, gap=gap, w_core_1=width, w_core_2=width + delta_width)
    b0 = Basis(mesh, ElementTriP0(), intorder=4)
    eps_both = np.full(mesh.nelements, clad_index**2)
    eps_both[mesh.subdomains["core_1"]] = core_index**2
    eps_1 = eps_both.copy()
    eps_both[mesh.subdomains["core_2"]] = core_index**2
    so = dict(solver_options or {})
    modes_both = compute_modes(b0, eps_both, wavelength, num_modes=2, order=order, solver_options=dict(so))
    modes_1 = compute_modes(b0, eps_1, wavelength, num_modes=1, order=order, solver_options=dict(so))
    a1 = abs(modes_1[0].calculate_overlap(modes_both[0])) ** 2
    a2 = abs(modes_1[0].calculate_overlap(modes_both[1])) ** 2
    c1, c2 = np.sqrt(a1 / (a1 + a2)), np.sqrt(a2 / (a1 + a2))
    crosstalk = 4 * c1**2 * c2**2  # crosstalk.py:246-262: worst-case EME power transfer 1 - (c1^4 + c2^4 - 2 c1^2 c2^2)
    k0 = 2 * np.pi / wavelength
    dn = np.real(modes_both[0].n_eff - modes_both[1].n_eff)
    lc = np.pi / (k0 * dn) if dn != 0 else np.inf  # coupled_mode_theory.py:127
    ro_iso = ro_sym = np.nan
    if reference_mode is not None:
        ro_iso = abs(modes_1[0].calculate_overlap(reference_mode))
        ro_sym = abs(modes_both[0].calculate_overlap(reference_mode))
    row = [np.real(modes_both[0].n_eff), np.real(modes_both[1].n_eff), np.real(modes_1[0].n_eff), lc, a1, a2,
           crosstalk, ro_iso, ro_sym]
    return row, modes_both, modes_1


def coupler_geometry_sweep(gaps, widths, *, n: int = 224, wavelength: float = 1.55, order: int = 2,
                           cross_mesh: bool = True, **kw) -> np.ndarray:
    """BASELINE config 4: ``len(gaps) x len(widths)`` coupled-waveguide geometries, block-partitioned over the ranks, one
    geometry = one mesh + 2 solves (3 modes) + 2 same-mesh overlaps (+ 2 cross-mesh overlaps against the isolated mode of
    the rank's first geometry when ``cross_mesh``), results all_gathered once.  Returns
    ``(n_points, len(COUPLER_COLUMNS) + 1)`` (last column: status; NaN rows for failed points)."""
    points = [(float(g), float(w)) for g in gaps for w in widths]
    state = {"ref": None}

    def solve_point(i, pt):
        row, _both, m1 = coupler_point(pt[0], pt[1], n=n, wavelength=wavelength, order=order,
                                       reference_mode=state["ref"] if cross_mesh else None, **kw)
        if cross_mesh and state["ref"] is None:
            state["ref"] = m1[0]  # later geometries of this rank are compared with it across meshes
        return row

    return sweep(points, solve_point, len(COUPLER_COLUMNS))
