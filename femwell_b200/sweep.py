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
