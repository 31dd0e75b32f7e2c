"""Shared fixtures/helpers for the CPU and GPU test-suites."""
import numpy as np

from femwell_b200.basis import Basis
from femwell_b200.element import (ElementDG, ElementTriN1, ElementTriN2, ElementTriP0, ElementTriP1, ElementTriP2,
                                  ElementVector)
from femwell_b200.mesh import MeshTri, waveguide_mesh

RECT_CASES = [  # docs/benchmarks/mode_solver_rectangle.py:38-45
    ((2.25, 1.0), ["left", "right"], 1.27627404),
    ((8.0, 1.0), ["left", "right"], 2.65679692),
    ((1.0, 2.25), ["left", "top"], 1.387926425),
    ((1.0, 8.0), ["left", "top"], 2.761465320),
]


def mode_element(order):
    return ElementTriN1() * ElementTriP1() if order == 1 else ElementTriN2() * ElementTriP2()


def rectangle_problem(n, case, jitter=0.0):
    (ec, ecl), bnd, ref = RECT_CASES[case]
    m = MeshTri.rectangle(n, jitter=jitter, frozen_x=[0.5], frozen_y=[0.5])
    c = m.centroids()
    core = (c[0] < 0.5) & (c[1] < 0.5)
    eps = np.where(core, ec, ecl).astype(np.float64)
    fsel = np.unique(np.concatenate([m.boundaries[b] for b in bnd]))
    return m, eps, bnd, fsel, ref


def strip_problem(n, slab_h=0.0, jitter=0.15):
    """C1/C2-like Si strip/rib in SiO2."""
    m = waveguide_mesh(n, slab_h=slab_h, jitter=jitter)
    eps = np.full(m.nelements, 1.444**2)
    eps[m.subdomains["core"]] = 3.4777**2
    if slab_h > 0:
        eps[m.subdomains["slab"]] = 3.4777**2
    return m, eps


def csr_rel_diff(a, b):
    """max |a-b| / max|b| over the union pattern + exact pattern equality flag."""
    a = a.tocsr().copy()
    b = b.tocsr().copy()
    a.sort_indices()
    b.sort_indices()
    same = a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    d = abs(a - b)
    scale = abs(b).max()
    return (d.max() / scale if d.nnz else 0.0), same
