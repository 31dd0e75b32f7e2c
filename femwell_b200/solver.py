"""Eigen-solver adapter in scikit-fem's EigenSolver protocol -- the B200 counterpart of
``femwell.solver.solver_eigen_slepc`` (/root/reference/femwell/solver.py:87-142).

``solver_eigen_b200(k=..., sigma=...)`` returns ``solver(K, M) -> (lams, xs)`` for ``K x = lam M x`` with
``K``, ``M`` scipy sparse matrices assembled by anyone (e.g. by scikit-fem); shift-invert Arnoldi around
``sigma`` with the sparse LU of ``K - sigma M`` on the GPU.  ``xs`` has shape ``(N, k)`` like scipy's ``eigs``.
"""
from __future__ import annotations

import numpy as np


def solver_eigen_b200(**kwargs):
    params = {"sigma": None, "k": 5, "coords": None, "device": None}
    params.update(kwargs)

    def solver(K, M, **solve_time_kwargs):
        from .maxwell.waveguide import get_context

        params.update(solve_time_kwargs)
        if params["sigma"] is None:
            raise ValueError("solver_eigen_b200 is a shift-invert solver: `sigma` is required")
        ctx = get_context(params["device"])
        extra = {k: params[k] for k in ("v0", "ncv", "tol", "maxiter", "refine", "leaf_elements") if k in params}
        lams, X, stats = ctx.eigs_csr(K, M, params["k"], params["sigma"], coords=params["coords"], **extra)
        solver.stats = stats
        return lams, np.ascontiguousarray(X.T)

    return solver
