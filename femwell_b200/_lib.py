"""ctypes binding of libfemwell_b200.so (C ABI in include/femwell_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present when a handle is created, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfemwell_b200.so")

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER


class FwSpace(C.Structure):
    _fields_ = [
        ("order", c_i32), ("n_vertices", c_i32), ("n_elements", c_i32), ("n_dofs", c_i32),
        ("nbfun", c_i32), ("nqp", c_i32),
        ("p", c_vp), ("t", c_vp), ("element_dofs", c_vp), ("kind", c_vp), ("vec", c_vp),
        ("sc", c_vp), ("X", c_vp), ("W", c_vp),
    ]


class FwEigsParams(C.Structure):
    _fields_ = [
        ("k", c_i32), ("ncv", c_i32), ("maxiter", c_i32), ("tol", c_f64),
        ("sigma_re", c_f64), ("sigma_im", c_f64),
        ("n_dirichlet", c_i32), ("dirichlet", c_vp), ("v0", c_vp),
        ("refine", c_i32), ("leaf_elements", c_i32),
    ]


class FwEigsStats(C.Structure):
    _fields_ = [
        ("nconv", c_i32), ("iterations", c_i32), ("restarts", c_i32), ("n_op", c_i32),
        ("pivots_perturbed", c_i32), ("lu_levels", c_i32), ("lu_fronts", c_i32),
        ("lu_factor_entries", c_i64),
        ("t_symbolic_ms", c_f64), ("t_factor_ms", c_f64), ("t_arnoldi_ms", c_f64),
        ("t_solve_ms_total", c_f64), ("max_residual", c_f64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class FwModesParams(C.Structure):
    _fields_ = [
        ("wavelength", c_f64), ("mu_r", c_f64), ("radius", c_f64),
        ("eps_kind", c_i32), ("eps_is_complex", c_i32), ("eps", c_vp),
        ("eigs", FwEigsParams), ("reuse_symbolic", c_i32),
    ]


class FwTimings(C.Structure):
    _fields_ = [
        ("upload_ms", c_f64), ("symbolic_ms", c_f64), ("element_kernel_ms", c_f64), ("reduce_ms", c_f64),
        ("lu_symbolic_ms", c_f64), ("lu_factor_ms", c_f64), ("arnoldi_ms", c_f64), ("postprocess_ms", c_f64),
        ("total_ms", c_f64), ("download_ms", c_f64), ("element_kernel_launches", c_i64),
        ("kernel_launches", c_i64), ("nnz_structural", c_i64),
        ("lu_solve_entries", c_i64), ("lu_factor_flops", c_f64), ("pcg_iterations", c_i64),
        ("pcg_rel_residual", c_f64), ("hfield_ms", c_f64), ("mass_direct_blocks", c_i64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class FemwellB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfemwell_b200 error {code}: {msg}")
        self.code = code


try:  # the reference's callers catch scipy's class (scipy/sparse/linalg/_eigen/arpack/arpack.py:303-329)
    from scipy.sparse.linalg import ArpackNoConvergence as _ArpackNoConvergence
except Exception:  # pragma: no cover - scipy is a hard dependency of the reference
    _ArpackNoConvergence = RuntimeError


class B200NoConvergence(FemwellB200Error, _ArpackNoConvergence):
    """FW_ERR_NOCONV.  Compatible with ``scipy.sparse.linalg.ArpackNoConvergence``: ``eigenvalues`` and
    ``eigenvectors`` hold the pairs that did converge (shapes ``(nconv,)`` and ``(N, nconv)``)."""

    def __init__(self, code, msg, eigenvalues=None, eigenvectors=None):
        RuntimeError.__init__(self, f"ARPACK error -1: {msg}" if not msg.startswith("ARPACK") else msg)
        self.code = code
        self.eigenvalues = np.zeros(0, dtype=np.complex128) if eigenvalues is None else eigenvalues
        self.eigenvectors = np.zeros((0, 0), dtype=np.complex128) if eigenvectors is None else eigenvectors


class B200SingularError(FemwellB200Error):
    """FW_ERR_SINGULAR: a ``RuntimeError`` with the message of scipy's ``splu`` (SuperLU) for this case."""

    def __init__(self, code, msg):
        RuntimeError.__init__(self, "Factor is exactly singular")
        self.code = code


ERR_NOCONV, ERR_SINGULAR = -3, -4


# every symbol include/femwell_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "fw_last_error", "fw_version", "fw_create", "fw_destroy", "fw_synchronize", "fw_set_option", "fw_host_alloc", "fw_host_free", "fw_set_space", "fw_set_element_subset", "fw_analysis_hint",
    "fw_symbolic", "fw_assemble_waveguide", "fw_matrix_nnz", "fw_matrix_export", "fw_coo_to_csr",
    "fw_coo_to_csr_export", "fw_matrix_spmv", "fw_bench_spmv", "fw_eigs", "fw_eigs_csr", "fw_lu_factor_shifted", "fw_lu_solve",
    "fw_postprocess_modes", "fw_hfield", "fw_functional", "fw_error_estimator", "fw_intensity", "fw_poynting", "fw_overlap_cross", "fw_energy_current_density", "fw_compute_modes", "fw_get_timings",
    "fw_host_dense_eig", "fw_debug_lu_symbolic", "fw_debug_lu_symbolic_get", "fw_debug_lu_analysis", "fw_debug_lu_analysis_get", "fw_debug_lu_get",
]

_lib = None


def load():
    """Load the shared library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m femwell_b200.build` "
                "(the b200 path has no CPU fallback)"
            )
        lib = C.CDLL(LIB_PATH)
        lib.fw_last_error.restype = C.c_char_p
        lib.fw_version.restype = C.c_int
        for name in EXPORTS:
            fn = getattr(lib, name)
            if name not in ("fw_last_error", "fw_host_alloc"):
                fn.restype = C.c_int
        lib.fw_host_alloc.restype = C.c_void_p
        lib.fw_host_alloc.argtypes = [C.c_size_t]
        lib.fw_host_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def check(code, partial=None):
    """Raise for a non-zero status.  ``partial() -> (eigenvalues, eigenvectors)`` supplies the converged pairs that
    a FW_ERR_NOCONV call has already written to the caller's buffers."""
    if code == 0:
        return
    msg = load().fw_last_error().decode()
    if code == ERR_NOCONV:
        lam, vec = partial() if partial is not None else (None, None)
        raise B200NoConvergence(code, msg, lam, vec)
    if code == ERR_SINGULAR:
        raise B200SingularError(code, msg)
    raise FemwellB200Error(code, msg)


class PinnedPool:
    """Page-locked result buffers (fw_host_alloc) recycled between solves: the D2H copy of E and H into them is one DMA
    at bus speed.  ``empty(shape, dtype)`` returns a numpy array backed by such a buffer; the buffer goes back to the
    pool when the array (and every view of it) has been garbage collected."""

    MAX_CACHED_PER_SIZE = 4

    class _Block:
        def __init__(self, pool, ptr, nbytes, shape, typestr):
            self._pool, self._ptr, self._nbytes = pool, ptr, nbytes
            self.__array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

        def __del__(self):
            try:
                self._pool._give_back(self._ptr, self._nbytes)
            except Exception:  # interpreter shutdown
                pass

    def __init__(self):
        self._free = {}
        self.closed = False

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if self.closed or nbytes < (1 << 20):
            return np.empty(shape, dtype=dtype)
        cached = self._free.get(nbytes)
        p = cached.pop() if cached else load().fw_host_alloc(nbytes)
        if not p:  # no lockable memory left: a pageable array works too (staged copy)
            return np.empty(shape, dtype=dtype)
        return np.asarray(PinnedPool._Block(self, p, nbytes, tuple(int(v) for v in shape), dtype.str))

    def _give_back(self, p, nbytes):
        cached = self._free.setdefault(nbytes, [])
        if self.closed or len(cached) >= self.MAX_CACHED_PER_SIZE:
            load().fw_host_free(p)
        else:
            cached.append(p)

    def close(self):
        self.closed = True
        for cached in self._free.values():
            for p in cached:
                load().fw_host_free(p)
        self._free.clear()


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
