"""Host-side owner of one library handle (= one GPU, one stream, one resident FE space).

Thin: marshals numpy arrays across the C ABI (include/femwell_b200.h); no arithmetic here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from .basis import Basis
from .element import eps_kind_of

MAT_A, MAT_B, MAT_MASS, MAT_C0, MAT_C1 = 0, 1, 2, 3, 4
FUN_OVERLAP, FUN_EX2_EY2, FUN_EX2_EY2_EALL, FUN_COUPLING, FUN_E2_E4, FUN_CONFINEMENT = range(6)


class Context:
    def __init__(self, device: int = 0):
        self.lib = L.load()
        self._h = C.c_void_p()
        L.check(self.lib.fw_create(C.byref(self._h), C.c_int(device)))
        self.device = device
        self.basis: Optional[Basis] = None
        self._space_key = None
        self._subset_key = None
        self.n = 0
        self.nnz_structural = None
        self._pinned = L.PinnedPool()

    def close(self):
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.fw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- space / symbolic
    @staticmethod
    def _space_struct(basis: Basis):
        """fw_space of a mode basis + the arrays that must stay alive while it is used"""
        tb = basis.tables()
        m = basis.mesh
        ed = basis.element_dofs
        s = L.FwSpace()
        s.order = basis.elem.order
        s.n_vertices, s.n_elements, s.n_dofs = m.nvertices, m.nelements, basis.N
        s.nbfun, s.nqp = tb.nbfun, tb.nqp
        keep = [L.as_f64(m.p), L.as_i32(m.t), L.as_i32(ed), L.as_i32(tb.kind), L.as_f64(tb.vec), L.as_f64(tb.sc),
                L.as_f64(tb.X), L.as_f64(tb.W)]
        (s.p, s.t, s.element_dofs, s.kind, s.vec, s.sc, s.X, s.W) = [a.ctypes.data for a in keep]
        return s, keep, tb

    def set_space(self, basis: Basis, solve_hint=None):
        """Upload mesh + FE space (a1).  ``basis`` must carry the mixed N x P element.
        ``solve_hint = (dirichlet_dofs or None, leaf_elements)`` announces the eigen-solve that follows, so that the host
        analysis of its sparse LU already runs while the mesh is uploaded."""
        s, keep, tb = self._space_struct(basis)
        m = basis.mesh
        if solve_hint is not None:
            d, leaf = solve_hint
            d = None if d is None else L.as_i32(d)
            L.check(self.lib.fw_analysis_hint(self._h, 0 if d is None else d.size, L.ptr(d), int(leaf or 0)))
        L.check(self.lib.fw_set_space(self._h, C.byref(s)))
        self.basis = basis
        self.n = basis.N
        self.nnz_structural = None
        self._space_key = self._key_of(basis)
        self._subset_key = None
        self.set_element_subset(basis.tind)

    @staticmethod
    def _key_of(basis: Basis):
        tb = basis.tables()
        return (id(basis.mesh), basis.elem.order, tb.X.tobytes(), tb.W.tobytes(), basis.tables_id)

    def set_element_subset(self, tind):
        """Restrict the forms to the elements of the basis (scikit-fem ``CellBasis.tind``, waveguide.py:574);
        ``None`` or all elements: no restriction."""
        ne = self.basis.mesh.nelements
        if tind is not None:
            tind = L.as_i32(np.unique(np.asarray(tind)))
            if tind.size == ne:
                tind = None
        key = None if tind is None else tind.tobytes()
        if key == self._subset_key:
            return
        L.check(self.lib.fw_set_element_subset(self._h, -1 if tind is None else tind.size, L.ptr(tind)))
        self._subset_key = key

    def set_option(self, option: int, value: int):
        L.check(self.lib.fw_set_option(self._h, int(option), int(value)))

    def bench_spmv(self, which: int = MAT_A, x_is_complex: bool = False, reps: int = 20) -> float:
        """ms per device-resident y = M x (measurement only)."""
        ms = C.c_double()
        L.check(self.lib.fw_bench_spmv(self._h, which, int(x_is_complex), reps, C.byref(ms)))
        return ms.value

    def matrix_nnz(self, which: int, eliminate_zeros: bool = True) -> int:
        L.check(self.lib.fw_set_option(self._h, 0, int(eliminate_zeros)))
        nnz, cx = C.c_int64(), C.c_int32()
        L.check(self.lib.fw_matrix_nnz(self._h, which, C.byref(nnz), C.byref(cx)))
        return nnz.value

    def same_space(self, basis: Basis) -> bool:
        """Same mesh, order, quadrature and reference tables (the element subset is handled separately)."""
        return self._space_key is not None and self._space_key == self._key_of(basis)

    def symbolic(self) -> int:
        nnz = C.c_int64()
        L.check(self.lib.fw_symbolic(self._h, C.byref(nnz)))
        self.nnz_structural = nnz.value
        return nnz.value

    # ---------------------------------------------------------------- assembly
    def assemble(self, eps_kind: int, epsilon_r: np.ndarray, k0: float, mu_r: float = 1.0, radius: float = np.inf):
        cx = np.iscomplexobj(epsilon_r)
        eps = L.as_c128(epsilon_r) if cx else L.as_f64(epsilon_r)
        ne = self.basis.mesh.nelements
        expect = ne if eps_kind == 0 else 3 * ne
        if eps.size != expect:
            raise ValueError(f"epsilon_r has {eps.size} dofs, expected {expect}")
        L.check(self.lib.fw_assemble_waveguide(self._h, eps_kind, int(cx), L.ptr(eps), C.c_double(k0),
                                               C.c_double(mu_r), C.c_double(radius)))

    def matrix(self, which: int, eliminate_zeros: bool = True) -> sp.csr_matrix:
        L.check(self.lib.fw_set_option(self._h, 0, int(eliminate_zeros)))
        nnz, cx = C.c_int64(), C.c_int32()
        L.check(self.lib.fw_matrix_nnz(self._h, which, C.byref(nnz), C.byref(cx)))
        indptr = np.empty(self.n + 1, dtype=np.int32)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.complex128 if cx.value else np.float64)
        L.check(self.lib.fw_matrix_export(self._h, which, L.ptr(indptr), L.ptr(indices), L.ptr(data)))
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.n))

    def spmv(self, which: int, x: np.ndarray) -> np.ndarray:
        cx = np.iscomplexobj(x)
        xx = L.as_c128(x) if cx else L.as_f64(x)
        nnz, mcx = C.c_int64(), C.c_int32()
        L.check(self.lib.fw_matrix_nnz(self._h, which, C.byref(nnz), C.byref(mcx)))
        y = np.empty(self.n, dtype=np.complex128 if (cx or mcx.value) else np.float64)
        L.check(self.lib.fw_matrix_spmv(self._h, which, int(cx), L.ptr(xx), L.ptr(y)))
        return y

    def coo_to_csr(self, n_rows: int, rows, cols, data) -> sp.csr_matrix:
        cx = np.iscomplexobj(data)
        d = L.as_c128(data) if cx else L.as_f64(data)
        r, c = L.as_i32(rows), L.as_i32(cols)
        nnz = C.c_int64()
        L.check(self.lib.fw_coo_to_csr(self._h, n_rows, C.c_int64(d.size), L.ptr(r), L.ptr(c), L.ptr(d), int(cx),
                                       C.byref(nnz)))
        indptr = np.empty(n_rows + 1, dtype=np.int32)
        indices = np.empty(nnz.value, dtype=np.int32)
        out = np.empty(nnz.value, dtype=d.dtype)
        L.check(self.lib.fw_coo_to_csr_export(self._h, L.ptr(indptr), L.ptr(indices), L.ptr(out)))
        return sp.csr_matrix((out, indices, indptr), shape=(n_rows, n_rows))

    def debug_lu_analysis(self, dirichlet=None, leaf_elements=0) -> dict:
        """tests: run the sparse-LU analysis of the current space on the device and download its tables"""
        d = None if dirichlet is None else L.as_i32(dirichlet)
        sizes = np.zeros(8, dtype=np.int64)
        L.check(self.lib.fw_debug_lu_analysis(self._h, 0 if d is None else d.size, L.ptr(d), int(leaf_elements),
                                              L.ptr(sizes)))
        na, nn, nl, nbl = (int(x) for x in sizes[:4])
        out = dict(n_active=na, n_nodes=nn, n_levels=nl, factor_entries=int(sizes[4]), max_m=int(sizes[6]),
                   max_ns=int(sizes[7]))
        for name, which, size, dt in [("perm", 0, na, np.int32), ("iperm", 1, self.n, np.int32),
                                      ("node_of", 2, na, np.int32), ("m", 7, nn, np.int32), ("ns", 8, nn, np.int32),
                                      ("own_start", 9, nn, np.int32), ("blist", 10, nbl, np.int32),
                                      ("rel", 11, nbl, np.int32), ("blist_off", 21, nn, np.int64)]:
            a = np.zeros(max(size, 1), dtype=dt)
            L.check(self.lib.fw_debug_lu_analysis_get(self._h, which, L.ptr(a)))
            out[name] = a[:size]
        return out

    # ---------------------------------------------------------------- LU / eigs
    def lu_factor(self, sigma, dirichlet=None, leaf_elements=0):
        d = None if dirichlet is None else L.as_i32(dirichlet)
        L.check(self.lib.fw_lu_factor_shifted(self._h, C.c_double(np.real(sigma)), C.c_double(np.imag(sigma)),
                                              0 if d is None else d.size, L.ptr(d), leaf_elements))

    def lu_solve(self, b: np.ndarray, refine: int = 1, force_real: bool = False) -> np.ndarray:
        if force_real:  # real factors + real rhs: the single right-hand-side kernels
            bb = L.as_f64(b)
            x = np.empty_like(bb)
            L.check(self.lib.fw_lu_solve(self._h, 0, L.ptr(bb), L.ptr(x), refine))
            return x
        # complex across the ABI: the factors may be complex even for a real right-hand side
        bb = L.as_c128(b)
        x = np.empty_like(bb)
        L.check(self.lib.fw_lu_solve(self._h, 1, L.ptr(bb), L.ptr(x), refine))
        return x

    def _eigs_params(self, k, sigma, dirichlet, v0, ncv, tol, maxiter, refine, leaf_elements, keep):
        p = L.FwEigsParams()
        p.k, p.ncv, p.maxiter, p.tol = int(k), int(ncv), int(maxiter), float(tol)
        p.sigma_re, p.sigma_im = float(np.real(sigma)), float(np.imag(sigma))
        if dirichlet is not None and len(dirichlet):
            d = L.as_i32(dirichlet)
            keep.append(d)
            p.n_dirichlet, p.dirichlet = d.size, d.ctypes.data
        if v0 is not None:
            v = L.as_f64(np.real(v0))
            keep.append(v)
            p.v0 = v.ctypes.data
        p.refine, p.leaf_elements = int(refine), int(leaf_elements)
        return p

    def eigs(self, k, sigma, dirichlet=None, v0=None, ncv=0, tol=0.0, maxiter=0, refine=0, leaf_elements=0):
        keep = []
        p = self._eigs_params(k, sigma, dirichlet, v0, ncv, tol, maxiter, refine, leaf_elements, keep)
        lams = np.empty(k, dtype=np.complex128)
        X = np.empty((k, self.n), dtype=np.complex128)
        st = L.FwEigsStats()
        L.check(self.lib.fw_eigs(self._h, C.byref(p), L.ptr(lams), L.ptr(X), C.byref(st)),
                partial=lambda: (lams[:st.nconv].copy(), X[:st.nconv].T.copy()))
        return lams, X, st.as_dict()

    def eigs_csr(self, K, M, k, sigma, coords=None, v0=None, ncv=0, tol=0.0, maxiter=0, refine=0, leaf_elements=0):
        """K x = lam M x on caller-assembled scipy CSR matrices (EigenSolver protocol, femwell/solver.py:87-142)."""
        K = K.tocsr()
        M = M.tocsr()
        n = K.shape[0]
        cx = np.iscomplexobj(K.data) or np.iscomplexobj(M.data)
        conv = L.as_c128 if cx else L.as_f64
        kd, md = conv(K.data), conv(M.data)
        kp, ki, mp, mi = L.as_i32(K.indptr), L.as_i32(K.indices), L.as_i32(M.indptr), L.as_i32(M.indices)
        keep = []
        self.n = n
        self._space_key = None
        self.nnz_structural = None
        p = self._eigs_params(k, sigma, None, v0, ncv, tol, maxiter, refine, leaf_elements, keep)
        co = None if coords is None else L.as_f64(coords)
        lams = np.empty(k, dtype=np.complex128)
        X = np.empty((k, n), dtype=np.complex128)
        st = L.FwEigsStats()
        L.check(self.lib.fw_eigs_csr(self._h, n, L.ptr(kp), L.ptr(ki), L.ptr(kd), L.ptr(mp), L.ptr(mi), L.ptr(md),
                                     int(cx), L.ptr(co), C.byref(p), L.ptr(lams), L.ptr(X), C.byref(st)),
                partial=lambda: (lams[:st.nconv].copy(), X[:st.nconv].T.copy()))
        return lams, X, st.as_dict()

    # ---------------------------------------------------------------- post-processing
    def postprocess(self, lams, X, k0, omega, mu0):
        k = len(lams)
        lams = L.as_c128(lams)
        X = L.as_c128(X)
        E = np.empty((k, self.n), dtype=np.complex128)
        H = np.empty((k, self.n), dtype=np.complex128)
        powers = np.empty(k, dtype=np.complex128)
        L.check(self.lib.fw_postprocess_modes(self._h, k, L.ptr(lams), L.ptr(X), C.c_double(k0), C.c_double(omega),
                                              C.c_double(mu0), L.ptr(E), L.ptr(H), L.ptr(powers)))
        return E, H, powers

    def hfield(self, E, beta, omega, mu0):
        E = L.as_c128(E)
        H = np.empty(self.n, dtype=np.complex128)
        L.check(self.lib.fw_hfield(self._h, L.ptr(E), C.c_double(np.real(beta)), C.c_double(np.imag(beta)),
                                   C.c_double(omega), C.c_double(mu0), L.ptr(H)))
        return H

    def intensity(self, E, H):
        E, H = L.as_c128(E), L.as_c128(H)
        out = np.empty(3 * self.basis.mesh.nelements, dtype=np.float64)
        L.check(self.lib.fw_intensity(self._h, L.ptr(E), L.ptr(H), L.ptr(out)))
        return out

    def overlap_cross(self, basis_j: Basis, E_i, H_i, E_j, H_j, mode=0, elements=None):
        """Cross-mesh overlap / scalar product (waveguide.py:737-782); the current space is basis_i."""
        s, keep, _ = self._space_struct(basis_j)
        f = [None if a is None else L.as_c128(a) for a in (E_i, H_i, E_j, H_j)]
        sel = None if elements is None else L.as_i32(elements)
        out = np.zeros(2)
        L.check(self.lib.fw_overlap_cross(self._h, C.byref(s), L.ptr(f[0]), L.ptr(f[1]), L.ptr(f[2]), L.ptr(f[3]),
                                          int(mode), 0 if sel is None else sel.size, L.ptr(sel), L.ptr(out)))
        return complex(out[0], out[1])

    def poynting(self, E, H):
        """complex128 [n_elements * 3 * nl], dof = e*3*nl + a*3 + c (waveguide.py:74-95)."""
        E, H = L.as_c128(E), L.as_c128(H)
        nl = 3 if self.basis.elem.order == 1 else 6
        out = np.empty(self.basis.mesh.nelements * 3 * nl, dtype=np.complex128)
        L.check(self.lib.fw_poynting(self._h, L.ptr(E), L.ptr(H), L.ptr(out)))
        return out

    def energy_current_density(self, xs):
        xs = L.as_c128(xs)
        out = np.empty(self.basis.mesh.nelements, dtype=np.complex128)
        L.check(self.lib.fw_energy_current_density(self._h, L.ptr(xs), L.ptr(out)))
        return out

    def error_estimator(self, E, eps_kind, epsilon_r):
        """eta_E [Ne] of Mode.eval_error_estimator (waveguide.py:473-502) for the current space."""
        b = self.basis
        m = b.mesh
        E = L.as_c128(E)
        cx = np.iscomplexobj(epsilon_r)
        eps = L.as_c128(epsilon_r) if cx else L.as_f64(epsilon_r)
        s, w, vec = b.facet_tables()
        t2f, f2t = L.as_i32(m.t2f), L.as_i32(m.f2t)
        out = np.empty(m.nelements, dtype=np.float64)
        L.check(self.lib.fw_error_estimator(self._h, L.ptr(E), int(eps_kind), int(cx), L.ptr(eps), m.nfacets,
                                            L.ptr(t2f), L.ptr(f2t), s.size, L.ptr(s), L.ptr(w), L.ptr(vec),
                                            L.ptr(out)))
        return out

    def functional(self, kind, fields, aux=None, aux_kind=0, elements=None, option=0):
        fs = [None if f is None else L.as_c128(f) for f in fields] + [None] * (4 - len(fields))
        a = None
        acx = 0
        if aux is not None:
            acx = int(np.iscomplexobj(aux))
            a = L.as_c128(aux) if acx else L.as_f64(aux)
        el = None if elements is None else L.as_i32(elements)
        out = np.zeros(3, dtype=np.complex128)
        L.check(self.lib.fw_functional(self._h, kind, L.ptr(fs[0]), L.ptr(fs[1]), L.ptr(fs[2]), L.ptr(fs[3]),
                                       aux_kind, acx, L.ptr(a), 0 if el is None else el.size, L.ptr(el), option,
                                       L.ptr(out)))
        return out

    def compute_modes(self, eps_kind, epsilon_r, wavelength, k, sigma, mu_r=1.0, radius=np.inf, dirichlet=None,
                      v0=None, ncv=0, tol=0.0, maxiter=0, refine=0, leaf_elements=0, reuse_symbolic=False):
        keep = []
        cx = np.iscomplexobj(epsilon_r)
        eps = L.as_c128(epsilon_r) if cx else L.as_f64(epsilon_r)
        p = L.FwModesParams()
        p.wavelength, p.mu_r, p.radius = float(wavelength), float(mu_r), float(radius)
        p.eps_kind, p.eps_is_complex, p.eps = int(eps_kind), int(cx), eps.ctypes.data
        p.eigs = self._eigs_params(k, sigma, dirichlet, v0, ncv, tol, maxiter, refine, leaf_elements, keep)
        p.reuse_symbolic = int(reuse_symbolic)
        lams = np.empty(k, dtype=np.complex128)
        E = self._pinned.empty((k, self.n), np.complex128)  # page-locked: the result download is one DMA each
        H = self._pinned.empty((k, self.n), np.complex128)
        powers = np.empty(k, dtype=np.complex128)
        st = L.FwEigsStats()
        L.check(self.lib.fw_compute_modes(self._h, C.byref(p), L.ptr(lams), L.ptr(E), L.ptr(H), L.ptr(powers),
                                          C.byref(st)),
                partial=lambda: (lams[:st.nconv].copy(), E[:st.nconv].T.copy()))
        return lams, E, H, powers, st.as_dict()

    def timings(self) -> dict:
        t = L.FwTimings()
        L.check(self.lib.fw_get_timings(self._h, C.byref(t)))
        return t.as_dict()
