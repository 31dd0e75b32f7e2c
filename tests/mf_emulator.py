"""numpy emulation of the GPU multifrontal numeric phase (lu_numeric.cu), driven by the symbolic
structures that the C library's host analysis produces.  TEST INFRASTRUCTURE: validates the
analysis (fronts, extend-add maps, permutation) and the windowed-pivoting algorithm on CPU."""
import ctypes as C

import numpy as np

from femwell_b200 import _lib as L

PW, HWIN = 32, 1 << 30  # the GPU pivots over every fully-summed row of the front


def lu_symbolic(n, edofs, cx, cy, active, leaf):
    lib = L.load()
    nb, ne = edofs.shape
    ed = L.as_i32(edofs)
    cx, cy = L.as_f64(cx), L.as_f64(cy)
    act = np.ascontiguousarray(active, dtype=np.uint8)
    sizes = np.zeros(8, dtype=np.int64)
    L.check(lib.fw_debug_lu_symbolic(n, nb, ne, L.ptr(ed), L.ptr(cx), L.ptr(cy), L.ptr(act), leaf, L.ptr(sizes)))
    na, nn, nl, nbl = (int(sizes[i]) for i in range(4))
    out = dict(n=n, n_active=na, n_nodes=nn, n_levels=nl, factor_entries=int(sizes[4]), max_m=int(sizes[6]),
               max_ns=int(sizes[7]))

    def get(which, size, dtype=np.int32):
        a = np.zeros(max(size, 1), dtype=dtype)
        L.check(lib.fw_debug_lu_symbolic_get(which, L.ptr(a)))
        return a[:size]

    for name, which, size in [("perm", 0, na), ("iperm", 1, n), ("node_of", 2, na), ("parent", 3, nn),
                              ("child0", 4, nn), ("child1", 5, nn), ("depth", 6, nn), ("m", 7, nn), ("ns", 8, nn),
                              ("own_start", 9, nn), ("blist", 10, nbl), ("rel", 11, nbl),
                              ("level_start", 12, nl + 1), ("level_nodes", 13, nn)]:
        out[name] = get(which, size)
    out["front_off"] = get(20, nn, np.int64)
    out["blist_off"] = get(21, nn, np.int64)
    return out


class Emulator:
    def __init__(self, S, OP, window=HWIN, pw=PW, thr=0.0):
        """OP: scipy CSR (n x n) in original dof numbering."""
        self.S, self.window, self.pw = S, window, pw
        dtype = OP.dtype
        nn = S["n_nodes"]
        self.F = [np.zeros((S["m"][v], S["m"][v]), dtype=dtype) for v in range(nn)]
        self.piv = [np.arange(S["ns"][v]) for v in range(nn)]
        self.perturbed = 0
        self.thr = thr
        self._scatter(OP.tocoo())
        self._factor()

    def _flist(self, v):
        S = self.S
        os_, ns, m = S["own_start"][v], S["ns"][v], S["m"][v]
        bo = S["blist_off"][v]
        return np.concatenate([np.arange(os_, os_ + ns), S["blist"][bo:bo + m - ns]])

    def _scatter(self, coo):
        S = self.S
        ip = S["iperm"]
        pr, pc = ip[coo.row], ip[coo.col]
        ok = (pr >= 0) & (pc >= 0)
        pr, pc, val = pr[ok], pc[ok], coo.data[ok]
        node = S["node_of"][np.minimum(pr, pc)]
        order = np.argsort(node, kind="stable")
        pr, pc, val, node = pr[order], pc[order], val[order], node[order]
        bounds = np.searchsorted(node, np.arange(S["n_nodes"] + 1))
        for v in range(S["n_nodes"]):
            a, b = bounds[v], bounds[v + 1]
            if a == b:
                continue
            fl = self._flist(v)
            li = np.searchsorted(fl, pr[a:b])
            lj = np.searchsorted(fl, pc[a:b])
            assert (fl[li] == pr[a:b]).all() and (fl[lj] == pc[a:b]).all(), "entry outside its front"
            self.F[v][li, lj] = val[a:b]

    def _factor_front(self, v):
        S = self.S
        F = self.F[v]
        ns, m = S["ns"][v], S["m"][v]
        piv = self.piv[v]
        for k0 in range(0, ns, self.pw):
            w = min(self.pw, ns - k0)
            hw = min(self.window, ns - k0)
            for j in range(w):
                col = F[k0 + j:k0 + hw, k0 + j]
                mag = np.abs(col.real) + np.abs(col.imag) if np.iscomplexobj(col) else np.abs(col)
                bi = int(np.argmax(mag))
                if not (mag[bi] >= self.thr) or mag[bi] == 0:
                    F[k0 + j + bi, k0 + j] = self.thr if self.thr > 0 else 1e-300
                    self.perturbed += 1
                p = k0 + j + bi
                piv[k0 + j] = p
                if p != k0 + j:
                    F[[k0 + j, p], :] = F[[p, k0 + j], :]
                # panel columns only get the rank-1 update now (rows below, all of them)
                F[k0 + j + 1:, k0 + j] /= F[k0 + j, k0 + j]
                F[k0 + j + 1:, k0 + j + 1:k0 + w] -= np.outer(F[k0 + j + 1:, k0 + j], F[k0 + j, k0 + j + 1:k0 + w])
            # U row block
            Lkk = np.tril(F[k0:k0 + w, k0:k0 + w], -1) + np.eye(w)
            F[k0:k0 + w, k0 + w:] = np.linalg.solve(Lkk, F[k0:k0 + w, k0 + w:])
            F[k0 + w:, k0 + w:] -= F[k0 + w:, k0:k0 + w] @ F[k0:k0 + w, k0 + w:]

    def _factor(self):
        S = self.S
        for lev in range(S["n_levels"] - 1, -1, -1):
            for v in S["level_nodes"][S["level_start"][lev]:S["level_start"][lev + 1]]:
                for c in (S["child0"][v], S["child1"][v]):
                    if c < 0:
                        continue
                    nsc, mc = S["ns"][c], S["m"][c]
                    bo = S["blist_off"][c]
                    rel = S["rel"][bo:bo + mc - nsc]
                    self.F[v][np.ix_(rel, rel)] += self.F[c][nsc:, nsc:]
                self._factor_front(v)

    def solve(self, b):
        S = self.S
        xp = b[S["perm"]].astype(np.result_type(b.dtype, self.F[0].dtype))
        work = [None] * S["n_nodes"]
        for lev in range(S["n_levels"] - 1, -1, -1):
            for v in S["level_nodes"][S["level_start"][lev]:S["level_start"][lev + 1]]:
                ns, m, os_ = S["ns"][v], S["m"][v], S["own_start"][v]
                w = np.zeros(m, dtype=xp.dtype)
                w[:ns] = xp[os_:os_ + ns]
                for c in (S["child0"][v], S["child1"][v]):
                    if c < 0:
                        continue
                    nsc, mc = S["ns"][c], S["m"][c]
                    bo = S["blist_off"][c]
                    rel = S["rel"][bo:bo + mc - nsc]
                    w[rel] += work[c][nsc:]
                F = self.F[v]
                for j in range(ns):
                    p = self.piv[v][j]
                    if p != j:
                        w[[j, p]] = w[[p, j]]
                L11 = np.tril(F[:ns, :ns], -1) + np.eye(ns)
                w[:ns] = np.linalg.solve(L11, w[:ns]) if ns else w[:ns]
                w[ns:] -= F[ns:, :ns] @ w[:ns]
                work[v] = w
        for lev in range(S["n_levels"]):
            for v in S["level_nodes"][S["level_start"][lev]:S["level_start"][lev + 1]]:
                ns, m, os_ = S["ns"][v], S["m"][v], S["own_start"][v]
                w = work[v]
                p = S["parent"][v]
                if p >= 0:
                    bo = S["blist_off"][v]
                    rel = S["rel"][bo:bo + m - ns]
                    w[ns:] = work[p][rel]
                F = self.F[v]
                if ns:
                    w[:ns] = np.linalg.solve(np.triu(F[:ns, :ns]), w[:ns] - F[:ns, ns:] @ w[ns:])
                xp[os_:os_ + ns] = w[:ns]
        x = np.zeros(S["n"], dtype=xp.dtype)
        x[S["perm"]] = xp
        return x
