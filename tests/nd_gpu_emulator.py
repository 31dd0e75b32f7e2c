"""numpy emulation of the DEVICE nested-dissection analysis (femwell_b200/csrc/lu_symbolic_gpu.cu), step by step as
the kernels do it.  TEST INFRASTRUCTURE: validates the algorithm against the host analysis (lu_symbolic.cpp) on CPU;
the GPU suite then compares the device arrays with the host arrays bit by bit."""
import numpy as np


def tree_topology(ne, leaf):
    """geometry-independent part (host in the product too): breadth-first node ids like lu_symbolic.cpp"""
    parent, c0, c1, depth, lo, hi = [-1], [-1], [-1], [0], [0], [ne]
    level = [0]
    while level:
        nxt = []
        for v in level:
            if hi[v] - lo[v] <= leaf:
                continue
            mid = (lo[v] + hi[v]) // 2
            for a, b in ((lo[v], mid), (mid, hi[v])):
                parent.append(v), c0.append(-1), c1.append(-1), depth.append(depth[v] + 1), lo.append(a), hi.append(b)
                nxt.append(len(parent) - 1)
            c0[v], c1[v] = nxt[-2], nxt[-1]
        level = nxt
    arr = lambda x: np.asarray(x, dtype=np.int64)
    return dict(parent=arr(parent), child0=arr(c0), child1=arr(c1), depth=arr(depth), lo=arr(lo), hi=arr(hi))


def postorder(T):
    post, st = [], [(0, 0)]
    while st:
        v, s = st.pop()
        if T["child0"][v] < 0:
            post.append(v)
        elif s == 0:
            st.append((v, 1)), st.append((T["child0"][v], 0))
        elif s == 1:
            st.append((v, 2)), st.append((T["child1"][v], 0))
        else:
            post.append(v)
    return np.asarray(post, dtype=np.int64)


def node_at(T, pos, max_depth):
    """descend from the root: deepest node of depth <= max_depth whose range holds the position(s)"""
    pos = np.atleast_1d(pos)
    v = np.zeros(pos.shape, dtype=np.int64)
    for _ in range(max_depth):
        c0 = T["child0"][v]
        go = (c0 >= 0) & (T["depth"][v] < max_depth)
        mid = (T["lo"][v] + T["hi"][v]) // 2
        nv = np.where(pos < mid, c0, T["child1"][v])
        v = np.where(go, nv, v)
    return v


def analyse(n, edofs, cx, cy, active, leaf):
    nb, ne = edofs.shape
    T = tree_topology(ne, leaf)
    nn = T["parent"].size
    maxd = int(T["depth"].max())
    # ranks along x and y (stable sort: ties by element index)
    rx = np.empty(ne, dtype=np.int64)
    ry = np.empty(ne, dtype=np.int64)
    rx[np.argsort(cx, kind="stable")] = np.arange(ne)
    ry[np.argsort(cy, kind="stable")] = np.arange(ne)
    eperm = np.arange(ne, dtype=np.int64)
    pos = np.arange(ne, dtype=np.int64)
    for L in range(maxd):
        v = node_at(T, pos, L)
        x0 = np.full(nn, np.inf), np.full(nn, -np.inf)
        y0 = np.full(nn, np.inf), np.full(nn, -np.inf)
        np.minimum.at(x0[0], v, cx[eperm]), np.maximum.at(x0[1], v, cx[eperm])
        np.minimum.at(y0[0], v, cy[eperm]), np.maximum.at(y0[1], v, cy[eperm])
        use_x = (x0[1] - x0[0]) >= (y0[1] - y0[0])
        splits = (T["child0"][v] >= 0) & (T["depth"][v] == L)
        sub = np.where(splits, np.where(use_x[v], rx[eperm], ry[eperm]), pos - T["lo"][v])
        key = T["lo"][v] * (1 << 32) + sub
        eperm = eperm[np.argsort(key, kind="stable")]
    epos = np.empty(ne, dtype=np.int64)
    epos[eperm] = np.arange(ne)
    leaf_of = node_at(T, epos, maxd)  # per element
    # owners: deepest node containing [minpos, maxpos] of the dof's elements
    minpos = np.full(n, ne, dtype=np.int64)
    maxpos = np.full(n, -1, dtype=np.int64)
    for l in range(nb):
        np.minimum.at(minpos, edofs[l], epos)
        np.maximum.at(maxpos, edofs[l], epos)
    act = (np.asarray(active) != 0) & (maxpos >= 0)
    owner = np.zeros(n, dtype=np.int64)
    for _ in range(maxd):
        c0 = T["child0"][owner]
        mid = (T["lo"][owner] + T["hi"][owner]) // 2
        left, right = maxpos < mid, minpos >= mid
        nv = np.where(left, c0, np.where(right, T["child1"][owner], owner))
        owner = np.where(c0 >= 0, nv, owner)
    owner[~act] = -1
    post = postorder(T)
    prank = np.empty(nn, dtype=np.int64)
    prank[post] = np.arange(nn)
    ns = np.bincount(owner[act], minlength=nn)
    own_start = np.zeros(nn, dtype=np.int64)
    own_start[post] = np.concatenate([[0], np.cumsum(ns[post])[:-1]])
    key = np.where(act, prank[np.maximum(owner, 0)], nn) * (1 << 32) + np.arange(n)
    order = np.argsort(key, kind="stable")
    n_active = int(act.sum())
    perm = order[:n_active]
    iperm = np.full(n, -1, dtype=np.int64)
    iperm[perm] = np.arange(n_active)
    node_of = owner[perm]
    # boundary lists: every (dof, element) pair emits (v, new index) for the nodes v on the path leaf(e) .. below owner
    keys = []
    for l in range(nb):
        d = edofs[l]
        ok = act[d]
        v = leaf_of.copy()
        o = owner[d]
        live = ok & (v != o)
        while live.any():
            keys.append(v[live] * (1 << 32) + iperm[d[live]])
            v = np.where(live, T["parent"][v], v)
            live = live & (v != o)
    keys = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.int64)
    bnode, blist = keys >> 32, keys & ((1 << 32) - 1)
    cnt = np.bincount(bnode, minlength=nn)
    blist_off = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    m = ns + cnt
    rel = np.full(blist.size, -1, dtype=np.int64)
    for i in range(blist.size):
        p = T["parent"][bnode[i]]
        x = blist[i]
        if x < own_start[p] + ns[p]:
            rel[i] = x - own_start[p]
        else:
            seg = blist[blist_off[p]:blist_off[p] + cnt[p]]
            k = np.searchsorted(seg, x)
            assert seg[k] == x
            rel[i] = ns[p] + k
    return dict(T=T, eperm=eperm, perm=perm, iperm=iperm, node_of=node_of, ns=ns, m=m, own_start=own_start,
                blist=blist, blist_off=blist_off, rel=rel, n_active=n_active)
