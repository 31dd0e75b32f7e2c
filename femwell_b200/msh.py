"""Native reader/writer for gmsh ``.msh`` files (ASCII and binary, format 2.2 and 4.1) <-> :class:`MeshTri`.

Stands in for the ``meshio.read(...)`` + ``skfem.io.meshio.from_meshio`` step that precedes the hot path in
every reference example (``femwell/mesh/mesh.py:440-453`` writes a ``.msh`` with gmsh and reads it back with
meshio; ``docs/photonics/examples/waveguide_modes.py:69-72`` fills epsilon from ``mesh.subdomains``).
SURVEY.md section 8f row 4.  Host-side bookkeeping only (integer/topology work, like scikit-fem's mesh
constructor): the element order of the file is kept, z is dropped, 2-D physical groups become
``mesh.subdomains[name]`` (element indices) and 1-D physical groups ``mesh.boundaries[name]`` (facet indices).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .mesh import MeshTri

_LINE, _TRI, _POINT = 1, 2, 15
_NODES_OF_TYPE = {1: 2, 2: 3, 3: 4, 4: 4, 8: 3, 9: 6, 15: 1}


def _sections(text: str) -> Dict[str, List[str]]:
    out: Dict[str, List[str]] = {}
    name, buf = None, []
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("$End"):
            out[name] = buf
            name, buf = None, []
        elif line.startswith("$"):
            name, buf = line[1:], []
        elif name is not None:
            buf.append(line)
    return out


def _physical_names(sec) -> Dict[Tuple[int, int], str]:
    names = {}
    if sec:
        for line in sec[1:]:
            dim, tag, name = line.split(maxsplit=2)
            names[(int(dim), int(tag))] = name.strip().strip('"')
    return names


def _read_v2(sec, names):
    nodes = sec["Nodes"]
    n = int(nodes[0])
    ids = np.empty(n, dtype=np.int64)
    xyz = np.empty((n, 3))
    for i, line in enumerate(nodes[1:1 + n]):
        parts = line.split()
        ids[i] = int(parts[0])
        xyz[i] = [float(v) for v in parts[1:4]]
    tris, tri_phys, lines, line_phys = [], [], [], []
    els = sec["Elements"]
    for line in els[1:1 + int(els[0])]:
        parts = [int(v) for v in line.split()]
        etype, ntags = parts[1], parts[2]
        phys = parts[3] if ntags > 0 else 0
        conn = parts[3 + ntags:]
        if etype == _TRI:
            tris.append(conn[:3])
            tri_phys.append(phys)
        elif etype == _LINE:
            lines.append(conn[:2])
            line_phys.append(phys)
    return ids, xyz, tris, [[p] if p else [] for p in tri_phys], lines, [[p] if p else [] for p in line_phys]


def _read_v4(sec, names):
    # entity -> physical tags
    ent_phys: Dict[Tuple[int, int], List[int]] = {}
    ent = sec.get("Entities")
    if ent:
        counts = [int(v) for v in ent[0].split()]
        row = 1
        for dim, cnt in enumerate(counts):
            for _ in range(cnt):
                parts = ent[row].split()
                row += 1
                tag = int(parts[0])
                off = 4 if dim == 0 else 7
                nphys = int(parts[off])
                ent_phys[(dim, tag)] = [abs(int(v)) for v in parts[off + 1:off + 1 + nphys]]
    nodes = sec["Nodes"]
    nblocks, ntotal = [int(v) for v in nodes[0].split()[:2]]
    ids = np.empty(ntotal, dtype=np.int64)
    xyz = np.empty((ntotal, 3))
    row, k = 1, 0
    for _ in range(nblocks):
        _dim, _tag, parametric, nb = [int(v) for v in nodes[row].split()]
        row += 1
        for i in range(nb):
            ids[k + i] = int(nodes[row + i])
        row += nb
        for i in range(nb):
            xyz[k + i] = [float(v) for v in nodes[row + i].split()[:3]]
        row += nb
        k += nb
    tris, tri_phys, lines, line_phys = [], [], [], []
    els = sec["Elements"]
    nblocks = int(els[0].split()[0])
    row = 1
    for _ in range(nblocks):
        dim, tag, etype, nb = [int(v) for v in els[row].split()]
        row += 1
        phys = ent_phys.get((dim, tag), [])
        for i in range(nb):
            parts = [int(v) for v in els[row + i].split()]
            if etype == _TRI:
                tris.append(parts[1:4])
                tri_phys.append(phys)
            elif etype == _LINE:
                lines.append(parts[1:3])
                line_phys.append(phys)
        row += nb
    return ids, xyz, tris, tri_phys, lines, line_phys


class _Cursor:
    """byte cursor over a binary .msh file: ASCII lines and packed little/big-endian records"""

    def __init__(self, data: bytes):
        self.d, self.i, self.e = data, 0, "<"

    def line(self) -> str:
        j = self.d.index(b"\n", self.i)
        out = self.d[self.i:j].decode("ascii", "replace").strip()
        self.i = j + 1
        return out

    def skip_ws(self):
        while self.i < len(self.d) and self.d[self.i:self.i + 1] in b" \r\n\t":
            self.i += 1

    def arr(self, dtype, count):
        dt = np.dtype(dtype).newbyteorder(self.e)
        out = np.frombuffer(self.d, dtype=dt, count=count, offset=self.i)
        self.i += dt.itemsize * count
        return out

    def find(self, marker: str) -> bool:
        j = self.d.find(("$" + marker).encode(), 0)
        if j < 0:
            return False
        self.i = j
        self.line()
        return True


def _read_binary(data: bytes, version: float):
    """binary 2.2 / 4.1: returns the same tuple as the ASCII readers plus the physical names"""
    c = _Cursor(data)
    c.find("MeshFormat")
    fmt = c.line().split()
    size_t = np.uint64 if int(fmt[2]) == 8 else np.uint32
    one = c.arr(np.int32, 1)[0]
    if one != 1:
        c.e = ">"
        if int(np.frombuffer(data, dtype=">i4", count=1, offset=c.i - 4)[0]) != 1:
            raise ValueError("binary .msh: cannot determine the byte order")
    names = {}
    if c.find("PhysicalNames"):
        for _ in range(int(c.line())):
            dim, tag, name = c.line().split(maxsplit=2)
            names[(int(dim), int(tag))] = name.strip().strip('"')
    tris, tri_phys, lines, line_phys = [], [], [], []
    if version < 4:
        c.find("Nodes")
        n = int(c.line())
        rec = c.arr(np.dtype([("id", "i4"), ("x", "f8", 3)]), n)
        ids, xyz = rec["id"].astype(np.int64), np.array(rec["x"], dtype=np.float64)
        c.find("Elements")
        total, done = int(c.line()), 0
        while done < total:
            etype, nfollow, ntags = (int(v) for v in c.arr(np.int32, 3))
            nn = _NODES_OF_TYPE.get(etype)
            if nn is None:
                raise ValueError(f"binary .msh: unsupported element type {etype}")
            blk = c.arr(np.int32, nfollow * (1 + ntags + nn)).reshape(nfollow, 1 + ntags + nn)
            phys = blk[:, 1] if ntags > 0 else np.zeros(nfollow, dtype=np.int32)
            conn = blk[:, 1 + ntags:]
            if etype == _TRI:
                tris += conn[:, :3].tolist()
                tri_phys += [[int(q)] if q else [] for q in phys]
            elif etype == _LINE:
                lines += conn[:, :2].tolist()
                line_phys += [[int(q)] if q else [] for q in phys]
            done += nfollow
        return names, ids, xyz, tris, tri_phys, lines, line_phys
    ent_phys: Dict[Tuple[int, int], List[int]] = {}
    if c.find("Entities"):
        counts = [int(v) for v in c.arr(size_t, 4)]
        for dim, cnt in enumerate(counts):
            for _ in range(cnt):
                tag = int(c.arr(np.int32, 1)[0])
                c.arr(np.float64, 3 if dim == 0 else 6)
                nphys = int(c.arr(size_t, 1)[0])
                ent_phys[(dim, tag)] = [abs(int(v)) for v in c.arr(np.int32, nphys)]
                if dim > 0:
                    c.arr(np.int32, int(c.arr(size_t, 1)[0]))
    c.find("Nodes")
    nblocks, ntotal = (int(v) for v in c.arr(size_t, 4)[:2])
    ids = np.empty(ntotal, dtype=np.int64)
    xyz = np.empty((ntotal, 3))
    k = 0
    for _ in range(nblocks):
        dim, _tag, parametric = (int(v) for v in c.arr(np.int32, 3))
        nb = int(c.arr(size_t, 1)[0])
        ids[k:k + nb] = c.arr(size_t, nb).astype(np.int64)
        width = 3 + (dim if parametric else 0)
        xyz[k:k + nb] = c.arr(np.float64, nb * width).reshape(nb, width)[:, :3]
        k += nb
    c.find("Elements")
    nblocks = int(c.arr(size_t, 4)[0])
    for _ in range(nblocks):
        dim, tag, etype = (int(v) for v in c.arr(np.int32, 3))
        nb = int(c.arr(size_t, 1)[0])
        nn = _NODES_OF_TYPE.get(etype)
        if nn is None:
            raise ValueError(f"binary .msh: unsupported element type {etype}")
        blk = c.arr(size_t, nb * (1 + nn)).reshape(nb, 1 + nn).astype(np.int64)
        phys = ent_phys.get((dim, tag), [])
        if etype == _TRI:
            tris += blk[:, 1:4].tolist()
            tri_phys += [phys] * nb
        elif etype == _LINE:
            lines += blk[:, 1:3].tolist()
            line_phys += [phys] * nb
    return names, ids, xyz, tris, tri_phys, lines, line_phys


def read_msh(path: str) -> MeshTri:
    """Triangles + named physical groups of a gmsh ``.msh`` file (2.2 or 4.1, ASCII or binary) as a :class:`MeshTri`."""
    with open(path, "rb") as f:
        data = f.read()
    head = data[:256].decode("ascii", "replace")
    if not head.lstrip().startswith("$MeshFormat"):
        raise ValueError(f"{path}: not a gmsh .msh file")
    fmt = head.split("\n", 2)[1].split()
    version, binary = float(fmt[0]), int(fmt[1])
    if binary:
        names, ids, xyz, tris, tri_phys, lines, line_phys = _read_binary(data, version)
        return _mesh_from_records(path, names, ids, xyz, tris, tri_phys, lines, line_phys)
    sec = _sections(data.decode("ascii", "replace"))
    names = _physical_names(sec.get("PhysicalNames"))
    if version < 4:
        ids, xyz, tris, tri_phys, lines, line_phys = _read_v2(sec, names)
    else:
        ids, xyz, tris, tri_phys, lines, line_phys = _read_v4(sec, names)
    return _mesh_from_records(path, names, ids, xyz, tris, tri_phys, lines, line_phys)


def _mesh_from_records(path, names, ids, xyz, tris, tri_phys, lines, line_phys) -> MeshTri:
    if not tris:
        raise ValueError(f"{path}: no 3-node triangles found")
    # compact vertex numbering in file order (only vertices used by triangles, like from_meshio's prune)
    t_file = np.asarray(tris, dtype=np.int64).T  # (3, Ne) node tags
    lookup = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    lookup[ids] = np.arange(ids.size)
    used = np.zeros(ids.size, dtype=bool)
    used[lookup[t_file]] = True
    new_index = np.cumsum(used) - 1
    p = np.ascontiguousarray(xyz[used, :2].T)
    t = new_index[lookup[t_file]]
    mesh = MeshTri(p, t)
    # sub-domains
    subdomains: Dict[str, List[int]] = {}
    for e, phys in enumerate(tri_phys):
        for tag in phys:
            subdomains.setdefault(names.get((2, tag), f"gmsh:physical_{tag}"), []).append(e)
    mesh.subdomains = {k: np.asarray(v, dtype=np.int32) for k, v in subdomains.items()}
    # boundaries: line elements matched to facets by their sorted vertex pair
    if lines:
        nv = mesh.nvertices
        key_of_facet = mesh.facets[0].astype(np.int64) * nv + mesh.facets[1].astype(np.int64)
        order = np.argsort(key_of_facet)
        ln = np.asarray(lines, dtype=np.int64)
        ok = used[lookup[ln]].all(axis=1)
        a = new_index[lookup[ln]]
        lo, hi = np.minimum(a[:, 0], a[:, 1]), np.maximum(a[:, 0], a[:, 1])
        keys = lo * nv + hi
        pos = np.searchsorted(key_of_facet[order], keys)
        pos = np.clip(pos, 0, order.size - 1)
        found = ok & (key_of_facet[order][pos] == keys)
        facet = order[pos]
        boundaries: Dict[str, List[int]] = {}
        for i, phys in enumerate(line_phys):
            if not found[i]:
                continue
            for tag in phys:
                boundaries.setdefault(names.get((1, tag), f"gmsh:physical_{tag}"), []).append(int(facet[i]))
        mesh.boundaries = {k: np.unique(np.asarray(v, dtype=np.int32)) for k, v in boundaries.items()}
    return mesh


def write_msh(mesh: MeshTri, path: str, version: str = "2.2", binary: bool = False) -> None:
    """Write a :class:`MeshTri` with its named sub-domains / boundaries as gmsh 2.2 or 4.1, ASCII or binary (element
    order kept within a sub-domain; 2.2 keeps the global element order; an element in several sub-domains is written
    with the first one)."""
    sub = list(mesh.subdomains.items())
    bnd = list(mesh.boundaries.items())
    tri_tag = np.zeros(mesh.nelements, dtype=np.int64)
    for k, (_, idx) in reversed(list(enumerate(sub))):
        tri_tag[np.asarray(idx, dtype=np.int64)] = 100 + k
    if version not in ("2.2", "4.1"):
        raise ValueError("version must be '2.2' or '4.1'")
    nv, ne = mesh.nvertices, mesh.nelements
    with open(path, "wb") as f:
        w = lambda txt: f.write(txt.encode("ascii"))  # noqa: E731
        w("$MeshFormat\n%s %d 8\n" % (version, 1 if binary else 0))
        if binary:
            f.write(np.array([1], dtype="<i4").tobytes())
            w("\n")
        w("$EndMeshFormat\n$PhysicalNames\n%d\n" % (len(sub) + len(bnd)))
        for k, (name, _) in enumerate(bnd):
            w('1 %d "%s"\n' % (1000 + k, name))
        for k, (name, _) in enumerate(sub):
            w('2 %d "%s"\n' % (100 + k, name))
        w("$EndPhysicalNames\n")
        xyz = np.zeros((nv, 3))
        xyz[:, :2] = mesh.p.T
        if version == "2.2":
            w("$Nodes\n%d\n" % nv)
            if binary:
                rec = np.empty(nv, dtype=np.dtype([("id", "<i4"), ("x", "<f8", 3)]))
                rec["id"], rec["x"] = np.arange(1, nv + 1), xyz
                f.write(rec.tobytes())
                w("\n")
            else:
                for i in range(nv):
                    w("%d %.17g %.17g 0\n" % (i + 1, mesh.p[0, i], mesh.p[1, i]))
            nlines = sum(len(idx) for _, idx in bnd)
            w("$EndNodes\n$Elements\n%d\n" % (nlines + ne))
            eid = 1
            for k, (_, idx) in enumerate(bnd):
                fi = np.asarray(idx, dtype=np.int64)
                if binary and fi.size:
                    blk = np.empty((fi.size, 5), dtype="<i4")
                    blk[:, 0] = eid + np.arange(fi.size)
                    blk[:, 1] = blk[:, 2] = 1000 + k
                    blk[:, 3:] = mesh.facets[:, fi].T + 1
                    f.write(np.array([1, fi.size, 2], dtype="<i4").tobytes() + blk.tobytes())
                elif not binary:
                    for j in fi:
                        a, b = mesh.facets[:, j]
                        w("%d 1 2 %d %d %d %d\n" % (eid + 0, 1000 + k, 1000 + k, a + 1, b + 1))
                        eid += 1
                    continue
                eid += fi.size
            if binary:
                # runs of equal tag keep the global element order
                start = 0
                while start < ne:
                    stop = start + 1
                    while stop < ne and tri_tag[stop] == tri_tag[start]:
                        stop += 1
                    cnt, tag = stop - start, int(tri_tag[start])
                    ntags = 2 if tag else 0
                    blk = np.empty((cnt, 1 + ntags + 3), dtype="<i4")
                    blk[:, 0] = eid + np.arange(cnt)
                    if tag:
                        blk[:, 1] = blk[:, 2] = tag
                    blk[:, 1 + ntags:] = mesh.t[:, start:stop].T + 1
                    f.write(np.array([2, cnt, ntags], dtype="<i4").tobytes() + blk.tobytes())
                    eid += cnt
                    start = stop
                w("\n")
            else:
                for e in range(ne):
                    a, b, c = mesh.t[:, e]
                    if tri_tag[e]:
                        w("%d 2 2 %d %d %d %d %d\n" % (eid, tri_tag[e], tri_tag[e], a + 1, b + 1, c + 1))
                    else:
                        w("%d 2 0 %d %d %d\n" % (eid, a + 1, b + 1, c + 1))
                    eid += 1
            w("$EndElements\n")
            return
        # ---- 4.1: one curve entity per named boundary, one surface entity per sub-domain (+ one for the rest)
        surf = [(100 + k, [100 + k], np.nonzero(tri_tag == 100 + k)[0]) for k in range(len(sub))]
        rest = np.nonzero(tri_tag == 0)[0]
        if rest.size:
            surf.append((99, [], rest))
        surf = [sf for sf in surf if sf[2].size]
        curves = [(1000 + k, [1000 + k], np.asarray(idx, dtype=np.int64)) for k, (_, idx) in enumerate(bnd) if len(idx)]
        lo, hi = xyz.min(axis=0), xyz.max(axis=0)
        w("$Entities\n")
        if binary:
            f.write(np.array([0, len(curves), len(surf), 0], dtype="<u8").tobytes())
            for tag, phys, _ in curves + surf:
                f.write(np.array([tag], dtype="<i4").tobytes() + np.concatenate([lo, hi]).astype("<f8").tobytes()
                        + np.array([len(phys)], dtype="<u8").tobytes() + np.array(phys, dtype="<i4").tobytes()
                        + np.array([0], dtype="<u8").tobytes())
            w("\n")
        else:
            w("0 %d %d 0\n" % (len(curves), len(surf)))
            for tag, phys, _ in curves + surf:
                w("%d %.17g %.17g %.17g %.17g %.17g %.17g %d %s0\n" % (tag, *lo, *hi, len(phys),
                                                                      "".join("%d " % q for q in phys)))
        w("$EndEntities\n$Nodes\n")
        if binary:
            f.write(np.array([1, nv, 1, nv], dtype="<u8").tobytes() + np.array([2, surf[0][0], 0], dtype="<i4").tobytes()
                    + np.array([nv], dtype="<u8").tobytes() + np.arange(1, nv + 1, dtype="<u8").tobytes()
                    + xyz.astype("<f8").tobytes())
            w("\n")
        else:
            w("1 %d 1 %d\n2 %d 0 %d\n" % (nv, nv, surf[0][0], nv))
            for i in range(nv):
                w("%d\n" % (i + 1))
            for i in range(nv):
                w("%.17g %.17g 0\n" % (mesh.p[0, i], mesh.p[1, i]))
        total = sum(c[2].size for c in curves) + ne
        w("$EndNodes\n$Elements\n")
        if binary:
            f.write(np.array([len(curves) + len(surf), total, 1, total], dtype="<u8").tobytes())
        else:
            w("%d %d 1 %d\n" % (len(curves) + len(surf), total, total))
        eid = 1
        for dim, etype, ents in ((1, 1, curves), (2, 2, surf)):
            for tag, _phys, idx in ents:
                conn = (mesh.facets[:, idx] if dim == 1 else mesh.t[:, idx]).T.astype(np.int64) + 1
                blk = np.concatenate([(eid + np.arange(idx.size))[:, None], conn], axis=1)
                if binary:
                    f.write(np.array([dim, tag, etype], dtype="<i4").tobytes() + np.array([idx.size], dtype="<u8").tobytes()
                            + blk.astype("<u8").tobytes())
                else:
                    w("%d %d %d %d\n" % (dim, tag, etype, idx.size))
                    for r in blk:
                        w(" ".join(str(int(v)) for v in r) + "\n")
                eid += idx.size
        if binary:
            w("\n")
        w("$EndElements\n")
