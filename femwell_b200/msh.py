"""Native reader for gmsh ``.msh`` files (ASCII, format 2.2 and 4.1) -> :class:`MeshTri`.

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


def read_msh(path: str) -> MeshTri:
    """Triangles + named physical groups of a gmsh ASCII ``.msh`` file (2.2 or 4.1) as a :class:`MeshTri`."""
    with open(path, "r") as f:
        sec = _sections(f.read())
    if "MeshFormat" not in sec:
        raise ValueError(f"{path}: not a gmsh .msh file")
    fmt = sec["MeshFormat"][0].split()
    version, binary = float(fmt[0]), int(fmt[1])
    if binary != 0:
        raise NotImplementedError("binary .msh files are not supported (write with Mesh.Binary = 0)")
    names = _physical_names(sec.get("PhysicalNames"))
    if version < 4:
        ids, xyz, tris, tri_phys, lines, line_phys = _read_v2(sec, names)
    else:
        ids, xyz, tris, tri_phys, lines, line_phys = _read_v4(sec, names)
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


def write_msh(mesh: MeshTri, path: str) -> None:
    """Write a :class:`MeshTri` with its named sub-domains / boundaries as gmsh 2.2 ASCII (element order kept;
    an element in several sub-domains is written with the first one)."""
    sub = list(mesh.subdomains.items())
    bnd = list(mesh.boundaries.items())
    tri_tag = np.zeros(mesh.nelements, dtype=np.int64)
    for k, (_, idx) in reversed(list(enumerate(sub))):
        tri_tag[np.asarray(idx, dtype=np.int64)] = 100 + k
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % (len(sub) + len(bnd)))
        for k, (name, _) in enumerate(bnd):
            f.write('1 %d "%s"\n' % (1000 + k, name))
        for k, (name, _) in enumerate(sub):
            f.write('2 %d "%s"\n' % (100 + k, name))
        f.write("$EndPhysicalNames\n$Nodes\n%d\n" % mesh.nvertices)
        for i in range(mesh.nvertices):
            f.write("%d %.17g %.17g 0\n" % (i + 1, mesh.p[0, i], mesh.p[1, i]))
        nlines = sum(len(idx) for _, idx in bnd)
        f.write("$EndNodes\n$Elements\n%d\n" % (nlines + mesh.nelements))
        eid = 1
        for k, (_, idx) in enumerate(bnd):
            for fi in np.asarray(idx, dtype=np.int64):
                a, b = mesh.facets[:, fi]
                f.write("%d 1 2 %d %d %d %d\n" % (eid, 1000 + k, 1000 + k, a + 1, b + 1))
                eid += 1
        for e in range(mesh.nelements):
            a, b, c = mesh.t[:, e]
            if tri_tag[e]:
                f.write("%d 2 2 %d %d %d %d %d\n" % (eid, tri_tag[e], tri_tag[e], a + 1, b + 1, c + 1))
            else:
                f.write("%d 2 0 %d %d %d\n" % (eid, a + 1, b + 1, c + 1))
            eid += 1
        f.write("$EndElements\n")
