"""ORACLE (test infrastructure, not product code) -- mesh front end in NumPy.

CPU restatement of the grid layer the reference scripts rely on:
  * `.ugx` reader            <- LoadDomain(dom, gridName)                      3d_admm.lua:108-109, 2d_admm.lua:131-132
  * regular refinement       <- util.refinement.CreateRegularHierarchy(...)    3d_admm.lua:186
  * P1 / P0 numbering        <- ApproximationSpace:add_fct(...)                3d_admm.lua:329-333, 367-370
  * Dirichlet / subset masks <- DirichletBoundary():add(0,"u1","inlet") ...    3d_admm.lua:445-457

The arithmetic of UG4 itself is NOT in /root/reference; parity is
therefore *unpinned* by the reference and this file states the conventions both
the oracle and the CUDA library follow (DESIGN.md "Mesh conventions"):

  level l -> l+1
    - edges      = unique (min,max) vertex pairs of all elements, sorted lexicographically
    - vertices   = [copies of the level-l vertices in order] + [edge midpoints in edge order]
    - triangle (a,b,c)     -> (a,mab,mca) (mab,b,mbc) (mca,mbc,c) (mab,mbc,mca)
    - tetrahedron (0,1,2,3)-> 4 corner tets + inner octahedron cut along its SHORTEST
                              diagonal (ties: first of m01-m23, m02-m13, m03-m12);
                              children re-oriented so that det > 0
    - a midpoint inherits the subset of its parent edge; an edge inherits the subset of the
      lowest-dimensional parent entity it lies in (edge / face / element)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

__all__ = ["Mesh", "load_ugx", "load_npz", "save_npz", "refine", "build_hierarchy", "p1_pattern"]


@dataclass
class Mesh:
    dim: int
    xyz: np.ndarray            # (nv, dim) float64 current vertex coordinates
    elems: np.ndarray          # (ne, dim+1) int32
    subset_names: list         # list[str]
    vsub: np.ndarray           # (nv,) int32 subset index of every vertex
    esub: np.ndarray           # (ne,) int32 subset index of every element
    sp_edges: np.ndarray       # (k,2) int32 "special" edges (subset != element subset), sorted (min,max)
    sp_edges_sub: np.ndarray   # (k,)
    sp_faces: np.ndarray       # (m,3) int32 special faces (3D only), each row sorted
    sp_faces_sub: np.ndarray   # (m,)
    # filled by refine() on the CHILD level
    parent_a: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    parent_b: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    nv_coarse: int = 0

    @property
    def nv(self):
        return self.xyz.shape[0]

    @property
    def ne(self):
        return self.elems.shape[0]

    def subset_index(self, name: str) -> int:
        return self.subset_names.index(name)

    def vertex_mask(self, names) -> np.ndarray:
        if isinstance(names, str):
            names = [n.strip() for n in names.split(",")]
        ids = [self.subset_index(n) for n in names]
        return np.isin(self.vsub, ids)


def _ints(txt):
    return np.array(txt.split(), dtype=np.int64) if txt and txt.strip() else np.zeros(0, np.int64)


def load_ugx(path: str) -> Mesh:
    """Parse the subset of the UGX XML format the two shipped grids use
    (grids/refined.ugx:1-28, grids/box_3D_elongated.ugx:1-34)."""
    s = open(path).read()
    m = re.search(r'<vertices coords="(\d)">([^<]*)</vertices>', s)
    nc = int(m.group(1))
    xyz3 = np.array(m.group(2).split(), dtype=np.float64).reshape(-1, nc)
    gpart = s[: s.index("<subset_handler")]

    def grab(tag):
        mm = re.search(r"<%s>([^<]*)</%s>" % (tag, tag), gpart)
        return _ints(mm.group(1)) if mm else np.zeros(0, np.int64)

    edges = grab("edges").reshape(-1, 2)
    tris = grab("triangles").reshape(-1, 3)
    tets = grab("tetrahedrons").reshape(-1, 4)
    dim = 3 if len(tets) else 2
    elems = (tets if dim == 3 else tris).astype(np.int32)
    xyz = np.ascontiguousarray(xyz3[:, :dim])
    nv = xyz.shape[0]

    names, vsub, esub = [], -np.ones(nv, np.int32), -np.ones(len(elems), np.int32)
    edge_sub = -np.ones(len(edges), np.int32)
    face_sub = -np.ones(len(tris), np.int32)
    sh = s[s.index("<subset_handler"):]
    for si, sm in enumerate(re.finditer(r'<subset name="([^"]+)"[^>]*>(.*?)</subset>', sh, re.S)):
        names.append(sm.group(1))
        body = sm.group(2)

        def sub(tag):
            mm = re.search(r"<%s>([^<]*)</%s>" % (tag, tag), body)
            return _ints(mm.group(1)) if mm else np.zeros(0, np.int64)

        vsub[sub("vertices")] = si
        edge_sub[sub("edges")] = si
        if dim == 3:
            face_sub[sub("faces")] = si
            esub[sub("volumes")] = si
        else:
            esub[sub("faces")] = si
    assert (vsub >= 0).all() and (esub >= 0).all()
    vol_sub = esub[0]
    assert (esub == vol_sub).all(), "single element subset expected ('outer')"
    e_sorted = np.sort(edges, axis=1).astype(np.int32)
    keep = edge_sub != vol_sub
    o = np.lexsort((e_sorted[keep][:, 1], e_sorted[keep][:, 0]))
    sp_e, sp_es = e_sorted[keep][o], edge_sub[keep][o]
    if dim == 3:
        f_sorted = np.sort(tris, axis=1).astype(np.int32)
        keepf = face_sub != vol_sub
        sp_f, sp_fs = f_sorted[keepf], face_sub[keepf]
    else:
        sp_f, sp_fs = np.zeros((0, 3), np.int32), np.zeros(0, np.int32)
    return Mesh(dim, xyz, elems, names, vsub, esub, sp_e, sp_es.astype(np.int32), sp_f, sp_fs.astype(np.int32))


def save_npz(mesh: Mesh, path: str):
    np.savez_compressed(path, dim=mesh.dim, xyz=mesh.xyz, elems=mesh.elems,
                        subset_names=np.array(mesh.subset_names), vsub=mesh.vsub, esub=mesh.esub,
                        sp_edges=mesh.sp_edges, sp_edges_sub=mesh.sp_edges_sub,
                        sp_faces=mesh.sp_faces, sp_faces_sub=mesh.sp_faces_sub)


def load_npz(path: str) -> Mesh:
    z = np.load(path)
    return Mesh(int(z["dim"]), z["xyz"].copy(), z["elems"].copy(), [str(x) for x in z["subset_names"]],
                z["vsub"].copy(), z["esub"].copy(), z["sp_edges"].copy(), z["sp_edges_sub"].copy(),
                z["sp_faces"].copy(), z["sp_faces_sub"].copy())


_LOCAL_EDGES = {2: [(0, 1), (1, 2), (0, 2)],
                3: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]}


def unique_edges(mesh: Mesh) -> np.ndarray:
    """Sorted unique (min,max) edges of all elements, lexicographic order."""
    le = _LOCAL_EDGES[mesh.dim]
    a = np.concatenate([mesh.elems[:, i] for i, _ in le]).astype(np.int64)
    b = np.concatenate([mesh.elems[:, j] for _, j in le]).astype(np.int64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo * mesh.nv + hi)
    return np.stack([key // mesh.nv, key % mesh.nv], axis=1).astype(np.int32)


def _edge_lookup(edges, nv):
    key = edges[:, 0].astype(np.int64) * nv + edges[:, 1]

    def find(a, b):
        lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
        k = lo * nv + hi
        pos = np.searchsorted(key, k)
        assert (key[pos] == k).all()
        return pos

    return find, key


def _sqdist(p, q):
    d = p - q
    acc = d[:, 0] * d[:, 0]
    for c in range(1, d.shape[1]):
        acc = acc + d[:, c] * d[:, c]      # fixed left-to-right order, no fused ops
    return acc


def refine(mesh: Mesh) -> Mesh:
    nv, dim = mesh.nv, mesh.dim
    edges = unique_edges(mesh)
    find, ekey = _edge_lookup(edges, nv)
    mid_xyz = 0.5 * (mesh.xyz[edges[:, 0]] + mesh.xyz[edges[:, 1]])
    xyz = np.concatenate([mesh.xyz, mid_xyz])
    vol_sub = mesh.esub[0]
    mid_sub = np.full(len(edges), vol_sub, np.int32)
    if len(mesh.sp_edges):
        pos = find(mesh.sp_edges[:, 0], mesh.sp_edges[:, 1])
        mid_sub[pos] = mesh.sp_edges_sub
    vsub = np.concatenate([mesh.vsub, mid_sub])

    def mid(a, b):
        return (nv + find(a, b)).astype(np.int32)

    E = mesh.elems
    if dim == 2:
        a, b, c = E[:, 0], E[:, 1], E[:, 2]
        mab, mbc, mca = mid(a, b), mid(b, c), mid(c, a)
        ch = np.stack([np.stack([a, mab, mca], 1), np.stack([mab, b, mbc], 1),
                       np.stack([mca, mbc, c], 1), np.stack([mab, mbc, mca], 1)], axis=1)  # (ne,4,3)
        elems = ch.reshape(-1, 3).astype(np.int32)
    else:
        v0, v1, v2, v3 = E[:, 0], E[:, 1], E[:, 2], E[:, 3]
        m01, m02, m03 = mid(v0, v1), mid(v0, v2), mid(v0, v3)
        m12, m13, m23 = mid(v1, v2), mid(v1, v3), mid(v2, v3)
        d0 = _sqdist(xyz[m01], xyz[m23])
        d1 = _sqdist(xyz[m02], xyz[m13])
        d2 = _sqdist(xyz[m03], xyz[m12])
        choice = np.zeros(len(E), np.int64)
        best = d0.copy()
        c1 = d1 < best
        choice[c1] = 1
        best[c1] = d1[c1]
        c2 = d2 < best
        choice[c2] = 2

        def sel(x0, x1, x2):
            return np.where(choice == 0, x0, np.where(choice == 1, x1, x2))

        p, q = sel(m01, m02, m03), sel(m23, m13, m12)
        c_0, c_1, c_2, c_3 = sel(m02, m01, m01), sel(m03, m03, m02), sel(m13, m23, m23), sel(m12, m12, m13)
        ch = np.stack([np.stack([v0, m01, m02, m03], 1), np.stack([m01, v1, m12, m13], 1),
                       np.stack([m02, m12, v2, m23], 1), np.stack([m03, m13, m23, v3], 1),
                       np.stack([p, q, c_0, c_1], 1), np.stack([p, q, c_1, c_2], 1),
                       np.stack([p, q, c_2, c_3], 1), np.stack([p, q, c_3, c_0], 1)], axis=1)  # (ne,8,4)
        elems = ch.reshape(-1, 4).astype(np.int32)
        X = xyz[elems]
        J = X[:, 1:, :] - X[:, :1, :]
        neg = np.linalg.det(J) < 0
        elems[neg, 2], elems[neg, 3] = elems[neg, 3].copy(), elems[neg, 2].copy()
    esub = np.repeat(mesh.esub, 2 ** dim)

    # special (boundary-subset) entities of the child level
    se, ses = mesh.sp_edges, mesh.sp_edges_sub
    new_e, new_es = [], []
    if len(se):
        m = mid(se[:, 0], se[:, 1])
        new_e += [np.stack([se[:, 0], m], 1), np.stack([se[:, 1], m], 1)]
        new_es += [ses, ses]
    sp_f, sp_fs = np.zeros((0, 3), np.int32), np.zeros(0, np.int32)
    if dim == 3 and len(mesh.sp_faces):
        f, fs = mesh.sp_faces, mesh.sp_faces_sub
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        mab, mbc, mca = mid(a, b), mid(b, c), mid(c, a)
        new_e += [np.stack([mab, mbc], 1), np.stack([mbc, mca], 1), np.stack([mca, mab], 1)]
        new_es += [fs, fs, fs]
        sp_f = np.sort(np.concatenate([np.stack([a, mab, mca], 1), np.stack([mab, b, mbc], 1),
                                       np.stack([mca, mbc, c], 1), np.stack([mab, mbc, mca], 1)]), axis=1).astype(np.int32)
        sp_fs = np.concatenate([fs, fs, fs, fs]).astype(np.int32)
    if new_e:
        ne_ = np.sort(np.concatenate(new_e), axis=1).astype(np.int32)
        nes = np.concatenate(new_es).astype(np.int32)
        o = np.lexsort((ne_[:, 1], ne_[:, 0]))
        ne_, nes = ne_[o], nes[o]
    else:
        ne_, nes = np.zeros((0, 2), np.int32), np.zeros(0, np.int32)
    child = Mesh(dim, xyz, elems, mesh.subset_names, vsub.astype(np.int32), esub.astype(np.int32), ne_, nes, sp_f, sp_fs)
    child.parent_a, child.parent_b, child.nv_coarse = edges[:, 0].copy(), edges[:, 1].copy(), nv
    return child


def build_hierarchy(mesh: Mesh, num_refs: int) -> list:
    levels = [mesh]
    for _ in range(num_refs):
        levels.append(refine(levels[-1]))
    return levels


def p1_pattern(mesh: Mesh):
    """CSR pattern (rowptr, colidx) of the P1 vertex graph incl. diagonal, columns ascending.
    Also returns `mid`: for every entry (i,j) the vertex id the midpoint of edge (i,j) gets on
    the NEXT level (nv + edge index); the diagonal entry maps to i itself (the copy)."""
    nv = mesh.nv
    edges = unique_edges(mesh)
    rows = np.concatenate([edges[:, 0], edges[:, 1], np.arange(nv, dtype=np.int32)])
    cols = np.concatenate([edges[:, 1], edges[:, 0], np.arange(nv, dtype=np.int32)])
    eid = np.arange(len(edges), dtype=np.int64) + nv
    mids = np.concatenate([eid, eid, np.arange(nv, dtype=np.int64)])
    o = np.lexsort((cols, rows))
    rows, cols, mids = rows[o], cols[o], mids[o]
    rowptr = np.zeros(nv + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), cols.astype(np.int32), mids.astype(np.int32)
