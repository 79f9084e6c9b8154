"""`.ugx` writer: SaveGridLevelToFile(dom:grid(), dom:subset_handler(), numRefs, "Mesh_lev<L>_step<k>.ugx")
(3d_admm.lua:795, 2d_admm.lua:788 -- the bDebugOutput branch of the optimisation loop).

The grid level is written the way UG4's own exporter lays a simplex grid out (the layout of grids/box_3D_elongated.ugx and
grids/refined.ugx): ONE <vertices coords="d"> list, ALL edges, all triangles (3D: every face of every tetrahedron; 2D: the
elements), the tetrahedra, and a <subset_handler name="defSH"> in which every vertex / edge / face / volume belongs to exactly one
subset -- boundary edges and faces keep the subset the refinement handed down to them, interior ones fall into the subset of
the volume elements.  Coordinates are the CURRENT ones (after TransformDomainByDisplacement, read back from the device by
Domain.get_level) and are printed with repr precision, so a file read back by LoadDomain reproduces the level bit for bit.

Host-side output format code: no device work, no oracle.
"""
from __future__ import annotations

import numpy as np

_EDGES = {2: [(0, 1), (1, 2), (0, 2)], 3: [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]}
_FACES = [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)]
_COLORS = ["0.6054 0.0210 0.0270 1", "0.3826 0.3327 0.8282 1", "0.4645 0.6613 0.1648 1", "0.6913 0.6074 0.1469 1",
           "0.0979 0.1758 0.4615 1", "0.8 0.5 0.2 1", "0.2 0.8 0.8 1", "0.8 0.2 0.8 1"]


def _unique_rows(a):
    """Sorted-unique rows of an integer array whose rows are already sorted ascending (lexicographic order)."""
    a = np.asarray(a, np.int64)
    if len(a) == 0:
        return a
    return np.unique(a, axis=0)


def _row_keys(rows, n):
    rows = np.asarray(rows, np.int64)
    key = np.zeros(len(rows), np.int64)
    for c in range(rows.shape[1]):
        key = key * n + rows[:, c]
    return key


def _assign(all_rows, special_rows, special_sub, default_sub, n):
    """Subset per row of all_rows: the special ones keep theirs, the rest get default_sub."""
    sub = np.full(len(all_rows), default_sub, np.int32)
    if len(special_rows):
        ak = _row_keys(all_rows, n)                       # ascending: all_rows is lexicographically sorted
        sk = _row_keys(np.sort(np.asarray(special_rows, np.int64), axis=1), n)
        pos = np.searchsorted(ak, sk)
        if pos.max(initial=0) >= len(ak) or not np.array_equal(ak[pos], sk):
            raise ValueError("a boundary edge/face of the subset tables is not a side of any element")
        sub[pos] = special_sub
    return sub


def _fmt_floats(a):
    return " ".join(repr(float(x)) for x in np.asarray(a, np.float64).ravel())


def _fmt_ints(a):
    return " ".join(map(str, np.asarray(a).ravel().tolist()))


def save_grid_level(g: dict, filename: str, grid_name: str = "defGrid") -> str:
    """Write the grid-level dictionary of Domain.get_grid_dict(level) (dim, xyz, elems, vsub, esub, sp_edges[_sub],
    sp_faces[_sub], subset_names) as a .ugx file.  Returns the file name."""
    dim = int(g["dim"])
    xyz = np.asarray(g["xyz"], np.float64)
    el = np.asarray(g["elems"], np.int64)
    nv = len(xyz)
    esub = np.asarray(g["esub"], np.int32)
    if len(el) == 0 or np.any(esub != esub[0]):
        raise ValueError("exactly one element subset expected")
    vol_sub = int(esub[0])
    edges = _unique_rows(np.sort(np.concatenate([el[:, [i, j]] for i, j in _EDGES[dim]]), axis=1))
    edge_sub = _assign(edges, g["sp_edges"], np.asarray(g["sp_edges_sub"], np.int32), vol_sub, nv)
    if dim == 3:
        faces = _unique_rows(np.sort(np.concatenate([el[:, list(f)] for f in _FACES]), axis=1))
        face_sub = _assign(faces, g["sp_faces"], np.asarray(g["sp_faces_sub"], np.int32), vol_sub, nv)
    else:
        faces, face_sub = el, esub
    vsub = np.asarray(g["vsub"], np.int32)
    names = list(g["subset_names"])
    out = ['<?xml version="1.0" encoding="utf-8"?>', '<grid name="%s">' % grid_name,
           '\t<vertices coords="%d">%s</vertices>' % (dim, _fmt_floats(xyz)),
           "\t<edges>%s</edges>" % _fmt_ints(edges),
           "\t<triangles>%s</triangles>" % _fmt_ints(faces)]
    if dim == 3:
        out.append("\t<tetrahedrons>%s</tetrahedrons>" % _fmt_ints(el))
    out.append('\t<subset_handler name="defSH">')
    for s, name in enumerate(names):
        out.append('\t\t<subset name="%s" color="%s" state="0">' % (name, _COLORS[s % len(_COLORS)]))
        for tag, sub in (("vertices", vsub), ("edges", edge_sub), ("faces", face_sub)) + ((("volumes", esub),) if dim == 3 else ()):
            ids = np.nonzero(sub == s)[0]
            if len(ids):
                out.append("\t\t\t<%s>%s</%s>" % (tag, _fmt_ints(ids), tag))
        out.append("\t\t</subset>")
    out.append("\t</subset_handler>")
    out.append("</grid>")
    with open(filename, "w") as f:
        f.write("\n".join(out) + "\n")
    return filename
