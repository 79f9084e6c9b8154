"""VTKOutput for deformation-space grid functions: `vtkWriter:select_nodal("u1,u2,u3","u"); vtkWriter:print("u", u, step, time, false)`
(3d_admm.lua:716, 985-987, 1099-1101, 1400-1406).  The grid function lives in HBM; it is downloaded once and written as an
XML UnstructuredGrid (.vtu, ASCII) on the CURRENT vertex coordinates of the top level -- one piece per rank plus a .pvtu
index on rank 0 when the grid is decomposed.  File naming follows UG4's VTKOutput [UPSTREAM-UNVERIFIED]: <name>_t<step:04d>.vtu,
with _p<rank:04d> inserted before _t for the pieces of a parallel run."""
from __future__ import annotations

import os

import numpy as np

_CELL_TYPE = {3: 5, 4: 10}        # VTK_TRIANGLE, VTK_TETRA


def write_vtu(path, xyz, elems, point_data):
    """xyz (nv, dim), elems (ne, dim+1), point_data: {name: array (nv,) or (nv, ncomp)}; vectors are padded to 3 components."""
    xyz = np.asarray(xyz, float)
    elems = np.asarray(elems)
    nv, dim = xyz.shape
    ne, nen = elems.shape
    pts = np.zeros((nv, 3))
    pts[:, :dim] = xyz

    def block(a, fmt):
        return "\n".join(" ".join(fmt % v for v in row) for row in np.atleast_2d(a))

    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n <UnstructuredGrid>\n')
        f.write('  <Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (nv, ne))
        f.write('   <Points>\n    <DataArray type="Float64" NumberOfComponents="3" format="ascii">\n%s\n    </DataArray>\n   </Points>\n' % block(pts, "%.17g"))
        f.write('   <Cells>\n    <DataArray type="Int32" Name="connectivity" format="ascii">\n%s\n    </DataArray>\n' % block(elems, "%d"))
        f.write('    <DataArray type="Int32" Name="offsets" format="ascii">\n%s\n    </DataArray>\n' % " ".join(str(nen * (i + 1)) for i in range(ne)))
        f.write('    <DataArray type="UInt8" Name="types" format="ascii">\n%s\n    </DataArray>\n   </Cells>\n' % " ".join([str(_CELL_TYPE[nen])] * ne))
        f.write('   <PointData>\n')
        for name, a in point_data.items():
            a = np.asarray(a, float).reshape(nv, -1)
            ncomp = a.shape[1]
            if ncomp in (2, 3):
                v = np.zeros((nv, 3))
                v[:, :ncomp] = a
                a, ncomp = v, 3
            f.write('    <DataArray type="Float64" Name="%s" NumberOfComponents="%d" format="ascii">\n%s\n    </DataArray>\n' % (name, ncomp, block(a, "%.17g")))
        f.write('   </PointData>\n  </Piece>\n </UnstructuredGrid>\n</VTKFile>\n')


def write_pvtu(path, pieces, point_data_spec):
    """point_data_spec: {name: ncomp}"""
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="0.1" byte_order="LittleEndian">\n <PUnstructuredGrid GhostLevel="0">\n')
        f.write('  <PPoints>\n   <PDataArray type="Float64" NumberOfComponents="3"/>\n  </PPoints>\n  <PPointData>\n')
        for name, ncomp in point_data_spec.items():
            f.write('   <PDataArray type="Float64" Name="%s" NumberOfComponents="%d"/>\n' % (name, 3 if ncomp in (2, 3) else ncomp))
        f.write('  </PPointData>\n')
        for p in pieces:
            f.write('  <Piece Source="%s"/>\n' % os.path.basename(p))
        f.write(' </PUnstructuredGrid>\n</VTKFile>\n')


def read_vtu(path):
    """Minimal reader of the files written above (tests): returns dict(points, connectivity, point_data)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    nv, ne = int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells"))
    arr = lambda e, t=float: np.array(e.text.split(), dtype=t)
    pts = arr(piece.find("Points/DataArray")).reshape(nv, 3)
    cells = {d.get("Name"): arr(d, int) for d in piece.findall("Cells/DataArray")}
    pd = {d.get("Name"): arr(d).reshape(nv, int(d.get("NumberOfComponents"))) for d in piece.findall("PointData/DataArray")}
    return dict(points=pts, connectivity=cells["connectivity"].reshape(ne, -1), offsets=cells["offsets"], types=cells["types"], point_data=pd)


class VTKOutput:
    """The subset of UG4's VTKOutput the drivers use on deformation-space functions: clear_selection / select_nodal / select_all / print."""

    def __init__(self, ug=None):
        self.ug = ug
        self.selection = []
        self.written = []

    def clear_selection(self):
        self.selection = []

    def select_nodal(self, fcts, name):
        self.selection.append(([f.strip() for f in fcts.split(",")], name))

    def select_all(self, flag):
        self.selection = [] if not flag else self.selection

    def print(self, filename, gf, step=None, time=None, make_consistent=False):
        space = gf.space
        if space.kind != 1:
            raise ValueError("VTKOutput: nodal output needs a P1 grid function")
        dom = space.dom
        if hasattr(dom, "get_level"):
            lv = dom.get_level(dom.num_levels() - 1)
        else:                                  # a backend that keeps its levels on the host (the NumPy twin used by the tests)
            lv = dict(xyz=dom.top.xyz, elems=dom.top.elems)
        nranks, rank = getattr(self.ug, "nranks", 1), getattr(self.ug, "rank", 0)
        decomposed = bool(getattr(dom, "decomposed", False))
        if make_consistent or (decomposed and gf.has_storage_type_additive()):
            gf.change_storage_type_to_consistent()
        vals = gf.to_numpy().reshape(len(lv["xyz"]), -1)
        sel = self.selection or [(list(space.names), "_".join(space.names))]
        data = {name: vals[:, [space.fct_index(f) for f in fcts]] for fcts, name in sel}
        tag = "" if step is None else "_t%04d" % int(step)
        if decomposed and nranks > 1:
            piece = "%s_p%04d%s.vtu" % (filename, rank, tag)
            write_vtu(piece, lv["xyz"], lv["elems"], data)
            if rank == 0:
                write_pvtu("%s%s.pvtu" % (filename, tag), ["%s_p%04d%s.vtu" % (filename, r, tag) for r in range(nranks)],
                           {n: a.shape[1] for n, a in data.items()})
            self.written.append(piece)
            return piece
        if nranks > 1 and rank != 0:          # undivided: every rank holds the same function, rank 0 writes it
            return None
        path = "%s%s.vtu" % (filename, tag)
        write_vtu(path, lv["xyz"], lv["elems"], data)
        self.written.append(path)
        return path
