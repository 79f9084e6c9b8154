#!/usr/bin/env python
"""Convert the reference's shipped .ugx grids into the compact .npz fixtures under grids/.

The GPU box has no /root/reference, so the two input grids (grids/refined.ugx,
grids/box_3D_elongated.ugx of the reference) travel as converted data fixtures
(vertex coordinates, element connectivity, subset assignment) -- no reference source code.
Run here (container with /root/reference):  python tools/convert_ugx.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mesh_np as M

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/grids"
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "grids")
for name in ("refined", "box_3D_elongated"):
    m = M.load_ugx(os.path.join(src, name + ".ugx"))
    M.save_npz(m, os.path.join(dst, name + ".npz"))
    print(name, m.dim, m.nv, m.ne)
