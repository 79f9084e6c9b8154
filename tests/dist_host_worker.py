"""Worker of test_host.py::test_partition_and_interfaces_gloo_world_size_2 -- run as one of N gloo ranks (CPU only).
Exercises the host side of the multi-GPU path: RCB partition, sub-grid extraction, native refinement of the local
sub-grid, candidate masks, coordinate matching over torch.distributed, the C-ABI interface registration and the
vertical-interface maps of the agglomerated coarse levels (local -> global vertices / matrix blocks on the gather level)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch.distributed as dist

from admm_optim_b200 import partition as P
from admm_optim_b200 import ug4

grid, refs = sys.argv[1], int(sys.argv[2])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


ug = ug4.Backend.host_only(rank, world, gather)      # host-only: no GPU context
dom = ug4.Domain(ug)
ug.LoadDomain(dom, grid)
ug._create_regular_hierarchy(dom, refs)        # decides: undivided, or refine + match + ab_domain_set_interface + gather maps
if not dom.decomposed:
    out = dict(decomposed=False, levels=[dom.level_info(l)["nv"] for l in range(dom.num_levels())])
    allres = gather(out)
    if rank == 0:
        print(json.dumps(allres))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
res = []
for level in range(refs + 1):
    lv = dom.get_level(level, elems=False)
    I = dom._iface[level]
    shared = {int(q): lv["xyz"][I["idx"][I["offsets"][k]:I["offsets"][k + 1]]].tolist() for k, q in enumerate(I["neigh"])}
    res.append(dict(nv=int(lv["xyz"].shape[0]), owned=int(I["owned"].sum()), shared=shared))
# vertical interface: every rank's element-incidence counts per local block, summed through gpos, must give the global counts
G = dom._gather
lg = G["level"]
ll, gl = dom.get_level(lg), dom._cdom.get_level(lg)
nvl, nvg = len(ll["xyz"]), G["nv_global"]
def incidence(elems, nv):
    el = np.asarray(elems, np.int64)
    n = el.shape[1]
    keys = np.concatenate([el[:, a] * nv + el[:, b] for a in range(n) for b in range(n)])
    return np.unique(keys, return_counts=True)
lk, lc = incidence(ll["elems"], nvl)
assert np.array_equal(lk, P.pattern_keys(ll["elems"], nvl)) and len(lk) == len(G["gpos"])
# the C++ pattern (what the device matrices use) has exactly this order
import ctypes as C
nn = C.c_int64()
ug4.call("ab_domain_level_pattern", dom.h, lg, C.byref(nn), None, None)
rp, ci = np.empty(nvl + 1, np.int32), np.empty(nn.value, np.int32)
ug4.call("ab_domain_level_pattern", dom.h, lg, C.byref(nn), rp.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)))
rows = np.repeat(np.arange(nvl, dtype=np.int64), np.diff(rp))
assert np.array_equal(rows * nvl + ci, lk), "numpy pattern order differs from build_pattern"
parts = gather((G["gpos"], lc, G["l2g"], ll["xyz"]))
ok = True
if rank == 0:
    gk, gc = incidence(gl["elems"], nvg)
    tot = np.zeros(len(gk), np.int64)
    for gpos, cnt, l2g, xyz in parts:
        assert len(np.unique(gpos)) == len(gpos)                   # injective per rank
        np.add.at(tot, gpos, cnt)
        assert np.array_equal(gl["xyz"][l2g], xyz)                 # vertex map hits identical coordinates
    ok = bool(np.array_equal(tot, gc))
    covered = np.zeros(nvg, bool)
    for _, _, l2g, _ in parts:
        covered[l2g] = True
    ok = ok and bool(covered.all())
allres = gather(dict(decomposed=True, gather_level=lg, levels=res, blocks_ok=ok))
if rank == 0:
    print(json.dumps(allres))
dist.barrier()
dist.destroy_process_group()
