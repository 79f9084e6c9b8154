"""Worker of test_host.py::test_partition_and_interfaces_gloo_world_size_2 -- run as one of N gloo ranks (CPU only).
Exercises the host side of the multi-GPU path: RCB partition, sub-grid extraction, native refinement of the local
sub-grid, candidate masks, coordinate matching over torch.distributed, and the C-ABI interface registration."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch.distributed as dist

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
ug._create_regular_hierarchy(dom, refs)        # refine + match + ab_domain_set_interface / set_global_coarse
res = []
for level in range(refs + 1):
    lv = dom.get_level(level, elems=False)
    I = dom._iface[level]
    shared = {int(q): lv["xyz"][I["idx"][I["offsets"][k]:I["offsets"][k + 1]]].tolist() for k, q in enumerate(I["neigh"])}
    res.append(dict(nv=int(lv["xyz"].shape[0]), owned=int(I["owned"].sum()), shared=shared))
allres = gather(res)
if rank == 0:
    print(json.dumps(allres))
dist.barrier()
dist.destroy_process_group()
