"""One ADMM iteration for an ncu launch list (warm caches: run under `ncu --cache-control none --clock-control none`).
    python tools/ncu_iter.py [numRefs] [graph 0|1] [spmv_variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_optim_b200 import ug4  # noqa: E402
from admm_optim_b200.driver import ObstacleOptim  # noqa: E402
import torch  # noqa: E402

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
stream = torch.cuda.Stream()
ug = ug4.Backend(device=0, stream=stream.cuda_stream)
ug.set_tuning("graph", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
if len(sys.argv) > 3:
    ug.set_tuning("spmv_variant", int(sys.argv[3]))
p = ObstacleOptim(ug, 3, numRefs=refs, grid="grids/box_3D_elongated.npz").setup()
p.set_sensitivity(p.synthetic_sensitivity(0.5))
p.begin_step()
n0 = ug.launch_count()
p.admm_iteration()
ug.synchronize()
print("launches in the iteration:", ug.launch_count() - n0)
