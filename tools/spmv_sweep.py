"""SpMV / V-cycle micro-benchmark sweep over kernel variants (run on the GPU box)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
import bench

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
waves = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
stream = torch.cuda.Stream()
ug = ug4.Backend(device=0, stream=stream.cuda_stream)
t = time.time()
big = ObstacleOptim(ug, 3, numRefs=refs, grid=bench.GRID3D).setup()
print("setup %.1fs" % (time.time() - t))
DD = big.DeformationEquation_DomainDisc
DD.assemble_jacobian(big.A_u_Hessian, big.u)
_, nb, nnzb = big.A_u_Hessian.info()
big.sigma.from_numpy(np.random.default_rng(1).standard_normal(nb * 3))
DD.adjust_solution(big.sigma)
B = bench.spmv_bytes(3, nb, nnzb)
peak, _ = bench.measured_peak()
def timeit(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
s = big.SmallProblemRHS_Solver
s.init(big.A_u_Hessian, big.sigma)
levels = [s.level_info(l) for l in range(refs + 1)]
BV = bench.vcycle_bytes(3, levels)
hints = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
for v in variants:
    for w in waves:
        for h in hints:
            ug.set_tuning("spmv_variant", v); ug.set_tuning("spmv_waves", w); ug.set_tuning("l2_hint", h)
            ts = timeit(lambda: big.A_u_Hessian.apply(big.Lu, big.sigma), 20)
            tv = timeit(lambda: s.vcycle(big.delta_u, big.sigma), 5)
            print("variant %d waves %d l2_hint %d: spmv %.1f us %.0f GB/s (%.3f of peak) | vcycle %.3f ms %.0f GB/s (%.3f)" %
                  (v, w, h, ts * 1e6, B / ts / 1e9, B / ts / 1e9 / peak, tv * 1e3, BV / tv / 1e9, BV / tv / 1e9 / peak))
# copy-bandwidth sanity: torch copy of the same byte volume
a = torch.empty(B // 16, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
with torch.cuda.stream(stream):
    tc = timeit(lambda: b.copy_(a), 10)
print("torch copy of %.2f GB: %.0f GB/s" % (B / 1e9, 2 * a.numel() * 8 / tc / 1e9))
