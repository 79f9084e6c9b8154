"""Large-refinement measurements (1 or N GPUs under torchrun):
  mode "solver": lean problem (deformation space + Hessian + solver only, no P0 tensors) -> SpMV, V-cycle, BiCGStab solve;
                 numRefs=6 (160.8 M DoFs, 61 GB matrix) fits ONE B200 this way, which gives the 1 -> 8 GPU strong-scaling ratio
                 at the largest refinement.
  mode "admm"  : one full ADMM iteration of the script replay (3d_admm.lua:875-1304) at the given refinement.
Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim, linear_solver

mode, refs = sys.argv[1], int(sys.argv[2])
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
GRID = bench.GRID3D if dim == 3 else bench.GRID2D
CMP = ["u1", "u2", "u3"][:dim]
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
stream = torch.cuda.Stream()
ug = ug4.Backend(device=local, stream=stream.cuda_stream, distributed=world > 1)

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def maxtime(t):
    x = torch.tensor([t], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x.item())

t0 = time.time()
out = {"mode": mode, "numRefs": refs, "n_gpus": world, "dim": dim}
if mode == "solver":
    ug.InitUG(dim, None)
    dom = ug.Domain(); ug.LoadDomain(dom, GRID)
    ug.util.refinement.CreateRegularHierarchy(dom, refs, False, None)
    DS = ug.ApproximationSpace(dom); DS.add_fct(",".join(CMP), "Lagrange", 1); DS.init_levels(); DS.init_top_surface()
    H = ug.DeformationEquation(",".join(CMP), "outer")
    Dir = ug.DirichletBoundary()
    for sub in ("inlet", "wall", "outlet"):
        for c in CMP: Dir.add(0, c, sub)
    DD = ug.DomainDiscretization(DS); DD.add(H); DD.add(Dir)
    A = ug.AssembledLinearOperator(DD)
    x, b, y, u = (ug.GridFunction(DS) for _ in range(4))
    out["setup_s"] = maxtime(time.time() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    os.environ["ADMM_B200_NO_CACHE"] = "1"                            # no operator sharing: the second request re-assembles in place
    barrier(); e0.record(stream); DD.assemble_jacobian(A, u); e1.record(stream); e1.synchronize()
    out["assemble_first_ms"] = maxtime(e0.elapsed_time(e1))           # includes the allocation of the matrix
    barrier(); e0.record(stream); DD.assemble_jacobian(A, u); e1.record(stream); e1.synchronize()   # the kernels alone
    out["assemble_ms"] = maxtime(e0.elapsed_time(e1))
    del os.environ["ADMM_B200_NO_CACHE"]
    n_loc = DS.num_dofs()
    x.from_numpy(np.random.default_rng(1 + rank).standard_normal(n_loc)); DD.adjust_solution(x)
    def timeit(fn, reps):
        for _ in range(2): fn()
        barrier(); e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream); e1.synchronize()
        return maxtime(e0.elapsed_time(e1) / reps)
    levels = bench.global_counts(refs, dim)
    nb, nnzb = levels[-1]
    peak, _ = bench.measured_peak()
    t_spmv = timeit(lambda: A.apply(y, x), 10)
    out.update(dofs=nb * dim, nnzb=nnzb, spmv_ms=t_spmv, spmv_gbs=bench.spmv_bytes(dim, nb, nnzb) / t_spmv / 1e6,
               spmv_frac_per_gpu=bench.spmv_bytes(dim, nb, nnzb) / t_spmv / 1e6 / peak / world)
    s = linear_solver(ug, DD, DS, False, dim)
    s.desc.verbose = 0
    te = time.time(); s.init(A, x); ug.synchronize(); out["gmg_init_first_ms"] = maxtime((time.time() - te) * 1e3)
    os.environ["ADMM_B200_NO_CACHE"] = "1"; s.init(A, x); ug.synchronize(); barrier()      # second hierarchy; the first becomes idle and is recycled next
    te = time.time(); s.init(A, x); ug.synchronize(); out["gmg_init_ms"] = maxtime((time.time() - te) * 1e3)
    del os.environ["ADMM_B200_NO_CACHE"]
    t_v = timeit(lambda: s.vcycle(y, x), 5)
    bv = bench.vcycle_bytes(dim, levels)
    out.update(vcycle_ms=t_v, vcycle_gbs=bv / t_v / 1e6, vcycle_frac=bv / t_v / 1e6 / peak / world)
    b.from_numpy(np.random.default_rng(7 + rank).standard_normal(n_loc), 2); DD.adjust_solution(b)
    y.set(0.0)
    barrier(); ts = time.perf_counter(); ok = s.apply(y, b); ug.synchronize()
    out.update(solve_ms=maxtime((time.perf_counter() - ts) * 1e3), solve_its=s.step(), solve_ok=bool(ok), solve_defect=s.defect())
else:
    p = ObstacleOptim(ug, dim, numRefs=refs, grid=GRID, admmSteps=1).setup()
    out["setup_s"] = maxtime(time.time() - t0)
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    p.begin_step()
    barrier(); ts = time.perf_counter()
    rec = p.admm_iteration()
    ug.synchronize()
    out["admm_iteration_first_s"] = maxtime(time.perf_counter() - ts)      # first of the loop: allocations, graph captures
    assert rec is not None and not p.p_solver_failure
    barrier(); ts = time.perf_counter()
    rec = p.admm_iteration()
    ug.synchronize()
    out["admm_iteration_s"] = maxtime(time.perf_counter() - ts)            # steady state
    assert rec is not None and not p.p_solver_failure
    out.update(dofs=bench.global_counts(refs, dim)[-1][0] * dim, newton_its=len(rec["newton"]), delta_lambda=[n["delta_lambda"] for n in rec["newton"]],
               bicgstab_its=[n["its"] for n in rec["newton"]], L_lambda=rec["L_lambda"], Lambda=[float(v) for v in rec["Lambda"]],
               u_diff=rec["u_diff"], lambda_inc=rec["lambda_inc"], reference_volume=p.ReferenceVolume)
    if world > 1:
        assert p.dom.p2p_status()["error"] == 0
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
sys.stdout.flush(); sys.stderr.flush()
os._exit(0)       # multi-rank tools end here: no interpreter-shutdown teardown order to depend on (every rank has passed the barrier)
