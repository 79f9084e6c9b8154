"""Building-block diagnostics of the multi-GPU path against a single-GPU run on rank 0 (matched by coordinates)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
grid = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "grids", "box_3D_elongated.npz")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ug = ug4.Backend(device=local, distributed=True)
p = ObstacleOptim(ug, 3, numRefs=refs, grid=grid).setup()
X = p.dom.get_level(refs, elems=False)["xyz"]
n = X.shape[0]
f = lambda Y: np.stack([np.sin(Y[:, 0]) + Y[:, 1], np.cos(Y[:, 1] * 0.7) * Y[:, 2], Y[:, 0] * Y[:, 2] + 0.3], 1)

def gather_global(vec_local):     # consistent local vector -> (coords, values) of owned vertices, gathered on all ranks
    owned = p.dom._iface[refs]["owned"].astype(bool)
    out = [None] * world
    dist.all_gather_object(out, (X[owned], vec_local.reshape(-1, 3)[owned]))
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])

res = {}
# (1) multiplicity: ones (additive) -> consistent = number of sharing ranks
p.sigma.from_numpy(np.ones(n * 3), 2)
p.sigma.change_storage_type_to_consistent()
mult = p.sigma.to_numpy().reshape(-1, 3)[:, 0]
res["mult"] = gather_global(p.sigma.to_numpy())
# (2) SpMV: y = A x (additive) -> consistent
DD = p.DeformationEquation_DomainDisc
DD.assemble_jacobian(p.A_u_Hessian, p.u)
p.sigma.from_numpy(f(X).ravel(), 1)
DD.adjust_solution(p.sigma)
p.A_u_Hessian.apply(p.Lu, p.sigma)
p.Lu.change_storage_type_to_consistent()
res["spmv"] = gather_global(p.Lu.to_numpy())
# (3) V-cycle on an additive rhs built from a consistent field (made additive by the owner mask)
p.Lu.from_numpy(f(X).ravel(), 1)
DD.adjust_solution(p.Lu)
p.Lu.change_storage_type_to_additive()
s = p.SmallProblemRHS_Solver
s.init(p.A_u_Hessian, p.sigma)
s.vcycle(p.delta_u, p.Lu)
res["vcycle"] = gather_global(p.delta_u.to_numpy())
# (4) full solve
p.sigma.set(0.0)
ok = s.apply(p.sigma, p.Lu)
res["solve"] = gather_global(p.sigma.to_numpy())
its = s.step()
# (5) the same solve to 1e-13 with the residual history (tolerances are fixed when the C solver object is created)
p2 = ObstacleOptim(ug, 3, numRefs=refs, grid=grid, solver_verbose=True).setup()
DD2 = p2.DeformationEquation_DomainDisc
DD2.assemble_jacobian(p2.A_u_Hessian, p2.u)
p2.Lu.from_numpy(f(X).ravel(), 1)
DD2.adjust_solution(p2.Lu)
p2.Lu.change_storage_type_to_additive()
s2 = p2.SmallProblemRHS_Solver
s2.desc.abs_tol = 1e-13
s2.desc.max_iterations = 25
s2.init(p2.A_u_Hessian, p2.sigma)
p2.sigma.set(0.0)
ok13 = s2.apply(p2.sigma, p2.Lu)
# true residual of the returned iterate
p2.A_u_Hessian.apply(p2.delta_u, p2.sigma)
ug.VecScaleAdd2(p2.delta_u, 1.0, p2.Lu, -1.0, p2.delta_u)
p2.delta_u.change_storage_type_to_consistent()
true_res = ug.VecNorm(p2.delta_u)
if rank == 0:
    print("1e-13 solve: ok", ok13, "its", s2.step(), "reported defect", s2.defect(), "true residual", true_res, flush=True)
if rank == 0:
    ug1 = ug4.Backend(device=local)
    q = ObstacleOptim(ug1, 3, numRefs=refs, grid=grid).setup()
    Xg = q.dom.get_level(refs, elems=False)["xyz"]
    order = lambda a: np.lexsort(tuple(a[:, c] for c in reversed(range(3))))
    og = order(Xg)
    DDq = q.DeformationEquation_DomainDisc
    DDq.assemble_jacobian(q.A_u_Hessian, q.u)
    q.sigma.from_numpy(f(Xg).ravel(), 1); DDq.adjust_solution(q.sigma)
    q.A_u_Hessian.apply(q.Lu, q.sigma)
    ref = {"spmv": q.Lu.to_numpy().reshape(-1, 3)}
    q.Lu.from_numpy(f(Xg).ravel(), 2); DDq.adjust_solution(q.Lu)
    sq = q.SmallProblemRHS_Solver; sq.init(q.A_u_Hessian, q.sigma); sq.vcycle(q.delta_u, q.Lu)
    ref["vcycle"] = q.delta_u.to_numpy().reshape(-1, 3)
    q.sigma.set(0.0); okq = sq.apply(q.sigma, q.Lu); ref["solve"] = q.sigma.to_numpy().reshape(-1, 3)
    print("ranks", world, "refs", refs, "solve ok", ok, "its", its, "| single ok", okq, "its", sq.step())
    Xd, m = res["mult"]; od = order(Xd)
    print("owned total", len(Xd), "global", len(Xg), "coords equal", len(Xd) == len(Xg) and np.array_equal(Xd[od], Xg[og]))
    print("multiplicity histogram", np.bincount(np.rint(m[:, 0]).astype(int)))
    for k in ("spmv", "vcycle", "solve"):
        Xd, v = res[k]; od = order(Xd)
        err = np.abs(v[od] - ref[k][og]).max() / np.abs(ref[k]).max()
        print("%-7s rel max err %.3e" % (k, err))
dist.barrier()
dist.destroy_process_group()
