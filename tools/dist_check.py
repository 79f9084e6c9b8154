"""Multi-GPU parity check (run under torchrun, one rank per GPU): the domain-decomposed ADMM iterations -- interface
exchanges over NVLink peer memory, agglomerated coarse levels on rank 0, graph-replayed distributed BiCGStab -- must reproduce
the single-GPU result (matched by vertex coordinates): deformation within 1e-9 relative L2, per-iteration scalars within 1e-8,
and the SAME Newton / BiCGStab iteration counts (+-1 per solve: the partition changes the summation order only).

  torchrun ... tools/dist_check.py <numRefs> <dim> [gather_dofs]
gather_dofs overrides ADMM_B200_GATHER_DOFS so that small test grids are decomposed at all (default here: 100 -> only level 0
is agglomerated)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get("DIST_CHECK_WATCHDOG_S", "180")), exit=True)   # a hang prints where, then ends the rank
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
os.environ["ADMM_B200_GATHER_DOFS"] = sys.argv[3] if len(sys.argv) > 3 else "100"
import numpy as np, torch, torch.distributed as dist
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim

grid = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "grids", "box_3D_elongated.npz" if dim == 3 else "refined.npz")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ug = ug4.Backend(device=local, distributed=True)
p = ObstacleOptim(ug, dim, numRefs=refs, grid=grid, admmSteps=2, verbose=(rank == 0 and os.environ.get("DIST_VERBOSE") == "1")).setup()
for s in [p.SmallProblemRHS_Solver, p.LargeProblem_Solver] + p.B_Solver:
    s.desc.abs_tol = 1e-13
J = p.synthetic_sensitivity(0.5)
p.set_sensitivity(J)
ug.synchronize(); dist.barrier()
t0 = time.time()
tr = p.run_admm()
ug.synchronize()
dt = time.time() - t0
assert not p.p_solver_failure
assert p.dom.p2p_status()["error"] == 0
u_loc = p.u.to_numpy().reshape(-1, dim)
X_loc = p.dom.get_level(refs, elems=False)["xyz"]
owned = p.dom._iface[refs]["owned"].astype(bool) if p.dom.decomposed else (np.ones(len(X_loc), bool) if rank == 0 else np.zeros(len(X_loc), bool))
out = [None] * world
dist.all_gather_object(out, (X_loc[owned], u_loc[owned], [{k: r[k] for k in ("u_diff", "lambda_inc", "max_norm", "Lambda")} for r in tr],
                            [[n["its"] for n in r["newton"]] for r in tr]))
# consistent copies must be bitwise identical on every rank sharing a vertex (rank-ordered interface sums)
if p.dom.decomposed:
    I = p.dom._iface[refs]
    mine = {int(q): u_loc[I["idx"][I["offsets"][k]:I["offsets"][k + 1]]].tobytes() for k, q in enumerate(I["neigh"])}
    every = [None] * world
    dist.all_gather_object(every, mine)
    for q, blob in mine.items():
        assert every[q][rank] == blob, "rank %d and %d hold different copies of their shared vertices" % (rank, q)
del ug, p                                    # this rank's distributed context goes before rank 0 builds the single-GPU reference
if rank == 0:
    X = np.concatenate([o[0] for o in out]); U = np.concatenate([o[1] for o in out])
    os.environ["ADMM_B200_GATHER_DOFS"] = "400000"
    ug1 = ug4.Backend(device=local)
    q = ObstacleOptim(ug1, dim, numRefs=refs, grid=grid, admmSteps=2).setup()
    for s in [q.SmallProblemRHS_Solver, q.LargeProblem_Solver] + q.B_Solver:
        s.desc.abs_tol = 1e-13
    q.set_sensitivity(q.synthetic_sensitivity(0.5))
    t0 = time.time(); tq = q.run_admm(); ug1.synchronize(); dt1 = time.time() - t0
    Xg = q.dom.get_level(refs, elems=False)["xyz"]; Ug = q.u.to_numpy().reshape(-1, dim)
    assert len(X) == len(Xg), (len(X), len(Xg))
    def order(a):
        return np.lexsort(tuple(a[:, c] for c in reversed(range(dim))))
    oa, ob = order(X), order(Xg)
    assert np.array_equal(X[oa], Xg[ob])
    rel = np.linalg.norm(U[oa] - Ug[ob]) / np.linalg.norm(Ug)
    print("ranks %d  numRefs %d  gather_dofs %s: rel L2 diff of u vs single GPU = %.3e ; time dist %.3fs single %.3fs" % (world, refs, sys.argv[3] if len(sys.argv) > 3 else "100", rel, dt, dt1))
    for a, b in zip(out[0][2], tq):
        for k in ("u_diff", "lambda_inc", "max_norm"):
            print("   %-11s dist %.14e single %.14e rel %.2e" % (k, a[k], b[k], abs(a[k] - b[k]) / max(abs(b[k]), 1e-300)))
            assert abs(a[k] - b[k]) <= 1e-8 * max(abs(b[k]), 1e-3)
    its_d = out[0][3]
    its_s = [[n["its"] for n in r["newton"]] for r in tq]
    print("   its dist  ", its_d[0][:2])
    print("   its single", its_s[0][:2])
    assert [len(a) for a in its_d] == [len(b) for b in its_s], "Newton iteration counts differ"
    flat = lambda its: [v for step in its for n in step for v in ([n["rhs"], n["large"]] + list(n["B"]))]
    worst = max(abs(a - b) for a, b in zip(flat(its_d), flat(its_s)))
    print("   max |BiCGStab its dist - single| over all solves: %d ; totals %d vs %d" % (worst, sum(flat(its_d)), sum(flat(its_s))))
    assert worst <= 1
    assert rel < 1e-9
    print("DIST CHECK OK")
sys.stdout.flush()
dist.barrier()
sys.stdout.flush(); sys.stderr.flush()
os._exit(0)       # multi-rank tools end here: no interpreter-shutdown teardown order to depend on (every rank has passed the barrier)
