"""CPU tests that pin the oracle (no GPU): known answers from the grids, analytic invariants of the restated
model, lua-matrix semantics.  The reference has no tests or golden vectors -- parity is
unpinned by the reference, these are the checks that stand in."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GRID2D, GRID3D
from oracle import fem_np as F
from oracle import mesh_np as M

# V' = V + E ...   (V, E, T) per level
COUNTS3D = [(338, 1786, 1216), (2124, 12786, 9728), (14910, 96476, 77824)]
COUNTS2D = [(160, 436, 276), (596, 1700, 1104), (2296, 6712, 4416), (9008, 26672, 17664)]


@pytest.fixture(scope="module")
def h3():
    return M.build_hierarchy(M.load_npz(GRID3D), 2)


@pytest.fixture(scope="module")
def h2():
    return M.build_hierarchy(M.load_npz(GRID2D), 3)


def test_entity_counts_and_volume(h3, h2):
    for lv, (v, e, t) in zip(h3, COUNTS3D):
        assert (lv.nv, len(M.unique_edges(lv)), lv.ne) == (v, e, t)
        _, vol, _ = F.geometry(lv)
        assert abs(vol.sum() - 719.0) < 1e-9                      # 20*6*6 - 1, exact under refinement
    for lv, (v, e, t) in zip(h2, COUNTS2D):
        assert (lv.nv, len(M.unique_edges(lv)), lv.ne) == (v, e, t)
        _, vol, _ = F.geometry(lv)
        assert abs(vol.sum() - 83.0) < 1e-10
    X = h3[2].xyz[h3[2].elems]
    assert (np.linalg.det(X[:, 1:] - X[:, :1]) > 0).all()         # children re-oriented
    X2 = h2[0].xyz[h2[0].elems]
    assert (np.linalg.det(X2[:, 1:] - X2[:, :1]) < 0).sum() == 134   # clockwise triangles of refined.ugx


def test_subset_inheritance(h3):
    # obstacle surface of the unit cube with n cells per edge has 6 n^2 + 2 vertices
    for lv, n in zip(h3, (4, 8, 16)):
        assert lv.vertex_mask("obstacle_surface").sum() == 6 * n * n + 2
        on = lv.xyz[lv.vertex_mask("obstacle_surface")]
        assert np.allclose(np.abs(on).max(axis=1), 0.5)
        inlet = lv.xyz[lv.vertex_mask("inlet")]
        assert np.allclose(inlet[:, 0], -10.0)


def test_barycenter_zero_and_constants(h3):
    m = h3[1]
    z = np.zeros(m.nv * 3)
    assert np.abs(F.barycenter_defect(m, z)).max() < 1e-10       # symmetric domain
    assert abs(F.volume_defect(m, z, 719.0)) < 1e-9
    # rigid translation: volume unchanged, barycentre moves by V * t
    t = np.tile([0.1, -0.2, 0.05], m.nv)
    assert abs(F.volume_defect(m, t, 719.0)) < 1e-9
    assert np.allclose(F.barycenter_defect(m, t), 719.0 * np.array([0.1, -0.2, 0.05]), atol=1e-9)


@pytest.mark.parametrize("which", ["3d", "2d"])
def test_constraint_derivatives_fd(h3, h2, which):
    m = h3[1] if which == "3d" else h2[2]
    d = m.dim
    rng = np.random.default_rng(0)
    u = 0.02 * rng.standard_normal(m.nv * d)
    v = rng.standard_normal(m.nv * d)
    h = 1e-6

    def g(uu):
        return np.concatenate([[F.volume_defect(m, uu, 0.0)], F.barycenter_defect(m, uu)])
    fd = (g(u + h * v) - g(u - h * v)) / (2 * h)
    for i in range(d + 1):
        w = np.zeros(d + 1)
        w[i] = 1.0
        gp = F.load_vector(m, u, None, w, 1.0)
        assert abs(gp @ v - fd[i]) < 1e-6 * max(1.0, abs(fd[i]))
        H = F.hessian_matrix(m, u, c=0.0, lam_vol=w[0], lam_bary=w[1:])
        fd2 = (F.load_vector(m, u + h * v, None, w, 1.0) - F.load_vector(m, u - h * v, None, w, 1.0)) / (2 * h)
        assert np.abs(H @ v - fd2).max() < 1e-6 * np.abs(fd2).max()
        assert abs(H - H.T).max() < 1e-12


def test_laplacian_kernel_and_dirichlet(h3):
    m = h3[1]
    A = F.hessian_matrix(m, None)
    assert np.abs(A @ np.ones(m.nv * 3)).max() < 1e-10            # constants in the kernel when Lambda = 0
    dm = F.dirichlet_dofs(m)
    Ad = F.hessian_matrix(m, None, dmask=dm)
    assert abs(Ad - Ad.T).max() == 0 or abs(Ad - Ad.T).max() < 1e-14
    assert np.allclose(Ad.diagonal()[dm], 1.0)
    assert dm.sum() == 3 * (m.vertex_mask("inlet").sum() + m.vertex_mask("wall").sum() + m.vertex_mask("outlet").sum())


def test_transfer_reproduces_linear_and_rap_is_rediscretisation(h3):
    P = F.prolongation(h3[2], 3)
    lin_c = (h3[1].xyz @ np.array([[1.0, 2, 3], [0.5, -1, 2], [2, 0, 1]])).ravel()
    lin_f = (h3[2].xyz @ np.array([[1.0, 2, 3], [0.5, -1, 2], [2, 0, 1]])).ravel()
    assert np.abs(P @ lin_c - lin_f).max() < 1e-12
    Af = F.hessian_matrix(h3[2], None)
    Ac = F.hessian_matrix(h3[1], None)
    assert abs(P.T @ Af @ P - Ac).max() < 1e-11                   # nested P1 spaces, constant coefficient


def test_gmg_bicgstab_smoothers_side_by_side(h3):
    dm = [F.dirichlet_dofs(l) for l in h3]
    A = F.hessian_matrix(h3[2], None, dmask=dm[2])
    b = np.random.default_rng(1).standard_normal(A.shape[0]) * (~dm[2])
    its = {}
    for sm in ("gs", "cheb", "jac"):
        g = F.GMG(h3, A, dm, smoother=sm, cheb_ratio=6.0)
        x, ok, n, r = F.bicgstab(A, b, np.zeros_like(b), g.apply, abs_tol=1e-10)
        assert ok and np.linalg.norm(b - A @ x) < 1e-9
        its[sm] = n
    assert its["gs"] <= its["cheb"] <= its["jac"] <= 12, its   # 6 / 7 / 9


def test_projections():
    rng = np.random.default_rng(2)
    q = rng.standard_normal(9 * 50)
    p = F.project_frobenius(q, 0.3, 3).reshape(-1, 9)
    assert (np.linalg.norm(p, axis=1) <= 0.3 + 1e-14).all()
    small = 0.01 * q
    assert np.array_equal(F.project_frobenius(small, 0.3, 3), small)
    q2 = rng.standard_normal(4 * 200)
    p2 = F.project_spectral(q2, 0.5).reshape(-1, 2, 2)
    Q2 = q2.reshape(-1, 2, 2)
    U, S, Vt = np.linalg.svd(Q2)
    ref = U @ (np.minimum(S, 0.5)[:, :, None] * Vt)
    assert np.allclose(p2, ref, atol=1e-12)


def test_lua_matrix_semantics():
    from admm_optim_b200.schur import Matrix
    rng = np.random.default_rng(3)
    for n in (3, 4):
        S = rng.standard_normal((n, n)) + n * np.eye(n)
        inv_o = np.array(F.lua_matrix_invert(S.tolist()))
        inv_p = np.array(Matrix(S.tolist()).invert())
        assert np.array_equal(inv_o, inv_p)                       # two independent restatements, same roundings
        assert np.allclose(inv_o, np.linalg.inv(S), rtol=1e-12, atol=1e-13)
        r = rng.standard_normal((n, 1))
        assert np.array_equal(np.array(F.lua_matrix_mul(inv_o.tolist(), r.tolist())), np.array(Matrix(inv_p.tolist()).mul(Matrix(r.tolist()))))
    # pivot rule: smallest non-zero magnitude in the column is swapped up (matrix.lua:422-442)
    S = [[4.0, 1.0], [0.5, 3.0]]
    assert np.allclose(np.array(Matrix(S).invert()), np.linalg.inv(np.array(S)))
    assert Matrix([[1.0, 2.0], [2.0, 4.0]]).invert() is None and F.lua_matrix_invert([[1.0, 2.0], [2.0, 4.0]]) is None


def test_newton_schur_replay_converges_quadratically():
    """The driver replay on the oracle backend: Newton on the KKT system contracts quadratically and the constraint
    residuals L_lambda vanish -- the sign conventions of DESIGN.md 'Signs' are self-consistent in 2D and 3D."""
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    for dim, grid, refs in ((3, GRID3D, 0), (2, GRID2D, 1)):
        p = ObstacleOptim(ug4_np.Backend(smoother="cheb"), dim, numRefs=refs, grid=grid, admmSteps=1).setup()
        p.set_sensitivity(p.synthetic_sensitivity(0.5))
        tr = p.run_admm()
        assert tr and not p.p_solver_failure
        dl = [n["delta_lambda"] for n in tr[0]["newton"]]
        assert dl[-1] <= 1e-9 and len(dl) <= 8
        assert all(dl[i + 1] < 0.5 * dl[i] for i in range(len(dl) - 1))
        assert max(abs(v) for v in tr[0]["L_lambda"]) < 1e-8
        assert tr[0]["u_diff"] > 0 and tr[0]["lambda_inc"] > 0


@pytest.mark.parametrize("name", ["3d_refs1", "2d_refs2"])
def test_oracle_reproduces_committed_golden_trace(name):
    """tests/golden/admm_trace_*.json (tools/make_golden.py): regression pin of the oracle's ADMM trace -- the columns the
    drivers write to __ADMMStats_step_*.txt (3d_admm.lua:1265-1276) plus the multiplier / Newton trace."""
    import json
    import os
    from conftest import ROOT
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "admm_trace_%s.json" % name)))
    p = ObstacleOptim(ug4_np.Backend(smoother="cheb"), gold["dim"], numRefs=gold["numRefs"], grid=os.path.join(ROOT, "grids", gold["grid"]),
                      admmSteps=1).setup()
    for s in [p.SmallProblemRHS_Solver, p.LargeProblem_Solver] + p.B_Solver:
        s.desc["convCheck"]["absolute"] = gold["abs_tol"]
    p.set_sensitivity(p.synthetic_sensitivity(gold["amplitude"]))
    tr = p.run_admm()                                            # first ADMM iteration only (keeps the CPU suite fast)
    g = gold["admm"][0]
    assert abs(p.ReferenceVolume - gold["reference_volume"]) < 1e-9
    assert len(tr[0]["newton"]) == g["newton_its"]
    for k in ("u_diff", "lambda_inc", "max_norm"):
        assert abs(tr[0][k] - g[k]) <= 1e-9 * max(abs(g[k]), 1e-3), k
    assert np.allclose(tr[0]["Lambda"], g["Lambda"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("n,K", [(37, 16), (64, 16), (594, 16), (10, 16)])
def test_blocked_sweep_inversion_is_the_inverse(n, K):
    """The algorithm of k_gauss_jordan_resident (coarse direct solve, replaces SuperLU(), obstacle_optim_3d_util.lua:21) restated
    in NumPy: in-place blocked Gauss-Jordan "sweep" without row exchanges.  Per block of K pivots P:
        Dinv = T[P,P]^-1 ; pivot rows: T[P,J] = Dinv T[P,J], T[P,P] = Dinv ; other rows i: w = T[i,P] Dinv,
        T[i,J] -= w T[P,J], T[i,P] = -w.      After all blocks T = A^-1 (checked against numpy.linalg.inv on an SPD matrix)."""
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    A = B @ B.T + n * np.eye(n)
    T = A.copy()
    for k0 in range(0, n, K):
        P = np.arange(k0, min(k0 + K, n))
        J = np.setdiff1d(np.arange(n), P)
        Dinv = np.linalg.inv(T[np.ix_(P, P)])
        R = T[np.ix_(P, J)].copy()
        W = T[np.ix_(J, P)] @ Dinv
        T[np.ix_(J, J)] -= W @ R
        T[np.ix_(J, P)] = -W
        T[np.ix_(P, J)] = Dinv @ R
        T[np.ix_(P, P)] = Dinv
    assert np.allclose(T, np.linalg.inv(A), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("grid,refs", [(GRID3D, 1), (GRID2D, 2)])
def test_compiled_assembly_matches_numpy_assembly(grid, refs):
    """bench.py's CPU legs assemble with the C element kernels (oracle_kernels.c: oracle_hessian_scatter / oracle_load_scatter,
    `ug4_np.Backend(fast_assembly=True)`); the NumPy functions remain the checker of every parity test.  Same formulas:
    the matrices and load vectors must agree to rounding for every combination the drivers use."""
    if F.c_kernels() is None or not hasattr(F.c_kernels(), "oracle_hessian_scatter"):
        pytest.skip("oracle C library not built")
    mesh = M.build_hierarchy(M.load_npz(grid), refs)[-1]
    d = mesh.dim
    rng = np.random.default_rng(7)
    u = 0.02 * rng.standard_normal(mesh.nv * d)
    dm = F.dirichlet_dofs(mesh)
    for uu in (u, None):
        for lamv, lb in ((0.0, [0.0, 0.0, 0.0]), (0.3, [-0.2, 0.1, 0.05])):
            A = F.hessian_matrix(mesh, uu, 1.3, lamv, lb[:d], dm)
            B = F.hessian_matrix_fast(mesh, uu, 1.3, lamv, lb[:d], dm)
            assert abs(A - B).max() <= 1e-13 * abs(A).max()
            A = F.hessian_matrix(mesh, uu, 0.7, lamv, lb[:d], None)
            B = F.hessian_matrix_fast(mesh, uu, 0.7, lamv, lb[:d], None)
            assert abs(A - B).max() <= 1e-13 * abs(A).max()
    lam = 0.1 * rng.standard_normal(mesh.ne * d * d)
    q = 0.1 * rng.standard_normal(mesh.ne * d * d)
    G, _, _ = F.geometry(mesh)
    w = np.array([0.3, -0.2, 0.1, 0.05][:d + 1])
    for uu in (u, None):
        S = lam.reshape(-1, d, d) + 0.7 * (F.grad_u(mesh, G, uu) - q.reshape(-1, d, d))
        for ww in (w, None):
            a = F.load_vector(mesh, uu, S, ww, -1.0)
            b = F.load_vector_fast(mesh, uu, lam, q, 0.7, ww, -1.0)
            assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()
        for k in range(d + 1):
            e = np.zeros(d + 1); e[k] = 1.0
            a = F.load_vector(mesh, uu, None, e, 1.0)
            b = F.load_vector_fast(mesh, uu, None, None, 0.0, e, 1.0)
            assert np.abs(a - b).max() <= 1e-13 * max(np.abs(a).max(), 1e-300)


def test_fast_assembly_backend_reproduces_the_admm_iteration():
    """The whole ADMM iteration with fast_assembly=True equals the NumPy-assembled one (3D, numRefs = 1)."""
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    recs = []
    for fast in (False, True):
        p = ObstacleOptim(ug4_np.Backend(smoother="gs", fast_assembly=fast), 3, numRefs=1, grid=GRID3D).setup()
        p.set_sensitivity(p.synthetic_sensitivity(0.5))
        p.begin_step()
        recs.append(p.admm_iteration())
    a, b = recs
    assert len(a["newton"]) == len(b["newton"])
    for k in ("u_diff", "lambda_inc", "max_norm"):
        assert abs(a[k] - b[k]) <= 1e-11 * max(abs(a[k]), 1e-3), k


@pytest.mark.parametrize("dim,grid,refs", [(3, GRID3D, 1), (2, GRID2D, 2)])
@pytest.mark.parametrize("smoother", ["gs", "cheb", "jac"])
def test_c_solve_path_matches_numpy_statement(dim, grid, refs, smoother):
    """oracle/solver_c.c (RAP chain, V(3,3) cycle, dense-LU base solve, BiCGStab in C/OpenMP: bench.py's multi-threaded CPU
    baseline) against the NumPy/SciPy statement of the same algorithm: one thread -> same V-cycle and solution to rounding and
    the same iteration count; several threads (Gauss-Seidel becomes block-Jacobi across the thread blocks) ->
    the same solution at the solver tolerance."""
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import fem_np, ug4_np
    if fem_np.c_kernels() is None or not hasattr(fem_np.c_kernels(), "oracle_gmg_create"):
        pytest.skip("oracle/liboracle_c.so not built")
    outs = {}
    for name, kw in (("numpy", dict(threads=1)), ("c1", dict(threads=1, c_solver=True)), ("c3", dict(threads=3, c_solver=True))):
        ug = ug4_np.Backend(smoother=smoother, **kw)
        p = ObstacleOptim(ug, dim, numRefs=refs, grid=grid).setup()
        p.Hessian_ElemDisc.set_lambda_vol(0.2)
        p.u.from_numpy(0.01 * np.random.default_rng(0).standard_normal(p.u.v.size))
        DD = p.DeformationEquation_DomainDisc
        DD.adjust_solution(p.u)
        DD.assemble_jacobian(p.A_u_Hessian, p.u)
        p.Lu.from_numpy(np.random.default_rng(5).standard_normal(p.u.v.size), 2)
        DD.adjust_solution(p.Lu)
        s = p.SmallProblemRHS_Solver
        s.desc["convCheck"]["absolute"] = 1e-12
        s.init(p.A_u_Hessian, p.sigma)
        assert s.apply(p.sigma, p.Lu)
        z = p.sigma.to_numpy()
        s.vcycle(p.delta_u, p.Lu)
        outs[name] = (z, s.step(), p.delta_u.to_numpy())
        if name == "numpy":
            A, b = p.A_u_Hessian.to_scipy(), p.Lu.to_numpy()
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert outs["c1"][1] == outs["numpy"][1]
    assert rel(outs["c1"][0], outs["numpy"][0]) < 1e-12 and rel(outs["c1"][2], outs["numpy"][2]) < 1e-12
    assert abs(outs["c3"][1] - outs["numpy"][1]) <= 2 and np.linalg.norm(A @ outs["c3"][0] - b) < 1e-11
