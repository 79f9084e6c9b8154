"""Replay of the reference drivers' hot path: the ADMM loop of 3d_admm.lua:875-1304 / 2d_admm.lua:868-1253
with its Newton/Schur loop (3d:940-1202, 2d:926-1171) and the object wiring of 3d:327-713 / 2d:343-690.

The reference scripts are Lua run by `ugshell`; neither exists in this image, so the call sequence is
restated here in Python against a backend object `ug` that offers the UG4 Lua-registered names
(admm_optim_b200.ug4.Backend on the GPU; the tests inject the NumPy oracle twin).  Statement order,
signs and quirks follow the scripts line by line (cited); everything outside the deformation /
extension subproblem (Navier-Stokes, adjoint, drag, VTK) is NOT here: J' (SensitivityGF) is an input.

This module must not import oracle/ (dependency injection only).
"""
from __future__ import annotations

import math

from . import gnuplot
from .schur import Matrix


def linear_solver(ug, domainDisc, approxSpace, vrb, dim):
    """util.oo.linear_solver  -- obstacle_optim_3d_util.lua:9-43 (3D) / obstacle_optim_util.lua:9-44 (2D)."""
    LinSolverDesc = {
        "type": "bicgstab",
        "precond": {
            "type": "gmg",
            "smoother": "gs",
            "adaptive": False,
            "approxSpace": approxSpace,
            "baseLevel": 0,
            "gatheredBaseSolverIfAmbiguous": False,
            "baseSolver": ug.SuperLU(),
            "cycle": "V",
            "discretization": domainDisc,
            "preSmooth": 3,
            "postSmooth": 3,
            "rap": True,
            "transfer": "std",
            "debug": False,
        },
        "convCheck": {
            "type": "standard",
            "iterations": 3000 if dim == 3 else 2000,       # u3:34 / u2:35
            "absolute": 1e-10 if dim == 3 else 1e-12,       # u3:35 / u2:36
            "reduction": 0.0,
            "verbose": True if dim == 3 else vrb,            # u3:37 / u2:38
        },
    }
    return ug.util.solver.CreateSolver(LinSolverDesc)


DEFAULTS_3D = dict(numRefs=2, admmSteps=2, sigma_threshold=0.3, scaling=1.0, admm_tolerance=1e-2, step_length=1.0, tau=1.0,
                   nsMaxIts=30, nsTol=1e-9, lambda_vol=0.0, lambda_x=0.0, lambda_y=0.0, lambda_z=0.0,
                   grid="./grids/box_3D_elongated.ugx")                                     # 3d_admm.lua:46-70
DEFAULTS_2D = dict(numRefs=3, admmSteps=1000, sigma_threshold=0.3, scaling=1.0, admm_tolerance=1e-2, step_length=1.0, tau=1.0,
                   nsMaxIts=30, nsTol=1e-9, lambda_vol=0.0, lambda_x=0.0, lambda_y=0.0, lambda_z=0.0,
                   admm_gradient_tolerance=0.05, normName="frobenius", nsRelLuTol=1e-12, nsAbsLuTol=1e-12,
                   nsRelLlambdaTol=1e-12, nsAbsLlambdaTol=1e-12, b2ndOrder=False,
                   grid="./grids/refined.ugx")                                              # 2d_admm.lua:43-87


class ObstacleOptim:
    """Hot-path half of one `ugshell -ex {2d,3d}_admm.lua` session."""

    def __init__(self, ug, dim, verbose=False, solver_verbose=False, trace_dir=None, newton_output=False, trace_first_row=1, **params):
        """trace_dir: directory that receives the reference's trace files (__ADMMStats_step_<k>_.txt, and with newton_output --
        the scripts' -bNewtonOutput -- __NewtonStats_step_<k>_.txt / __NewtonIterations_step_<k>_.txt), written by the same
        statements at the same places as in the scripts.  trace_first_row: see gnuplot.write_data."""
        self.ug, self.dim, self.verbose, self.solver_verbose = ug, dim, verbose, solver_verbose
        self.trace_dir, self.newton_output, self.trace_first_row = trace_dir, bool(newton_output), trace_first_row
        self.sensitivity_callback = None
        self.P = dict(DEFAULTS_3D if dim == 3 else DEFAULTS_2D)
        unknown = set(params) - set(self.P)
        if unknown:
            raise ValueError("unknown parameters: %s" % sorted(unknown))
        self.P.update(params)
        self.m = dim + 1                                                       # 3d:34
        self.step = 0
        self.p_solver_failure = False
        self.trace = []

    def log(self, *a):
        if self.verbose:
            print(*a)

    # ------------------------------------------------------------------------------------------
    # setup: 3d_admm.lua:105-186 (grid) and :327-713 (Lagrange-matrix + deformation objects)
    # ------------------------------------------------------------------------------------------
    def setup(self):
        ug, dim, P = self.ug, self.dim, self.P
        ug.InitUG(dim, ug.AlgebraType("CPU", 1))                               # 3d:105
        self.dom = dom = ug.Domain()                                           # 3d:108
        ug.LoadDomain(dom, P["grid"])                                          # 3d:109
        ug.util.refinement.CreateRegularHierarchy(dom, P["numRefs"], False, None)   # 3d:186

        ucmp = ["u1", "u2", "u3"][:dim]
        self.ucmps = ",".join(ucmp)
        lcmp = ["l%d" % (k + 1) for k in range(dim * dim)]
        self.lcmps = ",".join(lcmp)

        # LAGRANGE MATRIX space (3d:329-341)
        self.Lambda_ApproxSpace = LS = ug.ApproximationSpace(dom)
        LS.add_fct(self.lcmps, "Piecewise-Constant")
        LS.init_levels(); LS.init_top_surface()
        self.lambda_piecewise = ug.AdvancedGridFunction(LS); self.lambda_piecewise.set(0.0)
        self.q_piecewise = ug.AdvancedGridFunction(LS); self.q_piecewise.set(0.0)
        self.q_projected = ug.AdvancedGridFunction(LS); self.q_projected.set(0.0)
        self.rhs_piecewise = ug.AdvancedGridFunction(LS); self.rhs_piecewise.set(0.0)
        self.temp1_piecewise = ug.AdvancedGridFunction(LS); self.temp1_piecewise.set(0.0)
        lam_g = {(i, j): ug.GlobalGridFunctionNumberData(self.lambda_piecewise, lcmp[i * dim + j]) for i in range(dim) for j in range(dim)}   # 3d:343-351
        qpr_g = {(i, j): ug.GlobalGridFunctionNumberData(self.q_projected, lcmp[i * dim + j]) for i in range(dim) for j in range(dim)}       # 3d:355-363

        # DEFORMATION space (3d:367-389)
        self.DeformationSpace_ApproxSpace = DS = ug.ApproximationSpace(dom)
        DS.add_fct(self.ucmps, "Lagrange", 1)
        DS.init_levels(); DS.init_top_surface()
        gf = lambda: ug.AdvancedGridFunction(DS)
        for name in ("delta_u", "u", "u_converged", "u_diff", "u_old", "u_negative", "sigma", "Lu", "u_zeros"):
            f = gf(); f.set(0.0); setattr(self, name, f)
        u_grad = [ug.GlobalGridFunctionGradientData(self.u, c) for c in ucmp]
        u_val = [ug.GlobalGridFunctionNumberData(self.u, c) for c in ucmp]

        def bind_u(disc):
            for k in range(dim):
                getattr(disc, "set_deformation_d%d" % (k + 1))(u_val[k])
                getattr(disc, "set_deformation_vector_d%d" % (k + 1))(u_grad[k])

        def bind_tensor(disc, prefix, table):
            for (i, j), imp in table.items():
                getattr(disc, "%s%d%d" % (prefix, i, j))(imp)

        lam0 = (P["lambda_x"], P["lambda_y"], P["lambda_z"])
        # Hessian (3d:393-405)
        self.Hessian_ElemDisc = H = ug.DeformationEquation(self.ucmps, "outer")
        if dim == 3: H.set_quad_order(1)
        if dim == 2:
            H.set_second_order(P["b2ndOrder"]); H.set_scaling(P["scaling"]); H.set_high_order_scaling(1.0)   # 2d:389-394
        H.set_lambda_vol(P["lambda_vol"]); H.set_lambda_barycenter(*lam0); H.set_step_length(P["step_length"])
        bind_u(H)
        # RHS of the small problem (3d:407-442)
        self.RHS_ElemDisc = R = ug.DeformationEquationRHS(self.ucmps, "outer")
        if dim == 3: R.set_quad_order(1)
        R.set_lambda_vol(P["lambda_vol"]); R.set_lambda_barycenter(*lam0); R.set_step_length(P["step_length"]); R.set_tau(P["tau"])
        bind_u(R); bind_tensor(R, "set_lambda", lam_g); bind_tensor(R, "set_q", qpr_g)
        # Dirichlet (3d:445-457)
        self.Dirich = Dir = ug.DirichletBoundary()
        for subset in ("inlet", "wall", "outlet"):
            for c in ucmp:
                Dir.add(0, c, subset)
        # DomainDisc of the small problem (3d:460-467)
        self.DeformationEquation_DomainDisc = DD = ug.DomainDiscretization(DS)
        DD.add(H); DD.add(Dir); DD.add(R)
        DD.adjust_solution(self.sigma); DD.adjust_solution(self.u)
        self.A_u_Hessian = ug.AssembledLinearOperator(DD)
        # large problem (3d:472-517)
        self.LargeRHS_ElemDisc = LR = ug.DeformationEquationLargeProblemRHS(self.ucmps, "outer")
        if dim == 3: LR.set_quad_order(1)
        LR.set_tau(1.0); LR.set_lambda_vol(P["lambda_vol"]); LR.set_lambda_barycenter(*lam0); LR.set_step_length(P["step_length"])
        bind_u(LR); bind_tensor(LR, "set_lambda", lam_g); bind_tensor(LR, "set_q", qpr_g)
        self.Large_DomainDisc = LD = ug.DomainDiscretization(DS)
        LD.add(H); LD.add(LR); LD.add(Dir)
        self.A_Large = ug.AssembledLinearOperator(LD)
        LD.adjust_solution(self.delta_u); LD.adjust_solution(self.u)
        # J' (3d:545-547): assembled on the UG4/CPU side (Sensitivity ElemDisc, OUT OF SCOPE) -> input vector here
        self.SensitivityGF = gf()
        # constraint discs + their M-solve discretisations (3d:551-631)
        self.B_vector, self.B_DomainDisc, self.A_B, self.t_B = [], [], [], []
        for i in range(self.m):
            B = gf(); B.set(0.0)
            if i == 0:
                E = (ug.VolumeConstraintSecondDerivative if dim == 3 else ug.SecondDerivativeVolume)(self.ucmps, "outer")   # 3d:559 / 2d:564
            elif i == 3:
                E = ug.XBarycenterConstraintSecondDerivative(self.ucmps, "outer")   # 3d:616 (z uses another class name)
            else:
                E = ug.SecondDerivativeBarycenter(self.ucmps, "outer")              # 3d:577,596
            if dim == 3: E.set_quad_order(1)
            if i > 0: E.set_index(i)
            bind_u(E)
            BD = ug.DomainDiscretization(DS)
            BD.add(E); BD.add(H); BD.add(Dir)
            A = ug.AssembledLinearOperator(BD)
            t = gf(); t.set(0.0)
            BD.adjust_solution(t)
            self.B_vector.append(B); self.B_DomainDisc.append(BD); self.A_B.append(A); self.t_B.append(t)
        # Schur complement data (3d:634-647)
        m = self.m
        self.BTranspose_sigma = Matrix(m, 1); self.L_lambda = Matrix(m, 1); self.Lambda = Matrix(m, 1)
        for i, v in enumerate([P["lambda_vol"], P["lambda_x"], P["lambda_y"], P["lambda_z"]][:m]):
            self.Lambda[i][0] = v
        self.S = Matrix(m, m); self.rhs = Matrix(m, 1); self.DeltaLambda = Matrix(m, 1)
        self.MinusLu_BdeltaLambda = gf(); self.MinusLu_BdeltaLambda.set(0.0)
        # mass model (3d:652-674) and lambda update (3d:677-694)
        self.MassModel_ElemDisc = MM = ug.MassModel(self.lcmps, "outer")
        bind_u(MM); bind_tensor(MM, "set_lambda", lam_g)
        self.MassModel_DomainDisc = MD = ug.DomainDiscretization(LS)
        MD.add(MM)
        self.DiagQ = ug.AssembledLinearOperator(MD)
        self.LambdaUpdate_ElemDisc = LU = ug.LambdaUpdate(self.lcmps, "outer")
        for k in range(dim):
            getattr(LU, "set_deformation_vector_d%d" % (k + 1))(u_grad[k])
        bind_tensor(LU, "set_qproj", qpr_g)
        self.LambdaUpdate_DomainDisc = LUD = ug.DomainDiscretization(LS)
        LUD.add(LU)
        # solvers (3d:701-709)
        self.ADMMDiagonal_Solver = ug.CG()
        self.ADMMDiagonal_Solver.set_preconditioner(ug.Jacobi(0.66))
        self.ADMMDiagonal_Solver.set_convergence_check(ug.ConvCheck(2000, 1.0e-9, 0.0, self.solver_verbose))
        mk = lambda dd: self._mk_solver(dd)
        self.SmallProblemRHS_Solver = mk(DD)
        self.B_Solver = [mk(bd) for bd in self.B_DomainDisc]
        self.LargeProblem_Solver = mk(LD)
        # 3d:780  ReferenceVolume = VolumeDefect(u,0,...)
        self.ReferenceVolume = ug.VolumeDefect(self.u, 0, "outer", self.ucmps, 4, False, 1, False)
        self.maximum_norm = 0.0
        self.admm_steps = 0
        return self

    def _mk_solver(self, dd):
        s = linear_solver(self.ug, dd, self.DeformationSpace_ApproxSpace, False, self.dim)
        if hasattr(s, "desc") and not isinstance(s.desc, dict):
            s.desc.verbose = int(self.solver_verbose)
        elif hasattr(s, "desc"):
            s.desc["convCheck"]["verbose"] = self.solver_verbose
        return s

    # ------------------------------------------------------------------------------------------
    # J' enters here (stands for 3d:816-817: Jprime assemble_defect + SetZeroAwayFromSubset)
    # ------------------------------------------------------------------------------------------
    def set_sensitivity(self, jprime_host, scaling=None):
        """J' as assembled by the UG4/CPU side with Jprime_ElemDisc:set_step_length(scaling) (3d:816-817, 1286-1287)."""
        self.SensitivityGF.from_numpy(jprime_host, 2)      # additive, like assemble_defect output
        self.ug.SetZeroAwayFromSubset(self.SensitivityGF, self.ucmps, "obstacle_surface")
        self._jprime_host, self._jprime_scaling = jprime_host, (self.P["scaling"] if scaling is None else scaling)

    def refresh_sensitivity(self, scaling):
        """3d:1286-1288 / 2d:1234-1236: after a fake convergence the scripts re-assemble J' with the doubled step length
        (Jprime_ElemDisc:set_step_length(scaling); assemble_defect; SetZeroAwayFromSubset).  The Sensitivity ElemDisc lives on the
        UG4/CPU side (out of scope): `sensitivity_callback(scaling)` asks it for the new host vector; without a callback the stored
        J' is rescaled, which is what a J' linear in its step length gives."""
        if self.sensitivity_callback is not None:
            self.set_sensitivity(self.sensitivity_callback(scaling), scaling)
        else:
            import numpy as np
            self.set_sensitivity(np.asarray(self._jprime_host) * (scaling / self._jprime_scaling), scaling)

    def synthetic_sensitivity(self, amplitude=0.5):
        """Deterministic stand-in for the shape derivative J' (no Navier-Stokes here): a smooth normal
        traction on the obstacle surface, J'_v = amplitude * prof(x_v) * g_vol'(0)_v, where g_vol'(0)_v = int_Gamma phi_v nu ds
        is the area-weighted vertex normal -- obtained from the volume-constraint disc at u = 0 through the public API.
        Returns the host array (what the UG4/CPU side would hand over)."""
        import numpy as np
        d = self.dim
        self.u.set(0.0)
        self.B_vector[0].set(0.0)
        self.B_DomainDisc[0].assemble_defect(self.B_vector[0], self.u)
        g = self.B_vector[0].to_numpy().reshape(-1, d) * (-1.0 if d == 3 else 1.0)
        self.B_vector[0].set(0.0)
        top = self.dom.num_levels() - 1
        X = self.dom.get_level(top)["xyz"] if hasattr(self.dom, "get_level") else self.dom.top.xyz
        prof = 0.3 * np.sin(np.pi * X[:, 0]) + 0.5 * np.cos(2 * np.pi * X[:, 1]) + (0.4 * np.cos(2 * np.pi * X[:, d - 1]) if d == 3 else 0.2)
        return (amplitude * prof[:, None] * g).ravel()

    # 3d:843-874: start of a (repeated) optimisation step
    def begin_step(self):
        m = self.m
        for i in range(m):
            self.L_lambda[i][0] = 0.0
        if not self.p_solver_failure:
            self.ug.VecScaleAssign(self.u_converged, 1.0, self.u)
        self.u.set(0.0); self.lambda_piecewise.set(0.0)
        self.sigma.set(0.0)
        for i in range(m):
            self.Lambda[i][0] = 0.0
        self.p_solver_failure = False
        self.admm_steps = 0
        self.admm_trace = []
        # 3d:867-873: the tables behind __ADMMStats_step_<k>_.txt, indexed by admm_steps like the Lua tables
        self.vADMM = {k: {} for k in ("Step", "Scaling", "Sigma", "Udiff", "LambdaInc", "MaxFrobNorm", "SigmaMinusMaxNorm")}
        self.vNS = None

    # ------------------------------------------------------------------------------------------
    # one pass of the ADMM loop body, 3d_admm.lua:876-1303 (2d_admm.lua:869-1252)
    # returns a dict with the __ADMMStats columns + Newton diagnostics; sets self.p_solver_failure
    # ------------------------------------------------------------------------------------------
    def admm_iteration(self):
        ug, dim, P, m = self.ug, self.dim, self.P, self.m
        for i in range(m):
            self.L_lambda[i][0] = 0.0                                           # 3d:884-887
        for disc in (self.Hessian_ElemDisc, self.RHS_ElemDisc, self.LargeRHS_ElemDisc):   # 3d:889-894
            disc.set_lambda_vol(0.0); disc.set_lambda_barycenter(0.0, 0.0, 0.0)
        # ---- q-step: mass model solve (3d:897-905) ----
        self.q_piecewise.set(0.0); self.rhs_piecewise.set(0.0); self.temp1_piecewise.set(0.0)
        self.MassModel_DomainDisc.assemble_jacobian(self.DiagQ, self.u_negative)
        self.MassModel_DomainDisc.assemble_defect(self.rhs_piecewise, self.u_negative)
        self.ADMMDiagonal_Solver.init(self.DiagQ, self.q_piecewise)
        if not self.ADMMDiagonal_Solver.apply(self.q_piecewise, self.rhs_piecewise):
            self.log("Mass model solver failed at admm step", self.admm_steps)
            self.p_solver_failure = True
            return None
        ug.VecScaleAssign(self.q_piecewise, -1.0, self.q_piecewise)
        # ---- prox: projection (3d:910-916 / 2d:896-903) ----
        sigma_threshold = P["sigma_threshold"]
        if dim == 3 or P["normName"] == "frobenius":
            ug.Testing(self.q_projected, self.q_piecewise, self.lcmps, sigma_threshold)
            self.maximum_norm = ug.MaximumFrobeniusNorm(self.u_old, self.ucmps, "outer", 4)
        elif P["normName"] == "spectral":
            self.maximum_norm = ug.MaxSpectralNorm(self.u_old, self.ucmps, "outer", 4)
            ug.ProjectWithSpectralNorm(self.q_projected, self.q_piecewise, self.lcmps, sigma_threshold)
        self.q_projected.change_storage_type_to_consistent()
        # ---- u-step: Newton on the KKT system (3d:920-1202) ----
        newton = self.newton_loop()
        if self.p_solver_failure:
            self.log("ADMM LOOP::solver failure, breaking")
            return None
        # ---- lambda-step (3d:1219-1225) ----
        self.LambdaUpdate_DomainDisc.assemble_defect(self.temp1_piecewise, self.u_negative)
        ug.VecScaleAssign(self.temp1_piecewise, -1.0, self.temp1_piecewise)
        self.temp1_piecewise.change_storage_type_to_consistent()
        ug.VecScaleAdd2(self.lambda_piecewise, 1.0, self.lambda_piecewise, 1.0, self.temp1_piecewise)
        self.lambda_piecewise.change_storage_type_to_consistent()
        # ---- bookkeeping (3d:1231-1253) ----
        self.u_diff.set(0.0)
        ug.VecScaleAdd2(self.u_diff, 1.0, self.u, -1.0, self.u_old)
        ug.VecScaleAssign(self.u_old, 1.0, self.u)
        u_diff_norm = math.sqrt(sum(ug.L2Norm(self.u_diff, c, 4, "outer") ** 2 for c in self.ucmps.split(",")))
        lambda_inc_norm = math.sqrt(sum(ug.L2Norm(self.temp1_piecewise, c, 4, "outer") ** 2 for c in self.lcmps.split(",")))
        rec = dict(step=self.step, admm_step=self.admm_steps, scaling=P["scaling"], sigma=sigma_threshold, u_diff=u_diff_norm,
                   lambda_inc=lambda_inc_norm, max_norm=self.maximum_norm, sigma_minus_max=sigma_threshold - self.maximum_norm,
                   newton=newton, Lambda=[self.Lambda[i][0] for i in range(m)], L_lambda=[self.L_lambda[i][0] for i in range(m)])
        self.admm_trace.append(rec)
        self.log("ADMM LOOP::STEP=%d  MaxNorm=%.12g  u_diff=%.12g  lambda_inc=%.12g  newton its=%d" %
                 (self.admm_steps, self.maximum_norm, u_diff_norm, lambda_inc_norm, len(newton)))
        # ---- trace tables + file (3d:1265-1276 / 2d:1213-1223) ----
        if getattr(self, "vADMM", None) is not None:
            k = self.admm_steps
            for name, val in (("Step", k), ("Scaling", P["scaling"]), ("Sigma", sigma_threshold), ("Udiff", u_diff_norm), ("LambdaInc", lambda_inc_norm),
                              ("MaxFrobNorm", self.maximum_norm), ("SigmaMinusMaxNorm", sigma_threshold - self.maximum_norm)):
                self.vADMM[name][k] = val
            if self.trace_dir is not None:
                import os
                V = self.vADMM
                gnuplot.write_data(os.path.join(self.trace_dir, "__ADMMStats_step_%d_.txt" % self.step),
                                   [V["Step"], V["Scaling"], V["Sigma"], V["Udiff"], V["LambdaInc"], V["MaxFrobNorm"], V["SigmaMinusMaxNorm"]],
                                   False, first_row=self.trace_first_row)
        # ---- convergence check (3d:1279-1302) ----
        tol = P["admm_tolerance"]
        grad_tol = 0.05 if dim == 3 else P["admm_gradient_tolerance"]
        rec["converged"] = bool(lambda_inc_norm < tol and u_diff_norm < tol and (sigma_threshold - self.maximum_norm > -grad_tol * sigma_threshold))
        rec["fake"] = bool(rec["converged"] and (sigma_threshold - self.maximum_norm > grad_tol * sigma_threshold))
        self.admm_steps += 1
        return rec

    # ------------------------------------------------------------------------------------------
    # Newton / Schur loop, 3d_admm.lua:940-1202 (2d_admm.lua:926-1171)
    # ------------------------------------------------------------------------------------------
    def newton_loop(self):
        ug, dim, P, m = self.ug, self.dim, self.P, self.m
        three_d = dim == 3
        B, t_B = self.B_vector, self.t_B
        Lu, sigma, delta_u, u = self.Lu, self.sigma, self.delta_u, self.u
        jsign = 1.0 if three_d else -1.0                                        # 3d:976 (+J') vs 2d:956 (-J')
        lin_u = self.u if three_d else self.u_zeros                             # 3d:954 vs 2d:939 (argument only; imports rule)
        ns_i = 1
        recs = []
        Norm_Lu_0 = Norm_Llambda_0 = 0.0
        # 3d:923-934: the tables behind __NewtonStats_step_<k>_.txt / __NewtonIterations_step_<k>_.txt (re-created every ADMM iteration)
        self.vNS = {k: {} for k in ("Step", "NormSum", "NormDeltaUp", "NormDeltaLambda", "Lu2Norm", "RHS", "Large", "Bvol", "Bx", "By", "Bz")}
        while ns_i <= P["nsMaxIts"]:
            self.vNS["Step"][ns_i] = ns_i                                        # 3d:942
            self.MinusLu_BdeltaLambda.set(0.0); Lu.set(0.0)                      # 3d:944-948
            delta_u.set(0.0); sigma.set(0.0)
            for i in range(m):
                self.rhs[i][0] = 0.0; self.DeltaLambda[i][0] = 0.0; self.BTranspose_sigma[i][0] = 0.0
            for i in range(m):                                                   # 3d:952-959
                B[i].set(0.0)
            for i in range(m):
                self.B_DomainDisc[i].assemble_defect(B[i], lin_u)
            if three_d:
                for i in range(m):
                    ug.VecScaleAssign(B[i], -1.0, B[i])
            # (1) A sigma = L_u  (3d:971-982)
            DD = self.DeformationEquation_DomainDisc
            DD.adjust_solution(sigma)
            if not three_d: self.Hessian_ElemDisc.set_second_order(P["b2ndOrder"])
            DD.assemble_jacobian(self.A_u_Hessian, u)
            DD.assemble_defect(Lu, u if three_d else self.u_zeros)
            ug.VecScaleAdd2(Lu, 1.0, Lu, jsign, self.SensitivityGF)
            if not Lu.has_storage_type_additive():
                raise RuntimeError("CATASTROPHIC FAILURE::RHS NOT ADDITIVE")      # 3d:978
            self.SmallProblemRHS_Solver.init(self.A_u_Hessian, sigma)
            if not self.SmallProblemRHS_Solver.apply(sigma, Lu):
                self.log("A.sigma=Lu, solver failed, at step", ns_i); self.p_solver_failure = True; break
            if three_d:
                ug.VecScaleAssign(sigma, -1.0, sigma)                            # 3d:981
            Lu.change_storage_type_to_consistent()
            # (2) B.sigma and the small rhs (3d:993-1001)
            for i in range(m):
                self.BTranspose_sigma[i][0] = ug.VecProd(B[i], sigma)
            for i in range(m):
                self.rhs[i][0] = -1.0 * self.L_lambda[i][0] - self.BTranspose_sigma[i][0]
            # (3) constraint solves and Schur complement columns (3d:1005-1059)
            failed = False
            for i in range(m):
                B[i].set(0.0)
                self.B_DomainDisc[i].assemble_defect(B[i], lin_u)
                self.B_DomainDisc[i].assemble_jacobian(self.A_B[i], lin_u)
                self.B_Solver[i].init(self.A_B[i], t_B[i])
                if not self.B_Solver[i].apply(t_B[i], B[i]):
                    self.log("Solver for B[%d] failed" % (i + 1)); self.p_solver_failure = True; failed = True; break
                if three_d:
                    ug.VecScaleAssign(B[i], -1.0, B[i])                          # 3d:1012
                else:
                    ug.VecScaleAssign(t_B[i], -1.0, t_B[i])                      # 2d:989
                for r in range(m):
                    self.S[r][i] = ug.VecProd(B[r], t_B[i])                      # 3d:1014-1017
            if failed:
                break
            # (4) host Schur algebra with lua-matrix semantics (3d:1063-1078)
            inverse_S = self.S.invert()
            if inverse_S is None:
                self.log("Schur complement singular"); self.p_solver_failure = True; break
            self.DeltaLambda = inverse_S.mul(self.rhs)
            # (5) large problem (3d:1081-1107)
            LR = self.LargeRHS_ElemDisc
            LR.set_multiplier_vol(self.DeltaLambda[0][0]); LR.set_multiplier_bx(self.DeltaLambda[1][0]); LR.set_multiplier_by(self.DeltaLambda[2][0])
            if three_d: LR.set_multiplier_bz(self.DeltaLambda[3][0])
            self.MinusLu_BdeltaLambda.set(0.0)
            LD = self.Large_DomainDisc
            LD.adjust_solution(delta_u)
            LD.assemble_jacobian(self.A_Large, self.u_zeros)
            LD.assemble_defect(self.MinusLu_BdeltaLambda, self.u_zeros)
            ug.VecScaleAdd2(self.MinusLu_BdeltaLambda, 1.0, self.MinusLu_BdeltaLambda, jsign, self.SensitivityGF)
            self.LargeProblem_Solver.init(self.A_Large, delta_u)
            if not self.LargeProblem_Solver.apply_return_defect(delta_u, self.MinusLu_BdeltaLambda):
                self.log("Large problem solver failed at step", ns_i); self.p_solver_failure = True; break
            if three_d:
                self.MinusLu_BdeltaLambda.change_storage_type_to_consistent()    # 3d:1096
            LR.set_multiplier_vol(0.0); LR.set_multiplier_bx(0.0); LR.set_multiplier_by(0.0)
            if three_d: LR.set_multiplier_bz(0.0)
            LD.adjust_solution(delta_u)
            # (6) updates (3d:1109-1122)
            ug.VecScaleAdd2(u, 1.0, u, -1.0 if three_d else 1.0, delta_u)        # 3d:1109 vs 2d:1068
            for i in range(m):
                self.Lambda[i][0] = self.Lambda[i][0] + self.DeltaLambda[i][0]
            DD.adjust_solution(u)
            ns_i += 1
            if ns_i > P["nsMaxIts"]:
                self.log("NEWTON METHOD DID NOT CONVERGE"); self.p_solver_failure = True; break   # 3d:1126-1132
            # (7) diagnostics (3d:1137-1185)
            cm = self.ucmps.split(",")
            lu_norm_sum = math.sqrt(sum(ug.L2Norm(Lu, c, 4, "outer") ** 2 for c in cm))
            delta_u_norm_sum = math.sqrt(sum(ug.L2Norm(delta_u, c, 4, "outer") ** 2 for c in cm))
            delta_lambda_norm = math.sqrt(sum(self.DeltaLambda[j][0] ** 2 for j in range(m)))
            self.L_lambda[0][0] = ug.VolumeDefect(u, self.ReferenceVolume, "outer", self.ucmps, 4, False, 1, False)
            bary = ug.BarycenterDefect(u, self.ucmps, "outer", 4)
            for k in range(dim):
                self.L_lambda[1 + k][0] = bary[k]
            llambda_norm = math.sqrt(sum(self.L_lambda[i][0] ** 2 for i in range(m)))
            lam = [self.Lambda[i][0] for i in range(m)]
            for disc in (self.Hessian_ElemDisc, self.RHS_ElemDisc, self.LargeRHS_ElemDisc):
                disc.set_lambda_vol(lam[0])
                disc.set_lambda_barycenter(lam[1], lam[2], lam[3] if three_d else 0)
            recs.append(dict(ns=ns_i - 1, delta_u=delta_u_norm_sum, delta_lambda=delta_lambda_norm, lu=lu_norm_sum,
                             L_lambda=[self.L_lambda[i][0] for i in range(m)], DeltaLambda=[self.DeltaLambda[i][0] for i in range(m)],
                             S=[[self.S[r][c] for c in range(m)] for r in range(m)],
                             its=dict(rhs=self.SmallProblemRHS_Solver.step(), large=self.LargeProblem_Solver.step(),
                                      B=[s.step() for s in self.B_Solver])))
            V, its_ = self.vNS, recs[-1]["its"]                                  # 3d:1154-1164
            V["NormDeltaUp"][ns_i - 1] = delta_u_norm_sum; V["NormDeltaLambda"][ns_i - 1] = delta_lambda_norm
            V["NormSum"][ns_i - 1] = 0.0; V["Lu2Norm"][ns_i - 1] = lu_norm_sum
            V["RHS"][ns_i - 1] = its_["rhs"]; V["Large"][ns_i - 1] = its_["large"]; V["Bvol"][ns_i - 1] = its_["B"][0]
            V["Bx"][ns_i - 1] = its_["B"][1]; V["By"][ns_i - 1] = its_["B"][2]
            if three_d: V["Bz"][ns_i - 1] = its_["B"][3]
            self.log("#   %d DELTA_U INCREMENT NORM IS: %.6e   DELTA_LAMBDA INCREMENT NORM IS: %.6e   |Lu| %.6e  its %s" %
                     (ns_i, delta_u_norm_sum, delta_lambda_norm, lu_norm_sum, recs[-1]["its"]))
            # (8) stop (3d:1198 ; 2d:1163-1166)
            if ns_i - 1 == 1:
                Norm_Lu_0, Norm_Llambda_0 = lu_norm_sum, llambda_norm
            if delta_lambda_norm <= P["nsTol"]:
                break
            if not three_d:
                if lu_norm_sum < P["nsAbsLuTol"] and llambda_norm < P["nsAbsLlambdaTol"]:
                    break
                if Norm_Lu_0 > 0 and Norm_Llambda_0 > 0 and lu_norm_sum / Norm_Lu_0 < P["nsRelLuTol"] and llambda_norm / Norm_Llambda_0 < P["nsRelLlambdaTol"]:
                    break
        return recs

    # the surrounding ADMM loop control, 3d:875,1279-1311 (2d:868,1227-1259)
    def run_admm(self, max_steps=None):
        """The ADMM loop of one optimisation step: `while admm_steps < admmSteps` with the scripts' exits --
          * converged (3d:1279)                      -> break ("convergence break case")
          * fake convergence (3d:1281-1289)          -> scaling *= 2, admm_steps = 0 (then the increment at the end of the body),
                                                        J' refreshed for the new scaling (refresh_sensitivity), loop goes on
          * admm_steps == admmSteps (3d:1296-1301)   -> step marked for repetition (p_solver_failure).  In the scripts this test sits
                                                        inside `while admm_steps < admmSteps` before the increment, so it can never
                                                        fire; it is kept at the same place for fidelity
          * solver failure                           -> break
        followed by the Newton trace files of the LAST ADMM iteration (3d:1307-1311).  Returns the trace."""
        self.begin_step()
        limit = self.P["admmSteps"] if max_steps is None else max_steps
        while self.admm_steps < limit:
            rec = self.admm_iteration()                                          # increments admm_steps at its end (3d:1302)
            if rec is None:
                break
            self.admm_steps -= 1                                                 # ... so step back to where the scripts run their checks
            if rec["converged"]:
                if rec["fake"]:
                    self.P["scaling"] = self.P["scaling"] * 2.0                  # 3d:1284
                    self.admm_steps = 0                                          # 3d:1285
                    self.refresh_sensitivity(self.P["scaling"])                  # 3d:1286-1288
                    self.log("ADMM LOOP::t'was fake convergence, scaling= %g" % self.P["scaling"])
                else:
                    self.admm_steps += 1
                    break
            if self.admm_steps == limit:                                         # 3d:1296-1301 (unreachable, see above)
                self.p_solver_failure = True
                self.admm_steps = 0
                break
            self.admm_steps += 1                                                 # 3d:1302
        self.write_newton_traces()
        return self.admm_trace

    def write_deformation(self, directory="."):
        """`if bOutputMesh then ... vtkWriter:select_nodal("u1,u2,u3","u"); vtkWriter:print("u", u, step+1, step+1, false)` (3d:1400-1406):
        the accepted deformation of the step as <directory>/u_t<step+1>.vtu on the current coordinates."""
        import os
        from .vtk import VTKOutput
        w = self.ug.VTKOutput() if hasattr(self.ug, "VTKOutput") else VTKOutput(self.ug)
        w.clear_selection()
        w.select_nodal(self.ucmps, "u")
        return w.print(os.path.join(directory, "u"), self.u, self.step + 1, self.step + 1, False)

    def write_debug_mesh(self, directory="."):
        """`if bDebugOutput then SaveGridLevelToFile(dom:grid(), dom:subset_handler(), numRefs, "Mesh_lev"..numRefs.."_step"..step..".ugx")`
        (3d:795, 2d:788): the top grid level with its current (transformed) coordinates."""
        import os
        top = self.dom.num_levels() - 1
        return self.ug.SaveGridLevelToFile(self.dom.grid(), self.dom.subset_handler(), top,
                                           os.path.join(directory, "Mesh_lev%d_step%d.ugx" % (top, self.step)))

    def write_newton_traces(self):
        """3d:1307-1311: gnuplot.write_data of the Newton tables of the last ADMM iteration (only with -bNewtonOutput true)."""
        if not (self.newton_output and self.trace_dir is not None and self.vNS):
            return
        import os
        V = self.vNS
        gnuplot.write_data(os.path.join(self.trace_dir, "__NewtonStats_step_%d_.txt" % self.step),
                           [V["Step"], V["NormSum"], V["NormDeltaUp"], V["NormDeltaLambda"], V["Lu2Norm"]])
        gnuplot.write_data(os.path.join(self.trace_dir, "__NewtonIterations_step_%d_.txt" % self.step),
                           [V["Step"], V["RHS"], V["Bvol"], V["Bx"], V["By"], V["Large"]])
