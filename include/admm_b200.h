/* admm_b200.h -- C ABI of libadmm_b200.so: the B200-native backend for the hot path of
 * MultigridShapeOpt/admm_optim (P1 assembly -> GMG-preconditioned BiCGStab -> ADMM prox/dual).
 *
 * The reference's unchanged driver scripts call UG4 Lua-registered objects; the arithmetic behind those
 * objects is UG4 ugcore + the plugins FluidOptim/PLaplacian/ADMMOptim (3d_admm.lua:1-3), none of which is
 * in the reference tree.  Each entry point below names the Lua-level object/method (file:line of its call
 * site in the reference) that a UG4 plugin shim would bind to it.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ab_status on error; ab_last_error() gives the text.
 *     Nothing throws across this boundary.  Solver non-convergence is a VALUE (converged=0), not an error
 *     (script convention `if solver:apply(x,b) == false then ...`, 3d_admm.lua:980).
 *   - handles are opaque, owned by the caller, released with the matching *_destroy.
 *   - all floating point data is fp64, indices int32.  Host pointers unless the name says "device".
 *   - calls are serialised on one host thread per context (one Lua state per rank, 3d_admm.lua:25);
 *     all device work is enqueued on the stream given at context creation.
 *   - DoF layout: P1 index = vertex*dim + comp ; P0 index = element*ncomp + comp (l1..l9 row-major,
 *     3d_admm.lua:343-351).
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ab_context ab_context;       /* one GPU (+ optional communicator)                         */
typedef struct ab_domain ab_domain;         /* Domain + MultiGrid hierarchy      (3d_admm.lua:108-186)    */
typedef struct ab_space ab_space;           /* ApproximationSpace                (3d_admm.lua:329-333,367-370) */
typedef struct ab_vector ab_vector;         /* (Advanced)GridFunction            (3d_admm.lua:337-341,375-383) */
typedef struct ab_elemdisc ab_elemdisc;     /* element discretisations           (3d_admm.lua:393-694)    */
typedef struct ab_domaindisc ab_domaindisc; /* DomainDiscretization              (3d_admm.lua:460-463)    */
typedef struct ab_operator ab_operator;     /* AssembledLinearOperator           (3d_admm.lua:467)        */
typedef struct ab_solver ab_solver;         /* linear solvers                    (obstacle_optim_3d_util.lua:9-43, 3d_admm.lua:701-703) */

enum ab_status { AB_OK = 0, AB_ERR_ARG = -1, AB_ERR_CUDA = -2, AB_ERR_IO = -3, AB_ERR_STATE = -4, AB_ERR_UNSUPPORTED = -5 };
const char* ab_last_error(void);
int ab_version(void);

/* ---- context ------------------------------------------------------------------------------------------- */
/* InitUG(dim, AlgebraType("CPU",1)) (3d_admm.lua:105) selects the algebra backend; here: device + stream.
 * `stream` is a cudaStream_t (0 = legacy default stream).                                                   */
int ab_context_create(int device, void* stream, ab_context** out);
int ab_context_destroy(ab_context* ctx);
int ab_context_synchronize(ab_context* ctx);
/* number of kernels the library launched since context creation (bench.py's gpu_launches claim) */
int ab_context_launch_count(ab_context* ctx, int64_t* out);
/* backend tuning knobs (no reference counterpart): "spmv_variant" 0 (TMA tiles) | 1 (warp per row), "spmv_waves" >= 1,
   "graph" 0|1 (CUDA-graph replay of the BiCGStab iteration), "loop" 0|1 (whole loop as one graph with a conditional node), "pdl" 0|1 (programmatic dependent launch), "coarse_variant" 0|1, "spmv2d_lanes" 4|8 (lanes per 2x2-block row), "assembly_variant" 0 (row-owner gather) | 1 (atomic scatter),
   "l2_hint" 0|1, "tma_small_ctas" 1|2 */
int ab_context_set_tuning(ab_context* ctx, const char* key, int value);
/* Multi-GPU: one process per GPU (replaces UG4's pcl/MPI layer, `mpirun -np 4 ugshell ...` 3d_admm.lua:25).
 * `nccl_unique_id` is the 128-byte ncclUniqueId created on rank 0 and distributed by the host program.     */
int ab_context_init_comm(ab_context* ctx, int rank, int nranks, const void* nccl_unique_id);
int ab_nccl_unique_id(void* out128);
/* sum (max_op = 0) or max (max_op != 0) of n <= 64 host doubles over all ranks (host-side scalars of the driver) */
int ab_context_allreduce_host(ab_context* ctx, double* v, int n, int max_op);

/* ---- Domain / grid hierarchy ---------------------------------------------------------------------------- */
/* LoadDomain(dom, gridName)  3d_admm.lua:108-109 */
int ab_domain_load_ugx(ab_context* ctx, const char* path, ab_domain** out);
/* same domain from raw arrays (used when the grid is already in memory). `sp_edges`/`sp_faces` are the edges
 * (faces) whose subset differs from the element subset, rows sorted ascending.                              */
int ab_domain_create(ab_context* ctx, int dim, int nv, const double* xyz, int ne, const int32_t* elems,
                     int nsubsets, const char* const* subset_names, const int32_t* vsub, const int32_t* esub,
                     int n_sp_edges, const int32_t* sp_edges, const int32_t* sp_edges_sub,
                     int n_sp_faces, const int32_t* sp_faces, const int32_t* sp_faces_sub, ab_domain** out);
int ab_domain_destroy(ab_domain* dom);
/* util.refinement.CreateRegularHierarchy(dom, numRefs, false, balancerDesc)  3d_admm.lua:186 */
int ab_domain_refine(ab_domain* dom, int num_refs);
/* Multi-GPU (replaces the ParMETIS/pcl distribution of 3d_admm.lua:124-186): this rank's domain is its element
 * partition of the level-0 grid; after refinement the host program supplies, per level, the vertices shared with
 * each neighbour rank (neighbour ranks ascending, same canonical vertex order on both sides) and the owner mask.
 * A domain on a multi-rank context for which no interface is set stays a plain local domain (a problem too small
 * to be decomposed runs undivided).  Must be called before the first ApproximationSpace.                       */
int ab_domain_set_interface(ab_domain* dom, int level, int nneigh, const int32_t* neigh_ranks, const int32_t* offsets,
                            const int32_t* idx, const unsigned char* owned);
/* Hierarchical agglomeration of the coarse grid levels (the reference keeps level 0 on one process and widens the
 * process set level by level: balancerDesc.hierarchy, 3d_admm.lua:151-183): levels <= gather_level are held by rank 0
 * as one global hierarchy.  Every rank states the level; rank 0 also passes `coarse` (the GLOBAL grid refined
 * gather_level times, same context) and, concatenated over the ranks 0..nranks-1: the number of level-`gather_level`
 * vertices / matrix blocks of each rank, the local -> global vertex ids and the local block -> global block positions
 * (BSR order of ab_domain_level_pattern).  Other ranks pass NULL for all of them.  `coarse` stays owned by the caller
 * and must outlive `dom` (the hierarchies of `dom` keep pointers into its level structures).                       */
int ab_domain_set_gather(ab_domain* dom, int gather_level, ab_domain* coarse, const int32_t* nv_per_rank, const int32_t* l2g_cat,
                         const int64_t* nblk_per_rank, const int32_t* gpos_cat);
/* Matrix blocks shared with neighbour ranks on a decomposed level (both vertices on the interface, block present on both sides), in
 * an order both sides agree on: slots offsets[n]..offsets[n+1] of neighbour n refer to compact ids 0..nshared-1 (slot_block);
 * bpos / brow / mult give, per compact id, the local block position (BSR order), its row vertex and the number of ranks holding
 * the block.  With these lists solver:init computes the Gershgorin bound of the GLOBAL operator exactly (shared blocks are summed
 * before the absolute value), so the Chebyshev smoother does not depend on the partition; without them the bound is the (valid,
 * partition-dependent) sum of the per-rank row sums.  Optional; before the first ApproximationSpace.                      */
int ab_domain_set_block_interface(ab_domain* dom, int level, int nneigh, const int32_t* neigh_ranks, const int32_t* offsets,
                                  const int32_t* slot_block, int nshared, const int32_t* bpos, const int32_t* brow, const int32_t* mult);
/* P1 block pattern of a level (host side, no GPU needed): block rows = vertices, columns ascending, diagonal included;
 * nnzb = V + 2E.  rowptr (nv+1) / colidx (nnzb) may be NULL to query the size only.                              */
int ab_domain_level_pattern(ab_domain* dom, int level, int64_t* nnzb, int32_t* rowptr, int32_t* colidx);
/* vertex -> element incidence of a level (host side): ptr (nv+1), idx (ne*(dim+1)), elements ascending per vertex -- the fixed
 * summation order of the row-owner Hessian assembly (assemble_jacobian, 3d_admm.lua:972).                          */
int ab_domain_level_incidence(ab_domain* dom, int level, int32_t* ptr, int32_t* idx);
/* NVLink peer-to-peer interface sums (CUDA IPC; optional -- without it the exchanges use ncclSend/ncclRecv):
 * export this rank's receive window after the first ApproximationSpace exists, gather all handles / layouts on the
 * host, then connect.  remote_dst / remote_stride: one entry per (level, neighbour) in level-major, neighbour order. */
int ab_domain_p2p_export(ab_domain* dom, void* handle64, int64_t* level_base, int32_t* totals);
int ab_domain_p2p_connect(ab_domain* dom, const void* handles, const int64_t* remote_dst, const int64_t* remote_stride);
int ab_domain_p2p_status(ab_domain* dom, int* connected, int* error);
int ab_domain_num_levels(ab_domain* dom, int* out);
/* dom:domain_info() 3d_admm.lua:112,189 -- sizes of one level */
int ab_domain_level_info(ab_domain* dom, int level, int* dim, int* nv, int* ne, int* nedges, int* nv_coarse);
/* copy one level out (any pointer may be NULL): xyz[nv*dim] elems[ne*(dim+1)] vsub[nv] parent_a/b[nv-nv_coarse] */
int ab_domain_get_level(ab_domain* dom, int level, double* xyz, int32_t* elems, int32_t* vsub,
                        int32_t* parent_a, int32_t* parent_b);
int ab_domain_subset_index(ab_domain* dom, const char* name, int* out);
int ab_domain_subset_name(ab_domain* dom, int index, char* buf, int buflen);
/* boundary-subset ("special") edges / faces and element subsets of a level (needed to re-create a partition of a loaded grid) */
int ab_domain_special_info(ab_domain* dom, int level, int* n_sp_edges, int* n_sp_faces, int* nsubsets);
int ab_domain_get_special(ab_domain* dom, int level, int32_t* sp_edges, int32_t* sp_edges_sub, int32_t* sp_faces, int32_t* sp_faces_sub,
                          int32_t* esub);
/* TransformDomainByDisplacement(u, "u1,u2,u3")  3d_admm.lua:1333,1352 : vertex coordinates += u (all levels) */
int ab_transform_domain_by_displacement(ab_domain* dom, ab_vector* u);

/* ---- ApproximationSpace ---------------------------------------------------------------------------------- */
enum ab_space_kind { AB_SPACE_P0 = 0, AB_SPACE_P1 = 1 };
/* ApproximationSpace(dom):add_fct("u1,u2,u3","Lagrange",1) / ("l1..l9","Piecewise-Constant"); init_levels();
 * init_top_surface()   3d_admm.lua:329-333, 367-370 */
int ab_space_create(ab_domain* dom, int kind, int ncomp, ab_space** out);
int ab_space_destroy(ab_space* sp);
int ab_space_num_dofs(ab_space* sp, int64_t* out);

/* ---- GridFunction / vector algebra ------------------------------------------------------------------------- */
enum ab_storage { AB_PST_UNDEFINED = 0, AB_PST_CONSISTENT = 1, AB_PST_ADDITIVE = 2, AB_PST_UNIQUE = 4 };
int ab_vector_create(ab_space* sp, ab_vector** out);                 /* AdvancedGridFunction(space) 3d_admm.lua:375 */
int ab_vector_destroy(ab_vector* v);
int ab_vector_set(ab_vector* v, double c);                           /* gf:set(0.0) -> consistent, 3d_admm.lua:951 */
int ab_vector_upload(ab_vector* v, const double* host, int storage); /* host -> device (H2D)                      */
int ab_vector_download(ab_vector* v, double* host);                  /* device -> host (D2H)                      */
int ab_vector_device_ptr(ab_vector* v, void** out, int64_t* n);
int ab_vector_storage(ab_vector* v, int* out);                       /* has_storage_type_additive() 3d_admm.lua:978 */
int ab_vector_change_storage(ab_vector* v, int storage);             /* change_storage_type_to_consistent() 3d_admm.lua:912,982,1096 */
int ab_vec_scale_assign(ab_vector* dst, double a, ab_vector* src);   /* VecScaleAssign  3d_admm.lua:760,905,956 */
int ab_vec_scale_add2(ab_vector* dst, double a, ab_vector* x, double b, ab_vector* y); /* VecScaleAdd2 3d_admm.lua:976,1109 */
int ab_vec_prod(ab_vector* x, ab_vector* y, double* out);            /* VecProd  3d_admm.lua:994,1014-1059 */
int ab_vec_prod_multi(int n, ab_vector* const* xs, ab_vector* y, double* out); /* n VecProds against one y, one reduction */
int ab_vec_norm(ab_vector* x, double* out);                          /* VecNorm  2d_admm.lua:1131 */
int ab_l2norm(ab_vector* v, int comp, double* out);                  /* L2Norm(gf,"u1",4,"outer") 3d_admm.lua:1137-1146,1237-1251 */
int ab_l2norm_all(ab_vector* v, double* out_per_comp);               /* all components in one pass */

/* ---- element discretisations ------------------------------------------------------------------------------- */
enum ab_disc_kind {
    AB_DISC_DEFORMATION_EQUATION = 1,   /* DeformationEquation("u1,u2,u3","outer")                3d_admm.lua:393 */
    AB_DISC_DEFORMATION_RHS = 2,        /* DeformationEquationRHS                                  3d_admm.lua:407 */
    AB_DISC_DEFORMATION_LARGE_RHS = 3,  /* DeformationEquationLargeProblemRHS                      3d_admm.lua:472 */
    AB_DISC_VOLUME_CONSTRAINT = 4,      /* VolumeConstraintSecondDerivative / SecondDerivativeVolume 3d:559, 2d:564 */
    AB_DISC_BARYCENTER_CONSTRAINT = 5,  /* SecondDerivativeBarycenter / XBarycenterConstraintSecondDerivative + set_index 3d:577-617 */
    AB_DISC_MASS_MODEL = 6,             /* MassModel("l1..l9","outer")                             3d_admm.lua:652 */
    AB_DISC_LAMBDA_UPDATE = 7           /* LambdaUpdate("l1..l9","outer")                          3d_admm.lua:677 */
};
enum ab_disc_param {
    AB_PARAM_LAMBDA_VOL = 1,      /* set_lambda_vol            3d_admm.lua:394 */
    AB_PARAM_LAMBDA_BARY_X = 2,   /* set_lambda_barycenter     3d_admm.lua:395 */
    AB_PARAM_LAMBDA_BARY_Y = 3,
    AB_PARAM_LAMBDA_BARY_Z = 4,
    AB_PARAM_STEP_LENGTH = 5,     /* set_step_length           3d_admm.lua:396 */
    AB_PARAM_TAU = 6,             /* set_tau                   3d_admm.lua:411,473 */
    AB_PARAM_MULT_VOL = 7,        /* set_multiplier_vol        3d_admm.lua:1081 */
    AB_PARAM_MULT_BX = 8,         /* set_multiplier_bx/by/bz   3d_admm.lua:1082-1084 */
    AB_PARAM_MULT_BY = 9,
    AB_PARAM_MULT_BZ = 10,
    AB_PARAM_INDEX = 11,          /* set_index(1|2|3)          3d_admm.lua:578,597,617 */
    AB_PARAM_QUAD_ORDER = 12,     /* set_quad_order(1)         3d_admm.lua:393 (P1: every rule of order>=1 is exact) */
    AB_PARAM_SCALING = 13,        /* set_scaling               2d_admm.lua:393 (J'' terms, only with second_order) */
    AB_PARAM_HIGH_ORDER_SCALING = 14, /* set_high_order_scaling 2d_admm.lua:394 */
    AB_PARAM_SECOND_ORDER = 15    /* set_second_order(b2ndOrder) 2d_admm.lua:389 ; !=0 is AB_ERR_UNSUPPORTED (needs NS fields) */
};
enum ab_disc_import {
    AB_IMPORT_DEFORMATION = 1,    /* set_deformation_d1..3 + set_deformation_vector_d1..3  3d_admm.lua:399-405 */
    AB_IMPORT_LAMBDA = 2,         /* set_lambda00..22                                       3d_admm.lua:423-431 */
    AB_IMPORT_Q = 3               /* set_q00..22 / set_qproj00..22                          3d_admm.lua:434-442,683-691 */
};
int ab_elemdisc_create(ab_space* sp, int kind, ab_elemdisc** out);
int ab_elemdisc_destroy(ab_elemdisc* d);
int ab_elemdisc_set_param(ab_elemdisc* d, int param, double value);
int ab_elemdisc_get_param(ab_elemdisc* d, int param, double* value);
int ab_elemdisc_bind(ab_elemdisc* d, int import, ab_vector* v);

/* ---- DomainDiscretization ------------------------------------------------------------------------------------ */
int ab_domaindisc_create(ab_space* sp, ab_domaindisc** out);           /* DomainDiscretization(space) 3d_admm.lua:460 */
int ab_domaindisc_destroy(ab_domaindisc* dd);
int ab_domaindisc_add_elemdisc(ab_domaindisc* dd, ab_elemdisc* d);     /* dd:add(elemDisc)            3d_admm.lua:461,463 */
/* DirichletBoundary():add(value,"u1","inlet") + dd:add(dirichlet)   3d_admm.lua:445-462 ; value must be 0 */
int ab_domaindisc_add_dirichlet(ab_domaindisc* dd, const char* subset, int comp, double value);
int ab_domaindisc_assemble_jacobian(ab_domaindisc* dd, ab_operator* A, ab_vector* u); /* 3d_admm.lua:972,1008,1090,899 */
int ab_domaindisc_assemble_defect(ab_domaindisc* dd, ab_vector* d, ab_vector* u);     /* 3d_admm.lua:954,973,1091,900,1221 */
int ab_domaindisc_adjust_solution(ab_domaindisc* dd, ab_vector* u);                   /* 3d_admm.lua:465,971,1087,1122 */

/* ---- AssembledLinearOperator ----------------------------------------------------------------------------------- */
int ab_operator_create(ab_domaindisc* dd, ab_operator** out);          /* AssembledLinearOperator(dd) 3d_admm.lua:467 */
int ab_operator_destroy(ab_operator* A);
int ab_operator_apply(ab_operator* A, ab_vector* y, ab_vector* x);     /* y = A x (SpMV; used inside the solvers) */
/* BSR/diagonal download for tests: block size b, nb block rows, nnzb blocks. Pointers may be NULL to query sizes. */
int ab_operator_info(ab_operator* A, int* block, int64_t* nb, int64_t* nnzb);
int ab_operator_download(ab_operator* A, int32_t* rowptr, int32_t* colidx, double* vals);

/* ---- solvers ------------------------------------------------------------------------------------------------------ */
enum ab_smoother { AB_SMOOTHER_CHEBYSHEV = 1, AB_SMOOTHER_JACOBI = 2 };
typedef struct ab_gmg_desc {      /* util.oo.linear_solver descriptor, obstacle_optim_3d_util.lua:10-40 */
    int smoother;                 /* "gs" in the reference (u3:16): replaced by the stated equivalent */
    int pre_smooth, post_smooth;  /* preSmooth = 3, postSmooth = 3      u3:25-26 */
    int base_level;               /* baseLevel = 0, SuperLU base solver u3:19-21 */
    int rap;                      /* rap = true                         u3:27 (0 is AB_ERR_UNSUPPORTED) */
    int max_iterations;           /* iterations = 3000 (2D: 2000)       u3:34 / u2:35 */
    double abs_tol;               /* absolute = 1e-10 (2D: 1e-12)       u3:35 / u2:36 */
    double red_tol;               /* reduction = 0                      u3:36 */
    int verbose;                  /* u3:37 */
    double cheb_ratio;            /* Chebyshev interval [lmax/ratio, lmax]; 0 -> default 6 */
    double jacobi_damp;           /* 0 -> default 0.66 */
} ab_gmg_desc;
int ab_solver_create_bicgstab_gmg(ab_space* sp, const ab_gmg_desc* desc, ab_solver** out);
/* CG(); set_preconditioner(Jacobi(0.66)); set_convergence_check(ConvCheck(2000,1e-9,0,true))  3d_admm.lua:701-703 */
int ab_solver_create_cg_jacobi(ab_space* sp, double damp, int max_iterations, double abs_tol, double red_tol,
                               int verbose, ab_solver** out);
int ab_solver_destroy(ab_solver* s);
int ab_solver_init(ab_solver* s, ab_operator* A, ab_vector* x);              /* solver:init(A,x)   3d_admm.lua:979 */
int ab_solver_apply(ab_solver* s, ab_vector* x, ab_vector* b, int* converged);              /* :apply(x,b) 3d_admm.lua:980 */
int ab_solver_apply_return_defect(ab_solver* s, ab_vector* x, ab_vector* b, int* converged); /* 3d_admm.lua:1095 */
int ab_solver_step(ab_solver* s, int* out);                                   /* solver:step()      3d_admm.lua:1160 */
int ab_solver_last_defect(ab_solver* s, double* out);
/* one application of the GMG preconditioner z = M^-1 r (the "GMG V-cycle ms" metric of BASELINE.json) */
int ab_solver_vcycle(ab_solver* s, ab_vector* z, ab_vector* r);
/* level sizes of the hierarchy a solver was initialised on: nb / nnzb per level (bench roofline bytes) */
int ab_solver_level_info(ab_solver* s, int level, int64_t* nb, int64_t* nnzb);

/* ---- ADMM / plugin free functions -------------------------------------------------------------------------------- */
int ab_project_frobenius(ab_vector* q_projected, ab_vector* q, double sigma);   /* Testing(...)  3d_admm.lua:910 */
int ab_project_spectral(ab_vector* q_projected, ab_vector* q, double sigma);    /* ProjectWithSpectralNorm 2d_admm.lua:902 */
int ab_max_frobenius_norm(ab_vector* u, double* out);                           /* MaximumFrobeniusNorm 3d_admm.lua:916 */
int ab_max_spectral_norm(ab_vector* u, double* out);                            /* MaxSpectralNorm 2d_admm.lua:901 */
int ab_volume_defect(ab_vector* u, double reference_volume, double* out);       /* VolumeDefect 3d_admm.lua:780,1167 */
int ab_barycenter_defect(ab_vector* u, double* out_dim);                        /* BarycenterDefect 3d_admm.lua:1168 */
int ab_set_zero_away_from_subset(ab_vector* v, const char* subset);             /* SetZeroAwayFromSubset 3d_admm.lua:817 */

#ifdef __cplusplus
}
#endif
#endif /* ADMM_B200_H */
