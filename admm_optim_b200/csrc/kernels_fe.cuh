// Finite-element kernels (fp64, sm_100a): P1 Hessian / load-vector assembly on simplices with atomic
// scatter into BSR, Dirichlet handling, P0 tensor (ADMM prox / dual) kernels, element reductions.
// Formulas: DESIGN.md "Model" (restated independently in oracle/fem_np.py); the reference ships no
// element code, the call sites served are cited per kernel.
#pragma once
#include "common.cuh"

namespace ab {

template <int D>
struct Elem {
    int v[D + 1];
    double G[D + 1][D];   // P1 gradients
    double vol;           // |det J| / D!
    double xbar[D];       // centroid
};

template <int D>
__device__ __forceinline__ void elem_load(const int* __restrict__ elems, const double* __restrict__ xyz, int64_t e, Elem<D>& E) {
    double X[D + 1][D];
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        E.v[a] = elems[e * (D + 1) + a];
#pragma unroll
        for (int c = 0; c < D; ++c) X[a][c] = xyz[(int64_t)E.v[a] * D + c];
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double s = X[0][c];
#pragma unroll
        for (int a = 1; a <= D; ++a) s += X[a][c];
        E.xbar[c] = s / (D + 1);
    }
    if constexpr (D == 2) {
        const double a = X[1][0] - X[0][0], b = X[1][1] - X[0][1], c = X[2][0] - X[0][0], d = X[2][1] - X[0][1];
        const double det = a * d - c * b, inv = 1.0 / det;
        E.G[1][0] = d * inv;  E.G[1][1] = -c * inv;
        E.G[2][0] = -b * inv; E.G[2][1] = a * inv;
        E.vol = fabs(det) * 0.5;
    } else {
        double e1[3], e2[3], e3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { e1[c] = X[1][c] - X[0][c]; e2[c] = X[2][c] - X[0][c]; e3[c] = X[3][c] - X[0][c]; }
        double c23[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
        double c31[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
        double c12[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double det = e1[0] * c23[0] + e1[1] * c23[1] + e1[2] * c23[2], inv = 1.0 / det;
#pragma unroll
        for (int c = 0; c < 3; ++c) { E.G[1][c] = c23[c] * inv; E.G[2][c] = c31[c] * inv; E.G[D][c] = c12[c] * inv; }
        E.vol = fabs(det) * (1.0 / 6.0);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double s = 0.0;
#pragma unroll
        for (int a = 1; a <= D; ++a) s += E.G[a][c];
        E.G[0][c] = -s;
    }
}

// grad u (constant per element) and mean of u; u may be null (= 0)
template <int D>
__device__ __forceinline__ void elem_gradu(const Elem<D>& E, const double* __restrict__ u, double (&gu)[D][D], double (&ubar)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        ubar[i] = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) gu[i][j] = 0.0;
    }
    if (!u) return;
#pragma unroll
    for (int a = 0; a <= D; ++a) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double ua = u[(int64_t)E.v[a] * D + i];
            ubar[i] += ua;
#pragma unroll
            for (int j = 0; j < D; ++j) gu[i][j] += ua * E.G[a][j];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) ubar[i] *= 1.0 / (D + 1);
}

// F = I + grad u ; C = cof(F) ; returns det F
template <int D>
__device__ __forceinline__ double cof_det(const double (&gu)[D][D], double (&F)[D][D], double (&C)[D][D]) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) F[i][j] = gu[i][j] + (i == j ? 1.0 : 0.0);
    if constexpr (D == 2) {
        C[0][0] = F[1][1]; C[0][1] = -F[1][0];
        C[1][0] = -F[0][1]; C[1][1] = F[0][0];
        return F[0][0] * F[1][1] - F[0][1] * F[1][0];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            C[i][0] = F[i1][1] * F[i2][D - 1] - F[i1][D - 1] * F[i2][1];
            C[i][1] = F[i1][D - 1] * F[i2][0] - F[i1][0] * F[i2][D - 1];
            C[i][D - 1] = F[i1][0] * F[i2][1] - F[i1][1] * F[i2][0];
        }
        return F[0][0] * C[0][0] + F[0][1] * C[0][1] + F[0][D - 1] * C[0][D - 1];
    }
}

// position of block (v_a, v_b) in the BSR row of v_a, for every element  (setup, once per level)
template <int D>
__global__ void k_elem_pos(int64_t ne, const int* __restrict__ elems, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                           int* __restrict__ pos) {
    constexpr int N = D + 1;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ne * N * N; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (N * N);
        const int ab_ = (int)(t - e * N * N), a = ab_ / N, b = ab_ - a * N;
        const int va = elems[e * N + a], vb = elems[e * N + b];
        int lo = rowptr[va], hi = rowptr[va + 1] - 1;
        while (lo < hi) {
            int m = (lo + hi) >> 1;
            if (colidx[m] < vb) lo = m + 1; else hi = m;
        }
        pos[t] = lo;
    }
}

struct HessParams {
    double c;          // coefficient of the vector Laplacian (step_length)
    double lam_vol;    // Lambda_vol
    double lam_b[3];   // Lambda_barycenter
    int has_lam;
};

// DeformationEquation jacobian (3d_admm.lua:393-405; assemble_jacobian :972,:1008,:1024,:1039,:1052,:1090):
//   K[(a,i),(b,j)] = vol*( c d_ij G_a.G_b + wc * eps_ijk (G_a x G_b).F_k + sum_k Lam_k (d_ik (C G_b)_j + d_jk (C G_a)_i)/(D+1) )
// Dirichlet rows/columns are skipped here (symmetric elimination); k_dirichlet_diag sets the unit diagonal.
template <int D>
__global__ void __launch_bounds__(128) k_assemble_hessian(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                          const double* __restrict__ u, const int* __restrict__ pos,
                                                          const unsigned char* __restrict__ dirmask, HessParams P, double* __restrict__ vals) {
    constexpr int N = D + 1, DD = D * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        unsigned char m[N];
#pragma unroll
        for (int a = 0; a < N; ++a) m[a] = dirmask ? dirmask[E.v[a]] : 0;
        double F[D][D], C[D][D], CG[N][D], wc = 0.0;
        if (P.has_lam) {
            double gu[D][D], ubar[D];
            elem_gradu<D>(E, u, gu, ubar);
            cof_det<D>(gu, F, C);
            wc = P.lam_vol;
#pragma unroll
            for (int k = 0; k < D; ++k) wc += P.lam_b[k] * (E.xbar[k] + ubar[k]);
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) s += C[i][j] * E.G[a][j];
                    CG[a][i] = s;
                }
        }
#pragma unroll
        for (int a = 0; a < N; ++a) {
#pragma unroll
            for (int b = 0; b < N; ++b) {
                double gg = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) gg += E.G[a][c] * E.G[b][c];
                double B[D][D];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) B[i][j] = (i == j) ? P.c * gg : 0.0;
                if (P.has_lam) {
                    if constexpr (D == 2) {
                        const double dt = wc * (E.G[a][0] * E.G[b][1] - E.G[a][1] * E.G[b][0]);
                        B[0][1] += dt;
                        B[1][0] -= dt;
                    } else {
                        const double w0 = E.G[a][1] * E.G[b][D - 1] - E.G[a][D - 1] * E.G[b][1];
                        const double w1 = E.G[a][D - 1] * E.G[b][0] - E.G[a][0] * E.G[b][D - 1];
                        const double w2 = E.G[a][0] * E.G[b][1] - E.G[a][1] * E.G[b][0];
                        double t[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) t[k] = wc * (w0 * F[k % D][0] + w1 * F[k % D][1] + w2 * F[k % D][D - 1]);
                        B[0][1] += t[2];     B[0][D - 1] -= t[1];
                        B[1][0] -= t[2];     B[1][D - 1] += t[0];
                        B[D - 1][0] += t[1]; B[D - 1][1] -= t[0];
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const double lk = P.lam_b[k] * (1.0 / (D + 1));
#pragma unroll
                        for (int j = 0; j < D; ++j) B[k][j] += lk * CG[b][j];
#pragma unroll
                        for (int i = 0; i < D; ++i) B[i][k] += lk * CG[a][i];
                    }
                }
                double* dst = vals + (int64_t)pos[e * N * N + a * N + b] * DD;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        if (!P.has_lam && i != j) continue;
                        if (((m[a] >> i) & 1) || ((m[b] >> j) & 1)) continue;
                        atomicAdd(dst + i * D + j, E.vol * B[i][j]);
                    }
            }
        }
    }
}

// Row-owner assembly of the same operator without atomics (default): a group of LPR lanes owns one block row.
//   phase 1: lane i computes the geometry (and, with multipliers, F, cof F G_b, the weight wc) of the i-th element incident to the
//            row's vertex and parks the record in shared memory -- every element is evaluated once per incident vertex;
//   phase 2: lane j owns the j-th block (v, w_j) of the row: it walks the parked elements, picks those that contain w_j and adds
//            their (a, b) contribution to nine accumulators in a fixed order, then writes its block: the row is written once, fully
//            coalesced, never read -- no memset, no atomics, bitwise reproducible.
// Needs the vertex -> element incidence (v2e, elements ascending) instead of the per-element block positions of the scatter kernel.
template <int D>
struct HessRec {
    double G[D + 1][D];
    double CG[D + 1][D];
    double F[D][D];
    double vol, wc;
    int v[D + 1];
};

template <int D, int LPR>
__global__ void __launch_bounds__(128) k_assemble_hessian_rows(int nv, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                               const int* __restrict__ v2e_ptr, const int* __restrict__ v2e_idx,
                                                               const int* __restrict__ elems, const double* __restrict__ xyz,
                                                               const double* __restrict__ u, const unsigned char* __restrict__ dirmask,
                                                               HessParams P, double* __restrict__ vals) {
    constexpr int N = D + 1, DD = D * D;
    extern __shared__ __align__(16) unsigned char sm_hess[];
    HessRec<D>* recs = reinterpret_cast<HessRec<D>*>(sm_hess) + (threadIdx.x / LPR) * LPR;       // this group's LPR records
    const int gl = threadIdx.x % LPR;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR, ngroups = (int64_t)gridDim.x * blockDim.x / LPR;
    const int64_t rounds = (nv + ngroups - 1) / ngroups;                 // every lane of a warp runs the same number of rounds
    for (int64_t rd = 0; rd < rounds; ++rd) {
        const int64_t row = gid + rd * ngroups;
        const bool row_on = row < nv;
        const int es = row_on ? v2e_ptr[row] : 0, ee = row_on ? v2e_ptr[row + 1] : 0;
        const int bs = row_on ? rowptr[row] : 0, be = row_on ? rowptr[row + 1] : 0;
        int nel = ee - es, nbl = be - bs;
        // warp-uniform trip counts (groups of one warp own different rows)
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) { nel = max(nel, __shfl_xor_sync(0xffffffffu, nel, o)); nbl = max(nbl, __shfl_xor_sync(0xffffffffu, nbl, o)); }
        const unsigned char mrow = (row_on && dirmask) ? dirmask[row] : 0;
        for (int b0 = 0; b0 < nbl; b0 += LPR) {
            const int blk = bs + b0 + gl;
            const bool blk_on = blk < be;
            const int w = blk_on ? colidx[blk] : -1;
            double acc[DD];
#pragma unroll
            for (int k = 0; k < DD; ++k) acc[k] = 0.0;
            for (int e0 = 0; e0 < nel; e0 += LPR) {
                __syncwarp();                                            // the previous chunk's records have been consumed
                if (es + e0 + gl < ee) {
                    const int64_t e = v2e_idx[es + e0 + gl];
                    Elem<D> E;
                    elem_load<D>(elems, xyz, e, E);
                    HessRec<D>& R = recs[gl];
#pragma unroll
                    for (int a = 0; a < N; ++a) {
                        R.v[a] = E.v[a];
#pragma unroll
                        for (int c = 0; c < D; ++c) R.G[a][c] = E.G[a][c];
                    }
                    R.vol = E.vol;
                    R.wc = 0.0;
                    if (P.has_lam) {
                        double gu[D][D], ubar[D], F[D][D], C[D][D];
                        elem_gradu<D>(E, u, gu, ubar);
                        cof_det<D>(gu, F, C);
                        double wc = P.lam_vol;
#pragma unroll
                        for (int k = 0; k < D; ++k) wc += P.lam_b[k] * (E.xbar[k] + ubar[k]);
                        R.wc = wc;
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = 0; j < D; ++j) R.F[i][j] = F[i][j];
#pragma unroll
                        for (int a = 0; a < N; ++a)
#pragma unroll
                            for (int i = 0; i < D; ++i) {
                                double sC = 0.0;
#pragma unroll
                                for (int j = 0; j < D; ++j) sC += C[i][j] * E.G[a][j];
                                R.CG[a][i] = sC;
                            }
                    }
                }
                __syncwarp();
                const int cnt = min(LPR, ee - es - e0);                  // records of this group (<= 0: none)
                if (blk_on) {
                    for (int r = 0; r < cnt; ++r) {
                        const HessRec<D>& R = recs[r];
                        int a = -1, b = -1;
#pragma unroll
                        for (int q = 0; q < N; ++q) { if (R.v[q] == (int)row) a = q; if (R.v[q] == w) b = q; }
                        if (b < 0) continue;                             // the element does not contain w (a >= 0 always: it is incident to the row)
                        double Ga[D], Gb[D];
#pragma unroll
                        for (int c = 0; c < D; ++c) { Ga[c] = R.G[a][c]; Gb[c] = R.G[b][c]; }
                        double gg = 0.0;
#pragma unroll
                        for (int c = 0; c < D; ++c) gg += Ga[c] * Gb[c];
                        double B[D][D];
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = 0; j < D; ++j) B[i][j] = (i == j) ? P.c * gg : 0.0;
                        if (P.has_lam) {
                            const double wc = R.wc;
                            if constexpr (D == 2) {
                                const double dt = wc * (Ga[0] * Gb[1] - Ga[1] * Gb[0]);
                                B[0][1] += dt;
                                B[1][0] -= dt;
                            } else {
                                const double w0 = Ga[1] * Gb[D - 1] - Ga[D - 1] * Gb[1];
                                const double w1 = Ga[D - 1] * Gb[0] - Ga[0] * Gb[D - 1];
                                const double w2 = Ga[0] * Gb[1] - Ga[1] * Gb[0];
                                double t[3];
#pragma unroll
                                for (int k = 0; k < 3; ++k) t[k] = wc * (w0 * R.F[k % D][0] + w1 * R.F[k % D][1] + w2 * R.F[k % D][D - 1]);
                                B[0][1] += t[2];     B[0][D - 1] -= t[1];
                                B[1][0] -= t[2];     B[1][D - 1] += t[0];
                                B[D - 1][0] += t[1]; B[D - 1][1] -= t[0];
                            }
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const double lk = P.lam_b[k] * (1.0 / (D + 1));
#pragma unroll
                                for (int j = 0; j < D; ++j) B[k][j] += lk * R.CG[b][j];
#pragma unroll
                                for (int i = 0; i < D; ++i) B[i][k] += lk * R.CG[a][i];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = 0; j < D; ++j) acc[i * D + j] += R.vol * B[i][j];
                    }
                }
            }
            if (blk_on) {
                const unsigned char mw = dirmask ? dirmask[w] : 0;
                double* dst = vals + (int64_t)blk * DD;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const bool dead = ((mrow >> i) & 1) || ((mw >> j) & 1);      // Dirichlet rows / columns (unit diagonal: k_dirichlet_diag)
                        dst[i * D + j] = dead ? 0.0 : acc[i * D + j];
                    }
            }
        }
    }
}

// unit diagonal for Dirichlet dofs (DirichletBoundary adjust_jacobian, 3d_admm.lua:445-462)
template <int D>
__global__ void k_dirichlet_diag(int nv, const unsigned char* __restrict__ dirmask, const unsigned char* __restrict__ owned,
                                 const int* __restrict__ diagpos, double* __restrict__ vals) {
    // multi-GPU (owned != null): the operator is additive over the ranks, so only the owner of a shared vertex sets the 1
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nv * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t / D), c = (int)(t - (int64_t)v * D);
        if ((dirmask[v] >> c) & 1) vals[(int64_t)diagpos[v] * D * D + c * D + c] = (!owned || owned[v]) ? 1.0 : 0.0;
    }
}
// Dirichlet dofs of a vector -> 0 (adjust_defect / adjust_solution with value 0, 3d_admm.lua:465-466,971)
template <int D>
__global__ void k_zero_dirichlet(int nv, const unsigned char* __restrict__ dirmask, double* __restrict__ x) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nv * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t / D), c = (int)(t - (int64_t)v * D);
        if ((dirmask[v] >> c) & 1) x[t] = 0.0;
    }
}
// SetZeroAwayFromSubset(gf, cmps, "obstacle_surface")  3d_admm.lua:817,1288
__global__ void k_zero_away_from_subset(int64_t n, int D, const int* __restrict__ vsub, int subset, double* __restrict__ x) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        if (vsub[t / D] != subset) x[t] = 0.0;
}

struct LoadParams {
    int use_S;         // S = lam + tau (grad u - q)
    double tau;
    double w[4];       // (w_vol, w_1..w_D)
    int has_w;
    double sign;
};

// Generic P1 element load vector, scattered with atomics (assemble_defect, 3d_admm.lua:954,973,1006,1091):
//   f[(a,i)] = sign*vol*( ((S + wc C) G_a)_i + w_i det F/(D+1) )
template <int D>
__global__ void __launch_bounds__(128) k_assemble_load(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                       const double* __restrict__ u, const double* __restrict__ lam,
                                                       const double* __restrict__ q, LoadParams P, double* __restrict__ out) {
    constexpr int N = D + 1, DD = D * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        double gu[D][D], ubar[D], M[D][D];
        elem_gradu<D>(E, u, gu, ubar);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double s = 0.0;
                if (P.use_S) s = (lam ? lam[e * DD + i * D + j] : 0.0) + P.tau * (gu[i][j] - (q ? q[e * DD + i * D + j] : 0.0));
                M[i][j] = s;
            }
        double detF = 0.0;
        if (P.has_w) {
            double F[D][D], C[D][D];
            detF = cof_det<D>(gu, F, C);
            double wc = P.w[0];
#pragma unroll
            for (int k = 0; k < D; ++k) wc += P.w[1 + k] * (E.xbar[k] + ubar[k]);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) M[i][j] += wc * C[i][j];
        }
        const double sv = P.sign * E.vol;
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) s += M[i][j] * E.G[a][j];
                if (P.has_w) s += P.w[1 + i] * detF * (1.0 / (D + 1));
                atomicAdd(out + (int64_t)E.v[a] * D + i, sv * s);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// P0 tensor kernels (ADMM prox / dual): one thread per element, D*D contiguous doubles per element
// ---------------------------------------------------------------------------------------------
// MassModel (3d_admm.lua:652-674, :899-900): diag = |T| per P0 dof, defect = -|T| (grad u + lam)
template <int D>
__global__ void k_mass_model(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz, const double* __restrict__ u,
                             const double* __restrict__ lam, double* __restrict__ diag, double* __restrict__ rhs) {
    constexpr int DD = D * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        double gu[D][D], ubar[D];
        elem_gradu<D>(E, u, gu, ubar);
#pragma unroll
        for (int k = 0; k < DD; ++k) {
            if (diag) diag[e * DD + k] = E.vol;
            if (rhs) rhs[e * DD + k] = -E.vol * (gu[k / D][k % D] + (lam ? lam[e * DD + k] : 0.0));
        }
    }
}
// LambdaUpdate (3d_admm.lua:677-694, :1221): defect = -tau (grad u - q_proj)
template <int D>
__global__ void k_lambda_update(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz, const double* __restrict__ u,
                                const double* __restrict__ q, double tau, double* __restrict__ out) {
    constexpr int DD = D * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        double gu[D][D], ubar[D];
        elem_gradu<D>(E, u, gu, ubar);
#pragma unroll
        for (int k = 0; k < DD; ++k) out[e * DD + k] = -tau * (gu[k / D][k % D] - (q ? q[e * DD + k] : 0.0));
    }
}
// Testing(q_projected,q_piecewise,cmps,sigma) (3d_admm.lua:910): projection onto the Frobenius ball per element
template <int DD>
__global__ void k_project_frobenius(int64_t ne, double sigma, const double* __restrict__ q, double* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        double v[DD], s = 0.0;
#pragma unroll
        for (int k = 0; k < DD; ++k) { v[k] = q[e * DD + k]; s += v[k] * v[k]; }
        const double nrm = sqrt(s);
        const double f = nrm > sigma ? sigma / nrm : 1.0;
#pragma unroll
        for (int k = 0; k < DD; ++k) out[e * DD + k] = v[k] * f;
    }
}
// closed-form 2x2 SVD pieces: Q = R(phi) diag(s1,s2) R(theta), s1 >= |s2|
__device__ __forceinline__ void svd2(const double* Q, double& s1, double& s2, double& theta, double& phi) {
    const double E = 0.5 * (Q[0] + Q[3]), Fh = 0.5 * (Q[0] - Q[3]), Gh = 0.5 * (Q[2] + Q[1]), H = 0.5 * (Q[2] - Q[1]);
    const double q_ = hypot(E, H), r_ = hypot(Fh, Gh);
    s1 = q_ + r_;
    s2 = q_ - r_;
    const double a1 = atan2(Gh, Fh), a2 = atan2(H, E);
    theta = 0.5 * (a2 - a1);
    phi = 0.5 * (a2 + a1);
}
// ProjectWithSpectralNorm (2d_admm.lua:902): clip singular values of the 2x2 at sigma
__global__ void k_project_spectral(int64_t ne, double sigma, const double* __restrict__ q, double* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        double Q[4] = {q[e * 4], q[e * 4 + 1], q[e * 4 + 2], q[e * 4 + 3]};
        double s1, s2, th, ph;
        svd2(Q, s1, s2, th, ph);
        const double t1 = fmin(s1, sigma), t2 = fmin(fmax(s2, -sigma), sigma);
        const double cp = cos(ph), sp = sin(ph), ct = cos(th), st = sin(th);
        out[e * 4 + 0] = cp * t1 * ct - sp * t2 * st;
        out[e * 4 + 1] = -cp * t1 * st - sp * t2 * ct;
        out[e * 4 + 2] = sp * t1 * ct + cp * t2 * st;
        out[e * 4 + 3] = -sp * t1 * st + cp * t2 * ct;
    }
}

// ---------------------------------------------------------------------------------------------
// element reductions
// ---------------------------------------------------------------------------------------------
// MaximumFrobeniusNorm (3d_admm.lua:916) / MaxSpectralNorm (2d_admm.lua:901):  max_T |grad u_T|
template <int D, int SPECTRAL>
__global__ void __launch_bounds__(256) k_max_grad_norm(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                       const double* __restrict__ u, double* partials, unsigned int* ticket, double* out) {
    double mx[1] = {0.0};
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        double gu[D][D], ubar[D];
        elem_gradu<D>(E, u, gu, ubar);
        double val;
        if (SPECTRAL && D == 2) {
            double Q[4] = {gu[0][0], gu[0][1], gu[1][0], gu[1][1]}, s1, s2, th, ph;
            svd2(Q, s1, s2, th, ph);
            val = s1;
        } else {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) s += gu[i][j] * gu[i][j];
            val = sqrt(s);
        }
        mx[0] = fmax(mx[0], val);
    }
    grid_reduce<1, 1>(mx, partials, ticket, out);
}
// VolumeDefect + BarycenterDefect (3d_admm.lua:1167-1168): out[0] = sum vol det F ; out[1+k] = sum vol det F (xbar_k+ubar_k)
template <int D>
__global__ void __launch_bounds__(256) k_volume_barycenter(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                           const double* __restrict__ u, double* partials, unsigned int* ticket, double* out) {
    double acc[D + 1];
#pragma unroll
    for (int k = 0; k <= D; ++k) acc[k] = 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        double gu[D][D], ubar[D], F[D][D], C[D][D];
        elem_gradu<D>(E, u, gu, ubar);
        const double vd = E.vol * cof_det<D>(gu, F, C);
        acc[0] += vd;
#pragma unroll
        for (int k = 0; k < D; ++k) acc[1 + k] += vd * (E.xbar[k] + ubar[k]);
    }
    grid_reduce<D + 1, 0>(acc, partials, ticket, out);
}
// L2Norm^2 of every component of a P1 function (exact P1 mass matrix; 3d_admm.lua:1137-1146)
template <int D>
__global__ void __launch_bounds__(256) k_l2norm_p1(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                   const double* __restrict__ f, double* partials, unsigned int* ticket, double* out) {
    double acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
        const double w = E.vol / ((D + 1) * (D + 2));
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double s = 0.0, s2 = 0.0;
#pragma unroll
            for (int a = 0; a <= D; ++a) {
                const double fa = f[(int64_t)E.v[a] * D + c];
                s += fa;
                s2 += fa * fa;
            }
            acc[c] += w * (s2 + s * s);
        }
    }
    grid_reduce<D, 0>(acc, partials, ticket, out);
}
// L2Norm^2 of every component of a P0 tensor function (3d_admm.lua:1243-1251)
template <int D>
__global__ void __launch_bounds__(256) k_l2norm_p0(int64_t ne, const int* __restrict__ elems, const double* __restrict__ xyz,
                                                   const double* __restrict__ f, double* partials, unsigned int* ticket, double* out) {
    constexpr int DD = D * D;
    double acc[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) acc[k] = 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        Elem<D> E;
        elem_load<D>(elems, xyz, e, E);
#pragma unroll
        for (int k = 0; k < DD; ++k) {
            const double v = f[e * DD + k];
            acc[k] += E.vol * v * v;
        }
    }
    grid_reduce<DD, 0>(acc, partials, ticket, out);
}

}  // namespace ab
