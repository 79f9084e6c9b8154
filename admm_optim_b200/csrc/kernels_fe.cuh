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
