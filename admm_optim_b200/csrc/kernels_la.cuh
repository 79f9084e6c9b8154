// Linear-algebra kernels (fp64, sm_100a): BLAS-1 with fused reductions, BSR SpMV / residual /
// Chebyshev-Jacobi smoother step, P1 transfers, Galerkin RAP, dense coarse inverse.
// All are HBM-bandwidth bound; no tensor cores by design.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "gershgorin_dist.cuh"

namespace ab {
namespace cg = cooperative_groups;

// PDL rule (common.cuh): in a kernel launched with AB_LAUNCH_PDL, pointers to data written by preceding kernels of the chain
// are NOT `const __restrict__`: nvcc turns such loads into LDG.CONSTANT and is then free to hoist them above the
// griddepcontrol.wait (observed in SASS).  Only data that is static during a solve keeps the qualifier.

// ---------------------------------------------------------------------------------------------
// BLAS-1
// ---------------------------------------------------------------------------------------------
__global__ void k_fill(int64_t n, double c, double* __restrict__ x) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = c;
}

// out = a*x + b*y  (y may be null -> out = a*x). In-place allowed (out == x or out == y).
__global__ void k_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (y) {
        for (; i < n; i += stride) out[i] = a * x[i] + b * y[i];
    } else {
        for (; i < n; i += stride) out[i] = a * x[i];
    }
}

// NX dot products <x_k, y> in one pass over y (VecProd; S-column batches)
template <int NX>
__global__ void __launch_bounds__(256) k_dot_multi(int64_t n, const double* x0, const double* x1,
                                                   const double* x2, const double* x3, const double* __restrict__ y,
                                                   double* partials, unsigned int* ticket, double* out) {
    const double* xs[4] = {x0, x1, x2, x3};
    double v[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) v[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double yi = y[i];
#pragma unroll
        for (int k = 0; k < NX; ++k) v[k] += xs[k][i] * yi;
    }
    grid_reduce<NX, 0>(v, partials, ticket, out);
}

// ---------------------------------------------------------------------------------------------
// device scalars of the Krylov recurrences (no host round trip inside an iteration)
// ---------------------------------------------------------------------------------------------
enum { SC_RHO = 0, SC_RHO_OLD, SC_ALPHA, SC_OMEGA, SC_RV, SC_TS, SC_TT, SC_RR, SC_SS, SC_BETA, SC_PQ, SC_RZ, SC_RZ_OLD, SC_COUNT };

// p = r + beta (p - omega v),  beta = (rho/rho_old)(alpha/omega)   [BiCGStab]
__global__ void k_bicg_update_p(int64_t n, const double* sc, const double* r, const double* v,
                                double* __restrict__ p) {
    pdl_prologue();
    const double beta = (sc[SC_RHO] / sc[SC_RHO_OLD]) * (sc[SC_ALPHA] / sc[SC_OMEGA]);
    const double omega = sc[SC_OMEGA];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = r[i] + beta * (p[i] - omega * v[i]);
}
// loop-control block of a solve (device memory): the conditional-graph loop reads its stopping rule from here
enum { CTL_TOL2 = 0, CTL_RED2, CTL_RR0, CTL_MAXIT, CTL_IT, CTL_COUNT };

// Start of a BiCGStab solve in one pass: x_ws = x ; rh = r ; p = v = 0 ; |r|^2 -> scalars {rho = |r|^2, rho_old = alpha = omega = 1}
// and the loop-control block {tol^2, reduction^2, |r0|^2, max its, it = 0}.  owned != null (multi-GPU): |r|^2 over the owned copies only;
// the caller all-reduces out2[0] and lets k_bicg_init_fix put the global value in place.
__global__ void __launch_bounds__(256) k_bicg_init(int64_t n, const double* x_in, double* __restrict__ x_ws, const double* r, double* __restrict__ rh,
                                                   double* __restrict__ p, double* __restrict__ v, double* sc, double* partials,
                                                   unsigned int* ticket, double* out2, double* ctl, double tol2, double red2, int max_it,
                                                   const unsigned char* __restrict__ owned, int D) {
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = r[i];
        x_ws[i] = x_in[i];
        rh[i] = ri;
        p[i] = 0.0;
        v[i] = 0.0;
        if (!owned || owned[i / D]) acc[0] += ri * ri;
    }
    const bool last = grid_reduce<1, 0>(acc, partials, ticket, out2);
    if (last) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const double rr = out2[0];
            for (int k = 0; k < SC_COUNT; ++k) sc[k] = 0.0;
            sc[SC_RHO] = rr; sc[SC_RHO_OLD] = 1.0; sc[SC_ALPHA] = 1.0; sc[SC_OMEGA] = 1.0; sc[SC_RR] = rr;
            ctl[CTL_TOL2] = tol2; ctl[CTL_RED2] = red2; ctl[CTL_RR0] = rr; ctl[CTL_MAXIT] = (double)max_it; ctl[CTL_IT] = 0.0;
        }
    }
}
__global__ void k_bicg_init_fix(double* sc, double* ctl, const double* out2) {
    const double rr = out2[0];
    sc[SC_RHO] = rr; sc[SC_RR] = rr; ctl[CTL_RR0] = rr;
}
// alpha = rho / <rh,v>;  s = r - alpha v;  reduce |s|^2
__global__ void __launch_bounds__(256) k_bicg_s(int64_t n, double* __restrict__ sc, const double* r, const double* v,
                                                double* __restrict__ s, double* partials, unsigned int* ticket) {
    pdl_prologue();
    const double alpha = sc[SC_RHO] / sc[SC_RV];
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double si = r[i] - alpha * v[i];
        s[i] = si;
        acc[0] += si * si;
    }
    grid_reduce<1, 0>(acc, partials, ticket, sc + SC_SS);
}
// Fused with the first smoothing step of the V-cycle that follows (zero initial guess: d = c2 D^-1 rhs, x = d):
//   FIRST = 0:  p = r + beta (p - omega v)   ; d = c2 dinv p ; x = d
//   FIRST = 1:  s = r - alpha v              ; d = c2 dinv s ; x = d      (|s|^2 is not needed by the recurrence)
// identical arithmetic to k_bicg_update_p / k_bicg_s followed by k_smooth_first, one pass and one launch less each.
template <int FIRST>
__global__ void __launch_bounds__(256) k_bicg_fused_first(int64_t n, const double* sc, const double* r, const double* v, double* pv,
                                                          const double* __restrict__ dinv, const double* __restrict__ cf, double* d, double* x,
                                                          double* uq, const unsigned char* __restrict__ owned, int D) {
    // uq != null (multi-GPU): the preconditioner input is the unique (owner-only) form of the vector, written to uq; the first
    // smoothing step acts on it (its additive result is summed over the interfaces by the exchange kernel that follows)
    pdl_trigger();
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const double c2 = ld_static(cf + 1);                       // static during a solve: fetched before the dependency wait
    double di = i0 < n ? ld_static(dinv + i0) : 0.0;
    pdl_wait();
    double a, b;                                               // FIRST = 0: pn = r + a (pv - b v);  FIRST = 1: s = r - a v
    if (FIRST == 0) {
        a = (sc[SC_RHO] / sc[SC_RHO_OLD]) * (sc[SC_ALPHA] / sc[SC_OMEGA]);
        b = sc[SC_OMEGA];
    } else {
        a = sc[SC_RHO] / sc[SC_RV];
        b = 0.0;
    }
    for (int64_t i = i0; i < n; i += stride) {
        if (i != i0) di = dinv[i];
        const double pn = FIRST == 0 ? r[i] + a * (pv[i] - b * v[i]) : r[i] - a * v[i];
        pv[i] = pn;
        double rhs = pn;
        if (uq) {
            rhs = owned[i / D] ? pn : 0.0;
            uq[i] = rhs;
        }
        const double dn = c2 * di * rhs;
        d[i] = dn;
        x[i] = dn;
    }
}
// omega = <t,s>/<t,t>; x += alpha ph + omega sh; r = s - omega t; reduce |r|^2 and <rh,r> (next rho).
// ROLL >= 1: the block that finishes last also does the bookkeeping of k_bicg_roll (one launch less per iteration).
// ROLL == 2: ... and evaluates the ConvCheck on the device (same rule as the host loop of bicgstab_apply) and tells the
//            conditional WHILE node of the solve graph whether to run another iteration (cudaGraphSetConditional).
template <int ROLL>
__global__ void __launch_bounds__(256) k_bicg_xr(int64_t n, double* sc, const double* ph, const double* sh,
                                                 const double* s, const double* t, const double* rh,
                                                 double* __restrict__ x, double* __restrict__ r, double* partials, unsigned int* ticket, double* out2,
                                                 unsigned long long cond_handle, double* ctl, double* hist, int hist_cap,
                                                 const unsigned char* __restrict__ owned, int D) {
    // owned != null (multi-GPU, ROLL = 0): the two sums run over the owned copies only; the caller all-reduces out2
    pdl_prologue();
    const double rho = sc[SC_RHO];
    const double alpha = rho / sc[SC_RV];
    const double tt = sc[SC_TT];
    const double omega = tt > 0.0 ? sc[SC_TS] / tt : 0.0;
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * ph[i] + omega * sh[i];
        double ri = s[i] - omega * t[i];
        r[i] = ri;
        if (!owned || owned[i / D]) {
            acc[0] += ri * ri;
            acc[1] += rh[i] * ri;
        }
    }
    const bool last = grid_reduce<2, 0>(acc, partials, ticket, out2);   // out2[0] = |r|^2, out2[1] = <rh,r>
    if (ROLL && last) {          // every other block read the scalars before it took its ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            const double rr = out2[0], rho_new = out2[1];
            sc[SC_RHO_OLD] = rho;
            sc[SC_ALPHA] = alpha;
            sc[SC_OMEGA] = omega;
            sc[SC_RR] = rr;
            sc[SC_RHO] = rho_new;
            if (ROLL == 2) {
                const int it = (int)ctl[CTL_IT] + 1;
                ctl[CTL_IT] = (double)it;
                if (hist && it < hist_cap) hist[it] = rr;
                const bool finite = (rr == rr) && !isinf(rr);
                const bool conv = rr < ctl[CTL_TOL2] || (ctl[CTL_RED2] > 0.0 && rr < ctl[CTL_RED2] * ctl[CTL_RR0]);
                const bool stop = !finite || conv || it >= (int)ctl[CTL_MAXIT] || rho_new == 0.0 || omega == 0.0;
                cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, stop ? 0u : 1u);
            }
        }
    }
}
// bookkeeping between iterations: rho_old = rho; alpha, omega stored; rho = <rh,r>
__global__ void k_bicg_roll(double* sc, const double* out2) {
    pdl_prologue();
    const double alpha = sc[SC_RHO] / sc[SC_RV];
    const double tt = sc[SC_TT];
    const double omega = tt > 0.0 ? sc[SC_TS] / tt : 0.0;
    sc[SC_RHO_OLD] = sc[SC_RHO];
    sc[SC_ALPHA] = alpha;
    sc[SC_OMEGA] = omega;
    sc[SC_RR] = out2[0];
    sc[SC_RHO] = out2[1];
}
// point-Jacobi data from consistent diagonal / row sums with two row-sum variants: red[0] = max rows_a / a_ii, red[1] = max rows_b / a_ii
__global__ void __launch_bounds__(256) k_dinv_lmax2(int64_t n, double* __restrict__ diag_to_dinv, const double* __restrict__ rows_a,
                                                    const double* __restrict__ rows_b, double* partials, unsigned int* ticket, double* red) {
    double mx[2] = {0.0, 0.0};
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double aii = diag_to_dinv[t];
        diag_to_dinv[t] = 1.0 / aii;
        mx[0] = fmax(mx[0], rows_a[t] / aii);
        mx[1] = fmax(mx[1], rows_b[t] / aii);
    }
    grid_reduce<2, 1>(mx, partials, ticket, red);
}
// ---------------------------------------------------------------------------------------------
// BSR SpMV family (k_bsr_spmv_tma: default; k_bsr_spmv_warp: fallback when a row exceeds a tile).
//   MODE 0: y = A x
//   MODE 1: y = b - A x
//   MODE 2: Chebyshev/Jacobi step  r = b - A xin ; d = c1*d + c2*dinv*r ; xout = xin + d
// DOTS (MODE 0 only): 0 none, 1: <w,y> -> red[0], 2: <w,y>, <y,y> -> red[0..1]   (w given)
// ---------------------------------------------------------------------------------------------
// streaming (read-once) load: keep the matrix stream out of L1 so that the x gathers stay cached
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Warp-per-row BSR SpMV with a FIXED lane -> (block slot, r, c) map (fallback kernel).
// D*D lanes cover one block, 32/(D*D) blocks per step (3 for 3x3: 27 active lanes reading 216 contiguous
// bytes; 8 for 2x2).  Because a lane's (r,c) never changes there is no index division, no select and a
// single accumulator.  U steps are unrolled with all their loads in flight; the next row's extent is prefetched
// while the current row streams.  32 registers -> 64 resident warps per SM (measured: occupancy beats deeper
// per-warp pipelining for this kernel, profiles/r01_spmv_notes.md).
// Same MODE / DOTS semantics as above.
// ---------------------------------------------------------------------------------------------
template <int D, int MODE, int DOTS, int U>
__global__ void __launch_bounds__(256) k_bsr_spmv_warp(int nb, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                       const double* __restrict__ vals, const double* x,
                                                       const double* b, double* __restrict__ y,
                                                       const double* __restrict__ dinv, const double* dvec, double* dout, double c1, double c2,
                                                       const double* w, double* partials, unsigned int* ticket, double* red, const double* __restrict__ cf, int prefetch) {
    pdl_prologue();
    if (cf) { c1 = cf[0]; c2 = cf[1]; }      // smoother coefficients from device memory (graph-replayable launches)
    constexpr int DD = D * D;
    constexpr int BPS = 32 / DD;                        // blocks per step
    const int lane = threadIdx.x & 31;
    const int lb = lane / DD;                           // block slot of this lane within a step
    const int wq = lane - lb * DD;                      // position inside the block
    const int c = wq % D;
    const bool lane_on = lb < BPS;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double dot[2] = {0.0, 0.0};
    int s_n = 0, e_n = 0;
    if (wid < nb) { s_n = __ldg(rowptr + wid); e_n = __ldg(rowptr + wid + 1); }
    for (int64_t row = wid; row < nb; row += nwarps) {
        const int s = s_n, e = e_n;
        if (row + nwarps < nb) { s_n = __ldg(rowptr + row + nwarps); e_n = __ldg(rowptr + row + nwarps + 1); }
        double acc = 0.0;
        for (int blk0 = s; blk0 < e; blk0 += BPS * U) {
            double a[U];
            int col[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int blk = blk0 + u * BPS + lb;
                const bool valid = lane_on && blk < e;
                a[u] = valid ? ld_stream(vals + (int64_t)blk * DD + wq) : 0.0;
                col[u] = valid ? __ldg(colidx + blk) : 0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int blk = blk0 + u * BPS + lb;
                const bool valid = lane_on && blk < e;
                const double xv = valid ? x[(unsigned)(col[u] * D + c)] : 0.0;
                acc = fma(a[u], xv, acc);
            }
        }
        // reduce over the block slots, then over c
        double v;
        if constexpr (D == 3) {
            const double t1 = __shfl_down_sync(0xffffffffu, acc, 9);
            const double t2 = __shfl_down_sync(0xffffffffu, acc, 18);
            v = acc + t1 + t2;                                   // valid on lanes 0..8
            const double u1 = __shfl_down_sync(0xffffffffu, v, 1);
            const double u2 = __shfl_down_sync(0xffffffffu, v, 2);
            v = v + u1 + u2;                                     // row r on lane 3r
        } else {
            v = acc;
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_down_sync(0xffffffffu, v, 1);            // row r on lane 2r
        }
        if (lane < DD && c == 0) {
            const int r = lane / D;
            const int64_t i = row * D + r;
            if (MODE == 0) {
                y[i] = v;
                if (DOTS >= 1) dot[0] += w[i] * v;
                if (DOTS >= 2) dot[1] += v * v;
            } else if (MODE == 1) {
                y[i] = b[i] - v;
            } else {
                const double res = b[i] - v;
                const double dn = (c1 != 0.0 ? c1 * dvec[i] : 0.0) + c2 * dinv[i] * res;
                dout[i] = dn;
                y[i] = x[i] + dn;
            }
        }
    }
    if (DOTS > 0) grid_reduce<2, 0>(dot, partials, ticket, red);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged BSR SpMV.  A tile is a run of consecutive block rows with <= TB blocks; its values, block
// columns, row extents and the row slices of the vectors the epilogue needs are contiguous in global memory.
// One producer thread per CTA streams them into a ring of NSTAGE shared-memory stages with 1-D bulk async
// copies (cp.async.bulk, completion on an mbarrier) -- the matrix stream never touches the LSU/L1 path and
// the bytes in flight are set by the ring depth, not by the number of resident warps.  The consumer warps
// take rows of a landed tile (warp per row, same fixed lane -> (block slot, r, c) map as k_bsr_spmv_warp),
// read the matrix from shared memory, gather x through L1/L2 and apply the MODE epilogue.
// Same MODE / DOTS semantics as above.
// ---------------------------------------------------------------------------------------------
template <int D>
struct SpmvTma {
    static constexpr int DD = D * D;
    static constexpr int TB = D == 3 ? 360 : 512;            // max blocks per tile (3 stages x 2 CTAs fit the 227 KB of an SM)
    static constexpr int RMAX = TB / (D + 1);                 // max rows per tile (a P1 row has >= D+1 blocks)
    static constexpr int NSTAGE = 3;
    static constexpr int NT = 512;                            // 15 consumer warps + 1 producer warp
    static constexpr int NAUX = 4;
    static constexpr int VALS_B = TB * DD * 8 + 16;
    static constexpr int COLS_B = TB * 4 + 16;
    static constexpr int ROWS_B = ((RMAX + 1) * 4 + 16 + 15) / 16 * 16;
    static constexpr int AUX_B = (RMAX * D * 8 + 16 + 15) / 16 * 16;
    static constexpr int STAGE_B = VALS_B + COLS_B + ROWS_B + NAUX * AUX_B;
    static constexpr int META_B = 64;                         // per stage: r0, nr, b0, pre offsets
    static constexpr int SMEM_B = 128 + NSTAGE * (STAGE_B + META_B);
};

template <int D, int MODE, int DOTS, int U, int HL>
__global__ void __launch_bounds__(SpmvTma<D>::NT, 2) k_bsr_spmv_tma(int ntiles, const int* __restrict__ tile_info, const int* __restrict__ rowptr,
                                                                 const int* __restrict__ colidx, const double* __restrict__ vals,
                                                                 const double* x, const double* b,
                                                                 double* __restrict__ y, const double* __restrict__ dinv,
                                                                 const double* dvec, double* dout, double c1, double c2, const double* w,
                                                                 double* partials, unsigned int* ticket, double* red, const double* __restrict__ cf, int prefetch) {
    pdl_trigger();                           // dependents may be scheduled; this kernel waits for ITS predecessor below (pdl_wait)
    if (cf) { c1 = cf[0]; c2 = cf[1]; }      // smoother coefficients from device memory (graph-replayable launches; written at setup)
    using T = SpmvTma<D>;
    constexpr int DD = T::DD, BPS = 32 / DD, NS = T::NSTAGE, NW = T::NT / 32 - 1;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);              // [NS]
    uint64_t* empty = full + NS;                                       // [NS]
    unsigned char* stage0 = smem + 128;
    int* meta0 = reinterpret_cast<int*>(smem + 128 + NS * T::STAGE_B);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NW); }
        mbar_fence_init();
    }
    __syncthreads();
    const int n_my = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    double dot[2] = {0.0, 0.0};
    // which vectors does the epilogue read per row?  (aux slot -> global pointer)
    const double* aux_src[T::NAUX] = {nullptr, nullptr, nullptr, nullptr};
    if (MODE == 0 && DOTS >= 1) aux_src[0] = w;
    if (MODE == 1) aux_src[0] = b;
    if (MODE == 2) { aux_src[0] = b; aux_src[1] = dinv; aux_src[2] = dvec; aux_src[3] = x; }

    if (warp == NW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            // one tile into ring stage i % NS.  The matrix part (values, columns, row extents) is static during a solve; the
            // vector slices of the epilogue are written by preceding kernels of the chain.
            const bool stream_hint = (prefetch & 2) != 0;     // bit 1 of `prefetch`: matrix larger than L2
            const uint64_t pol = l2_policy_evict_first();
            auto produce = [&](int i, bool do_matrix, bool do_aux) {
                const int st = i % NS;
                const int tile = blockIdx.x + i * gridDim.x;
                const int r0 = __ldg(tile_info + 2 * tile), b0 = __ldg(tile_info + 2 * tile + 1);
                const int r1 = __ldg(tile_info + 2 * tile + 2), b1 = __ldg(tile_info + 2 * tile + 3);
                const int nr = r1 - r0, nblk = b1 - b0;
                unsigned char* sb = stage0 + (size_t)st * T::STAGE_B;
                int* meta = meta0 + st * (T::META_B / 4);
                // 16-byte alignment: start each copy at the aligned-down source, remember the byte offset
                const uintptr_t pv = (uintptr_t)(vals + (int64_t)b0 * DD), pc = (uintptr_t)(colidx + b0), pr = (uintptr_t)(rowptr + r0);
                const uint32_t ov = pv & 15, oc = pc & 15, orw = pr & 15;
                const uint32_t nv_b = (ov + (uint32_t)nblk * DD * 8 + 15) & ~15u, nc_b = (oc + (uint32_t)nblk * 4 + 15) & ~15u,
                               nr_b = (orw + (uint32_t)(nr + 1) * 4 + 15) & ~15u;
                uint32_t total = nv_b + nc_b + nr_b;
                uint32_t oa[T::NAUX], na[T::NAUX];
#pragma unroll
                for (int k = 0; k < T::NAUX; ++k) {
                    oa[k] = 0; na[k] = 0;
                    if (aux_src[k]) {
                        const uintptr_t pa = (uintptr_t)(aux_src[k] + (int64_t)r0 * D);
                        oa[k] = pa & 15;
                        na[k] = (oa[k] + (uint32_t)nr * D * 8 + 15) & ~15u;
                        total += na[k];
                    }
                }
                if (do_matrix) {
                    meta[0] = r0; meta[1] = nr; meta[2] = b0; meta[3] = ov; meta[4] = oc; meta[5] = orw;
#pragma unroll
                    for (int k = 0; k < T::NAUX; ++k) meta[6 + k] = oa[k];
                    mbar_arrive_expect_tx(full + st, total);
                    if (stream_hint) {       // read-once matrix stream: first in line for eviction from L2
                        bulk_g2s_hint(sb, (const void*)(pv - ov), nv_b, full + st, pol);
                        bulk_g2s_hint(sb + T::VALS_B, (const void*)(pc - oc), nc_b, full + st, pol);
                    } else {
                        bulk_g2s(sb, (const void*)(pv - ov), nv_b, full + st);
                        bulk_g2s(sb + T::VALS_B, (const void*)(pc - oc), nc_b, full + st);
                    }
                    bulk_g2s(sb + T::VALS_B + T::COLS_B, (const void*)(pr - orw), nr_b, full + st);
                }
                if (do_aux) {
#pragma unroll
                    for (int k = 0; k < T::NAUX; ++k)
                        if (aux_src[k]) {
                            unsigned char* da = sb + T::VALS_B + T::COLS_B + T::ROWS_B + k * T::AUX_B;
                            const void* sa_ = (const void*)((uintptr_t)(aux_src[k] + (int64_t)r0 * D) - oa[k]);
                            if (stream_hint && k != 3) bulk_g2s_hint(da, sa_, na[k], full + st, pol);   // b, D^-1, d: read once (slot 3 = x is also gathered)
                            else bulk_g2s(da, sa_, na[k], full + st);
                        }
                }
            };
            // first ring fill: the matrix stream starts before the dependency wait, the vector slices after it
            // (prefetch = 0: the caller cannot vouch that the matrix was final before the last stream synchronisation)
            const int npre = min(n_my, NS);
            if (!(prefetch & 1)) pdl_wait();
            for (int i = 0; i < npre; ++i) produce(i, true, false);
            pdl_wait();
            for (int i = 0; i < npre; ++i) produce(i, false, true);
            for (int i = npre; i < n_my; ++i) {
                mbar_wait(empty + i % NS, ((i / NS) - 1) & 1);
                produce(i, true, true);
            }
        } else {
            pdl_wait();
        }
    } else {
        // ---------------- consumers ----------------
        // HL lanes per block row (half-warp for 3x3 blocks; 2x2 blocks, where a row is only ~250 bytes: 4 lanes, i.e. 8 rows
        // per warp in flight -- the row count in flight, not the issue rate, bounds that case), lane <-> (block slot, column
        // c) of the block: one x gather and D value loads (column c of the DxD block, from shared memory) feed D independent
        // accumulators -- no idle work per entry, no index division; the 32/HL rows of a warp share every instruction.
        constexpr int RPW = 32 / HL;                    // rows per warp
        constexpr int BPH = HL / D;                     // blocks per step of a row group (5 for 3x3, 4 for 2x2)
        const int hl = lane % HL, half = lane / HL;
        const int lb = hl / D, c = hl - lb * D;
        const bool lane_on = lb < BPH;
        pdl_wait();                                     // x, b, d ... come from the preceding kernels
        for (int i = 0; i < n_my; ++i) {
            const int st = i % NS;
            mbar_wait(full + st, (i / NS) & 1);
            const unsigned char* sb = stage0 + (size_t)st * T::STAGE_B;
            const int* meta = meta0 + st * (T::META_B / 4);
            const int r0 = meta[0], nr = meta[1], b0 = meta[2];
            const double* sv = reinterpret_cast<const double*>(sb + meta[3]);
            const int* sc = reinterpret_cast<const int*>(sb + T::VALS_B + meta[4]);
            const int* sr = reinterpret_cast<const int*>(sb + T::VALS_B + T::COLS_B + meta[5]);
            const unsigned char* sa = sb + T::VALS_B + T::COLS_B + T::ROWS_B;
            for (int task = warp; RPW * task < nr; task += NW) {
                const int lr = RPW * task + half;
                const bool row_on = lr < nr;
                const int s = row_on ? sr[lr] - b0 : 0, e = row_on ? sr[lr + 1] - b0 : 0;
                int len = e - s;                                           // warp-uniform trip count = longest row of the warp
#pragma unroll
                for (int o = 16; o >= HL; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
                double acc[D];
#pragma unroll
                for (int r = 0; r < D; ++r) acc[r] = 0.0;
                for (int off = 0; off < len; off += BPH * U) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int blk = s + off + u * BPH + lb;
                        if (lane_on && blk < e) {
                            const int col = sc[blk];
                            const double xv = x[(unsigned)(col * D + c)];   // plain load: x is written by the preceding kernel (PDL rule)
                            const double* ap = sv + blk * DD + c;
#pragma unroll
                            for (int r = 0; r < D; ++r) acc[r] = fma(ap[r * D], xv, acc[r]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < D; ++r) {
#pragma unroll
                    for (int o = HL / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                }
                if (row_on && hl < D) {
                    double v = acc[0];
#pragma unroll
                    for (int r = 1; r < D; ++r) v = (hl == r) ? acc[r] : v;
                    const int li = lr * D + hl;
                    const int64_t gi = (int64_t)(r0 + lr) * D + hl;
                    if (MODE == 0) {
                        y[gi] = v;
                        if (DOTS >= 1) dot[0] += reinterpret_cast<const double*>(sa + meta[6])[li] * v;
                        if (DOTS >= 2) dot[1] += v * v;
                    } else if (MODE == 1) {
                        y[gi] = reinterpret_cast<const double*>(sa + meta[6])[li] - v;
                    } else {
                        const double res = reinterpret_cast<const double*>(sa + meta[6])[li] - v;
                        const double di = reinterpret_cast<const double*>(sa + T::AUX_B + meta[7])[li];
                        const double dold = c1 != 0.0 ? reinterpret_cast<const double*>(sa + 2 * T::AUX_B + meta[8])[li] : 0.0;
                        const double xi = reinterpret_cast<const double*>(sa + 3 * T::AUX_B + meta[9])[li];
                        const double dn = c1 * dold + c2 * di * res;
                        dout[gi] = dn;
                        y[gi] = xi + dn;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
        }
    }
    if (DOTS > 0) grid_reduce<2, 0>(dot, partials, ticket, red);
}

// first Chebyshev/Jacobi step from a zero initial guess: d = c2*dinv*b ; x = d   (no matrix pass)
__global__ void k_smooth_first(int64_t n, double c2, const double* __restrict__ dinv, const double* b,
                               double* __restrict__ d, double* __restrict__ x, const double* __restrict__ cf) {
    pdl_trigger();
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cf) c2 = ld_static(cf + 1);                            // static during a solve: fetched before the dependency wait
    double di = i0 < n ? ld_static(dinv + i0) : 0.0;
    pdl_wait();
    for (int64_t i = i0; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (i != i0) di = dinv[i];
        double dn = c2 * di * b[i];
        d[i] = dn;
        x[i] = dn;
    }
}

// point-Jacobi data: dinv_i = 1/a_ii and the Gershgorin bound max_i sum_j |a_ij| / a_ii  (-> red[0])
template <int D>
__global__ void __launch_bounds__(256) k_diag_gershgorin(int nb, const int* __restrict__ rowptr, const int* __restrict__ diagpos,
                                                         const double* __restrict__ vals, double* __restrict__ dinv, double* partials,
                                                         unsigned int* ticket, double* red) {
    constexpr int DD = D * D;
    double mx[1] = {0.0};
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nb * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(t / D), r = (int)(t - (int64_t)row * D);
        const int s = rowptr[row], e = rowptr[row + 1];
        double sum = 0.0;
        for (int k = s; k < e; ++k) {
#pragma unroll
            for (int c = 0; c < D; ++c) sum += fabs(vals[(int64_t)k * DD + r * D + c]);
        }
        const double aii = vals[(int64_t)diagpos[row] * DD + r * D + r];
        dinv[t] = 1.0 / aii;
        mx[0] = fmax(mx[0], sum / aii);
    }
    grid_reduce<1, 1>(mx, partials, ticket, red);
}

// ---------------------------------------------------------------------------------------------
// P1 transfers (StdTransfer, obstacle_optim_3d_util.lua:28): copies weight 1, edge midpoints 1/2,1/2
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void k_prolong_add(int nvc, int nvf, const int* __restrict__ pa, const int* __restrict__ pb,
                              const double* xc, const double* xin, double* xout) {
    pdl_trigger();
    bool waited = false;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nvf * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t / D), c = (int)(t - (int64_t)v * D);
        int ia = -1, ib = -1;
        if (v >= nvc) { ia = ld_static(pa + (v - nvc)); ib = ld_static(pb + (v - nvc)); }   // static parent table: fetched before the dependency wait
        if (!waited) { pdl_wait(); waited = true; }
        const double add = v < nvc ? xc[t] : 0.5 * (xc[(int64_t)ia * D + c] + xc[(int64_t)ib * D + c]);
        xout[t] = xin[t] + add;
    }
}

// rc = mask * P^T rf : gather over the coarse vertex graph, `mid` gives the fine midpoint of every coarse edge.
// Half-warp per coarse vertex: the lanes take the row's entries (one midpoint each), gather the D components and
// shuffle-reduce -- two dependent loads per vertex instead of a serial walk over its ~15 neighbours.
template <int D>
__global__ void __launch_bounds__(256) k_restrict(int nvc, const int* __restrict__ rowptr, const int* __restrict__ mid, const int* __restrict__ diagpos,
                                                  const unsigned char* __restrict__ dirmask, const double* rf, double* __restrict__ rc) {
    pdl_trigger();
    const int hl = threadIdx.x & 15;
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const int64_t rounds = (nvc + nhw - 1) / nhw;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t v = hw + it * nhw;
        const bool on = v < nvc;
        // static part (pattern, midpoint table, Dirichlet mask): for the first round fetched before the dependency wait
        const int s = on ? ld_static(rowptr + v) : 0, e = on ? ld_static(rowptr + v + 1) : 0, dp = on ? ld_static(diagpos + v) : -1;
        const int k0 = s + hl, k1 = k0 + 16;
        const int m0 = (k0 < e && k0 != dp) ? ld_static(mid + k0) : -1, m1 = (k1 < e && k1 != dp) ? ld_static(mid + k1) : -1;
        const bool fixed = on && hl < D && dirmask && ((dirmask[v] >> hl) & 1);
        if (it == 0) pdl_wait();
        double acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = 0.0;
        if (m0 >= 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] += rf[(int64_t)m0 * D + c];
        }
        if (m1 >= 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] += rf[(int64_t)m1 * D + c];
        }
        for (int k = k0 + 32; k < e; k += 16) {
            if (k == dp) continue;
            const int64_t m = (int64_t)mid[k] * D;
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] += rf[m + c];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        }
        if (on && hl < D) {
            double h = acc[0];
#pragma unroll
            for (int c = 1; c < D; ++c) h = (hl == c) ? acc[c] : h;
            rc[v * D + hl] = fixed ? 0.0 : rf[v * D + hl] + 0.5 * h;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Galerkin coarse operator Ac = Dir(P^T Af P)  (rap = true, obstacle_optim_3d_util.lua:27).
// Row-owner gather: one warp per coarse block row I accumulates its row in shared memory from the
// fine rows of its children (copy of I: weight 1; midpoints of the coarse edges at I: weight 1/2).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_rap(int nvc, int maxrow, const int* __restrict__ crowptr, const int* __restrict__ ccol,
                                             const int* __restrict__ cmid, const int* __restrict__ cdiag,
                                             const int* __restrict__ frowptr, const int* __restrict__ fcol, const double* __restrict__ fvals,
                                             const int* __restrict__ pa, const int* __restrict__ pb,
                                             const unsigned char* __restrict__ dirmask, const unsigned char* __restrict__ owned,
                                             double* __restrict__ cvals) {
    // one CTA per coarse block row, its warps take the children of the row in turn (the walk over one child's fine row is a
    // dependent chain of binary searches and shared-memory atomics: the children are the available parallelism)
    constexpr int DD = D * D;
    extern __shared__ double sm_rap[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* acc = sm_rap;                                   // maxrow x DD
    int* cols = (int*)(sm_rap + (size_t)maxrow * DD);       // maxrow
    for (int64_t I = blockIdx.x; I < nvc; I += gridDim.x) {
        const int cs = crowptr[I], ce = crowptr[I + 1], len = ce - cs;
        for (int k = threadIdx.x; k < len * DD; k += blockDim.x) acc[k] = 0.0;
        for (int k = threadIdx.x; k < len; k += blockDim.x) cols[k] = ccol[cs + k];
        __syncthreads();
        for (int ck = wib; ck < len; ck += nw) {                 // children of I
            const int child = cmid[cs + ck];
            const double wI = (cs + ck == cdiag[I]) ? 1.0 : 0.5;
            const int fs = frowptr[child], fe = frowptr[child + 1];
            for (int k = lane; k < (fe - fs) * DD; k += 32) {
                const int blk = k / DD, rc = k - blk * DD;
                const int j = fcol[fs + blk];
                const double a = wI * fvals[(int64_t)fs * DD + k];
                int J0, J1;
                double wJ;
                if (j < nvc) { J0 = j; J1 = -1; wJ = 1.0; }
                else { J0 = pa[j - nvc]; J1 = pb[j - nvc]; wJ = 0.5; }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int J = q == 0 ? J0 : J1;
                    if (J < 0) continue;
                    int lo = 0, hi = len - 1;                    // binary search J in cols
                    while (lo < hi) {
                        int m = (lo + hi) >> 1;
                        if (cols[m] < J) lo = m + 1; else hi = m;
                    }
                    atomicAdd(&acc[lo * DD + rc], wJ * a);
                }
            }
        }
        __syncthreads();
        const unsigned char mI = dirmask ? dirmask[I] : 0;
        for (int k = threadIdx.x; k < len * DD; k += blockDim.x) {
            const int blk = k / DD, rc = k - blk * DD, r = rc / D, c = rc - r * D;
            double val = acc[k];
            if (dirmask) {
                const unsigned char mJ = dirmask[cols[blk]];
                // unit diagonal of an eliminated dof; multi-GPU (owned != null): set by the owner of the vertex only (additive operator)
                if (((mI >> r) & 1) || ((mJ >> c) & 1)) val = (cols[blk] == (int)I && r == c && (!owned || owned[I])) ? 1.0 : 0.0;
            }
            cvals[(int64_t)cs * DD + k] = val;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// vertical interface of the agglomerated multigrid (multi-GPU, rank 0; lib.cu Gmg::vcycle_base_gathered / gmg_gather_operator)
// ---------------------------------------------------------------------------------------------
// dst[gpos[b]] += src[b] for DD-entry blocks; the positions of ONE rank are distinct, the ranks are added one launch after the other
__global__ void k_scatter_add_blocks(int64_t n, int DD, const int* __restrict__ gpos, const double* __restrict__ src, double* __restrict__ dst) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / DD;
        const int e = (int)(t - b * DD);
        dst[(int64_t)gpos[b] * DD + e] += src[t];
    }
}
// global additive -> summed right-hand side: bg[v] = sum of the staged local values of the ranks holding v, in rank order
__global__ void k_vgather(int nvg, int D, const int* __restrict__ ptr, const int* __restrict__ idx, const double* stage, double* __restrict__ bg) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nvg * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t / D), c = (int)(t - (int64_t)v * D);
        double s = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) s += stage[(int64_t)idx[e] * D + c];
        bg[t] = s;
    }
}
// every rank's (consistent) part of the global coarse solution, packed rank by rank
__global__ void k_vscatter(int64_t ntot, int D, const int* __restrict__ l2g, const double* xg, double* __restrict__ stage) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ntot * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / D;
        const int c = (int)(t - k * D);
        stage[t] = xg[(int64_t)l2g[k] * D + c];
    }
}

// ---------------------------------------------------------------------------------------------
// coarse-level direct solve (replaces SuperLU(), obstacle_optim_3d_util.lua:21):
// dense inverse of the free-DoF block by Gauss-Jordan with partial pivoting (cooperative kernel),
// applied as one GEMV per V-cycle.
// ---------------------------------------------------------------------------------------------
// scatter the BSR level-0 matrix into the augmented dense system [A_ff | I]  (n x 2n, row-major)
template <int D>
__global__ void k_bsr_to_dense(int nb, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
                               const int* __restrict__ dof2free, int ld, double* __restrict__ M) {
    constexpr int DD = D * D;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)rowptr[nb] * DD; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t blk = t / DD;
        const int rc = (int)(t - blk * DD), r = rc / D, c = rc - r * D;
        // find block row by binary search
        int lo = 0, hi = nb - 1;
        while (lo < hi) {
            int m = (lo + hi + 1) >> 1;
            if (rowptr[m] <= blk) lo = m; else hi = m - 1;
        }
        const int fi = dof2free[lo * D + r], fj = dof2free[colidx[blk] * D + c];
        if (fi >= 0 && fj >= 0) M[(int64_t)fi * ld + fj] = vals[t];
    }
}
__global__ void k_dense_identity(int n, double* __restrict__ M) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) M[(int64_t)i * 2 * n + n + i] = 1.0;
}

// in-place Gauss-Jordan on M = [A | I] -> rows hold [e_k-ish | A^-1 rows] up to a row permutation and scaling.
// used[] is kept redundantly in every block's shared memory; pivrow[k] (global) records the pivot row of step k.
__global__ void __launch_bounds__(256) k_gauss_jordan(int n, double* __restrict__ M, int* __restrict__ pivrow, int* __restrict__ fail) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ unsigned char sm_used[];          // n bytes
    __shared__ double s_val[8];
    __shared__ int s_idx[8];
    __shared__ int s_piv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = 2 * (int64_t)n;
    for (int i = tid; i < n; i += blockDim.x) sm_used[i] = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        // pivot search (redundant per block): max |M[i][k]| over unused rows, ties -> smallest i
        double best = -1.0;
        int bi = -1;
        for (int i = tid; i < n; i += blockDim.x) {
            if (!sm_used[i]) {
                double a = fabs(M[i * ld + k]);
                if (a > best) { best = a; bi = i; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = s_val[0];
            int ix = s_idx[0];
            for (int w = 1; w < 8; ++w)
                if (s_val[w] > b || (s_val[w] == b && s_idx[w] >= 0 && (ix < 0 || s_idx[w] < ix))) { b = s_val[w]; ix = s_idx[w]; }
            s_piv = ix;
            if (blockIdx.x == 0) {
                pivrow[k] = ix;
                if (!(b > 0.0)) *fail = 1;
            }
        }
        __syncthreads();
        const int p = s_piv;
        if (p < 0) break;                                // singular: every block takes this branch together
        sm_used[p] = 1;                                  // benign same-value race
        const double piv = M[p * ld + k];
        const double* prow = M + p * ld;
        for (int i = blockIdx.x; i < n; i += gridDim.x) {
            if (i == p) continue;
            double* row = M + i * ld;
            const double f = row[k] / piv;
            __syncthreads();                             // everyone has read row[k] before it is overwritten
            if (f != 0.0) {
                for (int c = k + 1 + tid; c < 2 * n; c += blockDim.x) row[c] -= f * prow[c];
                if (tid == 0) row[k] = 0.0;
            }
        }
        grid.sync();
    }
}
// Blocked Gauss-Jordan without row exchanges (the free-dof coarse operator is symmetric and, for the moderate
// multipliers of the ADMM loop, positive definite): K pivot columns per grid-wide barrier.  Every block reduces the
// K x 2n pivot panel redundantly in shared memory (the K x K diagonal block becomes the identity), then applies the
// rank-K update to the rows it owns.  A pivot that is tiny relative to the largest pivot seen raises *fail = 2 and
// the host falls back to the partially pivoted k_gauss_jordan.  On exit M = [I | A^-1].
template <int K>
__global__ void __launch_bounds__(1024) k_gauss_jordan_blocked(int n, double* __restrict__ M, int* __restrict__ pivrow, int* __restrict__ fail) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double panel[];                   // K x 2n
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int ld = 2 * n;
    double maxpiv = 0.0;
    if (tid == 0) s_bad = 0;
    for (int i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) pivrow[i] = i;
    for (int k0 = 0; k0 < n; k0 += K) {
        const int kk = min(K, n - k0);
        for (int j = 0; j < kk; ++j)
            for (int c = k0 + tid; c < ld; c += blockDim.x) panel[j * ld + c] = M[(int64_t)(k0 + j) * ld + c];
        grid.sync();                                    // every block holds the panel before its owner overwrites it
        for (int j = 0; j < kk; ++j) {
            const double piv = panel[j * ld + k0 + j];
            const double ap = fabs(piv);
            if (!(ap > 1e-10 * maxpiv) || !(ap > 0.0)) { if (tid == 0) s_bad = 1; }
            maxpiv = fmax(maxpiv, ap);
            const double inv = 1.0 / piv;
            double f[K];
#pragma unroll
            for (int j2 = 0; j2 < K; ++j2) f[j2] = (j2 < kk && j2 != j) ? panel[j2 * ld + k0 + j] : 0.0;
            __syncthreads();                             // pivot column read by everybody before it is overwritten
            for (int c = k0 + tid; c < ld; c += blockDim.x) {
                const double pj = panel[j * ld + c] * inv;
                panel[j * ld + c] = pj;
#pragma unroll
                for (int j2 = 0; j2 < K; ++j2)
                    if (j2 < kk && j2 != j) panel[j2 * ld + c] -= f[j2] * pj;
            }
            __syncthreads();
        }
        for (int i = blockIdx.x; i < n; i += gridDim.x) {
            double* row = M + (int64_t)i * ld;
            if (i >= k0 && i < k0 + kk) {
                for (int c = k0 + tid; c < ld; c += blockDim.x) row[c] = panel[(i - k0) * ld + c];
                continue;
            }
            double cf[K];
#pragma unroll
            for (int j = 0; j < K; ++j) cf[j] = j < kk ? row[k0 + j] : 0.0;
            __syncthreads();
            for (int c = k0 + tid; c < ld; c += blockDim.x) {
                double v = row[c];
#pragma unroll
                for (int j = 0; j < K; ++j) v -= cf[j] * panel[j * ld + c];
                row[c] = (c < k0 + kk) ? 0.0 : v;
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0 && s_bad) *fail = 2;
}

// Fast path of the coarse inverse: in-place blocked Gauss-Jordan ("sweep") inversion with every row RESIDENT in shared
// memory.  Row i lives in CTA i % gridDim.x for the whole kernel; per block step of K pivots only the K pivot rows travel
// (published by their owners into a ping-pong global panel, read back by everybody after ONE grid barrier):
//     Dinv = T[P,P]^-1                      (K x K, one warp, rows in registers, shuffle broadcasts)
//     pivot rows   p: T[p,J] = Dinv[p,:] T[P,J],             T[p,P] = Dinv[p,:]
//     other rows   i: w = T[i,P] Dinv ; T[i,J] -= w T[P,J] ; T[i,P] = -w
// After ceil(n/K) steps T = A^-1.  No row exchanges (as k_gauss_jordan_blocked); a pivot tiny relative to the largest
// one seen raises *fail = 2 and the host redoes the setup with the partially pivoted kernel.
template <int K>
__global__ void __launch_bounds__(640) k_gauss_jordan_resident(int n, const double* __restrict__ A, double* __restrict__ Ainv, double* panelbuf,
                                                               int* __restrict__ fail) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm_gj[];
    __shared__ int s_bad;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int rmax = (n + G - 1) / G;
    const int nrows = b < n ? (n - b + G - 1) / G : 0;     // rows b, b+G, b+2G, ...
    double* rows = sm_gj;                                  // rmax x n
    double* panel = rows + (size_t)rmax * n;               // K x n
    double* Dinv = panel + (size_t)K * n;                  // K x K
    double* w = Dinv + K * K;                              // rmax x K
    if (tid == 0) s_bad = 0;
    for (int r = 0; r < nrows; ++r)
        for (int c = tid; c < n; c += nt) rows[(size_t)r * n + c] = A[(size_t)(b + r * G) * n + c];
    __syncthreads();
    // publish the pivot rows of step 0
    for (int q = 0; q < min(K, n); ++q)
        if (q % G == b) {
            const int r = q / G;
            for (int c = tid; c < n; c += nt) panelbuf[(size_t)q * n + c] = rows[(size_t)r * n + c];
        }
    grid.sync();
    double maxpiv = 0.0;                                   // warp 0 only
    int step = 0;
    for (int k0 = 0; k0 < n; k0 += K, ++step) {
        const int kk = min(K, n - k0);
        const double* pb = panelbuf + (size_t)(step & 1) * K * n;
        {   // written by other CTAs: read through L2, eight independent loads in flight per thread
            const int tot = kk * n;
            for (int base = 0; base < tot; base += 8 * nt) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * nt + tid;
                    v[u] = idx < tot ? __ldcg(pb + idx) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * nt + tid;
                    if (idx < tot) panel[idx] = v[u];
                }
            }
        }
        __syncthreads();
        if (tid < 32) {      // invert the K x K pivot block: lane l holds row l (identity padding beyond kk)
            const int lane = tid;
            double d[K];
#pragma unroll
            for (int q = 0; q < K; ++q) d[q] = (lane < kk && q < kk) ? panel[(size_t)lane * n + k0 + q] : (q == lane ? 1.0 : 0.0);
            bool bad = false;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                double pj[K];
#pragma unroll
                for (int q = 0; q < K; ++q) pj[q] = __shfl_sync(0xffffffffu, d[q], j);
                const double piv = pj[j], ap = fabs(piv);
                if (!(ap > 1e-10 * maxpiv) || !(ap > 0.0)) bad = true;
                maxpiv = fmax(maxpiv, ap);
                const double inv = 1.0 / piv;
                if (lane == j) {
#pragma unroll
                    for (int q = 0; q < K; ++q) d[q] = (q == j) ? inv : pj[q] * inv;
                } else {
                    const double f = d[j] * inv;
#pragma unroll
                    for (int q = 0; q < K; ++q) d[q] = (q == j) ? -f : d[q] - f * pj[q];
                }
            }
            if (lane < K) {
#pragma unroll
                for (int q = 0; q < K; ++q) Dinv[lane * K + q] = d[q];
            }
            if (bad && lane == 0) s_bad = 1;
        }
        __syncthreads();
        // update coefficients per own row (negated pivot-block rows of Dinv for the pivot rows themselves)
        for (int idx = tid; idx < nrows * K; idx += nt) {
            const int r = idx / K, q = idx - r * K, i = b + r * G;
            double wv = 0.0;
            if (q < kk) {
                if (i >= k0 && i < k0 + kk) wv = -Dinv[(i - k0) * K + q];
                else
                    for (int q2 = 0; q2 < kk; ++q2) wv += rows[(size_t)r * n + k0 + q2] * Dinv[q2 * K + q];
            }
            w[idx] = wv;
        }
        __syncthreads();
        for (int c = tid; c < n; c += nt) {
            const bool inP = c >= k0 && c < k0 + kk;
            double pv[K];
#pragma unroll
            for (int q = 0; q < K; ++q) pv[q] = (!inP && q < kk) ? panel[(size_t)q * n + c] : 0.0;
            for (int r = 0; r < nrows; ++r) {
                const int i = b + r * G;
                double* row = rows + (size_t)r * n;
                if (inP) { row[c] = -w[r * K + (c - k0)]; continue; }
                double v = (i >= k0 && i < k0 + kk) ? 0.0 : row[c];
#pragma unroll
                for (int q = 0; q < K; ++q) v -= w[r * K + q] * pv[q];
                row[c] = v;
            }
        }
        __syncthreads();
        const int next = k0 + K;
        if (next < n) {      // owners publish the pivot rows of the next step into the other half of the panel buffer
            double* nb = panelbuf + (size_t)((step + 1) & 1) * K * n;
            const int kn = min(K, n - next);
            for (int q = 0; q < kn; ++q) {
                const int i = next + q;
                if (i % G == b) {
                    const int r = i / G;
                    for (int c = tid; c < n; c += nt) nb[(size_t)q * n + c] = rows[(size_t)r * n + c];
                }
            }
            grid.sync();
        }
    }
    for (int r = 0; r < nrows; ++r)
        for (int c = tid; c < n; c += nt) Ainv[(size_t)(b + r * G) * n + c] = rows[(size_t)r * n + c];
    if (tid == 0 && s_bad) *fail = 2;
}

// Ainv[k][c] = M[pivrow[k]][n+c] / M[pivrow[k]][k]
__global__ void k_extract_inverse(int n, const double* __restrict__ M, const int* __restrict__ pivrow, double* __restrict__ Ainv) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)n * n; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / n), c = (int)(t - (int64_t)k * n);
        const int p = pivrow[k];
        Ainv[t] = M[(int64_t)p * 2 * n + n + c] / M[(int64_t)p * 2 * n + k];
    }
}
// x0 = scatter(Ainv * gather(b0)) ; Dirichlet dofs: x = b (identity rows).  The compact right-hand side is staged in
// shared memory once per block, then one warp per free row streams its row of the inverse with independent loads.
__global__ void __launch_bounds__(256) k_coarse_solve(int n, int ndof, const double* __restrict__ Ainv, const int* __restrict__ free2dof,
                                                      const int* __restrict__ dof2free, const double* b, double* __restrict__ x, int one) {
    pdl_trigger();
    extern __shared__ double sb[];
    constexpr int PF = 20;                                  // the first 32*PF columns of a warp's first row wait in registers
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // static data (the inverse, the index maps) is fetched before the dependency wait
    double a[PF];
    int myfree = 0;
    for (int rep = 0; rep < one; ++rep) {                   // one == 1: a loop boundary keeps ptxas from moving the wait above these loads
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int c = lane + 32 * k;
            a[k] = (warp < n && c < n) ? ld_static(Ainv + warp * n + c) : 0.0;
        }
        myfree = (warp < n && lane == 0) ? ld_static(free2dof + warp) : 0;
    }
    pdl_wait();
    for (int c = threadIdx.x; c < n; c += blockDim.x) sb[c] = b[free2dof[c]];
    __syncthreads();
    if (warp < n) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int c = lane + 32 * k;
            if (c < n) acc += a[k] * sb[c];
        }
        const double* row = Ainv + warp * n;
        for (int c = lane + 32 * PF; c < n; c += 32) acc += row[c] * sb[c];
        acc = warp_sum(acc);
        if (lane == 0) x[myfree] = acc;
    }
    for (int64_t i = warp + nwarps; i < n; i += nwarps) {
        const double* row = Ainv + i * n;
        double acc = 0.0;
#pragma unroll 8
        for (int c = lane; c < n; c += 32) acc += row[c] * sb[c];
        acc = warp_sum(acc);
        if (lane == 0) x[free2dof[i]] = acc;
    }
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ndof; t += (int64_t)gridDim.x * blockDim.x)
        if (dof2free[t] < 0) x[t] = b[t];
}

// ---------------------------------------------------------------------------------------------
// diagonal operator (P0 mass matrix) kernels for CG + Jacobi (3d_admm.lua:701-703)
// ---------------------------------------------------------------------------------------------
// q = diag .* p ; reduce <p,q>
__global__ void __launch_bounds__(256) k_diag_apply_dot(int64_t n, const double* __restrict__ diag, const double* __restrict__ p,
                                                        double* __restrict__ q, double* partials, unsigned int* ticket, double* out) {
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double qi = diag[i] * p[i];
        q[i] = qi;
        acc[0] += p[i] * qi;
    }
    grid_reduce<1, 0>(acc, partials, ticket, out);
}
// r = b - diag .* x ; z = damp * r / diag ; p = z ; reduce |r|^2, <r,z>
__global__ void __launch_bounds__(256) k_cg_init(int64_t n, double damp, const double* __restrict__ diag, const double* __restrict__ b,
                                                 const double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                 double* __restrict__ p, double* partials, unsigned int* ticket, double* out2) {
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double ri = b[i] - diag[i] * x[i];
        double zi = damp * ri / diag[i];
        r[i] = ri; z[i] = zi; p[i] = zi;
        acc[0] += ri * ri;
        acc[1] += ri * zi;
    }
    grid_reduce<2, 0>(acc, partials, ticket, out2);
}
// a = rz/pq ; x += a p ; r -= a q ; z = damp r/diag ; reduce |r|^2, <r,z>
__global__ void __launch_bounds__(256) k_cg_step(int64_t n, double damp, const double* __restrict__ sc, const double* __restrict__ diag,
                                                 const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
                                                 double* __restrict__ r, double* __restrict__ z, double* partials, unsigned int* ticket,
                                                 double* out2) {
    const double a = sc[SC_RZ] / sc[SC_PQ];
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += a * p[i];
        double ri = r[i] - a * q[i];
        double zi = damp * ri / diag[i];
        r[i] = ri; z[i] = zi;
        acc[0] += ri * ri;
        acc[1] += ri * zi;
    }
    grid_reduce<2, 0>(acc, partials, ticket, out2);
}
// p = z + (rz_new/rz_old) p ; roll scalars
__global__ void k_cg_p(int64_t n, const double* __restrict__ sc, const double* __restrict__ z, double* __restrict__ p) {
    const double beta = sc[SC_RZ] / sc[SC_RZ_OLD];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}
__global__ void k_cg_roll(double* sc, const double* out2) {
    sc[SC_RZ_OLD] = sc[SC_RZ];
    sc[SC_RR] = out2[0];
    sc[SC_RZ] = out2[1];
}

}  // namespace ab
