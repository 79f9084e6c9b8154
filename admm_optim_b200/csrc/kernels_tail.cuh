// Coarse tail of the V-cycle (levels 1 and 0) as ONE kernel on one thread-block cluster.
//
// Levels 0 and 1 are the same few thousand vertices whatever the refinement of the top level; as separate launches they
// cost ten dependent kernels of 3-5 us each per cycle (smoother steps, residual, restriction, coarse solve, prolongation)
// that are pure latency.  Here one cluster of NC CTAs keeps the level-1 operator resident in its shared memory, split by
// block rows, for the whole tail:
//
//   pre-smoothing (Chebyshev / Jacobi steps) -> residual -> restriction -> coarse solve (GEMV with the dense inverse)
//   -> prolongation -> post-smoothing
//
// Every CTA holds a FULL copy of the vector the next sparse product gathers from (x, later the residual) in its own
// shared memory; after a step each CTA stores its slice into all NC copies through distributed shared memory and the
// hardware cluster barrier orders the exchange (two barriers per step: everybody has finished reading the old copy /
// the new copy is complete).  Same arithmetic, operand order and coefficients as the separate kernels it replaces
// (k_smooth_first, k_bsr_spmv_tma modes 1/2, k_restrict, k_coarse_solve, k_prolong_add).
//
// Launched with programmatic dependent launch: the static operands (matrix slice, diagonal, tables) are staged before the
// dependency wait, i.e. while the level-2 restriction that produces the right-hand side is still running.
//
// STATUS (round 1, B200, 44 730-DoF hierarchy, 16-CTA cluster): bit-compatible with the separate kernels (tests), but SLOWER:
// 63 us per tail against ~30 us for the ten chained launches -- a shared-memory sparse product of 133 rows takes 3.8 us
// (latency-bound dependent chain on 16 warps) and each exchange (barrier, DSMEM broadcast, barrier) 3.7 us, i.e. as much
// as a kernel boundary under PDL.  Kept as an opt-in experiment (ADMM_B200_TAIL=1 / tuning key "tail"), off by default;
// phase timings: ADMM_B200_TAIL_PROF=1 (tools/tail_prof.py, profiles/r01_tail_phases.md).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace ab {
namespace cgt = cooperative_groups;

struct TailArgs {
    // level 1 (fine level of the tail)
    int nv1, nvc;                       // vertices of level 1, of level 0 (= copied vertices of level 1)
    const int* rowptr1;
    const int* colidx1;
    const double* vals1;
    const double* dinv1;
    const int* part;                    // NC+1 first block rows of the CTA slices
    const double* cf_pre;               // (c1,c2) per pre-smoothing step, device
    const double* cf_post;
    int nu_pre, nu_post;
    // transfers
    const int* rowptr0;                 // level-0 vertex graph: entry k of row v <-> coarse edge (v, col), mid0[k] = fine midpoint
    const int* mid0;
    const int* diagpos0;
    const unsigned char* mask0;         // Dirichlet mask of level 0 (may be null)
    const int* pa1;                     // parents of the level-1 vertices nvc..nv1-1
    const int* pb1;
    // coarse solve
    int n_free, n0;
    const double* Ainv;
    const int* free2dof;
    const int* dof2free;
    // vectors
    const double* b;                    // right-hand side on level 1 (written by the preceding kernel)
    double* x;                          // result on level 1
    // shared-memory layout (in doubles / ints), computed on the host
    int max_rows, max_blocks;
    unsigned long long* prof;           // optional: %globaltimer stamps of the phases (rank 0, thread 0), diagnostics
};

template <int D>
struct TailSmem {
    // byte offsets for given sizes; everything 16-byte aligned
    static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
    size_t vals, cols, rp, f2d, vfull, b0, x0, bown, down, dinvown, xown, total;
    TailSmem(int nv1, int n0, int n_free, int max_rows, int max_blocks) {
        size_t o = 0;
        vals = o; o = align16(o + (size_t)max_blocks * D * D * 8);
        cols = o; o = align16(o + (size_t)max_blocks * 4);
        rp = o; o = align16(o + (size_t)(max_rows + 1) * 4);
        f2d = o; o = align16(o + (size_t)std::max(n_free, 1) * 4);
        vfull = o; o = align16(o + (size_t)nv1 * D * 8);
        b0 = o; o = align16(o + (size_t)n0 * 8);
        x0 = o; o = align16(o + (size_t)n0 * 8);
        bown = o; o = align16(o + (size_t)max_rows * D * 8);
        down = o; o = align16(o + (size_t)max_rows * D * 8);
        dinvown = o; o = align16(o + (size_t)max_rows * D * 8);
        xown = o; o = align16(o + (size_t)max_rows * D * 8);
        total = o;
    }
};

// y_slice = A_slice * vfull for the rows of this CTA; calls epi(local row, component, value) for every entry.
// HL lanes per block row, lane <-> (block slot, column) as in k_bsr_spmv_tma.
template <int D, typename Epi>
__device__ __forceinline__ void tail_spmv(int nr, const int* __restrict__ s_rp, const int* __restrict__ s_cols, const double* __restrict__ s_vals,
                                          const double* vfull, Epi epi) {
    constexpr int DD = D * D;
    constexpr int HL = D == 3 ? 16 : 8, RPW = 32 / HL, BPH = HL / D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int hl = lane % HL, half = lane / HL;
    const int lb = hl / D, c = hl - lb * D;
    const bool lane_on = lb < BPH;
    for (int base = warp * RPW; base < nr; base += nwarps * RPW) {
        const int lr = base + half;
        const bool row_on = lr < nr;
        const int s = row_on ? s_rp[lr] : 0, e = row_on ? s_rp[lr + 1] : 0;
        double acc[D];
#pragma unroll
        for (int r = 0; r < D; ++r) acc[r] = 0.0;
        if (lane_on) {
            for (int blk = s + lb; blk < e; blk += BPH) {
                const double xv = vfull[s_cols[blk] * D + c];
                const double* ap = s_vals + (size_t)blk * DD + c;
#pragma unroll
                for (int r = 0; r < D; ++r) acc[r] = fma(ap[r * D], xv, acc[r]);
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
            epi(lr, hl, v);
        }
    }
}

// store `cnt` doubles of this CTA (src, shared) into every CTA's copy of `dst_base` at element offset `off`
__device__ __forceinline__ void tail_broadcast(cgt::cluster_group& cl, double* dst_base, int off, const double* src, int cnt) {
    const int NC = (int)cl.num_blocks();
    for (int dst = 0; dst < NC; ++dst) {
        double* remote = cl.map_shared_rank(dst_base, dst);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) remote[off + i] = src[i];
    }
}

template <int D>
__global__ void __launch_bounds__(512, 1) k_vcycle_tail(const TailArgs a) {
    cgt::cluster_group cl = cgt::this_cluster();
    pdl_trigger();
    // diagnostics: %globaltimer stamps of the phases, kept in registers and written at the very end (nothing is stored to global
    // memory before the dependency wait, tools/check_pdl_sass.py)
    unsigned long long ts[13];
    int prof_i = 0;
    const bool prof_on = a.prof && cl.block_rank() == 0 && threadIdx.x == 0;
    auto stamp = [&]() {
        if (prof_on) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            ts[prof_i] = t;
        }
        ++prof_i;
    };
    stamp();   // 0: start
    constexpr int DD = D * D;
    extern __shared__ __align__(16) unsigned char smem_tail[];
    const int rank = (int)cl.block_rank(), NC = (int)cl.num_blocks();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n1 = a.nv1 * D;
    // ---- shared-memory carve-up (must match TailSmem on the host) ----
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char* p = smem_tail + o; o = (o + bytes + 15) & ~(size_t)15; return p; };
    double* s_vals = reinterpret_cast<double*>(take((size_t)a.max_blocks * DD * 8));
    int* s_cols = reinterpret_cast<int*>(take((size_t)a.max_blocks * 4));
    int* s_rp = reinterpret_cast<int*>(take((size_t)(a.max_rows + 1) * 4));
    int* s_f2d = reinterpret_cast<int*>(take((size_t)max(a.n_free, 1) * 4));
    double* vfull = reinterpret_cast<double*>(take((size_t)n1 * 8));
    double* b0full = reinterpret_cast<double*>(take((size_t)a.n0 * 8));
    double* x0full = reinterpret_cast<double*>(take((size_t)a.n0 * 8));
    double* s_b = reinterpret_cast<double*>(take((size_t)a.max_rows * D * 8));
    double* s_d = reinterpret_cast<double*>(take((size_t)a.max_rows * D * 8));
    double* s_dinv = reinterpret_cast<double*>(take((size_t)a.max_rows * D * 8));
    double* s_x = reinterpret_cast<double*>(take((size_t)a.max_rows * D * 8));

    // ---- static prologue: operator slice, diagonal, index maps (none of it is written by the kernels of the chain) ----
    const int r0 = a.part[rank], r1 = a.part[rank + 1], nr = r1 - r0, nd = nr * D;
    const int bs = a.rowptr1[r0], be = a.rowptr1[r1], nb = be - bs;
    for (int i = tid; i < nb * DD; i += nt) s_vals[i] = a.vals1[(size_t)bs * DD + i];
    for (int i = tid; i < nb; i += nt) s_cols[i] = a.colidx1[bs + i];
    for (int i = tid; i <= nr; i += nt) s_rp[i] = a.rowptr1[r0 + i] - bs;
    for (int i = tid; i < nd; i += nt) s_dinv[i] = a.dinv1[(size_t)r0 * D + i];
    for (int i = tid; i < a.n_free; i += nt) s_f2d[i] = a.free2dof[i];
    // slices of the coarse level
    const int va = (int)((int64_t)a.nvc * rank / NC), vb = (int)((int64_t)a.nvc * (rank + 1) / NC);           // coarse vertices
    const int fa = (int)((int64_t)a.n_free * rank / NC), fb = (int)((int64_t)a.n_free * (rank + 1) / NC);     // free coarse rows
    stamp();   // 1: static prologue issued
    cl.sync();                 // every CTA of the cluster is running: its shared memory may be written remotely from here on
    stamp();   // 2: first cluster barrier
    pdl_wait();                // the right-hand side comes from the preceding kernel
    stamp();   // 3: dependency wait

    // ---- pre-smoothing, step 0 from a zero guess: d = c2 D^-1 b ; x = d ----
    {
        const double c2 = a.cf_pre[1];
        for (int i = tid; i < nd; i += nt) {
            const double bi = a.b[(size_t)r0 * D + i];
            s_b[i] = bi;
            const double dn = c2 * s_dinv[i] * bi;
            s_d[i] = dn;
            s_x[i] = dn;
        }
    }
    __syncthreads();
    tail_broadcast(cl, vfull, r0 * D, s_x, nd);
    cl.sync();
    // ---- further smoothing steps: d = c1 d + c2 D^-1 (b - A x) ; x += d ----
    auto smooth_step = [&](double c1, double c2, bool publish) {
        tail_spmv<D>(nr, s_rp, s_cols, s_vals, vfull, [&](int lr, int comp, double v) {
            const int i = lr * D + comp;
            const double res = s_b[i] - v;
            const double dn = (c1 != 0.0 ? c1 * s_d[i] : 0.0) + c2 * s_dinv[i] * res;
            s_d[i] = dn;
            s_x[i] = vfull[r0 * D + i] + dn;
        });
        if (publish) {
            cl.sync();             // everybody has finished gathering from the old copy
            tail_broadcast(cl, vfull, r0 * D, s_x, nd);
            cl.sync();             // the new copy is complete
        } else {
            __syncthreads();
        }
    };
    stamp();   // 4: first step + broadcast
    for (int k = 1; k < a.nu_pre; ++k) smooth_step(a.cf_pre[2 * k], a.cf_pre[2 * k + 1], true);
    stamp();   // 5: pre-smoothing steps
    // ---- residual r = b - A x (into s_d: the post-smoother restarts its recurrence), published for the restriction ----
    tail_spmv<D>(nr, s_rp, s_cols, s_vals, vfull, [&](int lr, int comp, double v) {
        const int i = lr * D + comp;
        s_d[i] = s_b[i] - v;
    });
    stamp();   // 6: residual product
    cl.sync();
    tail_broadcast(cl, vfull, r0 * D, s_d, nd);
    cl.sync();
    stamp();   // 7: residual published
    // ---- restriction b0 = mask * P^T r for the coarse vertices [va, vb): half-warp per vertex ----
    {
        const int hl = tid & 15, hw = tid >> 4, nhw = nt >> 4;
        for (int base = va; base < vb; base += nhw) {
            const int v = base + hw;
            const bool on = v < vb;
            const int s = on ? a.rowptr0[v] : 0, e = on ? a.rowptr0[v + 1] : 0, dp = on ? a.diagpos0[v] : -1;
            double acc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = 0.0;
            for (int k = s + hl; k < e; k += 16) {
                if (k == dp) continue;
                const int m = a.mid0[k] * D;
#pragma unroll
                for (int c = 0; c < D; ++c) acc[c] += vfull[m + c];
            }
#pragma unroll
            for (int c = 0; c < D; ++c) {
#pragma unroll
                for (int o2 = 8; o2 > 0; o2 >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o2);
            }
            if (on && hl < D) {
                double h = acc[0];
#pragma unroll
                for (int c = 1; c < D; ++c) h = (hl == c) ? acc[c] : h;
                double val = vfull[v * D + hl] + 0.5 * h;
                if (a.mask0 && ((a.mask0[v] >> hl) & 1)) val = 0.0;
                for (int dst = 0; dst < NC; ++dst) cl.map_shared_rank(b0full, dst)[v * D + hl] = val;
            }
        }
    }
    cl.sync();
    stamp();   // 8: restriction
    // ---- coarse solve: x0 = scatter(Ainv * gather(b0)) ; Dirichlet dofs: x0 = b0 ----
    {
        const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
        const int n = a.n_free;
        constexpr int PF = 20;
        for (int i = fa + warp; i < fb; i += nwarps) {
            const double* row = a.Ainv + (size_t)i * n;
            double av[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int c = lane + 32 * k;
                av[k] = c < n ? row[c] : 0.0;
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int c = lane + 32 * k;
                if (c < n) acc += av[k] * b0full[s_f2d[c]];
            }
            for (int c = lane + 32 * PF; c < n; c += 32) acc += row[c] * b0full[s_f2d[c]];
            acc = warp_sum(acc);
            if (lane == 0) {
                const int dof = s_f2d[i];
                for (int dst = 0; dst < NC; ++dst) cl.map_shared_rank(x0full, dst)[dof] = acc;
            }
        }
        for (int t = va * D + tid; t < vb * D; t += nt)
            if (a.dof2free[t] < 0) {
                const double val = b0full[t];
                for (int dst = 0; dst < NC; ++dst) cl.map_shared_rank(x0full, dst)[t] = val;
            }
    }
    cl.sync();
    stamp();   // 9: coarse solve
    // ---- prolongation x += P x0 for the own rows, published for the post-smoother ----
    for (int i = tid; i < nd; i += nt) {
        const int lr = i / D, c = i - lr * D, v = r0 + lr;
        double add;
        if (v < a.nvc) add = x0full[v * D + c];
        else {
            const int k = v - a.nvc;
            add = 0.5 * (x0full[a.pa1[k] * D + c] + x0full[a.pb1[k] * D + c]);
        }
        s_x[i] += add;
    }
    __syncthreads();
    tail_broadcast(cl, vfull, r0 * D, s_x, nd);    // the residual copy is no longer read: every CTA passed the barrier after its restriction
    cl.sync();
    stamp();   // 10: prolongation published
    // ---- post-smoothing ----
    for (int k = 0; k < a.nu_post; ++k) smooth_step(a.cf_post[2 * k], a.cf_post[2 * k + 1], k + 1 < a.nu_post);
    stamp();   // 11: post-smoothing
    for (int i = tid; i < nd; i += nt) a.x[(size_t)r0 * D + i] = s_x[i];
    cl.sync();                 // no CTA exits while its shared memory may still be the target of a remote access
    stamp();   // 12: end
    if (prof_on) {
#pragma unroll
        for (int i = 0; i < 13; ++i) a.prof[i] = ts[i];
    }
}

}  // namespace ab
