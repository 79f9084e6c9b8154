// The NVLink peer-memory interface exchange: host-side slot tables + the fused kernel.  Kept free of every other dependency so
// that tests/cuda_host_shim/xchg_emulation.cpp can compile THIS source for the host (one std::thread per CUDA thread, C++ atomics
// for the release / acquire flag traffic, ThreadSanitizer on) and check the protocol -- epochs over many launches, parity double
// buffering with ranks one exchange apart, capped grids with grid-stride phases, rank-ordered sums -- without a GPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ab {

// unique interface vertices ordered by (first neighbour, slot); CSR vertex -> (slot, neighbour index), neighbours ascending.
// neigh: neighbour ranks ascending (this rank excluded); offset[n]..offset[n+1]: the slots of neighbour n in idx (local vertex ids
// in the order both sides of the pair agreed on).
struct IfaceCsr {
    std::vector<int> iv, ptr, slot, nb;
    int my_pos = 0;                     // number of neighbours with a rank below this one's
};
inline IfaceCsr build_iface_csr(int nv, int me, const std::vector<int>& neigh, const std::vector<int>& offset, const std::vector<int>& idx) {
    IfaceCsr C;
    for (int q : neigh)
        if (q < me) C.my_pos++;
    std::vector<int> pos((size_t)nv, -1);
    for (size_t n = 0; n < neigh.size(); ++n)
        for (int k = offset[n]; k < offset[n + 1]; ++k)
            if (pos[idx[k]] < 0) { pos[idx[k]] = (int)C.iv.size(); C.iv.push_back(idx[k]); }
    C.ptr.assign(C.iv.size() + 1, 0);
    for (int v : idx) C.ptr[pos[v] + 1]++;
    for (size_t k = 0; k < C.iv.size(); ++k) C.ptr[k + 1] += C.ptr[k];
    const int total = offset.empty() ? 0 : offset.back();
    std::vector<int> fill(C.ptr.begin(), C.ptr.end() - 1);
    C.slot.resize((size_t)total);
    C.nb.resize((size_t)total);
    for (size_t n = 0; n < neigh.size(); ++n)
        for (int k = offset[n]; k < offset[n + 1]; ++k) {
            const int e = fill[pos[idx[k]]]++;
            C.slot[e] = k; C.nb[e] = (int)n;
        }
    return C;
}

#ifndef AB_HOST_EMULATION
#define AB_XCHG_KERNEL __global__ void __launch_bounds__(256)
#define AB_XCHG_SHARED(type, name) __shared__ type name
__device__ __forceinline__ void ab_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ab_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#endif

// Fused interface sum over NVLink peer memory -- pack, transfer, synchronisation and unpack in ONE launch, no NCCL call, no
// grid-wide barrier.  One thread per (unique interface vertex, component):
//   1. put   : the thread stores its additive value into the window of every neighbour that shares the vertex;
//   2. signal: the CTA that arrives last publishes this exchange's epoch in every neighbour's flag word (st.release.sys);
//   3. wait  : every CTA polls the neighbours' flags for the epoch (ld.acquire.sys; bounded spin -> error flag, never a hang);
//   4. sum   : the thread adds what the neighbours wrote for ITS vertex, contributions taken in ascending rank order with its own
//              value at its own rank's place -- every rank sharing a vertex performs the same additions in the same order, so the
//              consistent copies are bitwise identical on all ranks, and the result is written by the thread that read the input
//              (no intra-grid hazard, no grid-wide barrier).
// The epoch lives in device memory (state[0]) and is advanced by the CTA that finishes last, so the launch carries no
// host-side counter and can be captured into a CUDA graph.  Windows are double-buffered by epoch parity (a neighbour is at
// most one exchange ahead: it cannot finish exchange e+1 before this rank has published e+1, i.e. finished reading e).
// On a timeout the sum phase is skipped (v keeps its additive value) and *err is set; the solvers read err with their scalars.
//   SMOOTH = false: v <- sum over the sharing ranks of v                                  (additive -> consistent)
//   SMOOTH = true : the Chebyshev/Jacobi interface fix-up fused around the sum (Gmg::smooth): the locally updated
//                   v = c1 d_in + c2 D^-1 r_local is reduced to its additive increment, summed, and d, x are rebuilt at the
//                   shared vertices:  v = c1 d_in + total,  x_out = x_in + v.
constexpr long long kP2PSpinLimit = 20000000ll;   // ~10-20 s of polling

template <bool SMOOTH>
AB_XCHG_KERNEL k_iface_xchg(int niv, int D, int nneigh, int my_pos, const int* __restrict__ iv, const int* __restrict__ iv_ptr,
                                                    const int* __restrict__ iv_slot, const int* __restrict__ iv_nb, const int* __restrict__ offset,
                                                    const int* __restrict__ neigh, const unsigned long long* __restrict__ peer_dst,
                                                    const unsigned long long* __restrict__ peer_stride, const unsigned long long* __restrict__ peer_flag,
                                                    int total, const double* my_recv, unsigned long long* my_flags, unsigned long long* state,
                                                    int* err, double* v, const double* cf, const double* din, const double* xin, double* xout) {
    // The grid is capped by the host (lib.cu launch_xchg: three quarters of the residency limit of this 32-register kernel, at least
    // kXchgCtasPerSm CTAs per SM): every
    // CTA of the grid is resident while it waits for the neighbours, so the CTAs that still have to store can always run -- a grid
    // larger than the device could hold would dead-lock against the neighbour's equally oversized grid.  The entries are therefore
    // walked with a grid-stride loop, in the put phase and again in the sum phase.
    AB_XCHG_SHARED(int, s_fail);
    const unsigned long long epoch = *(volatile unsigned long long*)state + 1ull;   // stable until the last CTA of THIS launch is done
    const int parity = (int)(epoch & 1ull);
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x, n_ent = niv * D;
    const double c1 = (SMOOTH && cf) ? cf[0] : 0.0;
    if (threadIdx.x == 0) s_fail = 0;
    for (int t = t0; t < n_ent; t += nthreads) {                          // 1. put
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        double own = v[i];
        // additive increment c2 D^-1 r_local.  Separate multiply and subtract (no FMA contraction): the sum phase below recomputes
        // this value and must get the same bits as the copy sent to the neighbours
        if (SMOOTH && c1 != 0.0 && din) own = __dsub_rn(own, __dmul_rn(c1, din[i]));
        for (int e = iv_ptr[k]; e < iv_ptr[k + 1]; ++e) {
            const int nb = iv_nb[e];
            double* dst = reinterpret_cast<double*>(peer_dst[nb] + (unsigned long long)parity * peer_stride[nb]) + (size_t)(iv_slot[e] - offset[nb]) * D + c;
            *dst = own;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {                                               // 2. signal
        const unsigned long long old = atomicAdd(state + 1, 1ull);
        if (old + 1 == gridDim.x) {                                       // every CTA's stores are out: publish the epoch
            __threadfence_system();
            for (int n = 0; n < nneigh; ++n) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(peer_flag[n]);
                ab_st_release_sys(f, epoch);
            }
        }
    }
    for (int n = threadIdx.x; n < nneigh; n += blockDim.x) {              // 3. wait: one polling thread per neighbour
        long long spins = 0;
        unsigned long long f;
        do {
            f = ab_ld_acquire_sys(my_flags + neigh[n]);
            if (f >= epoch) break;
            if (++spins > kP2PSpinLimit) { s_fail = 1; break; }
        } while (true);
    }
    __syncthreads();
    const bool fail = s_fail != 0;
    if (fail && threadIdx.x == 0) *err = 1;
    if (!fail) {
        const double* buf = my_recv + (size_t)parity * total * D;
        for (int t = t0; t < n_ent; t += nthreads) {                      // 4. sum: v[i] is read and written by this thread only
            const int k = t / D, c = t - k * D;
            const int64_t i = (int64_t)iv[k] * D + c;
            double own = v[i], dold = 0.0;
            if (SMOOTH && c1 != 0.0 && din) { dold = __dmul_rn(c1, din[i]); own = __dsub_rn(own, dold); }
            const int e1 = iv_ptr[k + 1];
            int e = iv_ptr[k];
            double tot = 0.0;
            for (int pos = 0; pos <= nneigh; ++pos) {                      // ascending rank order, own value at position my_pos
                if (pos == my_pos) { tot += own; continue; }
                const int nb = pos < my_pos ? pos : pos - 1;
                if (e < e1 && iv_nb[e] == nb) { tot += __ldcg(buf + (size_t)iv_slot[e] * D + c); ++e; }
            }
            if (SMOOTH) {
                const double dn = dold + tot;
                v[i] = dn;
                xout[i] = (xin ? xin[i] : 0.0) + dn;
            } else {
                v[i] = tot;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(state + 2, 1ull);
        if (old + 1 == gridDim.x) {                                       // last CTA out: every CTA has read the epoch and arrived
            state[1] = 0ull;
            state[2] = 0ull;
            __threadfence();
            *(volatile unsigned long long*)state = epoch;
        }
    }
}
constexpr int kXchgCtasPerSm = 2;

}  // namespace ab
