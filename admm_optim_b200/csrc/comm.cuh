// Multi-GPU communication layer: one process per GPU, NCCL over NVLink/NVSwitch.
// Replaces UG4's pcl/MPI layer (`mpirun -np 4 ugshell ...`, 3d_admm.lua:25; storage-type conversions
// 3d_admm.lua:912,982,1096,1222): interface sums (additive -> consistent) and scalar all-reduces.
// NCCL is resolved at run time from the library already loaded in the process (torch's bundled
// libnccl.so.2) so that libadmm_b200.so has no link-time dependency and never mixes two NCCL copies.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace ab {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get() {
        static NcclApi api;
        static bool loaded = false;
        if (!loaded) {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // torch's copy when already loaded
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            AB_REQUIRE(h, -4, std::string("cannot load libnccl.so.2: ") + dlerror());
#define AB_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); AB_REQUIRE(api.name, -4, "libnccl.so.2 lacks nccl" #name)
            AB_SYM(GetUniqueId); AB_SYM(CommInitRank); AB_SYM(CommDestroy); AB_SYM(CommAbort); AB_SYM(AllReduce); AB_SYM(Send); AB_SYM(Recv);
            AB_SYM(GroupStart); AB_SYM(GroupEnd); AB_SYM(GetErrorString);
#undef AB_SYM
            loaded = true;
        }
        return api;
    }
};

#define AB_NCCL(call)                                                                                                       \
    do {                                                                                                                    \
        ncclResult_t r__ = (call);                                                                                          \
        if (r__ != ncclSuccess) throw ab::Error(-2, std::string(#call) + " failed: " + ab::NcclApi::get().GetErrorString(r__)); \
    } while (0)

struct Comm {
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    // ncclCommDestroy is an intra-node collective (it blocks until every rank of the node calls it); a context may be torn down
    // by a garbage collector at a different time on every rank, or not at all on a rank that is exiting.  All work of the
    // context has been synchronised by then (ab_context_destroy), so the communicator is released with ncclCommAbort, which
    // frees the local resources without waiting for the peers.
    ~Comm() { if (comm) NcclApi::get().CommAbort(comm); }
    void allreduce(double* d, int n, bool max_op, cudaStream_t s) {
        AB_NCCL(NcclApi::get().AllReduce(d, d, (size_t)n, ncclFloat64, max_op ? ncclMax : ncclSum, comm, s));
    }
};

// interface (shared-vertex) lists of one level: for every neighbour rank the local vertex ids of the shared
// vertices in a canonical order both sides agree on (sorted by coordinates at setup)
struct Interface {
    std::vector<int> neigh, offset;   // neighbour ranks (ascending), offset[n]..offset[n+1] into idx
    DevBuf<int> idx;                  // concatenated per-neighbour local vertex ids ("slots")
    DevBuf<int> iv;                   // unique interface vertices, ordered by (first neighbour, slot): the stores into the first
                                      // neighbour's window are then consecutive
    DevBuf<int> iv_ptr, iv_slot, iv_nb;   // CSR unique vertex -> its slots and their neighbour index, neighbour ranks ascending
    DevBuf<unsigned char> owned;      // per local vertex: 1 when this rank is the lowest rank sharing it
    DevBuf<double> send, recv, save;  // NCCL fallback: packed buffers (total * maxcomp), smoother scratch (niv * 2 * D)
    int total = 0, niv = 0, my_pos = 0;   // my_pos: number of neighbours with a rank below mine
    // peer-to-peer path (NVLink, CUDA IPC): neighbours write straight into this rank's window
    DevBuf<int> d_offset, d_neigh;                 // device copies of offset / neigh
    DevBuf<unsigned long long> d_peer_dst;          // per neighbour: address (in this process) of my slot in the neighbour's window, parity 0
    DevBuf<unsigned long long> d_peer_stride;       // per neighbour: byte distance between the neighbour's two parity buffers
    DevBuf<unsigned long long> d_peer_flag;         // per neighbour: address of the neighbour's flag word for this rank
    double* win_recv = nullptr;                     // my two parity buffers inside the window (total*D doubles each)
    unsigned long long* win_flags = nullptr;        // my flag words (one per rank) for this level
    DevBuf<unsigned long long> state;               // device side: [0] epoch of the last completed exchange, [1] arrive counter, [2] done counter
};

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
__global__ void __launch_bounds__(256) k_iface_xchg(int niv, int D, int nneigh, int my_pos, const int* __restrict__ iv, const int* __restrict__ iv_ptr,
                                                    const int* __restrict__ iv_slot, const int* __restrict__ iv_nb, const int* __restrict__ offset,
                                                    const int* __restrict__ neigh, const unsigned long long* __restrict__ peer_dst,
                                                    const unsigned long long* __restrict__ peer_stride, const unsigned long long* __restrict__ peer_flag,
                                                    int total, const double* my_recv, unsigned long long* my_flags, unsigned long long* state,
                                                    int* err, double* v, const double* cf, const double* din, const double* xin, double* xout) {
    // The grid is capped by the host (kXchgCtasPerSm CTAs per SM, far below the residency limit of this 32-register kernel): every
    // CTA of the grid is resident while it waits for the neighbours, so the CTAs that still have to store can always run -- a grid
    // larger than the device could hold would dead-lock against the neighbour's equally oversized grid.  The entries are therefore
    // walked with a grid-stride loop, in the put phase and again in the sum phase.
    __shared__ int s_fail;
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
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
            }
        }
    }
    for (int n = threadIdx.x; n < nneigh; n += blockDim.x) {              // 3. wait: one polling thread per neighbour
        long long spins = 0;
        unsigned long long f;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(my_flags + neigh[n]) : "memory");
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

__global__ void k_iface_pack(int total, int D, const int* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        buf[t] = v[(int64_t)idx[k] * D + c];
    }
}
__global__ void k_iface_unpack_add(int total, int D, const int* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ v) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        atomicAdd(v + (int64_t)idx[k] * D + c, buf[t]);
    }
}
// zero the copies this rank does not own (consistent -> unique, a valid additive representation)
__global__ void k_zero_not_owned(int64_t n, int D, const unsigned char* __restrict__ owned, double* __restrict__ v) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        if (!owned[t / D]) v[t] = 0.0;
}
// dot products that count every shared vertex once: sum over owned vertices of x_k . y
template <int NX>
__global__ void __launch_bounds__(256) k_dot_owned(int64_t n, int D, const unsigned char* __restrict__ owned, const double* x0, const double* x1,
                                                   const double* __restrict__ y, double* partials, unsigned int* ticket, double* out) {
    const double* xs[2] = {x0, x1};
    double v[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) v[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (owned[i / D]) {
            const double yi = y[i];
#pragma unroll
            for (int k = 0; k < NX; ++k) v[k] += xs[k][i] * yi;
        }
    }
    grid_reduce<NX, 0>(v, partials, ticket, out);
}
// smoother interface fix-up (see Gmg::smooth): save d and x at the interface vertices before the fused step ...
__global__ void k_iface_save(int niv, int D, const int* __restrict__ iv, const double* __restrict__ d, const double* __restrict__ x,
                             double* __restrict__ save) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        save[t] = d ? d[i] : 0.0;
        save[niv * D + t] = x ? x[i] : 0.0;
    }
}
// ... turn the locally updated d into the additive increment inc = d_new_local - c1*d_old ...
__global__ void k_iface_inc(int niv, int D, const int* __restrict__ iv, double c1, const double* __restrict__ save, double* __restrict__ d) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        d[i] = d[i] - c1 * save[t];
    }
}
// ... and after the interface sum rebuild d = c1*d_old + inc_total, x_new = x_old + d
__global__ void k_iface_fix(int niv, int D, const int* __restrict__ iv, double c1, const double* __restrict__ save, double* __restrict__ d,
                            double* __restrict__ xnew) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        const double dn = c1 * save[t] + d[i];
        d[i] = dn;
        xnew[i] = save[niv * D + t] + dn;
    }
}

}  // namespace ab
