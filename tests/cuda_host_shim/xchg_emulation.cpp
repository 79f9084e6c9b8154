// HOST EMULATION of the NVLink peer-memory interface exchange (admm_optim_b200/csrc/iface_xchg.cuh) -- test infrastructure.
// The kernel SOURCE the library ships is compiled here for the CPU: one std::thread per CUDA thread, a std::barrier per CTA for
// __syncthreads, C++ atomics for atomicAdd / st.release.sys / ld.acquire.sys, the "GPUs" of R ranks living in one process with
// their receive windows as ordinary arrays.  What it checks (tests/test_host.py::test_interface_exchange_protocol_on_the_host):
//   * the slot tables of build_iface_csr + the index arithmetic of the put and the sum phase,
//   * the epoch protocol over many back-to-back launches, SMOOTH and plain alternating, with the ranks drifting up to one exchange
//     apart (random delays): parity double buffering, counter reset by the last CTA, flags only ever growing,
//   * capped grids: fewer threads than entries (grid-stride put and sum phases), more than one CTA per rank,
//   * rank-ordered sums: the consistent copies are BITWISE identical on all ranks and equal to the sum in ascending rank order,
//   * entries that are not on an interface are never touched,
//   * and, when built with -fsanitize=thread, that every window / vector access is ordered by the release-acquire edges of the
//     protocol (no data race under the C++ memory model, the relaxed device atomics + fences modelled as acq_rel operations).
// It does not check anything about the GPU memory system, NVLink or co-residency.
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

// ---- CUDA stand-ins ----------------------------------------------------------------------------------------------------
struct EmuDim { unsigned x = 0; };
static thread_local EmuDim blockIdx, threadIdx, blockDim, gridDim;
static thread_local std::barrier<>* emu_cta_barrier = nullptr;
static thread_local void* emu_cta_shared = nullptr;

#define AB_HOST_EMULATION 1
#define AB_XCHG_KERNEL void
#define AB_XCHG_SHARED(type, name) type& name = *static_cast<type*>(emu_cta_shared)
#define __restrict__
static inline void __syncthreads() { emu_cta_barrier->arrive_and_wait(); }
#if defined(__SANITIZE_THREAD__)      // ThreadSanitizer does not model stand-alone fences: the orderings they give on the device are carried by
static inline void __threadfence() {}          // the acq_rel / release / acquire operations below, which it does model
static inline void __threadfence_system() {}
#else
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
#endif
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline double __ldcg(const double* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }          // built with -ffp-contract=off: no fused multiply-add
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline void ab_st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ab_ld_acquire_sys(const unsigned long long* p) {
    const unsigned long long v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    std::this_thread::yield();
    return v;
}

#include "iface_xchg.cuh"

// ---- the emulated machine ----------------------------------------------------------------------------------------------
struct Rank {
    int me = 0, nv = 0, D = 0;
    std::vector<int> glob;                       // local vertex -> global vertex
    std::vector<int> neigh, offset, idx;          // what ab_domain_set_interface receives
    ab::IfaceCsr csr;
    int total = 0, niv = 0;
    std::vector<double> window;                   // two parity buffers of total * D doubles
    std::vector<unsigned long long> flags;        // one word per rank
    std::vector<unsigned long long> peer_dst, peer_stride, peer_flag;
    unsigned long long state[4] = {0, 0, 0, 0};
    int err = 0;
    std::vector<double> v, din, xin, xout, cf;
};

static double val(int kind, int r, int e, int g, int c) {      // reproducible pseudo-random data, exactly representable mixing
    unsigned long long h = 1469598103934665603ull;
    for (unsigned long long x : {(unsigned long long)kind, (unsigned long long)r, (unsigned long long)e, (unsigned long long)g, (unsigned long long)c}) {
        h ^= x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h *= 1099511628211ull;
    }
    return (double)(h >> 11) / 9007199254740992.0 - 0.5 + 1e-3 * kind;
}

template <bool SMOOTH>
static void launch(Rank& R, int grid, int block) {
    std::vector<std::thread> th;
    std::vector<std::barrier<>*> bars;
    std::vector<int*> shared;
    for (int b = 0; b < grid; ++b) { bars.push_back(new std::barrier<>(block)); shared.push_back(new int(0)); }
    for (int b = 0; b < grid; ++b)
        for (int t = 0; t < block; ++t)
            th.emplace_back([&, b, t]() {
                blockIdx.x = b; threadIdx.x = t; blockDim.x = block; gridDim.x = grid;
                emu_cta_barrier = bars[b];
                emu_cta_shared = shared[b];
                ab::k_iface_xchg<SMOOTH>(R.niv, R.D, (int)R.neigh.size(), R.csr.my_pos, R.csr.iv.data(), R.csr.ptr.data(), R.csr.slot.data(), R.csr.nb.data(),
                                         R.offset.data(), R.neigh.data(), R.peer_dst.data(), R.peer_stride.data(), R.peer_flag.data(), R.total,
                                         R.window.data(), R.flags.data(), R.state, &R.err, R.v.data(), SMOOTH ? R.cf.data() : nullptr,
                                         SMOOTH ? R.din.data() : nullptr, SMOOTH ? R.xin.data() : nullptr, SMOOTH ? R.xout.data() : nullptr);
            });
    for (auto& t : th) t.join();
    for (auto* b : bars) delete b;
    for (auto* s : shared) delete s;
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? std::atoi(argv[1]) : 4, D = argc > 2 ? std::atoi(argv[2]) : 3, nglob = argc > 3 ? std::atoi(argv[3]) : 60;
    const int epochs = argc > 4 ? std::atoi(argv[4]) : 8, grid = argc > 5 ? std::atoi(argv[5]) : 3, block = argc > 6 ? std::atoi(argv[6]) : 8;
    std::mt19937 rng(12345);
    // every global vertex lives on 1..3 ranks
    std::vector<std::vector<int>> owners(nglob);
    for (int g = 0; g < nglob; ++g) {
        const int k = 1 + (int)(rng() % 3);
        while ((int)owners[g].size() < std::min(k, R)) {
            const int r = (int)(rng() % R);
            bool have = false;
            for (int q : owners[g]) have |= q == r;
            if (!have) owners[g].push_back(r);
        }
        std::sort(owners[g].begin(), owners[g].end());
    }
    std::vector<Rank> ranks(R);
    for (int r = 0; r < R; ++r) {
        Rank& K = ranks[r];
        K.me = r; K.D = D;
        for (int g = 0; g < nglob; ++g)
            for (int q : owners[g]) if (q == r) K.glob.push_back(g);
        std::shuffle(K.glob.begin(), K.glob.end(), rng);              // local numbering unrelated to the global one
        K.nv = (int)K.glob.size();
        K.offset.push_back(0);
        for (int q = 0; q < R; ++q) {
            if (q == r) continue;
            std::vector<std::pair<int, int>> common;                  // (global id, local id): the pair's canonical order = ascending global id
            for (int l = 0; l < K.nv; ++l)
                for (int o : owners[K.glob[l]]) if (o == q) common.push_back({K.glob[l], l});
            if (common.empty()) continue;
            std::sort(common.begin(), common.end());
            K.neigh.push_back(q);
            for (auto& c : common) K.idx.push_back(c.second);
            K.offset.push_back((int)K.idx.size());
        }
        K.csr = ab::build_iface_csr(K.nv, r, K.neigh, K.offset, K.idx);
        K.total = K.offset.back(); K.niv = (int)K.csr.iv.size();
        K.window.assign((size_t)2 * std::max(K.total, 1) * D, -777.0);
        K.flags.assign(R, 0);
        K.v.resize((size_t)K.nv * D); K.din.resize(K.v.size()); K.xin.resize(K.v.size()); K.xout.resize(K.v.size()); K.cf.assign(2, 0.0);
    }
    for (int r = 0; r < R; ++r) {                                       // what ab_domain_p2p_connect wires
        Rank& K = ranks[r];
        for (size_t n = 0; n < K.neigh.size(); ++n) {
            Rank& Q = ranks[K.neigh[n]];
            size_t k = 0;
            while (Q.neigh[k] != r) ++k;
            if (Q.offset[k + 1] - Q.offset[k] != K.offset[n + 1] - K.offset[n]) { std::printf("FAIL: asymmetric interface\n"); return 1; }
            K.peer_dst.push_back((unsigned long long)(uintptr_t)(Q.window.data() + (size_t)Q.offset[k] * D));
            K.peer_stride.push_back((unsigned long long)((size_t)Q.total * D * sizeof(double)));
            K.peer_flag.push_back((unsigned long long)(uintptr_t)(Q.flags.data() + r));
        }
    }
    std::atomic<int> failures{0};
    auto rank_main = [&](int r) {
        Rank& K = ranks[r];
        std::mt19937 jitter(777 + r);
        for (int e = 1; e <= epochs; ++e) {
            const bool smooth = (e % 3) != 0;                          // two fused smoother exchanges, one plain sum, ...
            const double c1 = (e % 2) ? 0.37 : 0.0;                    // c1 = 0: the first smoother step (no d_in)
            K.cf[0] = c1;
            for (int l = 0; l < K.nv; ++l)
                for (int c = 0; c < D; ++c) {
                    const size_t i = (size_t)l * D + c;
                    const int g = K.glob[l];
                    K.din[i] = val(1, 0, e, g, c);                     // consistent inputs: the same on every rank
                    K.xin[i] = val(2, 0, e, g, c);
                    K.xout[i] = -555.0;
                    const double inc = val(3, r, e, g, c);             // this rank's additive part
                    K.v[i] = smooth ? (c1 != 0.0 ? c1 * K.din[i] + inc : inc) : inc;
                }
            const std::vector<double> v_in = K.v;
            if (jitter() % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(jitter() % 3000));
            if (K.niv > 0) { if (smooth) launch<true>(K, grid, block); else launch<false>(K, grid, block); }
            if (K.err) { std::printf("FAIL: rank %d epoch %d: timeout flag\n", r, e); failures++; return; }
            if (K.niv > 0 && K.state[0] != (unsigned long long)e) { std::printf("FAIL: rank %d: epoch counter %llu != %d\n", r, K.state[0], e); failures++; }
            std::vector<char> on_iface((size_t)K.nv, 0);
            for (int v : K.csr.iv) on_iface[v] = 1;
            for (int l = 0; l < K.nv; ++l)
                for (int c = 0; c < D; ++c) {
                    const size_t i = (size_t)l * D + c;
                    const int g = K.glob[l];
                    if (!on_iface[l]) {
                        if (std::memcmp(&K.v[i], &v_in[i], 8) != 0 || K.xout[i] != -555.0) { std::printf("FAIL: rank %d epoch %d: interior entry touched\n", r, e); failures++; }
                        continue;
                    }
                    double tot = 0.0;                                  // ascending rank order, every rank's own increment as that rank computed it
                    for (int q : owners[g]) {
                        const double inc_q = val(3, q, e, g, c);
                        double sent = inc_q;
                        if (smooth && c1 != 0.0) { const double loc = c1 * K.din[i] + inc_q; sent = loc - c1 * K.din[i]; }
                        tot += sent;
                    }
                    const double want_v = smooth ? ((c1 != 0.0 ? c1 * K.din[i] : 0.0) + tot) : tot;
                    if (std::memcmp(&K.v[i], &want_v, 8) != 0) {
                        std::printf("FAIL: rank %d epoch %d vertex %d comp %d: v = %.17g, expected %.17g\n", r, e, g, c, K.v[i], want_v);
                        failures++;
                    }
                    if (smooth) {
                        const double want_x = K.xin[i] + want_v;
                        if (std::memcmp(&K.xout[i], &want_x, 8) != 0) { std::printf("FAIL: rank %d epoch %d: x_out\n", r, e); failures++; }
                    } else if (K.xout[i] != -555.0) { std::printf("FAIL: rank %d epoch %d: plain sum wrote x_out\n", r, e); failures++; }
                }
        }
    };
    std::vector<std::thread> drivers;
    for (int r = 0; r < R; ++r) drivers.emplace_back(rank_main, r);
    for (auto& t : drivers) t.join();
    int shared3 = 0, entries = 0;
    for (int g = 0; g < nglob; ++g) shared3 += owners[g].size() >= 3;
    for (auto& K : ranks) entries = std::max(entries, K.niv * D);
    std::printf("%s: %d ranks, %d epochs, grid %d x %d threads for up to %d entries per rank, %d vertices on 3 ranks\n", failures ? "FAILED" : "OK", R, epochs, grid,
                block, entries, shared3);
    return failures ? 1 : 0;
}
