// libadmm_b200: objects behind the C ABI of include/admm_b200.h.
//   Domain/levels -> Space -> Vector / ElemDisc / DomainDisc / Operator / Solver (BiCGStab+GMG, CG+Jacobi)
// The call order served is the ADMM / Newton-Schur loop of 3d_admm.lua:875-1304 (2d_admm.lua:868-1253).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>

#include "../../include/admm_b200.h"
#include "comm.cuh"
#include "common.cuh"
#include "kernels_fe.cuh"
#include "kernels_la.cuh"
#include "mesh.hpp"

namespace ab {

static thread_local std::string g_last_error;
static uint64_t g_version_counter = 1;
static int g_live_contexts = 0, g_context_device = -1;   // one process drives one GPU (every entry point assumes the current device)
struct TraceTimer {   // ADMM_B200_TRACE=1: wall-clock of setup phases (synchronising; diagnostics only)
    cudaStream_t st; const char* what; bool on; double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    TraceTimer(cudaStream_t s, const char* w);
    ~TraceTimer() { if (on) { cudaStreamSynchronize(st); fprintf(stderr, "[ab trace] %-28s %9.3f ms\n", what, 1e3 * (now() - t0)); } }
};
static bool env_flag(const char* name) {
    const char* v = getenv(name);
    return v && *v && strcmp(v, "0") != 0;
}

TraceTimer::TraceTimer(cudaStream_t s, const char* w) : st(s), what(w), on(env_flag("ADMM_B200_TRACE")), t0(0) {
    if (on) { cudaStreamSynchronize(st); t0 = now(); }
}

// =============================================================================================
// Domain: host hierarchy + device-resident level structures
// =============================================================================================
struct LevelDev {
    int nv = 0, ne = 0, nvc = 0, maxrow = 0;
    int64_t nnzb = 0;
    DevBuf<int> rowptr, colidx, diagpos, mid;   // P1 vertex graph (BSR pattern) + midpoint ids on the next level
    DevBuf<int> tile_info;                       // TMA SpMV tiles: (first row, first block) per tile, ntiles+1 entries
    int ntiles = 0;
    DevBuf<int> pa, pb;                          // parents of vertices nvc..nv-1
    DevBuf<int> vsub;
    DevBuf<double> xyz;                          // top level only
    DevBuf<int> elems;                           // top level only
    DevBuf<int> v2e_ptr, v2e_idx;                // top level only: vertex -> incident elements (ascending), row-owner assembly
    DevBuf<int> elem_pos;                        // top level, built on demand: block positions per element (atomic-scatter assembly variant)
};

struct MatrixData;

struct Domain {
    Context* ctx = nullptr;
    HostMesh mesh;
    std::vector<LevelDev> dev;
    bool finalized = false;
    uint64_t coords_version = 1;
    bool host_xyz_stale = false;
    std::vector<std::weak_ptr<MatrixData>> live;   // assembled matrices, for signature sharing
    std::vector<std::shared_ptr<struct Gmg>> gmg_pool;   // hierarchies are recycled (buffers, BiCGStab workspace, iteration graph)
    // every ADMM iteration resets the multipliers (3d_admm.lua:884-894), so the first Newton iteration of each asks for the
    // SAME u-independent operator: the latest such matrix (and the hierarchies built on it) is kept alive for the signature cache
    std::shared_ptr<MatrixData> keep_lam0;
    // multi-GPU: this domain is one part of a decomposed global grid once interfaces were set (ab_domain_set_interface); a domain
    // without interfaces on a multi-rank context is a plain local (replicated) domain.  Shared-vertex interfaces per level
    // (host staging until finalize):
    bool dist_enabled = false;
    int xchg_ctas_per_sm = 0;          // grid cap of the interface-exchange kernel (launch_xchg), set at finalize
    struct HostIface { std::vector<int> neigh, offset, idx; std::vector<unsigned char> owned; };
    std::vector<HostIface> host_iface;
    std::vector<Interface> iface;
    // Matrix blocks shared with neighbour ranks (both vertices on the interface, block present on both sides), per decomposed
    // level: the exact Gershgorin bound of the additive operators needs their sum BEFORE the absolute value (gmg_setup_kernels).
    // Setup-only traffic: packed ncclSend/ncclRecv.
    struct BlockIface {
        std::vector<int> neigh, offset, h_idx, h_bpos, h_brow, h_mult;     // host staging until finalize
        int total = 0, nsb = 0;
        DevBuf<int> idx, bpos, brow, mult;     // slot -> compact block id; compact id -> local block position / row vertex / ranks holding it
        DevBuf<double> cv, send, recv;         // compact values (nsb x d*d), packed buffers (total x d*d)
    };
    std::vector<BlockIface> biface;
    // Hierarchical agglomeration (the reference keeps level 0 on one process and widens the process set level by level,
    // 3d_admm.lua:151-183): levels <= gather_level live on rank 0 as ONE global, non-distributed hierarchy (`cdom`, the global
    // grid refined gather_level times); the level-`gather_level` Galerkin operators of all ranks are summed into it at every
    // solver:init, each V-cycle crosses the vertical interface once down (additive right-hand sides) and once up (solution).
    int gather_level = -1;
    Domain* cdom = nullptr;                        // rank 0 only; owned by the caller (ab_domain handle)
    std::vector<int> g_nv;                         // rank 0: vertices of every rank on the gather level
    std::vector<int64_t> g_nblk;                   // rank 0: blocks of every rank on the gather level
    DevBuf<int> g_l2g;                             // rank 0: concatenated local -> global vertex ids
    DevBuf<int> g_csr_ptr, g_csr_idx;              // rank 0: global vertex -> positions in the staging buffer, rank order
    DevBuf<int> g_gpos;                            // rank 0: concatenated local block -> global block position
    DevBuf<double> g_vstage, g_mstage;             // rank 0: staging for the ranks' vectors / operators
    int64_t g_nv_total = 0, g_nblk_total = 0;
    // P2P window (CUDA IPC): [flags: nlevels*nranks u64][level 0: 2 parity buffers][level 1: ...]
    unsigned char* window = nullptr;
    size_t window_bytes = 0;
    std::vector<size_t> win_level_base;
    std::vector<void*> peer_windows;        // opened IPC mappings, by rank
    bool p2p_connected = false;
    DevBuf<int> p2p_err;
    bool distributed() const { return dist_enabled && ctx && ctx->comm && ctx->comm->nranks > 1; }
    int dim() const { return mesh.dim; }
    int top() const { return (int)mesh.levels.size() - 1; }
    void finalize();
    Domain() = default;
    Domain(const Domain&) = delete;
    ~Domain() {     // the peer-to-peer window and the opened CUDA IPC mappings are raw driver objects, not pool blocks
        if (window || !peer_windows.empty()) cudaDeviceSynchronize();   // not ctx->stream: the context may already be gone
        for (void* p : peer_windows) if (p) cudaIpcCloseMemHandle(p);
        if (window) cudaFree(window);
        cudaGetLastError();
    }
};

template <int D>
static void launch_elem_pos(Context* ctx, LevelDev& L) {
    const int64_t n = (int64_t)L.ne * (D + 1) * (D + 1);
    AB_LAUNCH(ctx, (k_elem_pos<D>), grid_for(n, 256, ctx->num_sms * 16), 256, 0, (int64_t)L.ne, L.elems.p, L.rowptr.p, L.colidx.p, L.elem_pos.p);
}

void Domain::finalize() {
    if (finalized) return;
    const int nl = (int)mesh.levels.size();
    dev.resize(nl);
    for (int l = 0; l < nl; ++l) {
        HostLevel& H = mesh.levels[l];
        HostPattern P;
        build_pattern(H, P);
        LevelDev& L = dev[l];
        L.nv = H.nv; L.ne = H.ne; L.nvc = H.nv_coarse; L.nnzb = (int64_t)P.colidx.size();
        int mr = 0;
        for (int i = 0; i < H.nv; ++i) mr = std::max(mr, P.rowptr[i + 1] - P.rowptr[i]);
        L.maxrow = mr;
        {   // TMA SpMV tiles: greedy runs of consecutive rows with <= budget blocks and <= RMAX rows.  The kernel runs one
            // persistent CTA pair per SM (G CTAs), CTA c takes tiles c, c+G, ...: the block budget is lowered from the
            // shared-memory capacity TB to the smallest value that still needs no more rounds, so that no CTA is left with a
            // straggler tile (at 44 730 DoFs: 900 tiles on 296 CTAs = 4 rounds for 12 CTAs -> 888 tiles = 3 rounds for all).
            const int TB = H.dim == 3 ? SpmvTma<3>::TB : SpmvTma<2>::TB, RMAX = H.dim == 3 ? SpmvTma<3>::RMAX : SpmvTma<2>::RMAX;
            std::vector<int> ti;
            bool ok = true;
            auto tile = [&](int budget, bool keep) {
                if (keep) { ti.clear(); ti.push_back(0); ti.push_back(0); }
                int start = 0, count = 0;
                for (int i = 0; i < H.nv; ++i) {
                    if (P.rowptr[i + 1] - P.rowptr[i] > budget) { if (budget == TB) ok = false; return INT32_MAX; }
                    if (P.rowptr[i + 1] - P.rowptr[start] > budget || i + 1 - start > RMAX) {
                        if (keep) { ti.push_back(i); ti.push_back(P.rowptr[i]); }
                        start = i; ++count;
                    }
                }
                if (keep) { ti.push_back(H.nv); ti.push_back(P.rowptr[H.nv]); }
                return count + 1;
            };
            int budget = TB;
            const int G = 2 * ctx->num_sms, t_full = tile(TB, false);
            if (ok && t_full > G) {
                const int64_t target = (int64_t)G * ((t_full + G - 1) / G);
                int lo = std::max(mr, (int)(L.nnzb / target)), hi = TB;       // smallest budget with tile count <= target
                while (lo < hi) {
                    const int m = (lo + hi) / 2;
                    if ((int64_t)tile(m, false) <= target) hi = m; else lo = m + 1;
                }
                budget = hi;
            }
            if (ok) tile(budget, true); else { ti.assign({0, 0, H.nv, P.rowptr[H.nv]}); }
            L.ntiles = ok ? (int)ti.size() / 2 - 1 : 0;      // 0: a row exceeds the tile (fall back to the warp kernel)
            L.tile_info.upload(ti, ctx->stream);
        }
        L.rowptr.upload(P.rowptr, ctx->stream);
        L.colidx.upload(P.colidx, ctx->stream);
        L.diagpos.upload(P.diagpos, ctx->stream);
        if (l < nl - 1) L.mid.upload(P.mid, ctx->stream);
        if (l > 0) { L.pa.upload(H.pa, ctx->stream); L.pb.upload(H.pb, ctx->stream); }
        L.vsub.upload(H.vsub, ctx->stream);
        if (l == nl - 1) {
            L.xyz.upload(H.xyz, ctx->stream);
            L.elems.upload(H.elems, ctx->stream);
            {   // vertex -> element incidence, elements ascending (fixed summation order of the row-owner assembly)
                const int N = H.dim + 1;
                AB_REQUIRE((int64_t)H.ne * N < (int64_t)1 << 31, AB_ERR_UNSUPPORTED, "level exceeds int32 incidence entries");
                std::vector<int32_t> vp, vi;
                build_v2e(H, vp, vi);
                L.v2e_ptr.upload(vp, ctx->stream);
                L.v2e_idx.upload(vi, ctx->stream);
            }
        }
        if (l < nl - 1) { H.edges.clear(); H.edges.shrink_to_fit(); H.have_edges = false; }
    }
    if (distributed()) {
        AB_REQUIRE((int)host_iface.size() == nl, AB_ERR_STATE, "multi-GPU: ab_domain_set_interface must be called for every level before the first ApproximationSpace");
        AB_REQUIRE(gather_level >= 0 && gather_level < nl - 1, AB_ERR_STATE, "multi-GPU: ab_domain_set_gather missing (the gather level must lie below the top level)");
        iface.resize(nl);
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iface_xchg<true>, 256, 0) != cudaSuccess) { occ = 0; (void)cudaGetLastError(); }
            xchg_ctas_per_sm = occ * 3 / 4;
        }
        const int me = ctx->comm->rank;
        for (int l = 0; l < nl; ++l) {
            HostIface& H = host_iface[l];
            Interface& I = iface[l];
            const int nv = mesh.levels[l].nv;
            AB_REQUIRE((int)H.owned.size() == nv, AB_ERR_ARG, "interface: owned mask size mismatch");
            I.neigh = H.neigh; I.offset = H.offset;
            I.total = H.offset.empty() ? 0 : H.offset.back();
            for (size_t n = 0; n < H.neigh.size(); ++n)
                AB_REQUIRE(H.neigh[n] != me && (n == 0 || H.neigh[n] > H.neigh[n - 1]), AB_ERR_ARG, "interface: neighbour ranks must be ascending and exclude this rank");
            const IfaceCsr csr = build_iface_csr(nv, me, H.neigh, H.offset, H.idx);      // iface_xchg.cuh
            I.my_pos = csr.my_pos;
            const std::vector<int>&iv = csr.iv, &ptr = csr.ptr, &slot = csr.slot, &nb = csr.nb;
            I.niv = (int)iv.size();
            I.idx.upload(H.idx, ctx->stream);
            I.iv.upload(iv, ctx->stream);
            I.iv_ptr.upload(ptr, ctx->stream);
            I.iv_slot.upload(slot, ctx->stream);
            I.iv_nb.upload(nb, ctx->stream);
            I.owned.upload(H.owned, ctx->stream);
            const int maxc = mesh.dim;     // P1 vectors with dim components
            I.send.alloc((size_t)std::max(I.total, 1) * maxc);
            I.recv.alloc((size_t)std::max(I.total, 1) * maxc);
            I.save.alloc((size_t)std::max(I.niv, 1) * 2 * maxc);
            I.state.alloc(4);
            I.state.zero(ctx->stream);
        }
        const int DD = mesh.dim * mesh.dim;
        for (BlockIface& B : biface) {
            if (B.nsb == 0) continue;
            B.idx.upload(B.h_idx, ctx->stream);
            B.bpos.upload(B.h_bpos, ctx->stream);
            B.brow.upload(B.h_brow, ctx->stream);
            B.mult.upload(B.h_mult, ctx->stream);
            B.cv.alloc((size_t)B.nsb * DD);
            B.send.alloc((size_t)std::max(B.total, 1) * DD);
            B.recv.alloc((size_t)std::max(B.total, 1) * DD);
        }
    }
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    finalized = true;
}

// interface sum: additive -> consistent for a P1 vector with D components on `level`
static void launch_xchg(Domain* dom, Interface& I, int D, bool smooth, double* v, const double* cf, const double* din, const double* xin, double* xout) {
    Context* ctx = dom->ctx;
    // capped grid: all CTAs are resident while they wait for the neighbours (see k_iface_xchg); larger interfaces are strided.
    // Cap = three quarters of what the device can hold of this kernel (occupancy query at finalize; 8 CTAs per SM -> 6), never below
    // kXchgCtasPerSm: interfaces up to ~ 227 k entries (numRefs 5 on 8 GPUs: 868 CTAs) still get one thread per entry.
    const int per_sm = std::max(kXchgCtasPerSm, dom->xchg_ctas_per_sm);      // Domain::finalize
    const int g = std::max(1, std::min((I.niv * D + 255) / 256, per_sm * ctx->num_sms));
    if (smooth)
        AB_LAUNCH(ctx, (k_iface_xchg<true>), g, 256, 0, I.niv, D, (int)I.neigh.size(), I.my_pos, I.iv.p, I.iv_ptr.p, I.iv_slot.p, I.iv_nb.p, I.d_offset.p, I.d_neigh.p,
                  I.d_peer_dst.p, I.d_peer_stride.p, I.d_peer_flag.p, I.total, I.win_recv, I.win_flags, I.state.p, dom->p2p_err.p, v, cf, din, xin, xout);
    else
        AB_LAUNCH(ctx, (k_iface_xchg<false>), g, 256, 0, I.niv, D, (int)I.neigh.size(), I.my_pos, I.iv.p, I.iv_ptr.p, I.iv_slot.p, I.iv_nb.p, I.d_offset.p, I.d_neigh.p,
                  I.d_peer_dst.p, I.d_peer_stride.p, I.d_peer_flag.p, I.total, I.win_recv, I.win_flags, I.state.p, dom->p2p_err.p, v, cf, din, xin, xout);
}
static void exchange_sum(Domain* dom, int level, double* v, int D) {
    if (!dom->distributed()) return;
    Context* ctx = dom->ctx;
    Interface& I = dom->iface[level];
    if (I.total == 0) return;
    if (dom->p2p_connected) {   // one fused kernel over NVLink peer memory
        launch_xchg(dom, I, D, false, v, nullptr, nullptr, nullptr, nullptr);
        return;
    }
    NcclApi& nc = NcclApi::get();
    AB_LAUNCH(ctx, k_iface_pack, grid_for((int64_t)I.total * D, 256, ctx->num_sms * 4), 256, 0, I.total, D, I.idx.p, v, I.send.p);
    AB_NCCL(nc.GroupStart());
    for (size_t n = 0; n < I.neigh.size(); ++n) {
        const size_t cnt = (size_t)(I.offset[n + 1] - I.offset[n]) * D;
        AB_NCCL(nc.Send(I.send.p + (size_t)I.offset[n] * D, cnt, ncclFloat64, I.neigh[n], ctx->comm->comm, ctx->stream));
        AB_NCCL(nc.Recv(I.recv.p + (size_t)I.offset[n] * D, cnt, ncclFloat64, I.neigh[n], ctx->comm->comm, ctx->stream));
    }
    AB_NCCL(nc.GroupEnd());
    AB_LAUNCH(ctx, k_iface_unpack_add, grid_for((int64_t)I.total * D, 256, ctx->num_sms * 4), 256, 0, I.total, D, I.idx.p, I.recv.p, v);
}
// smoother interface fix-up fused around the peer-to-peer sum (needs the P2P window): see k_iface_xchg<true>.  cf -> (c1, c2) of the
// step in device memory (the launch stays valid when a new solver:init changes the coefficients: graph replay)
static void smooth_exchange_p2p(Domain* dom, int level, int D, const double* cf, const double* din, double* dout, const double* xin, double* xout) {
    Interface& I = dom->iface[level];
    if (I.total == 0) return;
    launch_xchg(dom, I, D, true, dout, cf, din, xin, xout);
}
// global sum / max of device scalars over the ranks that hold the parts of a DECOMPOSED domain (no-op on one GPU and for an
// undivided domain on a multi-rank context: there every rank already holds the whole value)
static void allreduce_dev(Domain* dom, double* d, int n, bool max_op = false) {
    if (dom->distributed()) dom->ctx->comm->allreduce(d, n, max_op, dom->ctx->stream);
}

struct Space {
    Domain* dom;
    int kind, ncomp;
    int64_t ndofs;
};

struct Vector {
    Space* sp;
    DevBuf<double> d;
    int storage = AB_PST_CONSISTENT;
    uint64_t version = 0, id = 0;
    bool aliased = false;            // the device pointer was handed out (ab_vector_device_ptr): contents may change unseen -> never cached
    // L2Norm(gf, cmp) is called once per component by the scripts (3d_admm.lua:1137-1146, 1237-1251); one element pass yields
    // all of them, so the components are remembered until the vector or the coordinates change
    uint64_t l2_version = 0, l2_coords = 0;
    int l2_storage = 0;
    double l2_vals[9];
    int64_t n() const { return sp->ndofs; }
    void touch() { version = ++g_version_counter; }
};

// VecProd results remembered per (x, y, versions).  The Newton/Schur loop asks for <B_r, y> with the SAME y for r = 1..m in
// consecutive calls (3d_admm.lua:993-995, 1014-1059): on a miss the products of y with the last few distinct x operands are
// computed in the same pass over y (one launch, one read-back instead of m), bitwise equal to the separate products.
struct DotCache {
    struct Entry { uint64_t xid, xver, yid, yver; int xst, yst; double val; };
    std::vector<Entry> entries;      // small ring
    size_t next = 0;
    std::vector<Vector*> recent_x;   // most recent first, at most 4
    bool lookup(const Vector* x, const Vector* y, double* out) const {
        for (const Entry& e : entries)
            if (e.xid == x->id && e.xver == x->version && e.yid == y->id && e.yver == y->version && e.xst == x->storage && e.yst == y->storage) { *out = e.val; return true; }
        return false;
    }
    void store(const Vector* x, const Vector* y, double val) {
        Entry e{x->id, x->version, y->id, y->version, x->storage, y->storage, val};
        if (entries.size() < 32) entries.push_back(e);
        else { entries[next] = e; next = (next + 1) % entries.size(); }
    }
    void used(Vector* x) {
        recent_x.erase(std::remove(recent_x.begin(), recent_x.end(), x), recent_x.end());
        recent_x.insert(recent_x.begin(), x);
        if (recent_x.size() > 4) recent_x.pop_back();
    }
    void forget(Vector* v) { recent_x.erase(std::remove(recent_x.begin(), recent_x.end(), v), recent_x.end()); }
};
static std::map<Context*, DotCache> g_dot_cache;

struct ElemDisc {
    Space* sp;
    int kind;
    double params[16];
    Vector *imp_u = nullptr, *imp_lam = nullptr, *imp_q = nullptr;
};

struct DomainDisc {
    Space* sp;
    std::vector<ElemDisc*> discs;
    std::vector<std::pair<int, int>> dir;          // (subset index, comp)
    std::vector<DevBuf<unsigned char>> masks;      // per level, built lazily
    bool masks_valid = false;
    const unsigned char* mask(int level);
};

const unsigned char* DomainDisc::mask(int level) {
    if (dir.empty()) return nullptr;
    Domain* dom = sp->dom;
    if (!masks_valid) {
        masks.clear();
        masks.resize(dom->mesh.levels.size());
        for (size_t l = 0; l < dom->mesh.levels.size(); ++l) {
            const HostLevel& H = dom->mesh.levels[l];
            std::vector<unsigned char> m((size_t)H.nv, 0);
            for (auto& e : dir)
                for (int v = 0; v < H.nv; ++v)
                    if (H.vsub[v] == e.first) m[v] |= (unsigned char)(1u << e.second);
            masks[l].upload(m, dom->ctx->stream);
        }
        masks_valid = true;
    }
    return masks[level].p;
}

// =============================================================================================
// matrices
// =============================================================================================
struct Signature {
    int kind = 0;                                  // 1 = P1 Hessian BSR, 2 = P0 diagonal mass
    double c = 0, lam_vol = 0, lam_b[3] = {0, 0, 0};
    uint64_t u_id = 0, u_version = 0, coords_version = 0;
    std::vector<std::pair<int, int>> dir;
    bool operator==(const Signature& o) const {
        return kind == o.kind && c == o.c && lam_vol == o.lam_vol && lam_b[0] == o.lam_b[0] && lam_b[1] == o.lam_b[1] &&
               lam_b[2] == o.lam_b[2] && u_id == o.u_id && u_version == o.u_version && coords_version == o.coords_version && dir == o.dir;
    }
};

struct Gmg;
struct MatrixData {
    Domain* dom = nullptr;
    int kind = 0;                                  // 1 BSR (top level pattern), 2 diagonal
    DevBuf<double> vals;
    Signature sig;
    bool assembled = false;
    std::map<std::string, std::shared_ptr<Gmg>> gmg;   // hierarchies built on this matrix, keyed by descriptor
};

struct Operator {
    DomainDisc* dd;
    std::shared_ptr<MatrixData> data;
};

// =============================================================================================
// small launch helpers
// =============================================================================================
static inline int ew_grid(Context* ctx, int64_t n) { return grid_for(n, 256, ctx->num_sms * 8); }
static inline int red_grid(Context* ctx, int64_t n) { return grid_for(n, 256, Context::kMaxBlocks); }

static void dev_fill(Context* ctx, double* x, int64_t n, double c) {
    if (c == 0.0) { AB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream)); return; }
    AB_LAUNCH(ctx, k_fill, ew_grid(ctx, n), 256, 0, n, c, x);
}
static void dev_axpby(Context* ctx, int64_t n, double a, const double* x, double b, const double* y, double* out) {
    AB_LAUNCH(ctx, k_axpby, ew_grid(ctx, n), 256, 0, n, a, x, b, y, out);
}
static void dev_copy(Context* ctx, int64_t n, const double* src, double* dst) {
    AB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
}
// read `cnt` device doubles (synchronises the stream)
static void read_back(Context* ctx, const double* dptr, int cnt, double* out) {
    AB_CUDA(cudaMemcpyAsync(ctx->h_results, dptr, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < cnt; ++i) out[i] = ctx->h_results[i];
}
static void dev_dots(Context* ctx, int64_t n, int nx, const double* const* xs, const double* y, double* dout) {
    const double* x[4] = {xs[0], nx > 1 ? xs[1] : xs[0], nx > 2 ? xs[2] : xs[0], nx > 3 ? xs[3] : xs[0]};
    const int g = red_grid(ctx, n);
    switch (nx) {
        case 1: AB_LAUNCH(ctx, (k_dot_multi<1>), g, 256, 0, n, x[0], x[1], x[2], x[3], y, ctx->d_partials, ctx->d_tickets, dout); break;
        case 2: AB_LAUNCH(ctx, (k_dot_multi<2>), g, 256, 0, n, x[0], x[1], x[2], x[3], y, ctx->d_partials, ctx->d_tickets, dout); break;
        case 3: AB_LAUNCH(ctx, (k_dot_multi<3>), g, 256, 0, n, x[0], x[1], x[2], x[3], y, ctx->d_partials, ctx->d_tickets, dout); break;
        default: AB_LAUNCH(ctx, (k_dot_multi<4>), g, 256, 0, n, x[0], x[1], x[2], x[3], y, ctx->d_partials, ctx->d_tickets, dout); break;
    }
}

// SpMV family dispatch. mode 0: y=Ax (dots: 0/1/2 with w), 1: y=b-Ax, 2: smoother step
template <int D, int U, int HL>
static void spmv_tma_launch(Context* ctx, const LevelDev& L, const double* vals, int mode, int dots, const double* x, const double* b, double* y,
                            const double* dinv, const double* dvec, double* dout, double c1, double c2, const double* w, double* red, const double* cf, int prefetch) {
    using T = SpmvTma<D>;
    // persistent CTAs: two per SM; levels that live in L2 may run one per SM (tuning key "tma_small_ctas") so that the next
    // kernel of the PDL chain finds room to become resident and prefetch while this one runs
    const bool small = L.nnzb * (int64_t)(D * D * 8 + 4) <= ((int64_t)48 << 20);
    const int g = std::max(1, std::min(L.ntiles, (small ? ctx->tma_small_ctas : 2) * ctx->num_sms));
    // bit 1: the level does not fit L2 anyway -> the matrix stream is marked evict-first so that it does not flush the gathered vector
    if (ctx->l2_hint && L.nnzb * (int64_t)(D * D * 8 + 4) > ((int64_t)48 << 20)) prefetch |= 2;
#define AB_SPMV(MODE, DOTS)                                                                                                          \
    do {                                                                                                                              \
        AB_LAUNCH_PDL(ctx, (k_bsr_spmv_tma<D, MODE, DOTS, U, HL>), g, T::NT, T::SMEM_B, L.ntiles, L.tile_info.p, L.rowptr.p, L.colidx.p, vals, x, b, y, dinv, dvec, dout, c1, c2, w, ctx->d_partials, ctx->d_tickets, red, cf, prefetch); \
    } while (0)
    if (mode == 0) {
        if (dots == 0) AB_SPMV(0, 0);
        else if (dots == 1) AB_SPMV(0, 1);
        else AB_SPMV(0, 2);
    } else if (mode == 1) AB_SPMV(1, 0);
    else AB_SPMV(2, 0);
#undef AB_SPMV
}
template <int D, int U>
static void spmv_warp_launch(Context* ctx, const LevelDev& L, const double* vals, int mode, int dots, const double* x, const double* b, double* y,
                             const double* dinv, const double* dvec, double* dout, double c1, double c2, const double* w, double* red, const double* cf, int prefetch) {
    const int64_t want = ((int64_t)L.nv + 7) / 8;        // 8 warps (rows) per CTA
    const int g = (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min(ctx->num_sms * ctx->spmv_waves, (int)Context::kMaxBlocks)));
#define AB_SPMV(MODE, DOTS) \
    AB_LAUNCH_PDL(ctx, (k_bsr_spmv_warp<D, MODE, DOTS, U>), g, 256, 0, L.nv, L.rowptr.p, L.colidx.p, vals, x, b, y, dinv, dvec, dout, c1, c2, w, ctx->d_partials, ctx->d_tickets, red, cf, prefetch)
    if (mode == 0) {
        if (dots == 0) AB_SPMV(0, 0);
        else if (dots == 1) AB_SPMV(0, 1);
        else AB_SPMV(0, 2);
    } else if (mode == 1) AB_SPMV(1, 0);
    else AB_SPMV(2, 0);
#undef AB_SPMV
}
// opt in to the large dynamic shared memory of every TMA SpMV instantiation once, up front (not lazily inside a launch
// that may be under stream capture)
template <int D, int U, int HL>
static void spmv_prepare_dim() {
    using T = SpmvTma<D>;
    AB_CUDA(cudaFuncSetAttribute(k_bsr_spmv_tma<D, 0, 0, U, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_B));
    AB_CUDA(cudaFuncSetAttribute(k_bsr_spmv_tma<D, 0, 1, U, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_B));
    AB_CUDA(cudaFuncSetAttribute(k_bsr_spmv_tma<D, 0, 2, U, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_B));
    AB_CUDA(cudaFuncSetAttribute(k_bsr_spmv_tma<D, 1, 0, U, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_B));
    AB_CUDA(cudaFuncSetAttribute(k_bsr_spmv_tma<D, 2, 0, U, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_B));
}
static void spmv_prepare_kernels() {
    spmv_prepare_dim<2, 4, 4>();
    spmv_prepare_dim<2, 2, 8>();
    spmv_prepare_dim<3, 3, 16>();
}
static void spmv(Context* ctx, int dim, const LevelDev& L, const double* vals, int mode, int dots, const double* x, const double* b, double* y,
                 const double* dinv = nullptr, double* dvec = nullptr, double c1 = 0, double c2 = 0, const double* w = nullptr, double* red = nullptr,
                 double* dout = nullptr, const double* cf = nullptr, int prefetch = 0) {
    // prefetch = 1 (solver / multigrid path only): the matrix and its pattern were final before the last stream
    // synchronisation, so the kernel may stream them before its programmatic-dependency wait (common.cuh, PDL)
    if (!dout) dout = dvec;
    // tuning knob "spmv_variant": 0 = TMA-staged tiles (default), 1 = warp per row through the LSU path (the fallback when
    // a row exceeds a tile)
    const int variant = (ctx->spmv_variant == 0 && L.ntiles == 0) ? 1 : ctx->spmv_variant;
    switch (variant) {
        case 0:
            // 2x2 blocks: 4 lanes per row (8 rows per warp in flight); tuning key "spmv2d_lanes" = 8 selects the quarter-warp mapping
            if (dim == 2 && ctx->spmv2d_lanes == 8) spmv_tma_launch<2, 2, 8>(ctx, L, vals, mode, dots, x, b, y, dinv, dvec, dout, c1, c2, w, red, cf, prefetch);
            else if (dim == 2) spmv_tma_launch<2, 4, 4>(ctx, L, vals, mode, dots, x, b, y, dinv, dvec, dout, c1, c2, w, red, cf, prefetch);
            else spmv_tma_launch<3, 3, 16>(ctx, L, vals, mode, dots, x, b, y, dinv, dvec, dout, c1, c2, w, red, cf, prefetch);
            return;
        default:
            if (dim == 2) spmv_warp_launch<2, 1>(ctx, L, vals, mode, dots, x, b, y, dinv, dvec, dout, c1, c2, w, red, cf, prefetch);
            else spmv_warp_launch<3, 3>(ctx, L, vals, mode, dots, x, b, y, dinv, dvec, dout, c1, c2, w, red, cf, prefetch);
            return;
    }
}

// =============================================================================================
// geometric multigrid hierarchy on one assembled matrix (obstacle_optim_3d_util.lua:13-31)
// =============================================================================================
struct GmgLevel {
    DevBuf<double> vals_own;          // Galerkin operator (levels below top)
    const double* vals = nullptr;
    DevBuf<double> dinv, x, b, r, d, x2, d2;   // d2: second increment buffer (multi-GPU P2P smoother ping-pong)
    double lmax = 0;
    std::vector<std::pair<double, double>> coef_pre, coef_post;   // (c1,c2) per smoothing step
    const double *cf_pre = nullptr, *cf_post = nullptr;           // the same pairs in device memory (Gmg::coefs): launches stay graph-replayable
    const unsigned char* mask = nullptr;
};

// BiCGStab work vectors and scalars.  They live with the hierarchy, not with the solver object: the six solvers of one
// Newton iteration share one hierarchy (operator sharing), so one BiCGStab iteration is captured
// ONCE into an executable CUDA graph and replayed by all of them (x is copied in and out of the workspace).
struct KrylovWs {
    DevBuf<double> r, rh, p, v, s, t, ph, sh, x, sc, out2;
    DevBuf<double> uq, bq;                 // multi-GPU: unique (owner-only) form of the preconditioner input / of a consistent right-hand side
    DevBuf<double> ctl, hist;              // loop-control block and |r|^2 history of the device-side loop
    cudaGraphExec_t exec_loop = nullptr;   // the WHOLE iteration loop: conditional WHILE node around one iteration + D2H of the result
    std::vector<const void*> key_loop;
    int64_t nodes_loop = 0;
    bool loop_failed = false;              // the driver refused the conditional graph once: stay on per-iteration replay
    bool graph_failed = false;             // multi-GPU: the capture (kernels + NCCL calls) was refused once: stay on stream launches
    cudaGraphExec_t exec = nullptr;        // one full BiCGStab iteration (2 V-cycles, 2 SpMV, recurrences, D2H of the scalars)
    std::vector<const void*> key;          // every pointer baked into the captured launches
    int64_t nodes = 0;                     // kernel launches per replay
    KrylovWs() = default;
    KrylovWs(const KrylovWs&) = delete;
    ~KrylovWs() { if (exec) cudaGraphExecDestroy(exec); if (exec_loop) cudaGraphExecDestroy(exec_loop); }
};

struct Gmg {
    Domain* dom = nullptr;
    ab_gmg_desc desc;
    std::vector<GmgLevel> L;
    // lowest level of THIS hierarchy: 0 on one GPU (dense inverse); distributed: the gather level, which is solved on rank 0 by
    // `coarse`, a plain single-GPU hierarchy over the global coarse grid (Domain::cdom)
    int base = 0;
    std::vector<std::pair<int, int>> dir;               // Dirichlet set (subset, component), sorted
    std::vector<DevBuf<unsigned char>> masks;           // per level
    // coarse level: dense inverse on the free dofs
    int n_free = 0, n0 = 0;
    DevBuf<double> Ainv, Mwork, gjpanel;
    DevBuf<int> free2dof, dof2free, pivrow, fail;
    bool force_pivoting = false;      // set when the unpivoted blocked elimination met a tiny pivot
    std::shared_ptr<Gmg> coarse;      // multi-GPU, rank 0
    DevBuf<double> gvals, bg, xg;     // multi-GPU, rank 0: summed operator of the gather level, global coarse right-hand side / solution
    DevBuf<double> coefs;             // smoother coefficients, [level][pre|post][step][c1,c2]
    int coef_stride = 0;              // doubles per (level, pre|post) slot
    std::unique_ptr<KrylovWs> ws;     // BiCGStab workspace + iteration graph
    std::string pool_key;             // descriptor + Dirichlet set (Domain::gmg_pool)
    void setup(const double* top_vals, const std::vector<std::pair<int, int>>& dir_sorted);
    void vcycle(int l, const double* b, double* x, bool first_done = false);
    void vcycle_base_gathered(const double* b, double* x);
    void smooth(int l, const double* b, double* x, int nu, bool zero_guess, const std::vector<std::pair<double, double>>& coef, const double* cf,
                bool first_done = false);
    // first_done: the caller already performed the first pre-smoothing step of the top level (fused into its own kernel)
    void apply(const double* r, double* z, bool first_done = false) { vcycle((int)L.size() - 1, r, z, first_done); }
    // where the first pre-smoothing step (zero initial guess) of a cycle writing its result to z puts d and x on the top level;
    // false when that step cannot be fused by the caller (single level, no pre-smoothing, NCCL-fallback interface fix-ups)
    bool first_step_targets(double* z, double** d, double** x) {
        const int top = (int)L.size() - 1;
        if (top <= base || desc.pre_smooth < 1 || (dom->distributed() && !dom->p2p_connected) || !L[top].cf_pre) return false;
        *d = L[top].d.p;
        *x = ((desc.pre_smooth - 1) % 2 == 0) ? z : L[top].x2.p;
        return true;
    }
};

static std::string gmg_key(const ab_gmg_desc& d) {
    char buf[160];
    snprintf(buf, sizeof buf, "%d/%d/%d/%d/%.17g/%.17g", d.smoother, d.pre_smooth, d.post_smooth, d.base_level, d.cheb_ratio, d.jacobi_damp);
    return buf;
}

static void smoother_coefs(const ab_gmg_desc& desc, double lmax, int nu, std::vector<std::pair<double, double>>& out) {
    out.clear();
    if (desc.smoother == AB_SMOOTHER_JACOBI) {
        for (int k = 0; k < nu; ++k) out.push_back({0.0, desc.jacobi_damp});
        return;
    }
    const double lmin = lmax / desc.cheb_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
    double rho = 1.0 / sigma1;
    out.push_back({0.0, 1.0 / theta});
    for (int k = 1; k < nu; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        out.push_back({rho_new * rho, 2.0 * rho_new / delta});
        rho = rho_new;
    }
}

// multi-GPU: the additive gather-level operators of all ranks are summed into the global operator on rank 0 (rank order:
// reproducible), which then builds its single-GPU hierarchy below
template <int D>
static void gmg_gather_operator(Gmg& G) {
    Domain* dom = G.dom;
    Context* ctx = dom->ctx;
    Comm& cm = *ctx->comm;
    NcclApi& nc = NcclApi::get();
    constexpr int DD = D * D;
    const LevelDev& Lb = dom->dev[G.base];
    const double* mine = G.L[G.base].vals;
    if (cm.rank != 0) {
        AB_NCCL(nc.Send(mine, (size_t)Lb.nnzb * DD, ncclFloat64, 0, cm.comm, ctx->stream));
        return;
    }
    std::vector<int64_t> off((size_t)cm.nranks + 1, 0);
    for (int r = 0; r < cm.nranks; ++r) off[r + 1] = off[r] + dom->g_nblk[r];
    AB_REQUIRE(dom->g_nblk[0] == Lb.nnzb, AB_ERR_STATE, "gather maps do not match the local pattern");
    AB_NCCL(nc.GroupStart());
    for (int r = 1; r < cm.nranks; ++r)
        AB_NCCL(nc.Recv(dom->g_mstage.p + (size_t)off[r] * DD, (size_t)dom->g_nblk[r] * DD, ncclFloat64, r, cm.comm, ctx->stream));
    AB_NCCL(nc.GroupEnd());
    AB_CUDA(cudaMemsetAsync(G.gvals.p, 0, G.gvals.n * sizeof(double), ctx->stream));
    for (int r = 0; r < cm.nranks; ++r) {
        const int64_t cnt = dom->g_nblk[r] * DD;
        const double* src = r == 0 ? mine : dom->g_mstage.p + (size_t)off[r] * DD;
        AB_LAUNCH(ctx, k_scatter_add_blocks, grid_for(cnt, 256, ctx->num_sms * 8), 256, 0, cnt, DD, dom->g_gpos.p + off[r], src, G.gvals.p);
    }
    G.coarse->setup(G.gvals.p, G.dir);
}

template <int D>
static void gmg_setup_kernels(Gmg& G) {
    Domain* dom = G.dom;
    Context* ctx = dom->ctx;
    const int top = dom->top(), base = G.base;
    const bool dist = dom->distributed();
    constexpr int DD = D * D;
    // Galerkin coarse operators, top-down
    for (int l = top; l > base; --l) {
        TraceTimer tt(ctx->stream, "gmg: rap level");
        const LevelDev& F = dom->dev[l];
        const LevelDev& C = dom->dev[l - 1];
        GmgLevel& gc = G.L[l - 1];
        if (gc.vals_own.n != (size_t)C.nnzb * DD) gc.vals_own.alloc((size_t)C.nnzb * DD);   // pointers stay put across setups (graph replay)
        gc.vals = gc.vals_own.p;
        const size_t smem = (size_t)C.maxrow * (DD * sizeof(double) + sizeof(int));      // one coarse row per CTA
        static bool attr_set[2] = {false, false};
        if (smem > 48 * 1024 && !attr_set[D - 2]) {
            AB_CUDA(cudaFuncSetAttribute(k_rap<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[D - 2] = true;
        }
        AB_REQUIRE(smem <= 200 * 1024, AB_ERR_UNSUPPORTED, "coarse row too long for k_rap shared memory");
        const int grid = (int)std::min<int64_t>((int64_t)C.nv, (int64_t)ctx->num_sms * 8);
        // distributed: the unit diagonal of an eliminated dof is set by the owner of the vertex only (additive operators)
        AB_LAUNCH(ctx, (k_rap<D>), std::max(grid, 1), 256, smem, C.nv, C.maxrow, C.rowptr.p, C.colidx.p, C.mid.p, C.diagpos.p, F.rowptr.p,
                  F.colidx.p, G.L[l].vals, F.pa.p, F.pb.p, gc.mask, dist ? dom->iface[l - 1].owned.p : (const unsigned char*)nullptr, gc.vals_own.p);
    }
    // smoother data on the levels above the base
    for (int l = base + 1; l <= top; ++l) {
        const LevelDev& Ld = dom->dev[l];
        GmgLevel& g = G.L[l];
        const int64_t n = (int64_t)Ld.nv * D;
        if (!dist) {
            AB_LAUNCH(ctx, (k_diag_gershgorin<D>), red_grid(ctx, n), 256, 0, Ld.nv, Ld.rowptr.p, Ld.diagpos.p, g.vals, g.dinv.p,
                      ctx->d_partials, ctx->d_tickets, ctx->d_results + l);
        } else {
            // additive rows: diagonal and |row| sums are made consistent across the interfaces.  Summing the per-rank |a_ij| over-
            // estimates sum_j |a_ij| wherever a block is shared (|x| + |y| >= |x + y|): the bound -- and with it the Chebyshev
            // interval and the iteration counts -- would depend on the partition.  With the shared-block lists (Domain::biface) the
            // shared blocks are summed first and every rank adds |sum| / (ranks holding the block) instead of its own |part|: the
            // interface sum of the rows is then EXACTLY the row sum of the global operator.  The loose sum is kept beside it (g.d is
            // free during setup) and only serves as a plausibility bracket for the exact one (Gmg::setup).
            AB_LAUNCH(ctx, (k_diag_rowabs<D>), ew_grid(ctx, n), 256, 0, Ld.nv, Ld.rowptr.p, Ld.diagpos.p, g.vals, g.dinv.p, g.r.p);
            dev_copy(ctx, n, g.r.p, g.d.p);
            if (l < (int)dom->biface.size() && dom->biface[l].nsb > 0) {
                Domain::BlockIface& B = dom->biface[l];
                NcclApi& nc = NcclApi::get();
                AB_LAUNCH(ctx, k_pack_blocks, ew_grid(ctx, (int64_t)B.nsb * DD), 256, 0, (int64_t)B.nsb * DD, DD, B.bpos.p, g.vals, B.cv.p);
                if (B.total > 0) {
                    AB_LAUNCH(ctx, k_iface_pack, grid_for((int64_t)B.total * DD, 256, ctx->num_sms * 4), 256, 0, B.total, DD, B.idx.p, B.cv.p, B.send.p);
                    AB_NCCL(nc.GroupStart());
                    for (size_t q = 0; q < B.neigh.size(); ++q) {
                        const size_t cnt = (size_t)(B.offset[q + 1] - B.offset[q]) * DD;
                        if (cnt == 0) continue;                      // symmetric: the common block list is empty on both sides
                        AB_NCCL(nc.Send(B.send.p + (size_t)B.offset[q] * DD, cnt, ncclFloat64, B.neigh[q], ctx->comm->comm, ctx->stream));
                        AB_NCCL(nc.Recv(B.recv.p + (size_t)B.offset[q] * DD, cnt, ncclFloat64, B.neigh[q], ctx->comm->comm, ctx->stream));
                    }
                    AB_NCCL(nc.GroupEnd());
                    AB_LAUNCH(ctx, k_iface_unpack_add, grid_for((int64_t)B.total * DD, 256, ctx->num_sms * 4), 256, 0, B.total, DD, B.idx.p, B.recv.p, B.cv.p);
                }
                AB_LAUNCH(ctx, (k_rowabs_fix<D>), ew_grid(ctx, (int64_t)B.nsb * D), 256, 0, B.nsb, B.bpos.p, B.brow.p, B.mult.p, g.vals, B.cv.p, g.r.p);
            }
            exchange_sum(dom, l, g.dinv.p, D);
            exchange_sum(dom, l, g.r.p, D);
            exchange_sum(dom, l, g.d.p, D);
            AB_LAUNCH(ctx, k_dinv_lmax2, red_grid(ctx, n), 256, 0, n, g.dinv.p, g.r.p, g.d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results + 2 * l);
        }
    }
    if (dist) {
        if (top > base) allreduce_dev(dom, ctx->d_results + 2 * (base + 1), 2 * (top - base), true);   // (exact, loose) per level
        AB_CUDA(cudaMemsetAsync(G.fail.p, 0, sizeof(int), ctx->stream));
        TraceTimer tt(ctx->stream, "gmg: gather + coarse setup");
        gmg_gather_operator<D>(G);
        return;
    }
    // dense inverse of level 0 (free dofs)
    {
        TraceTimer tt(ctx->stream, "gmg: dense inverse");
        const LevelDev& L0 = dom->dev[0];
        const int n = G.n_free;
        AB_CUDA(cudaMemsetAsync(G.fail.p, 0, sizeof(int), ctx->stream));
        if (n > 0) {
            constexpr int KR = 16;                                 // pivots per barrier of the shared-memory-resident kernel
            const int grid_r = std::min(ctx->num_sms, n);
            const int rmax = (n + grid_r - 1) / grid_r;
            const size_t smem_r = ((size_t)rmax * n + (size_t)KR * n + KR * KR + (size_t)rmax * KR) * sizeof(double);
            const bool resident = !G.force_pivoting && ctx->coarse_variant == 0 && smem_r <= 200 * 1024;
            const int ld = resident ? n : 2 * n;                   // resident: A (n x n) in, A^-1 straight into Ainv; otherwise [A | I]
            AB_CUDA(cudaMemsetAsync(G.Mwork.p, 0, (size_t)n * ld * sizeof(double), ctx->stream));
            AB_LAUNCH(ctx, (k_bsr_to_dense<D>), grid_for(L0.nnzb * DD, 256, ctx->num_sms * 8), 256, 0, L0.nv, L0.rowptr.p, L0.colidx.p, G.L[0].vals,
                      G.dof2free.p, ld, G.Mwork.p);
            int nn = n;
            double* M = G.Mwork.p;
            int* piv = G.pivrow.p;
            int* fail = G.fail.p;
            if (resident) {
                static bool attr = false;
                if (!attr) { AB_CUDA(cudaFuncSetAttribute(k_gauss_jordan_resident<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
                if (G.gjpanel.n < (size_t)2 * KR * n) G.gjpanel.alloc((size_t)2 * KR * n);
                double* ainv = G.Ainv.p;
                double* pbuf = G.gjpanel.p;
                void* rargs[] = {&nn, &M, &ainv, &pbuf, &fail};
                AB_CUDA(cudaLaunchCooperativeKernel((void*)k_gauss_jordan_resident<KR>, dim3(grid_r), dim3(640), rargs, smem_r, ctx->stream));
                ctx->launches++;
            } else {
                AB_LAUNCH(ctx, k_dense_identity, grid_for(n, 256, 64), 256, 0, n, G.Mwork.p);
                void* args[] = {&nn, &M, &piv, &fail};
                int grid = std::min(ctx->num_sms, n);
                constexpr int KB = 8;
                const size_t panel_bytes = (size_t)KB * 2 * n * sizeof(double);
                if (!G.force_pivoting && panel_bytes <= 200 * 1024) {      // blocked, unpivoted, rows in global memory
                    static bool attr = false;
                    if (!attr) { AB_CUDA(cudaFuncSetAttribute(k_gauss_jordan_blocked<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
                    AB_CUDA(cudaLaunchCooperativeKernel((void*)k_gauss_jordan_blocked<KB>, dim3(grid), dim3(1024), args, panel_bytes, ctx->stream));
                } else {
                    AB_CUDA(cudaLaunchCooperativeKernel((void*)k_gauss_jordan, dim3(grid), dim3(256), args, (size_t)n, ctx->stream));
                }
                ctx->launches++;
                AB_LAUNCH(ctx, k_extract_inverse, grid_for((int64_t)n * n, 256, ctx->num_sms * 8), 256, 0, n, G.Mwork.p, G.pivrow.p, G.Ainv.p);
            }
        }
    }
}

void Gmg::setup(const double* top_vals, const std::vector<std::pair<int, int>>& dir_sorted) {
    Context* ctx = dom->ctx;
    TraceTimer tt(ctx->stream, "gmg: setup total");
    dom->finalize();
    const int top = dom->top(), dim = dom->dim();
    const bool dist = dom->distributed();
    const bool first = L.empty();
    if (first) {
        dir = dir_sorted;
        base = dist ? dom->gather_level : 0;
        L.resize(top + 1);
        masks.resize(top + 1);
        for (int l = base; l <= top; ++l) {
            const int64_t n = (int64_t)dom->dev[l].nv * dim;
            GmgLevel& g = L[l];
            if (l < top) { g.x.alloc(n); g.b.alloc(n); }
            if (l > base) { g.r.alloc(n); g.d.alloc(n); g.x2.alloc(n); g.dinv.alloc(n); if (dist) g.d2.alloc(n); }
            if (!dir.empty()) {
                const HostLevel& H = dom->mesh.levels[l];
                std::vector<unsigned char> m((size_t)H.nv, 0);
                for (auto& e : dir)
                    for (int v = 0; v < H.nv; ++v)
                        if (H.vsub[v] == e.first) m[v] |= (unsigned char)(1u << e.second);
                masks[l].upload(m, ctx->stream);
                g.mask = masks[l].p;
            }
        }
        fail.alloc(1);
        if (!dist) {       // free-dof numbering of level 0 (dense base solver)
            const HostLevel& H0 = dom->mesh.levels[0];
            n0 = H0.nv * dim;
            std::vector<unsigned char> m((size_t)H0.nv, 0);
            for (auto& e : dir)
                for (int v = 0; v < H0.nv; ++v)
                    if (H0.vsub[v] == e.first) m[v] |= (unsigned char)(1u << e.second);
            std::vector<int> f2d, d2f((size_t)n0, -1);
            for (int v = 0; v < H0.nv; ++v)
                for (int c = 0; c < dim; ++c)
                    if (!((m[v] >> c) & 1)) { d2f[(size_t)v * dim + c] = (int)f2d.size(); f2d.push_back(v * dim + c); }
            n_free = (int)f2d.size();
            free2dof.upload(f2d, ctx->stream);
            dof2free.upload(d2f, ctx->stream);
            if (n_free) { Ainv.alloc((size_t)n_free * n_free); Mwork.alloc((size_t)n_free * 2 * n_free); }
            pivrow.alloc(std::max(n_free, 1));
        } else if (ctx->comm->rank == 0) {   // the global coarse hierarchy below the gather level
            AB_REQUIRE(dom->cdom, AB_ERR_STATE, "multi-GPU: rank 0 has no coarse domain (ab_domain_set_gather)");
            dom->cdom->finalize();
            const LevelDev& Cg = dom->cdom->dev[dom->cdom->top()];
            gvals.alloc((size_t)Cg.nnzb * dim * dim);
            bg.alloc((size_t)Cg.nv * dim);
            xg.alloc((size_t)Cg.nv * dim);
            coarse = std::make_shared<Gmg>();
            coarse->dom = dom->cdom;
            coarse->desc = desc;
        }
        coef_stride = 2 * std::max(1, std::max(desc.pre_smooth, desc.post_smooth));
        coefs.alloc((size_t)(top + 1) * 2 * coef_stride);
    }
    L[top].vals = top_vals;
    if (dim == 2) gmg_setup_kernels<2>(*this); else gmg_setup_kernels<3>(*this);
    // ONE synchronisation per setup: Gershgorin bounds and the singularity flag come back together (pinned memory)
    int* h_fail = reinterpret_cast<int*>(ctx->h_results + Context::kResultSlots - 1);
    auto fetch = [&]() {
        if (dist) {          // (exact, loose) per decomposed level, levels base+1 .. top
            if (top > base) AB_CUDA(cudaMemcpyAsync(ctx->h_results, ctx->d_results + 2 * (base + 1), 2 * (top - base) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        } else if (top >= 1) AB_CUDA(cudaMemcpyAsync(ctx->h_results, ctx->d_results + 1, top * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        AB_CUDA(cudaMemcpyAsync(h_fail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        AB_CUDA(cudaStreamSynchronize(ctx->stream));
        return *h_fail;
    };
    int hfail = fetch();
    if (hfail == 2 && !force_pivoting) {          // tiny pivot without row exchanges: redo everything with partial pivoting
        force_pivoting = true;
        if (dim == 2) gmg_setup_kernels<2>(*this); else gmg_setup_kernels<3>(*this);
        hfail = fetch();
    }
    if (top > base) {
        std::vector<double> hc((size_t)(top + 1) * 2 * coef_stride, 0.0);
        for (int l = base + 1; l <= top; ++l) {
            if (dist) {
                // the exact bound (shared blocks summed before the absolute value) lies in (0.5 loose, loose]; anything else means the
                // shared-block lists do not match the operator -- fall back to the loose (always valid, partition-dependent) bound
                const double exact = ctx->h_results[2 * (l - base - 1)], loose = ctx->h_results[2 * (l - base - 1) + 1];
                L[l].lmax = (exact > 0.5 * loose && exact <= loose * (1.0 + 1e-9)) ? exact : loose;
            } else {
                L[l].lmax = ctx->h_results[l - 1];
            }
            smoother_coefs(desc, L[l].lmax, desc.pre_smooth, L[l].coef_pre);
            smoother_coefs(desc, L[l].lmax, desc.post_smooth, L[l].coef_post);
            double* hp = hc.data() + (size_t)l * 2 * coef_stride;
            for (size_t k = 0; k < L[l].coef_pre.size(); ++k) { hp[2 * k] = L[l].coef_pre[k].first; hp[2 * k + 1] = L[l].coef_pre[k].second; }
            for (size_t k = 0; k < L[l].coef_post.size(); ++k) { hp[coef_stride + 2 * k] = L[l].coef_post[k].first; hp[coef_stride + 2 * k + 1] = L[l].coef_post[k].second; }
            L[l].cf_pre = coefs.p + (size_t)l * 2 * coef_stride;
            L[l].cf_post = L[l].cf_pre + coef_stride;
        }
        // stream-ordered before every later launch (a kernel that follows a copy waits for it, PDL or not); pageable source:
        // the call returns once the buffer is staged, so `hc` may go out of scope
        AB_CUDA(cudaMemcpyAsync(coefs.p, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    AB_REQUIRE(hfail == 0, AB_ERR_STATE, "coarse-level matrix is singular");
    if (dist && dom->p2p_connected) {      // the setup's interface sums (diagonal, row sums) must have arrived as well
        int herr = 0;
        AB_CUDA(cudaMemcpyAsync(&herr, dom->p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        AB_CUDA(cudaStreamSynchronize(ctx->stream));
        AB_REQUIRE(herr == 0, AB_ERR_CUDA, "peer-to-peer interface exchange timed out during solver:init (a neighbour rank did not arrive)");
    }
}

void Gmg::smooth(int l, const double* b, double* x, int nu, bool zero_guess, const std::vector<std::pair<double, double>>& coef, const double* cf,
                 bool first_done) {
    if (nu <= 0) { if (zero_guess) dev_fill(dom->ctx, x, (int64_t)dom->dev[l].nv * dom->dim(), 0.0); return; }
    Context* ctx = dom->ctx;
    const LevelDev& Ld = dom->dev[l];
    GmgLevel& g = L[l];
    const int64_t n = (int64_t)Ld.nv * dom->dim();
    double* bufs[2] = {x, g.x2.p};
    int cur;   // buffer holding the current iterate
    int k = 0;
    const bool dist = dom->distributed() && dom->iface[l].niv > 0;
    Interface* I = dist ? &dom->iface[l] : nullptr;
    const int D = dom->dim();
    const bool p2p = dist && dom->p2p_connected;
    double* dbuf[2] = {g.d.p, p2p ? g.d2.p : g.d.p};     // P2P path ping-pongs d so that d_old survives the fused step
    int dcur = 0;
    if (zero_guess) {
        cur = ((nu - 1) % 2 == 0) ? 0 : 1;
        if (!first_done) AB_LAUNCH_PDL(ctx, k_smooth_first, ew_grid(ctx, n), 256, 0, n, coef[0].second, g.dinv.p, b, dbuf[dcur], bufs[cur], cf);
        if (p2p) {          // d = c2 D^-1 b is additive at the interfaces: sum it, x = d there -- one fused launch (c1 = 0 in cf[0])
            smooth_exchange_p2p(dom, l, D, cf, nullptr, dbuf[dcur], nullptr, bufs[cur]);
        } else if (dist) {
            AB_LAUNCH(ctx, k_iface_save, grid_for((int64_t)I->niv * D, 256, ctx->num_sms), 256, 0, I->niv, D, I->iv.p, (const double*)nullptr, (const double*)nullptr, I->save.p);
            exchange_sum(dom, l, g.d.p, D);
            AB_LAUNCH(ctx, k_iface_fix, grid_for((int64_t)I->niv * D, 256, ctx->num_sms), 256, 0, I->niv, D, I->iv.p, 0.0, I->save.p, g.d.p, bufs[cur]);
        }
        k = 1;
    } else {
        cur = 0;   // caller guarantees: iterate is in x if nu even, in x2 if nu odd  (see vcycle)
        if (nu % 2 == 1) cur = 1;
    }
    for (; k < nu; ++k) {
        // multi-GPU: the local step yields d_new = c1 d_old + c2 D^-1 r_local at shared vertices; the increment
        // c2 D^-1 r_local is additive -> sum it over the interfaces and rebuild d, x there (comm.cuh)
        const double c1 = coef[k].first, c2 = coef[k].second;
        const double* cfk = cf ? cf + 2 * k : nullptr;
        if (p2p) {
            spmv(ctx, dom->dim(), Ld, g.vals, 2, 0, bufs[cur], b, bufs[1 - cur], g.dinv.p, dbuf[dcur], c1, c2, nullptr, nullptr, dbuf[1 - dcur], cfk, 1);
            smooth_exchange_p2p(dom, l, D, cfk, dbuf[dcur], dbuf[1 - dcur], bufs[cur], bufs[1 - cur]);
            dcur = 1 - dcur;
        } else {
            if (dist) AB_LAUNCH(ctx, k_iface_save, grid_for((int64_t)I->niv * D, 256, ctx->num_sms), 256, 0, I->niv, D, I->iv.p, c1 != 0.0 ? g.d.p : (const double*)nullptr, (const double*)bufs[cur], I->save.p);
            spmv(ctx, dom->dim(), Ld, g.vals, 2, 0, bufs[cur], b, bufs[1 - cur], g.dinv.p, g.d.p, c1, c2, nullptr, nullptr, nullptr, cfk, 1);
            if (dist) {
                AB_LAUNCH(ctx, k_iface_inc, grid_for((int64_t)I->niv * D, 256, ctx->num_sms), 256, 0, I->niv, D, I->iv.p, c1, I->save.p, g.d.p);
                exchange_sum(dom, l, g.d.p, D);
                AB_LAUNCH(ctx, k_iface_fix, grid_for((int64_t)I->niv * D, 256, ctx->num_sms), 256, 0, I->niv, D, I->iv.p, c1, I->save.p, g.d.p, bufs[1 - cur]);
            }
        }
        cur = 1 - cur;
    }
}

// multi-GPU base level: the additive right-hand sides of all ranks cross the vertical interface to rank 0, which runs the
// single-GPU cycle of the global coarse hierarchy and hands every rank its (consistent) part of the solution
void Gmg::vcycle_base_gathered(const double* b, double* x) {
    Context* ctx = dom->ctx;
    Comm& cm = *ctx->comm;
    NcclApi& nc = NcclApi::get();
    const int D = dom->dim();
    const size_t n_loc = (size_t)dom->dev[base].nv * D;
    if (cm.rank != 0) {
        AB_NCCL(nc.Send(b, n_loc, ncclFloat64, 0, cm.comm, ctx->stream));
        AB_NCCL(nc.Recv(x, n_loc, ncclFloat64, 0, cm.comm, ctx->stream));
        return;
    }
    std::vector<int64_t> off((size_t)cm.nranks + 1, 0);
    for (int r = 0; r < cm.nranks; ++r) off[r + 1] = off[r] + dom->g_nv[r];
    double* st = dom->g_vstage.p;
    dev_copy(ctx, (int64_t)n_loc, b, st);
    AB_NCCL(nc.GroupStart());
    for (int r = 1; r < cm.nranks; ++r) AB_NCCL(nc.Recv(st + (size_t)off[r] * D, (size_t)dom->g_nv[r] * D, ncclFloat64, r, cm.comm, ctx->stream));
    AB_NCCL(nc.GroupEnd());
    const int nvg = dom->cdom->dev[dom->cdom->top()].nv;
    AB_LAUNCH(ctx, k_vgather, ew_grid(ctx, (int64_t)nvg * D), 256, 0, nvg, D, dom->g_csr_ptr.p, dom->g_csr_idx.p, st, bg.p);
    coarse->apply(bg.p, xg.p);
    AB_LAUNCH(ctx, k_vscatter, ew_grid(ctx, dom->g_nv_total * D), 256, 0, dom->g_nv_total, D, dom->g_l2g.p, xg.p, st);
    AB_NCCL(nc.GroupStart());
    for (int r = 1; r < cm.nranks; ++r) AB_NCCL(nc.Send(st + (size_t)off[r] * D, (size_t)dom->g_nv[r] * D, ncclFloat64, r, cm.comm, ctx->stream));
    AB_NCCL(nc.GroupEnd());
    dev_copy(ctx, (int64_t)n_loc, st, x);
}

void Gmg::vcycle(int l, const double* b, double* x, bool first_done) {
    Context* ctx = dom->ctx;
    const int dim = dom->dim();
    if (l == base) {
        if (dom->distributed()) { vcycle_base_gathered(b, x); return; }
        const int n = n_free;
        const int grid = std::max(1, std::min((n + 7) / 8, ctx->num_sms * 8));
        AB_LAUNCH_PDL(ctx, k_coarse_solve, grid, 256, (size_t)n * sizeof(double), n, n0, Ainv.p, free2dof.p, dof2free.p, b, x, 1);
        return;
    }
    const LevelDev& Ld = dom->dev[l];
    const LevelDev& Lc = dom->dev[l - 1];
    GmgLevel& g = L[l];
    GmgLevel& gc = L[l - 1];
    const int64_t n = (int64_t)Ld.nv * dim;
    smooth(l, b, x, desc.pre_smooth, true, g.coef_pre, g.cf_pre, first_done);
    spmv(ctx, dim, Ld, g.vals, 1, 0, x, b, g.r.p, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 1);
    // the transfers are gather kernels (two dependent loads per entry): more CTAs in flight, not more work per thread, hide that
    // latency (ncu, numRefs 5: prolongation 259 us / restriction 132 us at 1.5 / 2.5 TB/s with 8 CTAs per SM)
    const int rg = grid_for((int64_t)Lc.nv * 16, 256, ctx->num_sms * 32);
    const int pg = grid_for(n, 256, ctx->num_sms * 32);
    if (dim == 2) AB_LAUNCH_PDL(ctx, (k_restrict<2>), rg, 256, 0, Lc.nv, Lc.rowptr.p, Lc.mid.p, Lc.diagpos.p, gc.mask, g.r.p, gc.b.p);
    else AB_LAUNCH_PDL(ctx, (k_restrict<3>), rg, 256, 0, Lc.nv, Lc.rowptr.p, Lc.mid.p, Lc.diagpos.p, gc.mask, g.r.p, gc.b.p);
    vcycle(l - 1, gc.b.p, gc.x.p);
    // x_out = x + P xc ; written to x2 when the post-smoother runs an odd number of steps so that it ends in x
    double* target = (desc.post_smooth % 2 == 1) ? g.x2.p : x;
    if (dim == 2) AB_LAUNCH_PDL(ctx, (k_prolong_add<2>), pg, 256, 0, Lc.nv, Ld.nv, Ld.pa.p, Ld.pb.p, gc.x.p, (const double*)x, target);
    else AB_LAUNCH_PDL(ctx, (k_prolong_add<3>), pg, 256, 0, Lc.nv, Ld.nv, Ld.pa.p, Ld.pb.p, gc.x.p, (const double*)x, target);
    smooth(l, b, x, desc.post_smooth, false, g.coef_post, g.cf_post);
}

// =============================================================================================
// solvers
// =============================================================================================
struct Solver {
    Space* sp = nullptr;
    int type = 0;                      // 1 = BiCGStab + GMG, 2 = CG + Jacobi (diagonal operators)
    ab_gmg_desc desc{};
    double damp = 0.66;
    std::shared_ptr<MatrixData> A;
    std::shared_ptr<Gmg> gmg;
    DevBuf<double> r, p, v, s, sc, out2;   // CG + Jacobi work vectors (BiCGStab: the workspace lives in the shared hierarchy, KrylovWs)
    int last_steps = 0;
    double last_defect = 0;
    void ensure_vectors();
};

void Solver::ensure_vectors() {
    const int64_t n = sp->ndofs;
    if (sc.n == 0) { sc.alloc(SC_COUNT + 3); out2.alloc(2); }
    if (type == 2 && r.n == 0) { r.alloc(n); p.alloc(n); v.alloc(n); s.alloc(n); }
}

static void solver_init(Solver* S, Operator* A) {
    AB_REQUIRE(A->data && A->data->assembled, AB_ERR_STATE, "solver:init called with an operator that was never assembled");
    S->A = A->data;
    S->ensure_vectors();
    if (S->type == 1) {
        AB_REQUIRE(A->data->kind == 1, AB_ERR_ARG, "BiCGStab+GMG needs a P1 (BSR) operator");
        const std::string key = gmg_key(S->desc);
        auto it = A->data->gmg.find(key);
        if (it != A->data->gmg.end() && !env_flag("ADMM_B200_NO_CACHE")) { S->gmg = it->second; return; }
        // recycle an idle hierarchy of this domain (same descriptor and Dirichlet set): its level buffers, BiCGStab workspace
        // and captured iteration graph keep their addresses, so in steady state nothing is reallocated or re-instantiated
        std::string pool_key = key + "|";
        for (auto& e : A->data->sig.dir) pool_key += std::to_string(e.first) + "." + std::to_string(e.second) + ",";
        Domain* dom = S->sp->dom;
        S->gmg.reset();
        std::shared_ptr<Gmg> G;
        for (auto& g : dom->gmg_pool)
            if (g.use_count() == 1 && g->pool_key == pool_key) { G = g; break; }
        if (!G) {
            auto& pool = dom->gmg_pool;
            if (pool.size() >= 8) pool.erase(std::remove_if(pool.begin(), pool.end(), [](const std::shared_ptr<Gmg>& g) { return g.use_count() == 1; }), pool.end());
            G = std::make_shared<Gmg>();
            G->pool_key = pool_key;
            pool.push_back(G);
        }
        G->dom = dom;
        G->desc = S->desc;
        G->setup(A->data->vals.p, A->data->sig.dir);
        A->data->gmg[key] = G;
        S->gmg = G;
    } else {
        AB_REQUIRE(A->data->kind == 2, AB_ERR_ARG, "CG+Jacobi is implemented for diagonal (P0 mass) operators");
    }
}

// Multi-GPU BiCGStab.  The matrix is additive, so every SpMV result is made consistent by one interface sum; the Krylov
// vectors (r, p, v, s, t) are then all CONSISTENT -- bitwise identical on every rank sharing a vertex, because the
// interface sum adds the ranks' contributions in rank order -- and every inner product is taken over the owned copies only
// (owner mask): mathematically the global dot product, free of the large cancelling per-rank parts an additive residual
// carries at shared vertices.  The preconditioner input is the unique (owner-only) form, a valid additive representation.
// 2 interface sums + 3 all-reduces per iteration on top of the single-GPU sequence.
static void dot_owned(Domain* dom, int64_t n, const double* x0, const double* x1, const double* y, int nx, double* out) {
    Context* ctx = dom->ctx;
    const unsigned char* owned = dom->iface[dom->top()].owned.p;
    if (nx == 1) AB_LAUNCH(ctx, (k_dot_owned<1>), red_grid(ctx, n), 256, 0, n, dom->dim(), owned, x0, x0, y, ctx->d_partials, ctx->d_tickets, out);
    else AB_LAUNCH(ctx, (k_dot_owned<2>), red_grid(ctx, n), 256, 0, n, dom->dim(), owned, x0, x1, y, ctx->d_partials, ctx->d_tickets, out);
    allreduce_dev(dom, out, nx);
}
static void to_unique(Domain* dom, int64_t n, const double* src, double* dst) {
    dev_copy(dom->ctx, n, src, dst);
    AB_LAUNCH(dom->ctx, k_zero_not_owned, ew_grid(dom->ctx, n), 256, 0, n, dom->dim(), dom->iface[dom->top()].owned.p, dst);
}

// one full BiCGStab iteration on the workspace vectors; ends with the D2H copy of the scalars the ConvCheck reads.
// The vector updates that feed a V-cycle are fused with that cycle's first smoothing step, the scalar bookkeeping
// with the last reduction: 41 launches per iteration at three levels instead of 44.
static void bicg_iteration(Domain* dom, const double* Av, Gmg* G, KrylovWs& W, int64_t n, bool in_loop = false, unsigned long long cond_handle = 0) {
    Context* ctx = dom->ctx;
    const int dim = dom->dim(), top = dom->top();
    const LevelDev& Lt = dom->dev[top];
    const bool dist = dom->distributed();
    const unsigned char* owned = dist ? dom->iface[top].owned.p : nullptr;
    double* uq = dist ? W.uq.p : nullptr;
    double* sc = W.sc.p;
    double *fd = nullptr, *fx = nullptr;
    const bool fuse = G->first_step_targets(W.ph.p, &fd, &fx);
    const GmgLevel& gt = G->L.back();
    if (fuse) {
        AB_LAUNCH_PDL(ctx, k_bicg_fused_first<0>, ew_grid(ctx, n), 256, 0, n, sc, W.r.p, W.v.p, W.p.p, gt.dinv.p, gt.cf_pre, fd, fx, uq, owned, dim);
        G->apply(dist ? uq : W.p.p, W.ph.p, true);
    } else {
        AB_LAUNCH_PDL(ctx, k_bicg_update_p, ew_grid(ctx, n), 256, 0, n, sc, W.r.p, W.v.p, W.p.p);
        if (dist) to_unique(dom, n, W.p.p, uq);
        G->apply(dist ? uq : W.p.p, W.ph.p);
    }
    if (!dist) {
        spmv(ctx, dim, Lt, Av, 0, 1, W.ph.p, nullptr, W.v.p, nullptr, nullptr, 0, 0, W.rh.p, sc + SC_RV, nullptr, nullptr, 1);
    } else {
        spmv(ctx, dim, Lt, Av, 0, 0, W.ph.p, nullptr, W.v.p, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 1);
        exchange_sum(dom, top, W.v.p, dim);
        dot_owned(dom, n, W.rh.p, nullptr, W.v.p, 1, sc + SC_RV);
    }
    if (fuse) {
        G->first_step_targets(W.sh.p, &fd, &fx);
        AB_LAUNCH_PDL(ctx, k_bicg_fused_first<1>, ew_grid(ctx, n), 256, 0, n, sc, W.r.p, W.v.p, W.s.p, gt.dinv.p, gt.cf_pre, fd, fx, uq, owned, dim);
        G->apply(dist ? uq : W.s.p, W.sh.p, true);
    } else {
        AB_LAUNCH_PDL(ctx, k_bicg_s, red_grid(ctx, n), 256, 0, n, sc, W.r.p, W.v.p, W.s.p, ctx->d_partials, ctx->d_tickets);   // its local |s|^2 is not used
        if (dist) to_unique(dom, n, W.s.p, uq);
        G->apply(dist ? uq : W.s.p, W.sh.p);
    }
    if (!dist) {
        spmv(ctx, dim, Lt, Av, 0, 2, W.sh.p, nullptr, W.t.p, nullptr, nullptr, 0, 0, W.s.p, sc + SC_TS, nullptr, nullptr, 1);
    } else {
        spmv(ctx, dim, Lt, Av, 0, 0, W.sh.p, nullptr, W.t.p, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 1);
        exchange_sum(dom, top, W.t.p, dim);
        dot_owned(dom, n, W.s.p, W.t.p, W.t.p, 2, sc + SC_TS);               // <s,t>, <t,t>
    }
    if (dist) {      // owner-masked local sums, one all-reduce, then the scalar bookkeeping
        AB_LAUNCH_PDL(ctx, k_bicg_xr<0>, red_grid(ctx, n), 256, 0, n, sc, W.ph.p, W.sh.p, W.s.p, W.t.p, W.rh.p, W.x.p, W.r.p, ctx->d_partials,
                      ctx->d_tickets, W.out2.p, 0ull, (double*)nullptr, (double*)nullptr, 0, owned, dim);
        allreduce_dev(dom, W.out2.p, 2);
        AB_LAUNCH(ctx, k_bicg_roll, 1, 1, 0, sc, W.out2.p);
    } else if (in_loop) {       // body of the conditional WHILE node: the ConvCheck runs on the device, nothing goes to the host
        AB_LAUNCH_PDL(ctx, k_bicg_xr<2>, red_grid(ctx, n), 256, 0, n, sc, W.ph.p, W.sh.p, W.s.p, W.t.p, W.rh.p, W.x.p, W.r.p, ctx->d_partials,
                      ctx->d_tickets, W.out2.p, cond_handle, W.ctl.p, W.hist.p, (int)W.hist.n, owned, dim);
        return;
    } else {
        AB_LAUNCH_PDL(ctx, k_bicg_xr<1>, red_grid(ctx, n), 256, 0, n, sc, W.ph.p, W.sh.p, W.s.p, W.t.p, W.rh.p, W.x.p, W.r.p, ctx->d_partials,
                      ctx->d_tickets, W.out2.p, 0ull, (double*)nullptr, (double*)nullptr, 0, owned, dim);
    }
    AB_CUDA(cudaMemcpyAsync(ctx->h_results, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
}

static std::vector<const void*> iteration_key(Context* ctx, const double* Av, Gmg* G) {
    std::vector<const void*> key;
    key.push_back(Av);
    key.push_back(G->Ainv.p);
    key.push_back(G->coefs.p);
    key.push_back((const void*)(intptr_t)(G->n_free * 256 + ctx->spmv2d_lanes * 8 + ctx->tma_small_ctas * 4 + ctx->spmv_variant * 2 + (ctx->use_pdl ? 1 : 0)));
    for (const GmgLevel& g : G->L) { key.push_back(g.vals); key.push_back(g.mask); key.push_back(g.dinv.p); key.push_back(g.x.p); key.push_back(g.r.p); }
    if (G->coarse) {
        key.push_back(G->coarse->Ainv.p);
        key.push_back(G->coarse->coefs.p);
        for (const GmgLevel& g : G->coarse->L) { key.push_back(g.vals); key.push_back(g.dinv.p); key.push_back(g.x.p); }
    }
    return key;
}

// The whole iteration loop as ONE graph launch: a conditional WHILE node whose body is one BiCGStab iteration; the last kernel
// of the body evaluates the ConvCheck on the device and sets the loop condition (cudaGraphSetConditional), so the host does not
// see the solve again before it has ended (no launch + synchronisation per iteration).  The graph ends with the D2H copies of
// the scalars and of the iteration count.  Returns false (once and for all for this workspace) if the driver refuses any step:
// the caller then replays the per-iteration graph.  Single GPU only (distributed solves are not launch-bound: small problems
// are not decomposed at all, see ug4.Backend).
static bool ensure_loop_graph(Domain* dom, const double* Av, Gmg* G, KrylovWs& W, int64_t n) {
    Context* ctx = dom->ctx;
    if (W.loop_failed) return false;
    std::vector<const void*> key = iteration_key(ctx, Av, G);
    if (W.exec_loop && W.key_loop == key) return true;
    TraceTimer tt(ctx->stream, "bicgstab: loop graph build");
    if (W.exec_loop) { cudaGraphExecDestroy(W.exec_loop); W.exec_loop = nullptr; }
    cudaGraph_t graph = nullptr;
    bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess;
    cudaGraphConditionalHandle handle = 0;
    ok = ok && cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    cudaGraphNode_t wnode = nullptr;
    cudaGraph_t body = nullptr;
    if (ok) {
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        ok = cudaGraphAddNode(&wnode, graph, nullptr, 0, &cp) == cudaSuccess && cp.conditional.phGraph_out != nullptr;
        if (ok) body = cp.conditional.phGraph_out[0];
    }
    const int64_t l0 = ctx->launches;
    if (ok && cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        try {
            bicg_iteration(dom, Av, G, W, n, true, (unsigned long long)handle);
        } catch (...) {
            ok = false;
        }
        cudaGraph_t ended = nullptr;
        if (cudaStreamEndCapture(ctx->stream, &ended) != cudaSuccess) ok = false;
    } else {
        ok = false;
    }
    W.nodes_loop = ctx->launches - l0;
    ctx->launches = l0;
    if (ok) {
        cudaGraphNode_t m1 = nullptr, m2 = nullptr;
        ok = cudaGraphAddMemcpyNode1D(&m1, graph, &wnode, 1, ctx->h_results, W.sc.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
             cudaGraphAddMemcpyNode1D(&m2, graph, &m1, 1, ctx->h_results + 16, W.ctl.p, CTL_COUNT * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (ok) ok = cudaGraphInstantiate(&W.exec_loop, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        if (W.exec_loop) { cudaGraphExecDestroy(W.exec_loop); W.exec_loop = nullptr; }
        W.loop_failed = true;
        if (env_flag("ADMM_B200_TRACE")) fprintf(stderr, "[ab trace] conditional loop graph unavailable, per-iteration replay\n");
        return false;
    }
    W.key_loop = key;
    return true;
}

// (re)capture the iteration when any pointer baked into its launches changed since the last capture.  Multi-GPU: the capture
// holds the interface-exchange kernels (device-side epochs) and the NCCL calls of the iteration; if the driver or NCCL refuses
// it, the workspace stays on stream launches (returns false).
static bool ensure_iteration_graph(Domain* dom, const double* Av, Gmg* G, KrylovWs& W, int64_t n) {
    Context* ctx = dom->ctx;
    if (W.graph_failed) return false;
    std::vector<const void*> key = iteration_key(ctx, Av, G);
    if (W.exec && W.key == key) return true;
    TraceTimer tt(ctx->stream, "bicgstab: graph capture");
    const bool dist = dom->distributed();
    cudaGraph_t graph = nullptr;
    const int64_t l0 = ctx->launches;
    AB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    bool ok = true;
    try {
        bicg_iteration(dom, Av, G, W, n);
    } catch (...) {
        cudaStreamEndCapture(ctx->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->launches = l0;
        if (!dist) throw;
        ok = false;
        graph = nullptr;
    }
    if (ok && cudaStreamEndCapture(ctx->stream, &graph) != cudaSuccess) {
        cudaGetLastError();
        if (!dist) throw ab::Error(-2, "cudaStreamEndCapture failed for the BiCGStab iteration");
        ok = false;
    }
    if (!ok) {
        W.graph_failed = true;
        if (env_flag("ADMM_B200_TRACE")) fprintf(stderr, "[ab trace] multi-GPU iteration graph unavailable, stream launches\n");
        return false;
    }
    if (const char* dot = getenv("ADMM_B200_GRAPH_DOT")) cudaGraphDebugDotPrint(graph, dot, cudaGraphDebugDotFlagsVerbose);   // diagnostics
    W.nodes = ctx->launches - l0;
    ctx->launches = l0;                      // capturing launches nothing
    if (W.exec) {                            // same topology, new pointers: update in place (cheaper than instantiating)
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(W.exec, graph, &info) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphExecDestroy(W.exec);
            W.exec = nullptr;
        }
    }
    if (!W.exec && cudaGraphInstantiate(&W.exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        W.exec = nullptr;
        cudaGraphDestroy(graph);
        if (!dist) throw ab::Error(-2, "cudaGraphInstantiate failed for the BiCGStab iteration");
        W.graph_failed = true;
        return false;
    }
    AB_CUDA(cudaGraphDestroy(graph));
    W.key = key;
    return true;
}

static bool bicgstab_apply(Solver* S, Vector* x, Vector* b, bool return_defect) {
    Domain* dom = S->sp->dom;
    Context* ctx = dom->ctx;
    const int dim = dom->dim(), top = dom->top();
    const LevelDev& Lt = dom->dev[top];
    const int64_t n = S->sp->ndofs;
    const double* Av = S->A->vals.p;
    const bool dist = dom->distributed();
    const unsigned char* owned = dist ? dom->iface[top].owned.p : nullptr;
    Gmg* G = S->gmg.get();
    if (!G->ws) G->ws = std::make_unique<KrylovWs>();
    KrylovWs& W = *G->ws;
    if (W.r.n != (size_t)n) {
        W.r.alloc(n); W.rh.alloc(n); W.p.alloc(n); W.v.alloc(n); W.s.alloc(n); W.t.alloc(n); W.ph.alloc(n); W.sh.alloc(n); W.x.alloc(n);
        if (dist) { W.uq.alloc(n); W.bq.alloc(n); }
        W.sc.alloc(SC_COUNT + 3); W.out2.alloc(2);
        W.ctl.alloc(CTL_COUNT + 1); W.hist.alloc(4097);
        W.key.clear(); W.key_loop.clear();
    }
    double* sc = W.sc.p;
    const double tol2 = S->desc.abs_tol * S->desc.abs_tol;
    double h[SC_COUNT];
    const double* bp = b->d.p;
    if (dist) {
        // storage types at the solver boundary (UG4 converts here; 3d_admm.lua:978 only guards one call): the right-hand side
        // must be additive (a consistent one would be counted once per sharing rank), the start iterate consistent
        if (!(b->storage & (AB_PST_ADDITIVE | AB_PST_UNIQUE))) { to_unique(dom, n, b->d.p, W.bq.p); bp = W.bq.p; }
        if (!(x->storage & AB_PST_CONSISTENT)) { exchange_sum(dom, top, x->d.p, dim); x->storage = AB_PST_CONSISTENT; }
    }
    // r = b - A x ; then in one pass: workspace x = x, rh = r, p = v = 0, rho = <r,r> and the other scalars
    spmv(ctx, dim, Lt, Av, 1, 0, x->d.p, bp, W.r.p);
    if (dist) exchange_sum(dom, top, W.r.p, dim);                              // additive -> consistent
    AB_LAUNCH(ctx, k_bicg_init, red_grid(ctx, n), 256, 0, n, x->d.p, W.x.p, W.r.p, W.rh.p, W.p.p, W.v.p, sc, ctx->d_partials, ctx->d_tickets, W.out2.p,
              W.ctl.p, tol2, S->desc.red_tol > 0 ? S->desc.red_tol * S->desc.red_tol : 0.0, S->desc.max_iterations, owned, dim);
    if (dist) {
        allreduce_dev(dom, W.out2.p, 1);
        AB_LAUNCH(ctx, k_bicg_init_fix, 1, 1, 0, sc, W.ctl.p, W.out2.p);
    }
    read_back(ctx, sc, SC_COUNT, h);
    const double rr0 = h[SC_RR];
    double rr = rr0;
    bool ok = rr < tol2;
    int it = 0;
    const bool verbose = S->desc.verbose && (!dist || ctx->comm->rank == 0);
    if (verbose) printf("  BiCGStab+GMG: iter %4d  defect %.6e\n", 0, std::sqrt(rr));
    bool graph = ctx->use_graph && ctx->stream != nullptr && (!dist || dom->p2p_connected);
    const bool want_iter = !ok && S->desc.max_iterations > 0;
    if (graph && want_iter && !dist && ctx->use_loop && ensure_loop_graph(dom, Av, G, W, n)) {
        // the whole loop in one launch (conditional WHILE node); the device applied the same ConvCheck as the host loop below
        AB_CUDA(cudaGraphLaunch(W.exec_loop, ctx->stream));
        AB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < SC_COUNT; ++i) h[i] = ctx->h_results[i];
        it = (int)ctx->h_results[16 + CTL_IT];
        ctx->launches += W.nodes_loop * it;
        rr = h[SC_RR];
        const bool finite = (rr == rr) && !std::isinf(rr);
        ok = finite && (rr < tol2 || (S->desc.red_tol > 0 && rr < S->desc.red_tol * S->desc.red_tol * rr0));
        if (verbose) {
            std::vector<double> hh((size_t)std::min<int>(it + 1, (int)W.hist.n), 0.0);
            AB_CUDA(cudaMemcpy(hh.data(), W.hist.p, hh.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (size_t k = 1; k < hh.size(); ++k) printf("  BiCGStab+GMG: iter %4d  defect %.6e\n", (int)k, std::sqrt(hh[k]));
        }
    } else {
        if (graph && want_iter) graph = ensure_iteration_graph(dom, Av, G, W, n);
        while (!ok && it < S->desc.max_iterations) {
            ++it;
            if (graph) {
                AB_CUDA(cudaGraphLaunch(W.exec, ctx->stream));
                ctx->launches += W.nodes;
            } else {
                bicg_iteration(dom, Av, G, W, n);
            }
            AB_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < SC_COUNT; ++i) h[i] = ctx->h_results[i];
            rr = h[SC_RR];
            if (verbose) printf("  BiCGStab+GMG: iter %4d  defect %.6e\n", it, std::sqrt(rr));
            if (!(rr == rr) || std::isinf(rr)) break;                       // NaN/Inf: breakdown
            if (rr < tol2 || (S->desc.red_tol > 0 && rr < S->desc.red_tol * S->desc.red_tol * rr0)) { ok = true; break; }
            if (h[SC_RHO] == 0.0 || h[SC_OMEGA] == 0.0) break;               // breakdown
        }
    }
    if (it > 0) dev_copy(ctx, n, W.x.p, x->d.p);
    x->touch();
    S->last_steps = it;
    S->last_defect = std::sqrt(std::max(rr, 0.0));
    if (return_defect) {
        if (dist) { to_unique(dom, n, W.r.p, b->d.p); b->storage = AB_PST_ADDITIVE; }   // unique = additive form of the defect
        else dev_copy(ctx, n, W.r.p, b->d.p);
        b->touch();
    }
    if (dist && dom->p2p_connected) {      // a lost or stalled peer shows up as an error flag of the exchange kernels, never as a wrong sum
        int herr = 0;
        AB_CUDA(cudaMemcpyAsync(&herr, dom->p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        AB_CUDA(cudaStreamSynchronize(ctx->stream));
        AB_REQUIRE(herr == 0, AB_ERR_CUDA, "peer-to-peer interface exchange timed out (a neighbour rank did not arrive)");
    }
    return ok;
}

static bool cg_jacobi_apply(Solver* S, Vector* x, Vector* b, bool return_defect) {
    Context* ctx = S->sp->dom->ctx;
    const int64_t n = S->sp->ndofs;
    const double* diag = S->A->vals.p;
    double* sc = S->sc.p;
    double *r = S->r.p, *z = S->s.p, *p = S->p.p, *q = S->v.p;
    const double tol2 = S->desc.abs_tol * S->desc.abs_tol;
    double h[SC_COUNT];
    AB_LAUNCH(ctx, k_cg_init, red_grid(ctx, n), 256, 0, n, S->damp, diag, b->d.p, x->d.p, r, z, p, ctx->d_partials, ctx->d_tickets, S->out2.p);
    allreduce_dev(S->sp->dom, S->out2.p, 2);      // P0 dofs are element-local: plain sums over ranks
    AB_LAUNCH(ctx, k_cg_roll, 1, 1, 0, sc, S->out2.p);
    read_back(ctx, sc, SC_COUNT, h);
    double rr = h[SC_RR];
    const double rr0 = rr;
    bool ok = rr < tol2;
    int it = 0;
    if (S->desc.verbose) printf("  CG+Jacobi: iter %4d  defect %.6e\n", 0, std::sqrt(rr));
    while (!ok && it < S->desc.max_iterations) {
        ++it;
        AB_LAUNCH(ctx, k_diag_apply_dot, red_grid(ctx, n), 256, 0, n, diag, p, q, ctx->d_partials, ctx->d_tickets, sc + SC_PQ);
        allreduce_dev(S->sp->dom, sc + SC_PQ, 1);
        AB_LAUNCH(ctx, k_cg_step, red_grid(ctx, n), 256, 0, n, S->damp, sc, diag, p, q, x->d.p, r, z, ctx->d_partials, ctx->d_tickets, S->out2.p);
        allreduce_dev(S->sp->dom, S->out2.p, 2);
        AB_LAUNCH(ctx, k_cg_roll, 1, 1, 0, sc, S->out2.p);
        AB_LAUNCH(ctx, k_cg_p, ew_grid(ctx, n), 256, 0, n, sc, z, p);
        read_back(ctx, sc, SC_COUNT, h);
        rr = h[SC_RR];
        if (S->desc.verbose) printf("  CG+Jacobi: iter %4d  defect %.6e\n", it, std::sqrt(rr));
        if (!(rr == rr)) break;
        if (rr < tol2 || (S->desc.red_tol > 0 && rr < S->desc.red_tol * S->desc.red_tol * rr0)) { ok = true; break; }
    }
    x->touch();
    S->last_steps = it;
    S->last_defect = std::sqrt(rr);
    if (return_defect) { dev_copy(ctx, n, r, b->d.p); b->touch(); }
    return ok;
}

// =============================================================================================
// assembly
// =============================================================================================
static Vector* import_u(ElemDisc* d, Vector* arg) { return d->imp_u ? d->imp_u : arg; }

template <int D>
static void assemble_hessian_kernels(Domain* dom, DomainDisc* dd, ElemDisc* h, const double* u, double* vals) {
    Context* ctx = dom->ctx;
    const LevelDev& L = dom->dev[dom->top()];
    HessParams P;
    P.c = h->params[AB_PARAM_STEP_LENGTH];
    P.lam_vol = h->params[AB_PARAM_LAMBDA_VOL];
    P.lam_b[0] = h->params[AB_PARAM_LAMBDA_BARY_X];
    P.lam_b[1] = h->params[AB_PARAM_LAMBDA_BARY_Y];
    P.lam_b[2] = D == 3 ? h->params[AB_PARAM_LAMBDA_BARY_Z] : 0.0;
    P.has_lam = (P.lam_vol != 0.0 || P.lam_b[0] != 0.0 || P.lam_b[1] != 0.0 || P.lam_b[2] != 0.0) ? 1 : 0;
    const unsigned char* mask = dd->mask(dom->top());
    if (ctx->assembly_variant == 0) {      // row-owner gather: every block row is written once, no memset, no atomics
        constexpr int LPR = D == 3 ? 32 : 8;
        const int64_t threads = (int64_t)L.nv * LPR;
        AB_LAUNCH(ctx, (k_assemble_hessian_rows<D, LPR>), grid_for(threads, 128, ctx->num_sms * 16), 128, 128 * sizeof(HessRec<D>), L.nv, L.rowptr.p, L.colidx.p,
                  L.v2e_ptr.p, L.v2e_idx.p, L.elems.p, L.xyz.p, u, mask, P, vals);
    } else {                               // atomic scatter per element (ADMM_B200_ASSEMBLY=atomic; kept for the A/B measurement)
        LevelDev& Lm = dom->dev[dom->top()];
        if (Lm.elem_pos.n == 0) {
            Lm.elem_pos.alloc((size_t)L.ne * (D + 1) * (D + 1));
            launch_elem_pos<D>(ctx, Lm);
        }
        AB_CUDA(cudaMemsetAsync(vals, 0, (size_t)L.nnzb * D * D * sizeof(double), ctx->stream));
        AB_LAUNCH(ctx, (k_assemble_hessian<D>), grid_for(L.ne, 128, ctx->num_sms * 16), 128, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u, L.elem_pos.p, mask, P, vals);
    }
    if (mask) AB_LAUNCH(ctx, (k_dirichlet_diag<D>), ew_grid(ctx, (int64_t)L.nv * D), 256, 0, L.nv, mask,
                        dom->distributed() ? dom->iface[dom->top()].owned.p : (const unsigned char*)nullptr, L.diagpos.p, vals);
}

static void assemble_jacobian(DomainDisc* dd, Operator* A, Vector* uarg) {
    Domain* dom = dd->sp->dom;
    Context* ctx = dom->ctx;
    const int dim = dom->dim();
    ElemDisc* jac = nullptr;
    for (ElemDisc* d : dd->discs)
        if (d->kind == AB_DISC_DEFORMATION_EQUATION || d->kind == AB_DISC_MASS_MODEL) {
            AB_REQUIRE(!jac, AB_ERR_UNSUPPORTED, "more than one jacobian-contributing ElemDisc in a DomainDiscretization");
            jac = d;
        }
    AB_REQUIRE(jac, AB_ERR_STATE, "assemble_jacobian: DomainDiscretization has no jacobian-contributing ElemDisc");
    Signature sig;
    sig.coords_version = dom->coords_version;
    sig.dir = dd->dir;
    std::sort(sig.dir.begin(), sig.dir.end());
    Vector* u = import_u(jac, uarg);
    const LevelDev& L = dom->dev[dom->top()];
    if (jac->kind == AB_DISC_DEFORMATION_EQUATION) {
        AB_REQUIRE(dd->sp->kind == AB_SPACE_P1 && dd->sp->ncomp == dim, AB_ERR_ARG, "DeformationEquation needs a P1 space with dim components");
        AB_REQUIRE(jac->params[AB_PARAM_SECOND_ORDER] == 0.0, AB_ERR_UNSUPPORTED,
                   "set_second_order(true) (J'' terms, 2d_admm.lua:389) needs the Navier-Stokes fields and is outside the hot path");
        sig.kind = 1;
        sig.c = jac->params[AB_PARAM_STEP_LENGTH];
        sig.lam_vol = jac->params[AB_PARAM_LAMBDA_VOL];
        sig.lam_b[0] = jac->params[AB_PARAM_LAMBDA_BARY_X];
        sig.lam_b[1] = jac->params[AB_PARAM_LAMBDA_BARY_Y];
        sig.lam_b[2] = dim == 3 ? jac->params[AB_PARAM_LAMBDA_BARY_Z] : 0.0;
        const bool has_lam = sig.lam_vol != 0 || sig.lam_b[0] != 0 || sig.lam_b[1] != 0 || sig.lam_b[2] != 0;
        if (has_lam && u) {
            // an aliased import (ab_vector_device_ptr) may change without a version bump: never matches a cached signature
            if (u->aliased) u->touch();
            sig.u_id = u->id; sig.u_version = u->version;
        }
    } else {
        AB_REQUIRE(dd->sp->kind == AB_SPACE_P0 && dd->sp->ncomp == dim * dim, AB_ERR_ARG, "MassModel needs a P0 space with dim*dim components");
        sig.kind = 2;
    }
    const bool cache = !env_flag("ADMM_B200_NO_CACHE");
    if (cache && A->data && A->data->assembled && A->data->sig == sig) return;
    if (cache) {
        auto& live = dom->live;
        live.erase(std::remove_if(live.begin(), live.end(), [](const std::weak_ptr<MatrixData>& w) { return w.expired(); }), live.end());
        for (auto& w : live) {
            auto m = w.lock();
            if (m && m->assembled && m->sig == sig) { A->data = m; return; }
        }
    }
    if (!A->data || A->data.use_count() > 1 || A->data->kind != sig.kind) {
        auto m = std::make_shared<MatrixData>();
        m->dom = dom;
        m->kind = sig.kind;
        m->vals.alloc(sig.kind == 1 ? (size_t)L.nnzb * dim * dim : (size_t)dd->sp->ndofs);
        A->data = m;
        dom->live.push_back(m);
    }
    TraceTimer tt(ctx->stream, "assemble_jacobian (miss)");
    MatrixData& M = *A->data;
    M.assembled = false;
    M.gmg.clear();
    if (sig.kind == 1) {
        const double* up = u ? u->d.p : nullptr;
        if (dim == 2) assemble_hessian_kernels<2>(dom, dd, jac, up, M.vals.p);
        else assemble_hessian_kernels<3>(dom, dd, jac, up, M.vals.p);
    } else {
        if (dim == 2) AB_LAUNCH(ctx, (k_mass_model<2>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, (const double*)nullptr, (const double*)nullptr, M.vals.p, (double*)nullptr);
        else AB_LAUNCH(ctx, (k_mass_model<3>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, (const double*)nullptr, (const double*)nullptr, M.vals.p, (double*)nullptr);
    }
    M.sig = sig;
    M.assembled = true;
    if (cache && sig.kind == 1 && sig.u_id == 0 && M.vals.n * sizeof(double) <= ((size_t)4 << 30)) dom->keep_lam0 = A->data;
}

template <int D>
static void load_launch(Context* ctx, const LevelDev& L, const double* u, const double* lam, const double* q, const LoadParams& P, double* out) {
    AB_LAUNCH(ctx, (k_assemble_load<D>), grid_for(L.ne, 128, ctx->num_sms * 16), 128, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u, lam, q, P, out);
}

static void assemble_defect(DomainDisc* dd, Vector* dvec, Vector* uarg) {
    Domain* dom = dd->sp->dom;
    Context* ctx = dom->ctx;
    const int dim = dom->dim();
    const LevelDev& L = dom->dev[dom->top()];
    AB_REQUIRE(dvec->sp->kind == dd->sp->kind && dvec->sp->ncomp == dd->sp->ncomp, AB_ERR_ARG, "assemble_defect: vector/space mismatch");
    dev_fill(ctx, dvec->d.p, dvec->n(), 0.0);
    for (ElemDisc* d : dd->discs) {
        Vector* u = import_u(d, uarg);
        const double* up = u ? u->d.p : nullptr;
        const double* lam = d->imp_lam ? d->imp_lam->d.p : nullptr;
        const double* q = d->imp_q ? d->imp_q->d.p : nullptr;
        LoadParams P;
        memset(&P, 0, sizeof P);
        switch (d->kind) {
            case AB_DISC_DEFORMATION_EQUATION: continue;   // jacobian only (DESIGN.md "Model")
            case AB_DISC_DEFORMATION_RHS:
            case AB_DISC_DEFORMATION_LARGE_RHS: {
                P.use_S = 1;
                P.tau = d->params[AB_PARAM_TAU];
                P.w[0] = d->params[AB_PARAM_LAMBDA_VOL];
                P.w[1] = d->params[AB_PARAM_LAMBDA_BARY_X];
                P.w[2] = d->params[AB_PARAM_LAMBDA_BARY_Y];
                P.w[3] = dim == 3 ? d->params[AB_PARAM_LAMBDA_BARY_Z] : 0.0;
                if (d->kind == AB_DISC_DEFORMATION_LARGE_RHS) {
                    P.w[0] += d->params[AB_PARAM_MULT_VOL];
                    P.w[1] += d->params[AB_PARAM_MULT_BX];
                    P.w[2] += d->params[AB_PARAM_MULT_BY];
                    if (dim == 3) P.w[3] += d->params[AB_PARAM_MULT_BZ];
                }
                P.sign = dim == 3 ? 1.0 : -1.0;    // sign conventions of the 3D / 2D plugin generations, DESIGN.md "Signs"
                break;
            }
            case AB_DISC_VOLUME_CONSTRAINT: P.w[0] = 1.0; P.sign = dim == 3 ? -1.0 : 1.0; break;
            case AB_DISC_BARYCENTER_CONSTRAINT: {
                const int k = (int)d->params[AB_PARAM_INDEX];
                AB_REQUIRE(k >= 1 && k <= dim, AB_ERR_ARG, "barycenter constraint: set_index must be 1..dim");
                P.w[k] = 1.0;
                P.sign = dim == 3 ? -1.0 : 1.0;
                break;
            }
            case AB_DISC_MASS_MODEL:
                if (dim == 2) AB_LAUNCH(ctx, (k_mass_model<2>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, up, lam, (double*)nullptr, dvec->d.p);
                else AB_LAUNCH(ctx, (k_mass_model<3>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, up, lam, (double*)nullptr, dvec->d.p);
                continue;
            case AB_DISC_LAMBDA_UPDATE:
                if (dim == 2) AB_LAUNCH(ctx, (k_lambda_update<2>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, up, q, d->params[AB_PARAM_TAU], dvec->d.p);
                else AB_LAUNCH(ctx, (k_lambda_update<3>), ew_grid(ctx, L.ne), 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, up, q, d->params[AB_PARAM_TAU], dvec->d.p);
                continue;
            default: AB_REQUIRE(false, AB_ERR_ARG, "unknown ElemDisc kind");
        }
        AB_REQUIRE(dd->sp->kind == AB_SPACE_P1, AB_ERR_ARG, "P1 ElemDisc in a non-P1 DomainDiscretization");
        P.has_w = (P.w[0] != 0 || P.w[1] != 0 || P.w[2] != 0 || P.w[3] != 0) ? 1 : 0;
        if (dim == 2) load_launch<2>(ctx, L, up, lam, q, P, dvec->d.p); else load_launch<3>(ctx, L, up, lam, q, P, dvec->d.p);
    }
    if (dd->sp->kind == AB_SPACE_P1) {
        const unsigned char* mask = dd->mask(dom->top());
        if (mask) {
            if (dim == 2) AB_LAUNCH(ctx, (k_zero_dirichlet<2>), ew_grid(ctx, dvec->n()), 256, 0, L.nv, mask, dvec->d.p);
            else AB_LAUNCH(ctx, (k_zero_dirichlet<3>), ew_grid(ctx, dvec->n()), 256, 0, L.nv, mask, dvec->d.p);
        }
    }
    dvec->storage = AB_PST_ADDITIVE;
    dvec->touch();
}

}  // namespace ab

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ab;

struct ab_context : Context {};
struct ab_domain : Domain {};
struct ab_space : Space {};
struct ab_vector : Vector {};
struct ab_elemdisc : ElemDisc {};
struct ab_domaindisc : DomainDisc {};
struct ab_operator : Operator {};
struct ab_solver : Solver {};

#define AB_TRY try {
#define AB_CATCH                                                         \
    }                                                                    \
    catch (const ab::Error& e) { g_last_error = e.what(); return e.code; } \
    catch (const std::exception& e) { g_last_error = e.what(); return AB_ERR_STATE; } \
    return AB_OK;
#define AB_CHECK_LAUNCH(ctx) AB_CUDA(cudaGetLastError())

extern "C" {

const char* ab_last_error(void) { return g_last_error.c_str(); }
int ab_version(void) { return 100; }

int ab_context_create(int device, void* stream, ab_context** out) {
    AB_TRY
    AB_REQUIRE(out, AB_ERR_ARG, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    AB_REQUIRE(e == cudaSuccess && ndev > 0, AB_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                                                            " (libadmm_b200 has no CPU fallback)");
    AB_REQUIRE(device >= 0 && device < ndev, AB_ERR_ARG, "device index out of range");
    AB_REQUIRE(g_live_contexts == 0 || g_context_device == device, AB_ERR_UNSUPPORTED,
               "a context on device " + std::to_string(g_context_device) + " is alive: one process drives one GPU (run one process per GPU)");
    AB_CUDA(cudaSetDevice(device));
    auto* c = new ab_context();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    if (!c->stream) {   // the legacy default stream cannot be captured into a graph: use an own blocking stream (still ordered with it)
        AB_CUDA(cudaStreamCreate(&c->stream));
        c->own_stream = true;
    }
    if (const char* v = getenv("ADMM_B200_GRAPH")) c->use_graph = atoi(v) != 0;
    if (env_flag("ADMM_B200_NO_CACHE")) c->use_cache = false;
    if (const char* v = getenv("ADMM_B200_PDL")) c->use_pdl = atoi(v) != 0;
    if (const char* v = getenv("ADMM_B200_LOOP")) c->use_loop = atoi(v) != 0;
    if (const char* v = getenv("ADMM_B200_L2_HINT")) c->l2_hint = atoi(v) != 0;
    if (const char* v = getenv("ADMM_B200_COARSE_VARIANT")) c->coarse_variant = atoi(v);
    if (const char* v = getenv("ADMM_B200_SPMV2D_LANES")) c->spmv2d_lanes = atoi(v) == 8 ? 8 : 4;
    if (const char* v = getenv("ADMM_B200_ASSEMBLY")) c->assembly_variant = (strcmp(v, "atomic") == 0 || strcmp(v, "1") == 0) ? 1 : 0;
    spmv_prepare_kernels();
    cudaDeviceProp prop;
    AB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (const char* v = getenv("ADMM_B200_SPMV_VARIANT")) c->spmv_variant = atoi(v);
    if (const char* v = getenv("ADMM_B200_SPMV_WAVES")) c->spmv_waves = std::max(1, atoi(v));
    AB_CUDA(cudaMalloc((void**)&c->d_partials, sizeof(double) * Context::kMaxBlocks * Context::kMaxVals));
    AB_CUDA(cudaMalloc((void**)&c->d_tickets, sizeof(unsigned int) * 4));
    AB_CUDA(cudaMemset(c->d_tickets, 0, sizeof(unsigned int) * 4));
    AB_CUDA(cudaMalloc((void**)&c->d_results, sizeof(double) * Context::kResultSlots));
    AB_CUDA(cudaMallocHost((void**)&c->h_results, sizeof(double) * Context::kResultSlots));
    g_context_device = device;
    ++g_live_contexts;
    *out = c;
    AB_CATCH
}
int ab_context_destroy(ab_context* ctx) {
    AB_TRY
    if (!ctx) return AB_OK;
    cudaStreamSynchronize(ctx->stream);
    g_dot_cache.erase(ctx);
    cudaFree(ctx->d_partials); cudaFree(ctx->d_tickets); cudaFree(ctx->d_results); cudaFreeHost(ctx->h_results);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    const int dev = ctx->device;
    delete ctx;
    if (--g_live_contexts <= 0) { g_live_contexts = 0; DevicePool::get().trim(dev); }   // last context gone: cached blocks go back to the driver
    AB_CATCH
}
int ab_context_synchronize(ab_context* ctx) {
    AB_TRY
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    AB_CATCH
}
int ab_context_launch_count(ab_context* ctx, int64_t* out) {
    AB_TRY
    *out = ctx->launches;
    AB_CATCH
}
int ab_context_set_tuning(ab_context* ctx, const char* key, int value) {
    AB_TRY
    const std::string k = key;
    if (k == "spmv_variant") ctx->spmv_variant = value;
    else if (k == "spmv_waves") ctx->spmv_waves = std::max(1, value);
    else if (k == "graph") ctx->use_graph = value != 0;
    else if (k == "coarse_variant") ctx->coarse_variant = value;
    else if (k == "spmv2d_lanes") ctx->spmv2d_lanes = value == 8 ? 8 : 4;
    else if (k == "assembly_variant") ctx->assembly_variant = value != 0 ? 1 : 0;
    else if (k == "pdl") ctx->use_pdl = value != 0;
    else if (k == "loop") ctx->use_loop = value != 0;
    else if (k == "l2_hint") ctx->l2_hint = value != 0;
    else if (k == "tma_small_ctas") ctx->tma_small_ctas = std::max(1, std::min(2, value));
    else AB_REQUIRE(false, AB_ERR_ARG, "unknown tuning key '" + k + "'");
    AB_CATCH
}
int ab_context_init_comm(ab_context* ctx, int rank, int nranks, const void* nccl_unique_id) {
    AB_TRY
    AB_REQUIRE(ctx && nranks >= 1 && rank >= 0 && rank < nranks, AB_ERR_ARG, "bad rank / nranks");
    if (nranks == 1) return AB_OK;
    AB_REQUIRE(nccl_unique_id, AB_ERR_ARG, "nccl_unique_id is NULL");
    auto c = std::make_shared<Comm>();
    c->rank = rank; c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    AB_CUDA(cudaSetDevice(ctx->device));
    AB_NCCL(NcclApi::get().CommInitRank(&c->comm, nranks, id, rank));
    ctx->comm = c;
    AB_CATCH
}
int ab_nccl_unique_id(void* out128) {
    AB_TRY
    AB_REQUIRE(out128, AB_ERR_ARG, "NULL argument");
    ncclUniqueId id;
    AB_NCCL(NcclApi::get().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    AB_CATCH
}
int ab_context_allreduce_host(ab_context* ctx, double* v, int n, int max_op) {
    AB_TRY
    if (ctx->comm && ctx->comm->nranks > 1) {
        AB_REQUIRE(n <= Context::kResultSlots, AB_ERR_ARG, "too many values");
        AB_CUDA(cudaMemcpyAsync(ctx->d_results, v, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->comm->allreduce(ctx->d_results, n, max_op != 0, ctx->stream);
        read_back(ctx, ctx->d_results, n, v);
    }
    AB_CATCH
}

// ---- domain ---------------------------------------------------------------------------------
int ab_domain_load_ugx(ab_context* ctx, const char* path, ab_domain** out) {
    AB_TRY
    AB_REQUIRE(path && out, AB_ERR_ARG, "NULL argument");   // ctx may be NULL for host-only use (no ApproximationSpace)
    auto d = std::make_unique<ab_domain>();
    d->ctx = ctx;
    std::string err;
    AB_REQUIRE(load_ugx(path, d->mesh, err), AB_ERR_IO, err);
    *out = d.release();
    AB_CATCH
}
int ab_domain_create(ab_context* ctx, int dim, int nv, const double* xyz, int ne, const int32_t* elems, int nsubsets,
                     const char* const* subset_names, const int32_t* vsub, const int32_t* esub, int n_sp_edges, const int32_t* sp_edges,
                     const int32_t* sp_edges_sub, int n_sp_faces, const int32_t* sp_faces, const int32_t* sp_faces_sub, ab_domain** out) {
    AB_TRY
    AB_REQUIRE(xyz && elems && vsub && esub && out, AB_ERR_ARG, "NULL argument");
    AB_REQUIRE(dim == 2 || dim == 3, AB_ERR_ARG, "dim must be 2 or 3");
    auto d = std::make_unique<ab_domain>();
    d->ctx = ctx;
    d->mesh.dim = dim;
    for (int i = 0; i < nsubsets; ++i) d->mesh.subset_names.push_back(subset_names[i]);
    HostLevel L;
    L.dim = dim; L.nv = nv; L.ne = ne;
    L.xyz.assign(xyz, xyz + (size_t)nv * dim);
    L.elems.assign(elems, elems + (size_t)ne * (dim + 1));
    L.vsub.assign(vsub, vsub + nv);
    L.esub.assign(esub, esub + ne);
    if (n_sp_edges) { L.sp_edges.assign(sp_edges, sp_edges + 2 * (size_t)n_sp_edges); L.sp_edges_sub.assign(sp_edges_sub, sp_edges_sub + n_sp_edges); }
    if (n_sp_faces) { L.sp_faces.assign(sp_faces, sp_faces + 3 * (size_t)n_sp_faces); L.sp_faces_sub.assign(sp_faces_sub, sp_faces_sub + n_sp_faces); }
    for (int e = 0; e < ne * (dim + 1); ++e) AB_REQUIRE(elems[e] >= 0 && elems[e] < nv, AB_ERR_ARG, "element vertex index out of range");
    d->mesh.levels.push_back(std::move(L));
    *out = d.release();
    AB_CATCH
}
int ab_domain_destroy(ab_domain* dom) {
    AB_TRY
    delete dom;
    AB_CATCH
}
int ab_domain_refine(ab_domain* dom, int num_refs) {
    AB_TRY
    AB_REQUIRE(dom, AB_ERR_ARG, "NULL domain");
    AB_REQUIRE(!dom->finalized, AB_ERR_STATE, "the hierarchy is frozen once an ApproximationSpace exists");
    AB_REQUIRE(num_refs >= 0, AB_ERR_ARG, "numRefs < 0");
    for (int r = 0; r < num_refs; ++r) {
        HostLevel& C = dom->mesh.levels.back();
        ensure_edges(C);
        AB_REQUIRE((int64_t)C.nv + C.nedges() < (int64_t)1 << 31, AB_ERR_UNSUPPORTED, "level exceeds int32 vertex ids");
        HostLevel F;
        refine_level(C, F);
        dom->mesh.levels.push_back(std::move(F));
    }
    AB_CATCH
}
int ab_domain_set_interface(ab_domain* dom, int level, int nneigh, const int32_t* neigh_ranks, const int32_t* offsets, const int32_t* idx,
                            const unsigned char* owned) {
    AB_TRY
    AB_REQUIRE(dom && !dom->finalized, AB_ERR_STATE, "interfaces must be set before the first ApproximationSpace");
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    if (dom->host_iface.size() < dom->mesh.levels.size()) dom->host_iface.resize(dom->mesh.levels.size());
    Domain::HostIface& H = dom->host_iface[level];
    H.neigh.assign(neigh_ranks, neigh_ranks + nneigh);
    H.offset.assign(offsets, offsets + nneigh + 1);
    H.idx.assign(idx, idx + (nneigh ? offsets[nneigh] : 0));
    const int nv = dom->mesh.levels[level].nv;
    H.owned.assign(owned, owned + nv);
    for (int v : H.idx) AB_REQUIRE(v >= 0 && v < nv, AB_ERR_ARG, "interface vertex index out of range");
    dom->dist_enabled = true;
    AB_CATCH
}
int ab_domain_set_gather(ab_domain* dom, int gather_level, ab_domain* coarse, const int32_t* nv_per_rank, const int32_t* l2g_cat,
                         const int64_t* nblk_per_rank, const int32_t* gpos_cat) {
    AB_TRY
    AB_REQUIRE(dom && !dom->finalized, AB_ERR_STATE, "must be set before the first ApproximationSpace");
    AB_REQUIRE(dom->ctx && dom->ctx->comm && dom->ctx->comm->nranks > 1, AB_ERR_STATE, "ab_domain_set_gather needs a multi-rank context");
    AB_REQUIRE(gather_level >= 0 && gather_level < (int)dom->mesh.levels.size() - 1, AB_ERR_ARG, "the gather level must lie below the top level");
    dom->gather_level = gather_level;
    Context* ctx = dom->ctx;
    if (ctx->comm->rank != 0) return AB_OK;
    const int nr = ctx->comm->nranks, D = dom->dim();
    AB_REQUIRE(coarse && nv_per_rank && l2g_cat && nblk_per_rank && gpos_cat, AB_ERR_ARG, "rank 0 needs the coarse domain and the gather maps");
    AB_REQUIRE(coarse->ctx == ctx && (int)coarse->mesh.levels.size() == gather_level + 1 && coarse->dim() == D, AB_ERR_ARG,
               "the coarse domain must be the global grid refined gather_level times on the same context");
    dom->cdom = coarse;
    dom->g_nv.assign(nv_per_rank, nv_per_rank + nr);
    dom->g_nblk.assign(nblk_per_rank, nblk_per_rank + nr);
    AB_REQUIRE(dom->g_nv[0] == dom->mesh.levels[gather_level].nv, AB_ERR_ARG, "gather maps: rank 0's own vertex count does not match");
    int64_t nvt = 0, nbt = 0;
    for (int r = 0; r < nr; ++r) { nvt += dom->g_nv[r]; nbt += dom->g_nblk[r]; }
    dom->g_nv_total = nvt; dom->g_nblk_total = nbt;
    const int nvg = coarse->mesh.levels[gather_level].nv;
    std::vector<int> ptr((size_t)nvg + 1, 0);
    for (int64_t k = 0; k < nvt; ++k) {
        AB_REQUIRE(l2g_cat[k] >= 0 && l2g_cat[k] < nvg, AB_ERR_ARG, "gather maps: global vertex id out of range");
        ptr[l2g_cat[k] + 1]++;
    }
    for (int v = 0; v < nvg; ++v) {
        AB_REQUIRE(ptr[v + 1] > 0, AB_ERR_ARG, "gather maps: a global coarse vertex belongs to no rank");
        ptr[v + 1] += ptr[v];
    }
    std::vector<int> fill(ptr.begin(), ptr.end() - 1), idx((size_t)nvt);
    for (int64_t k = 0; k < nvt; ++k) idx[fill[l2g_cat[k]]++] = (int)k;     // k ascends rank by rank: every list is in rank order
    dom->g_l2g.upload(std::vector<int>(l2g_cat, l2g_cat + nvt), ctx->stream);
    dom->g_csr_ptr.upload(ptr, ctx->stream);
    dom->g_csr_idx.upload(idx, ctx->stream);
    dom->g_gpos.upload(std::vector<int>(gpos_cat, gpos_cat + nbt), ctx->stream);
    dom->g_vstage.alloc((size_t)nvt * D);
    dom->g_mstage.alloc((size_t)nbt * D * D);
    AB_CATCH
}
int ab_domain_set_block_interface(ab_domain* dom, int level, int nneigh, const int32_t* neigh_ranks, const int32_t* offsets,
                                  const int32_t* slot_block, int nshared, const int32_t* bpos, const int32_t* brow, const int32_t* mult) {
    AB_TRY
    AB_REQUIRE(dom && !dom->finalized, AB_ERR_STATE, "block interfaces must be set before the first ApproximationSpace");
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    AB_REQUIRE(nneigh >= 0 && nshared >= 0 && offsets, AB_ERR_ARG, "bad block interface");
    if (dom->biface.size() < dom->mesh.levels.size()) dom->biface.resize(dom->mesh.levels.size());
    Domain::BlockIface& B = dom->biface[level];
    B.neigh.assign(neigh_ranks, neigh_ranks + nneigh);
    B.offset.assign(offsets, offsets + nneigh + 1);
    B.total = nneigh ? offsets[nneigh] : 0;
    B.nsb = nshared;
    B.h_idx.assign(slot_block, slot_block + B.total);
    B.h_bpos.assign(bpos, bpos + nshared);
    B.h_brow.assign(brow, brow + nshared);
    B.h_mult.assign(mult, mult + nshared);
    const int nv = dom->mesh.levels[level].nv;
    for (int k : B.h_idx) AB_REQUIRE(k >= 0 && k < nshared, AB_ERR_ARG, "block interface: slot refers to no shared block");
    for (int k = 0; k < nshared; ++k) AB_REQUIRE(B.h_bpos[k] >= 0 && B.h_brow[k] >= 0 && B.h_brow[k] < nv && B.h_mult[k] >= 2, AB_ERR_ARG, "block interface: bad shared block");
    AB_CATCH
}
int ab_domain_level_pattern(ab_domain* dom, int level, int64_t* nnzb, int32_t* rowptr, int32_t* colidx) {
    AB_TRY
    AB_REQUIRE(dom && level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    HostPattern P;
    build_pattern(dom->mesh.levels[level], P);
    if (nnzb) *nnzb = (int64_t)P.colidx.size();
    if (rowptr) memcpy(rowptr, P.rowptr.data(), P.rowptr.size() * sizeof(int32_t));
    if (colidx) memcpy(colidx, P.colidx.data(), P.colidx.size() * sizeof(int32_t));
    AB_CATCH
}
int ab_domain_level_incidence(ab_domain* dom, int level, int32_t* ptr, int32_t* idx) {
    AB_TRY
    AB_REQUIRE(dom && ptr && idx && level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    std::vector<int32_t> vp, vi;
    build_v2e(dom->mesh.levels[level], vp, vi);
    memcpy(ptr, vp.data(), vp.size() * sizeof(int32_t));
    memcpy(idx, vi.data(), vi.size() * sizeof(int32_t));
    AB_CATCH
}
int ab_domain_p2p_export(ab_domain* dom, void* handle64, int64_t* level_base /* nlevels */, int32_t* totals /* nlevels */) {
    AB_TRY
    AB_REQUIRE(dom->finalized && dom->distributed(), AB_ERR_STATE, "p2p window needs a finalised multi-GPU domain");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const int nl = (int)dom->iface.size(), nr = dom->ctx->comm->nranks, D = dom->dim();
    if (!dom->window) {
        size_t off = (size_t)nl * nr * sizeof(unsigned long long);
        off = (off + 255) & ~(size_t)255;
        dom->win_level_base.resize(nl);
        for (int l = 0; l < nl; ++l) {
            dom->win_level_base[l] = off;
            off += ((size_t)2 * dom->iface[l].total * D * sizeof(double) + 255) & ~(size_t)255;
        }
        dom->window_bytes = std::max<size_t>(off, 256);
        AB_CUDA(cudaMalloc((void**)&dom->window, dom->window_bytes));      // a plain allocation: exported whole through CUDA IPC
        AB_CUDA(cudaMemset(dom->window, 0, dom->window_bytes));
        dom->p2p_err.alloc(1); dom->p2p_err.zero(dom->ctx->stream);
        AB_CUDA(cudaStreamSynchronize(dom->ctx->stream));
    }
    cudaIpcMemHandle_t h;
    AB_CUDA(cudaIpcGetMemHandle(&h, dom->window));
    memcpy(handle64, &h, 64);
    for (int l = 0; l < nl; ++l) { level_base[l] = (int64_t)dom->win_level_base[l]; totals[l] = dom->iface[l].total; }
    AB_CATCH
}
// handles: nranks x 64 bytes; per level l and neighbour slot k (concatenated over levels in neighbour order):
// remote_dst[.] = byte offset inside the neighbour's window of this rank's slot (parity 0), remote_stride[.] = the neighbour's
// parity stride in bytes
int ab_domain_p2p_connect(ab_domain* dom, const void* handles, const int64_t* remote_dst, const int64_t* remote_stride) {
    AB_TRY
    AB_REQUIRE(dom->window && !dom->p2p_connected, AB_ERR_STATE, "ab_domain_p2p_export first");
    Context* ctx = dom->ctx;
    const int nl = (int)dom->iface.size(), nr = ctx->comm->nranks, me = ctx->comm->rank, D = dom->dim();
    dom->peer_windows.assign(nr, nullptr);
    std::vector<char> used(nr, 0);
    for (int l = 0; l < nl; ++l) for (int q : dom->iface[l].neigh) used[q] = 1;
    for (int q = 0; q < nr; ++q) {
        if (!used[q] || q == me) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + (size_t)q * 64, 64);
        AB_CUDA(cudaIpcOpenMemHandle(&dom->peer_windows[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    size_t pos = 0;
    for (int l = 0; l < nl; ++l) {
        Interface& I = dom->iface[l];
        const size_t nn = I.neigh.size();
        std::vector<unsigned long long> dst(nn), stride(nn), flag(nn);
        for (size_t k = 0; k < nn; ++k, ++pos) {
            const int q = I.neigh[k];
            unsigned char* base = (unsigned char*)dom->peer_windows[q];
            dst[k] = (unsigned long long)(uintptr_t)(base + remote_dst[pos]);
            stride[k] = (unsigned long long)remote_stride[pos];
            flag[k] = (unsigned long long)(uintptr_t)(base + ((size_t)l * nr + me) * sizeof(unsigned long long));
        }
        if (nn) {
            I.d_peer_dst.upload(dst, ctx->stream); I.d_peer_stride.upload(stride, ctx->stream); I.d_peer_flag.upload(flag, ctx->stream);
            I.d_offset.upload(I.offset, ctx->stream); I.d_neigh.upload(I.neigh, ctx->stream);
        }
        I.win_recv = (double*)(dom->window + dom->win_level_base[l]);
        I.win_flags = (unsigned long long*)dom->window + (size_t)l * nr;
        (void)D;
    }
    dom->p2p_connected = true;
    AB_CATCH
}
int ab_domain_p2p_status(ab_domain* dom, int* connected, int* error) {
    AB_TRY
    if (connected) *connected = dom->p2p_connected ? 1 : 0;
    if (error) {
        *error = 0;
        if (dom->p2p_err.p) { AB_CUDA(cudaStreamSynchronize(dom->ctx->stream)); AB_CUDA(cudaMemcpy(error, dom->p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost)); }
    }
    AB_CATCH
}
int ab_domain_num_levels(ab_domain* dom, int* out) {
    AB_TRY
    *out = (int)dom->mesh.levels.size();
    AB_CATCH
}
int ab_domain_level_info(ab_domain* dom, int level, int* dim, int* nv, int* ne, int* nedges, int* nv_coarse) {
    AB_TRY
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    HostLevel& L = dom->mesh.levels[level];
    if (dim) *dim = L.dim;
    if (nv) *nv = L.nv;
    if (ne) *ne = L.ne;
    if (nv_coarse) *nv_coarse = L.nv_coarse;
    if (nedges) {
        if (dom->finalized) *nedges = (int)((dom->dev[level].nnzb - L.nv) / 2);
        else { ensure_edges(L); *nedges = (int)L.nedges(); }
    }
    AB_CATCH
}
int ab_domain_get_level(ab_domain* dom, int level, double* xyz, int32_t* elems, int32_t* vsub, int32_t* parent_a, int32_t* parent_b) {
    AB_TRY
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    HostLevel& L = dom->mesh.levels[level];
    if (xyz) {
        if (dom->host_xyz_stale) {
            HostLevel& T = dom->mesh.levels.back();
            AB_CUDA(cudaMemcpy(T.xyz.data(), dom->dev[dom->top()].xyz.p, T.xyz.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int l = dom->top() - 1; l >= 0; --l) {
                HostLevel& H = dom->mesh.levels[l];
                std::copy(T.xyz.begin(), T.xyz.begin() + H.xyz.size(), H.xyz.begin());
            }
            dom->host_xyz_stale = false;
        }
        memcpy(xyz, L.xyz.data(), L.xyz.size() * sizeof(double));
    }
    if (elems) memcpy(elems, L.elems.data(), L.elems.size() * sizeof(int32_t));
    if (vsub) memcpy(vsub, L.vsub.data(), L.vsub.size() * sizeof(int32_t));
    if (parent_a && !L.pa.empty()) memcpy(parent_a, L.pa.data(), L.pa.size() * sizeof(int32_t));
    if (parent_b && !L.pb.empty()) memcpy(parent_b, L.pb.data(), L.pb.size() * sizeof(int32_t));
    AB_CATCH
}
int ab_domain_special_info(ab_domain* dom, int level, int* n_sp_edges, int* n_sp_faces, int* nsubsets) {
    AB_TRY
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    const HostLevel& L = dom->mesh.levels[level];
    if (n_sp_edges) *n_sp_edges = (int)L.sp_edges_sub.size();
    if (n_sp_faces) *n_sp_faces = (int)L.sp_faces_sub.size();
    if (nsubsets) *nsubsets = (int)dom->mesh.subset_names.size();
    AB_CATCH
}
int ab_domain_get_special(ab_domain* dom, int level, int32_t* sp_edges, int32_t* sp_edges_sub, int32_t* sp_faces, int32_t* sp_faces_sub, int32_t* esub) {
    AB_TRY
    AB_REQUIRE(level >= 0 && level < (int)dom->mesh.levels.size(), AB_ERR_ARG, "level out of range");
    const HostLevel& L = dom->mesh.levels[level];
    if (sp_edges && !L.sp_edges.empty()) memcpy(sp_edges, L.sp_edges.data(), L.sp_edges.size() * sizeof(int32_t));
    if (sp_edges_sub && !L.sp_edges_sub.empty()) memcpy(sp_edges_sub, L.sp_edges_sub.data(), L.sp_edges_sub.size() * sizeof(int32_t));
    if (sp_faces && !L.sp_faces.empty()) memcpy(sp_faces, L.sp_faces.data(), L.sp_faces.size() * sizeof(int32_t));
    if (sp_faces_sub && !L.sp_faces_sub.empty()) memcpy(sp_faces_sub, L.sp_faces_sub.data(), L.sp_faces_sub.size() * sizeof(int32_t));
    if (esub) memcpy(esub, L.esub.data(), L.esub.size() * sizeof(int32_t));
    AB_CATCH
}
int ab_domain_subset_name(ab_domain* dom, int index, char* buf, int buflen) {
    AB_TRY
    AB_REQUIRE(index >= 0 && index < (int)dom->mesh.subset_names.size() && buf && buflen > 0, AB_ERR_ARG, "subset index out of range");
    snprintf(buf, buflen, "%s", dom->mesh.subset_names[index].c_str());
    AB_CATCH
}
int ab_domain_subset_index(ab_domain* dom, const char* name, int* out) {
    AB_TRY
    const int i = dom->mesh.subset_index(name);
    AB_REQUIRE(i >= 0, AB_ERR_ARG, std::string("unknown subset '") + name + "'");
    *out = i;
    AB_CATCH
}
int ab_transform_domain_by_displacement(ab_domain* dom, ab_vector* u) {
    AB_TRY
    AB_REQUIRE(dom->finalized, AB_ERR_STATE, "domain has no ApproximationSpace yet");
    AB_REQUIRE(u->sp->kind == AB_SPACE_P1 && u->sp->ncomp == dom->dim() && u->sp->dom == dom, AB_ERR_ARG, "displacement must be a P1 dim-vector on this domain");
    LevelDev& L = dom->dev[dom->top()];
    if (dom->distributed() && !(u->storage & AB_PST_CONSISTENT)) {
        // every rank must move its copy of a shared vertex by the same (summed) displacement
        DevBuf<double> tmp((size_t)u->n());
        dev_copy(dom->ctx, u->n(), u->d.p, tmp.p);
        exchange_sum(dom, dom->top(), tmp.p, dom->dim());
        dev_axpby(dom->ctx, u->n(), 1.0, L.xyz.p, 1.0, tmp.p, L.xyz.p);
        AB_CUDA(cudaStreamSynchronize(dom->ctx->stream));   // tmp is released below
    } else {
        dev_axpby(dom->ctx, u->n(), 1.0, L.xyz.p, 1.0, u->d.p, L.xyz.p);
    }
    dom->coords_version = ++g_version_counter;
    dom->host_xyz_stale = true;
    AB_CHECK_LAUNCH(dom->ctx);
    AB_CATCH
}

// ---- space ----------------------------------------------------------------------------------
int ab_space_create(ab_domain* dom, int kind, int ncomp, ab_space** out) {
    AB_TRY
    AB_REQUIRE(dom && out, AB_ERR_ARG, "NULL argument");
    AB_REQUIRE(kind == AB_SPACE_P0 || kind == AB_SPACE_P1, AB_ERR_ARG, "unknown space kind");
    AB_REQUIRE(ncomp >= 1 && ncomp <= 9, AB_ERR_ARG, "ncomp out of range");
    AB_REQUIRE(dom->ctx, AB_ERR_STATE, "domain was created without a context (host-only); ApproximationSpace needs a GPU context");
    dom->finalize();
    auto* s = new ab_space();
    s->dom = dom; s->kind = kind; s->ncomp = ncomp;
    const HostLevel& T = dom->mesh.levels.back();
    s->ndofs = (int64_t)(kind == AB_SPACE_P1 ? T.nv : T.ne) * ncomp;
    *out = s;
    AB_CATCH
}
int ab_space_destroy(ab_space* sp) {
    AB_TRY
    delete sp;
    AB_CATCH
}
int ab_space_num_dofs(ab_space* sp, int64_t* out) {
    AB_TRY
    *out = sp->ndofs;
    AB_CATCH
}

// ---- vectors --------------------------------------------------------------------------------
int ab_vector_create(ab_space* sp, ab_vector** out) {
    AB_TRY
    auto v = std::make_unique<ab_vector>();
    v->sp = sp;
    v->d.alloc((size_t)sp->ndofs);
    v->d.zero(sp->dom->ctx->stream);
    v->id = ++g_version_counter;
    v->touch();
    *out = v.release();
    AB_CATCH
}
int ab_vector_destroy(ab_vector* v) {
    AB_TRY
    if (v)
        for (auto& kv : g_dot_cache) kv.second.forget(v);   // no dereference of the space / domain: they may already be gone
    delete v;
    AB_CATCH
}
int ab_vector_set(ab_vector* v, double c) {
    AB_TRY
    dev_fill(v->sp->dom->ctx, v->d.p, v->n(), c);
    v->storage = AB_PST_CONSISTENT;
    v->touch();
    AB_CHECK_LAUNCH(v->sp->dom->ctx);
    AB_CATCH
}
int ab_vector_upload(ab_vector* v, const double* host, int storage) {
    AB_TRY
    Context* ctx = v->sp->dom->ctx;
    AB_CUDA(cudaMemcpyAsync(v->d.p, host, v->n() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    v->storage = storage ? storage : AB_PST_CONSISTENT;
    v->touch();
    AB_CATCH
}
int ab_vector_download(ab_vector* v, double* host) {
    AB_TRY
    Context* ctx = v->sp->dom->ctx;
    AB_CUDA(cudaMemcpyAsync(host, v->d.p, v->n() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    AB_CATCH
}
int ab_vector_device_ptr(ab_vector* v, void** out, int64_t* n) {
    AB_TRY
    *out = v->d.p;
    if (n) *n = v->n();
    v->aliased = true;
    v->touch();   // the caller may write through the pointer
    AB_CATCH
}
int ab_vector_storage(ab_vector* v, int* out) {
    AB_TRY
    *out = v->storage;
    AB_CATCH
}
int ab_vector_change_storage(ab_vector* v, int storage) {
    AB_TRY
    // single GPU: additive, consistent and unique representations coincide; only the flag changes.
    // multi-GPU (P1): additive -> consistent is the interface sum; consistent -> additive/unique keeps the owner's copy.
    Domain* dom = v->sp->dom;
    if (dom->distributed() && v->sp->kind == AB_SPACE_P1 && storage != v->storage) {
        AB_REQUIRE(v->sp->ncomp == dom->dim(), AB_ERR_UNSUPPORTED, "storage conversion: P1 spaces with ncomp != dim");
        const int top = dom->top();
        if ((v->storage & AB_PST_ADDITIVE) && (storage & AB_PST_CONSISTENT)) exchange_sum(dom, top, v->d.p, v->sp->ncomp);
        else if ((v->storage & AB_PST_CONSISTENT) && !(storage & AB_PST_CONSISTENT))
            AB_LAUNCH(dom->ctx, k_zero_not_owned, ew_grid(dom->ctx, v->n()), 256, 0, v->n(), v->sp->ncomp, dom->iface[top].owned.p, v->d.p);
        v->touch();
    }
    v->storage = storage;
    AB_CATCH
}
int ab_vec_scale_assign(ab_vector* dst, double a, ab_vector* src) {
    AB_TRY
    AB_REQUIRE(dst->n() == src->n(), AB_ERR_ARG, "VecScaleAssign: size mismatch");
    dev_axpby(dst->sp->dom->ctx, dst->n(), a, src->d.p, 0.0, nullptr, dst->d.p);
    dst->storage = src->storage;
    dst->touch();
    AB_CHECK_LAUNCH(dst->sp->dom->ctx);
    AB_CATCH
}
int ab_vec_scale_add2(ab_vector* dst, double a, ab_vector* x, double b, ab_vector* y) {
    AB_TRY
    AB_REQUIRE(dst->n() == x->n() && dst->n() == y->n(), AB_ERR_ARG, "VecScaleAdd2: size mismatch");
    dev_axpby(dst->sp->dom->ctx, dst->n(), a, x->d.p, b, y->d.p, dst->d.p);
    dst->storage = (x->storage == y->storage) ? x->storage : (x->storage & y->storage ? (x->storage & y->storage) : x->storage);
    dst->touch();
    AB_CHECK_LAUNCH(dst->sp->dom->ctx);
    AB_CATCH
}
int ab_vec_prod_multi(int n, ab_vector* const* xs, ab_vector* y, double* out) {
    AB_TRY
    AB_REQUIRE(n >= 1 && n <= 4, AB_ERR_ARG, "ab_vec_prod_multi: n must be 1..4");
    Domain* dom = y->sp->dom;
    Context* ctx = dom->ctx;
    const double* p[4];
    for (int i = 0; i < n; ++i) { AB_REQUIRE(xs[i]->n() == y->n(), AB_ERR_ARG, "VecProd: size mismatch"); p[i] = xs[i]->d.p; }
    if (!dom->distributed() || y->sp->kind == AB_SPACE_P0) {
        dev_dots(ctx, y->n(), n, p, y->d.p, ctx->d_results);
    } else {
        // <additive, consistent> is the plain local sum (3d_admm.lua:991-992); <consistent, consistent> counts every
        // shared vertex once (owner mask); <additive, additive> first makes a consistent copy of y
        const int top = dom->top(), D = y->sp->ncomp;
        const bool y_add = (y->storage & AB_PST_ADDITIVE) != 0;
        const double* ycons = nullptr;
        DevBuf<double> tmp;
        for (int i = 0; i < n; ++i) {
            const bool x_add = (xs[i]->storage & AB_PST_ADDITIVE) != 0;
            if (x_add != y_add) {
                const double* one[1] = {p[i]};
                dev_dots(ctx, y->n(), 1, one, y->d.p, ctx->d_results + i);
            } else if (!x_add) {
                AB_LAUNCH(ctx, (k_dot_owned<1>), red_grid(ctx, y->n()), 256, 0, y->n(), D, dom->iface[top].owned.p, p[i], p[i], y->d.p, ctx->d_partials,
                          ctx->d_tickets, ctx->d_results + i);
            } else {
                if (!ycons) {
                    tmp.alloc((size_t)y->n());
                    dev_copy(ctx, y->n(), y->d.p, tmp.p);
                    exchange_sum(dom, top, tmp.p, D);
                    ycons = tmp.p;
                }
                const double* one[1] = {p[i]};
                dev_dots(ctx, y->n(), 1, one, ycons, ctx->d_results + i);
            }
        }
        AB_CUDA(cudaStreamSynchronize(ctx->stream));   // tmp is released below
    }
    allreduce_dev(dom, ctx->d_results, n);
    read_back(ctx, ctx->d_results, n, out);
    AB_CATCH
}
int ab_vec_prod(ab_vector* x, ab_vector* y, double* out) {
    Domain* dom = y->sp->dom;
    Context* ctx = dom->ctx;
    const bool cacheable = ctx->use_cache && !env_flag("ADMM_B200_NO_CACHE") && !dom->distributed() && !x->aliased && !y->aliased && x->n() == y->n();
    if (!cacheable) {
        ab_vector* xs[1] = {x};
        return ab_vec_prod_multi(1, xs, y, out);
    }
    DotCache& dc = g_dot_cache[ctx];
    if (dc.lookup(x, y, out)) { dc.used(x); return AB_OK; }
    ab_vector* xs[4] = {x, nullptr, nullptr, nullptr};
    int nx = 1;
    for (Vector* c : dc.recent_x)
        if (nx < 4 && c != x && c != y && !c->aliased && c->sp == x->sp) xs[nx++] = static_cast<ab_vector*>(c);
    double vals[4];
    const int rc = ab_vec_prod_multi(nx, xs, y, vals);
    if (rc != AB_OK) return rc;
    for (int i = 0; i < nx; ++i) dc.store(xs[i], y, vals[i]);
    dc.used(x);
    *out = vals[0];
    return AB_OK;
}
int ab_vec_norm(ab_vector* x, double* out) {
    double v = 0;
    int rc = ab_vec_prod(x, x, &v);
    if (rc == AB_OK) *out = std::sqrt(v);
    return rc;
}
int ab_l2norm_all(ab_vector* v, double* out) {
    AB_TRY
    Domain* dom = v->sp->dom;
    Context* ctx = dom->ctx;
    const LevelDev& L = dom->dev[dom->top()];
    const int dim = dom->dim();
    const int g = red_grid(ctx, L.ne);
    int nc;
    DevBuf<double> tmp;
    if (v->sp->kind == AB_SPACE_P1) {
        AB_REQUIRE(v->sp->ncomp == dim, AB_ERR_UNSUPPORTED, "L2Norm: P1 spaces with ncomp != dim are not supported");
        nc = dim;
        const double* vp = v->d.p;
        if (dom->distributed() && (v->storage & AB_PST_ADDITIVE)) {   // the FE function is the consistent representation
            tmp.alloc((size_t)v->n());
            dev_copy(ctx, v->n(), v->d.p, tmp.p);
            exchange_sum(dom, dom->top(), tmp.p, dim);
            vp = tmp.p;
        }
        if (dim == 2) AB_LAUNCH(ctx, (k_l2norm_p1<2>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, vp, ctx->d_partials, ctx->d_tickets, ctx->d_results);
        else AB_LAUNCH(ctx, (k_l2norm_p1<3>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, vp, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    } else {
        AB_REQUIRE(v->sp->ncomp == dim * dim, AB_ERR_UNSUPPORTED, "L2Norm: P0 spaces with ncomp != dim*dim are not supported");
        nc = dim * dim;
        if (dim == 2) AB_LAUNCH(ctx, (k_l2norm_p0<2>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, v->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
        else AB_LAUNCH(ctx, (k_l2norm_p0<3>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, v->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    }
    allreduce_dev(dom, ctx->d_results, nc);
    read_back(ctx, ctx->d_results, nc, out);
    for (int i = 0; i < nc; ++i) out[i] = std::sqrt(out[i]);
    AB_CATCH
}
int ab_l2norm(ab_vector* v, int comp, double* out) {
    if (comp < 0 || comp >= v->sp->ncomp || comp >= 9) { g_last_error = "L2Norm: component out of range"; return AB_ERR_ARG; }
    Domain* dom = v->sp->dom;
    const bool cacheable = dom->ctx->use_cache && !env_flag("ADMM_B200_NO_CACHE") && !v->aliased;
    if (cacheable && v->l2_version == v->version && v->l2_coords == dom->coords_version && v->l2_storage == v->storage) {
        *out = v->l2_vals[comp];
        return AB_OK;
    }
    double tmp[9];
    int rc = ab_l2norm_all(v, tmp);
    if (rc != AB_OK) return rc;
    for (int i = 0; i < v->sp->ncomp && i < 9; ++i) v->l2_vals[i] = tmp[i];
    v->l2_version = v->version; v->l2_coords = dom->coords_version; v->l2_storage = v->storage;
    *out = tmp[comp];
    return AB_OK;
}

// ---- elem discs -----------------------------------------------------------------------------
int ab_elemdisc_create(ab_space* sp, int kind, ab_elemdisc** out) {
    AB_TRY
    AB_REQUIRE(kind >= AB_DISC_DEFORMATION_EQUATION && kind <= AB_DISC_LAMBDA_UPDATE, AB_ERR_ARG, "unknown ElemDisc kind");
    auto* d = new ab_elemdisc();
    d->sp = sp; d->kind = kind;
    for (double& p : d->params) p = 0.0;
    d->params[AB_PARAM_STEP_LENGTH] = 1.0;
    d->params[AB_PARAM_TAU] = 1.0;
    d->params[AB_PARAM_INDEX] = 1.0;
    d->params[AB_PARAM_QUAD_ORDER] = 1.0;
    *out = d;
    AB_CATCH
}
int ab_elemdisc_destroy(ab_elemdisc* d) {
    AB_TRY
    delete d;
    AB_CATCH
}
int ab_elemdisc_set_param(ab_elemdisc* d, int param, double value) {
    AB_TRY
    AB_REQUIRE(param >= 1 && param <= 15, AB_ERR_ARG, "unknown ElemDisc parameter");
    d->params[param] = value;
    AB_CATCH
}
int ab_elemdisc_get_param(ab_elemdisc* d, int param, double* value) {
    AB_TRY
    AB_REQUIRE(param >= 1 && param <= 15, AB_ERR_ARG, "unknown ElemDisc parameter");
    *value = d->params[param];
    AB_CATCH
}
int ab_elemdisc_bind(ab_elemdisc* d, int import, ab_vector* v) {
    AB_TRY
    const int dim = d->sp->dom->dim();
    if (import == AB_IMPORT_DEFORMATION) {
        AB_REQUIRE(!v || (v->sp->kind == AB_SPACE_P1 && v->sp->ncomp == dim), AB_ERR_ARG, "deformation import must be a P1 dim-vector");
        d->imp_u = v;
    } else if (import == AB_IMPORT_LAMBDA || import == AB_IMPORT_Q) {
        AB_REQUIRE(!v || (v->sp->kind == AB_SPACE_P0 && v->sp->ncomp == dim * dim), AB_ERR_ARG, "tensor import must be a P0 dim*dim function");
        (import == AB_IMPORT_LAMBDA ? d->imp_lam : d->imp_q) = v;
    } else AB_REQUIRE(false, AB_ERR_ARG, "unknown import");
    AB_CATCH
}

// ---- domain disc ----------------------------------------------------------------------------
int ab_domaindisc_create(ab_space* sp, ab_domaindisc** out) {
    AB_TRY
    auto* dd = new ab_domaindisc();
    dd->sp = sp;
    *out = dd;
    AB_CATCH
}
int ab_domaindisc_destroy(ab_domaindisc* dd) {
    AB_TRY
    delete dd;
    AB_CATCH
}
int ab_domaindisc_add_elemdisc(ab_domaindisc* dd, ab_elemdisc* d) {
    AB_TRY
    AB_REQUIRE(d->sp->dom == dd->sp->dom, AB_ERR_ARG, "ElemDisc and DomainDiscretization live on different domains");
    dd->discs.push_back(d);
    AB_CATCH
}
int ab_domaindisc_add_dirichlet(ab_domaindisc* dd, const char* subset, int comp, double value) {
    AB_TRY
    AB_REQUIRE(value == 0.0, AB_ERR_UNSUPPORTED, "only homogeneous Dirichlet values are used on the hot path (3d_admm.lua:447-457)");
    AB_REQUIRE(dd->sp->kind == AB_SPACE_P1 && comp >= 0 && comp < dd->sp->ncomp, AB_ERR_ARG, "Dirichlet component out of range");
    const int si = dd->sp->dom->mesh.subset_index(subset);
    AB_REQUIRE(si >= 0, AB_ERR_ARG, std::string("unknown subset '") + subset + "'");
    dd->dir.push_back({si, comp});
    dd->masks_valid = false;
    AB_CATCH
}
int ab_domaindisc_assemble_jacobian(ab_domaindisc* dd, ab_operator* A, ab_vector* u) {
    AB_TRY
    assemble_jacobian(dd, A, u);
    AB_CHECK_LAUNCH(dd->sp->dom->ctx);
    AB_CATCH
}
int ab_domaindisc_assemble_defect(ab_domaindisc* dd, ab_vector* d, ab_vector* u) {
    AB_TRY
    assemble_defect(dd, d, u);
    AB_CHECK_LAUNCH(dd->sp->dom->ctx);
    AB_CATCH
}
int ab_domaindisc_adjust_solution(ab_domaindisc* dd, ab_vector* u) {
    AB_TRY
    Domain* dom = dd->sp->dom;
    const unsigned char* mask = dd->mask(dom->top());
    if (mask) {
        const LevelDev& L = dom->dev[dom->top()];
        if (dom->dim() == 2) AB_LAUNCH(dom->ctx, (k_zero_dirichlet<2>), ew_grid(dom->ctx, u->n()), 256, 0, L.nv, mask, u->d.p);
        else AB_LAUNCH(dom->ctx, (k_zero_dirichlet<3>), ew_grid(dom->ctx, u->n()), 256, 0, L.nv, mask, u->d.p);
        u->touch();
    }
    AB_CHECK_LAUNCH(dom->ctx);
    AB_CATCH
}

// ---- operator -------------------------------------------------------------------------------
int ab_operator_create(ab_domaindisc* dd, ab_operator** out) {
    AB_TRY
    auto* A = new ab_operator();
    A->dd = dd;
    *out = A;
    AB_CATCH
}
int ab_operator_destroy(ab_operator* A) {
    AB_TRY
    delete A;
    AB_CATCH
}
int ab_operator_apply(ab_operator* A, ab_vector* y, ab_vector* x) {
    AB_TRY
    AB_REQUIRE(A->data && A->data->assembled, AB_ERR_STATE, "operator not assembled");
    Domain* dom = A->data->dom;
    Context* ctx = dom->ctx;
    if (A->data->kind == 1) spmv(ctx, dom->dim(), dom->dev[dom->top()], A->data->vals.p, 0, 0, x->d.p, nullptr, y->d.p);
    else AB_LAUNCH(ctx, k_diag_apply_dot, red_grid(ctx, x->n()), 256, 0, x->n(), A->data->vals.p, x->d.p, y->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    y->storage = AB_PST_ADDITIVE;
    y->touch();
    AB_CHECK_LAUNCH(ctx);
    AB_CATCH
}
int ab_operator_info(ab_operator* A, int* block, int64_t* nb, int64_t* nnzb) {
    AB_TRY
    AB_REQUIRE(A->data, AB_ERR_STATE, "operator not assembled");
    Domain* dom = A->data->dom;
    if (A->data->kind == 1) {
        if (block) *block = dom->dim();
        if (nb) *nb = dom->dev[dom->top()].nv;
        if (nnzb) *nnzb = dom->dev[dom->top()].nnzb;
    } else {
        if (block) *block = 1;
        if (nb) *nb = (int64_t)A->data->vals.n;
        if (nnzb) *nnzb = (int64_t)A->data->vals.n;
    }
    AB_CATCH
}
int ab_operator_download(ab_operator* A, int32_t* rowptr, int32_t* colidx, double* vals) {
    AB_TRY
    AB_REQUIRE(A->data && A->data->assembled, AB_ERR_STATE, "operator not assembled");
    Domain* dom = A->data->dom;
    AB_CUDA(cudaStreamSynchronize(dom->ctx->stream));
    if (A->data->kind == 1) {
        const LevelDev& L = dom->dev[dom->top()];
        if (rowptr) AB_CUDA(cudaMemcpy(rowptr, L.rowptr.p, ((size_t)L.nv + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        if (colidx) AB_CUDA(cudaMemcpy(colidx, L.colidx.p, (size_t)L.nnzb * sizeof(int), cudaMemcpyDeviceToHost));
    } else {
        const int64_t n = (int64_t)A->data->vals.n;
        if (rowptr) for (int64_t i = 0; i <= n; ++i) rowptr[i] = (int32_t)i;
        if (colidx) for (int64_t i = 0; i < n; ++i) colidx[i] = (int32_t)i;
    }
    if (vals) AB_CUDA(cudaMemcpy(vals, A->data->vals.p, A->data->vals.n * sizeof(double), cudaMemcpyDeviceToHost));
    AB_CATCH
}

// ---- solvers --------------------------------------------------------------------------------
int ab_solver_create_bicgstab_gmg(ab_space* sp, const ab_gmg_desc* desc, ab_solver** out) {
    AB_TRY
    AB_REQUIRE(sp && desc && out, AB_ERR_ARG, "NULL argument");
    AB_REQUIRE(sp->kind == AB_SPACE_P1 && sp->ncomp == sp->dom->dim(), AB_ERR_ARG, "BiCGStab+GMG needs the P1 deformation space");
    AB_REQUIRE(desc->rap != 0, AB_ERR_UNSUPPORTED, "rap=false (re-discretised coarse operators) is not implemented; the reference uses rap=true (u3:27)");
    AB_REQUIRE(desc->base_level == 0, AB_ERR_UNSUPPORTED, "baseLevel must be 0 (u3:19)");
    AB_REQUIRE(desc->smoother == AB_SMOOTHER_CHEBYSHEV || desc->smoother == AB_SMOOTHER_JACOBI, AB_ERR_ARG, "unknown smoother");
    auto* s = new ab_solver();
    s->sp = sp; s->type = 1; s->desc = *desc;
    if (s->desc.cheb_ratio <= 1.0) s->desc.cheb_ratio = 6.0;
    if (s->desc.jacobi_damp <= 0.0) s->desc.jacobi_damp = 0.66;
    *out = s;
    AB_CATCH
}
int ab_solver_create_cg_jacobi(ab_space* sp, double damp, int max_iterations, double abs_tol, double red_tol, int verbose, ab_solver** out) {
    AB_TRY
    auto* s = new ab_solver();
    s->sp = sp; s->type = 2; s->damp = damp;
    s->desc.max_iterations = max_iterations; s->desc.abs_tol = abs_tol; s->desc.red_tol = red_tol; s->desc.verbose = verbose;
    *out = s;
    AB_CATCH
}
int ab_solver_destroy(ab_solver* s) {
    AB_TRY
    delete s;
    AB_CATCH
}
int ab_solver_init(ab_solver* s, ab_operator* A, ab_vector*) {
    AB_TRY
    solver_init(s, A);
    AB_CHECK_LAUNCH(s->sp->dom->ctx);
    AB_CATCH
}
static int solver_apply_impl(ab_solver* s, ab_vector* x, ab_vector* b, int* converged, bool ret_def) {
    AB_TRY
    AB_REQUIRE(s->A, AB_ERR_STATE, "solver:apply before solver:init");
    AB_REQUIRE(x->n() == s->sp->ndofs && b->n() == s->sp->ndofs, AB_ERR_ARG, "solver:apply: vector size mismatch");
    const bool ok = s->type == 1 ? bicgstab_apply(s, x, b, ret_def) : cg_jacobi_apply(s, x, b, ret_def);
    x->storage = AB_PST_CONSISTENT;
    if (converged) *converged = ok ? 1 : 0;
    AB_CHECK_LAUNCH(s->sp->dom->ctx);
    AB_CATCH
}
int ab_solver_apply(ab_solver* s, ab_vector* x, ab_vector* b, int* converged) { return solver_apply_impl(s, x, b, converged, false); }
int ab_solver_apply_return_defect(ab_solver* s, ab_vector* x, ab_vector* b, int* converged) { return solver_apply_impl(s, x, b, converged, true); }
int ab_solver_step(ab_solver* s, int* out) {
    AB_TRY
    *out = s->last_steps;
    AB_CATCH
}
int ab_solver_last_defect(ab_solver* s, double* out) {
    AB_TRY
    *out = s->last_defect;
    AB_CATCH
}
int ab_solver_vcycle(ab_solver* s, ab_vector* z, ab_vector* r) {
    AB_TRY
    AB_REQUIRE(s->type == 1 && s->gmg, AB_ERR_STATE, "ab_solver_vcycle needs an initialised BiCGStab+GMG solver");
    s->gmg->apply(r->d.p, z->d.p);
    z->touch();
    AB_CHECK_LAUNCH(s->sp->dom->ctx);
    AB_CATCH
}
int ab_solver_level_info(ab_solver* s, int level, int64_t* nb, int64_t* nnzb) {
    AB_TRY
    Domain* dom = s->sp->dom;
    AB_REQUIRE(level >= 0 && level <= dom->top(), AB_ERR_ARG, "level out of range");
    if (nb) *nb = dom->dev[level].nv;
    if (nnzb) *nnzb = dom->dev[level].nnzb;
    AB_CATCH
}

// ---- ADMM free functions ----------------------------------------------------------------------
static int project_impl(ab_vector* qp, ab_vector* q, double sigma, bool spectral) {
    AB_TRY
    Domain* dom = q->sp->dom;
    Context* ctx = dom->ctx;
    const int dim = dom->dim();
    AB_REQUIRE(q->sp->kind == AB_SPACE_P0 && q->sp->ncomp == dim * dim && qp->n() == q->n(), AB_ERR_ARG, "projection needs P0 dim*dim tensors");
    const int64_t ne = dom->dev[dom->top()].ne;
    if (spectral) {
        AB_REQUIRE(dim == 2, AB_ERR_UNSUPPORTED, "ProjectWithSpectralNorm exists for 2D only (2d_admm.lua:902)");
        AB_LAUNCH(ctx, k_project_spectral, ew_grid(ctx, ne), 256, 0, ne, sigma, q->d.p, qp->d.p);
    } else if (dim == 2) AB_LAUNCH(ctx, (k_project_frobenius<4>), ew_grid(ctx, ne), 256, 0, ne, sigma, q->d.p, qp->d.p);
    else AB_LAUNCH(ctx, (k_project_frobenius<9>), ew_grid(ctx, ne), 256, 0, ne, sigma, q->d.p, qp->d.p);
    qp->storage = q->storage;
    qp->touch();
    AB_CHECK_LAUNCH(ctx);
    AB_CATCH
}
int ab_project_frobenius(ab_vector* qp, ab_vector* q, double sigma) { return project_impl(qp, q, sigma, false); }
int ab_project_spectral(ab_vector* qp, ab_vector* q, double sigma) { return project_impl(qp, q, sigma, true); }

static int max_norm_impl(ab_vector* u, double* out, bool spectral) {
    AB_TRY
    Domain* dom = u->sp->dom;
    Context* ctx = dom->ctx;
    const LevelDev& L = dom->dev[dom->top()];
    AB_REQUIRE(u->sp->kind == AB_SPACE_P1 && u->sp->ncomp == dom->dim(), AB_ERR_ARG, "needs the P1 deformation function");
    const int g = red_grid(ctx, L.ne);
    if (dom->dim() == 2) {
        if (spectral) AB_LAUNCH(ctx, (k_max_grad_norm<2, 1>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
        else AB_LAUNCH(ctx, (k_max_grad_norm<2, 0>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    } else {
        AB_REQUIRE(!spectral, AB_ERR_UNSUPPORTED, "MaxSpectralNorm exists for 2D only (2d_admm.lua:901)");
        AB_LAUNCH(ctx, (k_max_grad_norm<3, 0>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    }
    allreduce_dev(dom, ctx->d_results, 1, true);
    read_back(ctx, ctx->d_results, 1, out);
    AB_CATCH
}
int ab_max_frobenius_norm(ab_vector* u, double* out) { return max_norm_impl(u, out, false); }
int ab_max_spectral_norm(ab_vector* u, double* out) { return max_norm_impl(u, out, true); }

static int vol_bary_impl(ab_vector* u, double* out4) {
    AB_TRY
    Domain* dom = u->sp->dom;
    Context* ctx = dom->ctx;
    const LevelDev& L = dom->dev[dom->top()];
    AB_REQUIRE(u->sp->kind == AB_SPACE_P1 && u->sp->ncomp == dom->dim(), AB_ERR_ARG, "needs the P1 deformation function");
    const int g = red_grid(ctx, L.ne);
    if (dom->dim() == 2) AB_LAUNCH(ctx, (k_volume_barycenter<2>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    else AB_LAUNCH(ctx, (k_volume_barycenter<3>), g, 256, 0, (int64_t)L.ne, L.elems.p, L.xyz.p, u->d.p, ctx->d_partials, ctx->d_tickets, ctx->d_results);
    allreduce_dev(dom, ctx->d_results, dom->dim() + 1);
    read_back(ctx, ctx->d_results, dom->dim() + 1, out4);
    AB_CATCH
}
int ab_volume_defect(ab_vector* u, double reference_volume, double* out) {
    double t[4];
    int rc = vol_bary_impl(u, t);
    if (rc == AB_OK) *out = t[0] - reference_volume;
    return rc;
}
int ab_barycenter_defect(ab_vector* u, double* out_dim) {
    double t[4];
    int rc = vol_bary_impl(u, t);
    if (rc == AB_OK) for (int k = 0; k < u->sp->dom->dim(); ++k) out_dim[k] = t[1 + k];
    return rc;
}
int ab_set_zero_away_from_subset(ab_vector* v, const char* subset) {
    AB_TRY
    Domain* dom = v->sp->dom;
    Context* ctx = dom->ctx;
    AB_REQUIRE(v->sp->kind == AB_SPACE_P1, AB_ERR_ARG, "SetZeroAwayFromSubset needs a P1 function");
    const int si = dom->mesh.subset_index(subset);
    AB_REQUIRE(si >= 0, AB_ERR_ARG, std::string("unknown subset '") + subset + "'");
    AB_LAUNCH(ctx, k_zero_away_from_subset, ew_grid(ctx, v->n()), 256, 0, v->n(), v->sp->ncomp, dom->dev[dom->top()].vsub.p, si, v->d.p);
    v->touch();
    AB_CHECK_LAUNCH(ctx);
    AB_CATCH
}

}  // extern "C"
