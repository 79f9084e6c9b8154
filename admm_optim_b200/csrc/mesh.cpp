// Host-side grid hierarchy (see mesh.hpp).  Compiled with -ffp-contract=off so that the
// shortest-diagonal rule evaluates exactly the same doubles as the NumPy oracle.
#include "mesh.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#if defined(_OPENMP)
#include <omp.h>
#include <parallel/algorithm>
#define AB_SORT(b, e) __gnu_parallel::sort((b), (e))
#else
#define AB_SORT(b, e) std::sort((b), (e))
#endif

namespace ab {

int HostMesh::subset_index(const std::string& name) const {
    for (size_t i = 0; i < subset_names.size(); ++i)
        if (subset_names[i] == name) return (int)i;
    return -1;
}

static const int LE2[3][2] = {{0, 1}, {1, 2}, {0, 2}};
static const int LE3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

static inline uint64_t ekey(int32_t a, int32_t b) {
    uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    return ((uint64_t)lo << 32) | hi;
}

// Unique edges (lo, hi), sorted lexicographically, from a generator of candidate pairs (duplicates allowed).  Bucket sort by the
// lower vertex instead of one comparison sort over all candidates (239 M element edges at level 5): count per lo, scatter the hi
// ends into the buckets (order inside a bucket is irrelevant), then sort + unique every bucket (a few dozen entries) and compact.
// The result does not depend on the scatter order.  gen(i, emit) calls emit(a, b) for every candidate of outer index i.
namespace {
template <class Gen>
void unique_edges(int nv, int64_t n_outer, const Gen& gen, std::vector<int32_t>& edges) {
    std::vector<int64_t> start((size_t)nv + 1, 0);
    {
        std::vector<int32_t> cnt((size_t)nv, 0);
#pragma omp parallel for schedule(dynamic, 16384)
        for (int64_t i = 0; i < n_outer; ++i)
            gen(i, [&](int32_t a, int32_t b) { __atomic_fetch_add(&cnt[a < b ? a : b], 1, __ATOMIC_RELAXED); });
        for (int v = 0; v < nv; ++v) start[v + 1] = start[v] + cnt[v];
    }
    std::vector<int32_t> his((size_t)start[nv]);
    {
        std::vector<int32_t> fill((size_t)nv, 0);
#pragma omp parallel for schedule(dynamic, 16384)
        for (int64_t i = 0; i < n_outer; ++i)
            gen(i, [&](int32_t a, int32_t b) {
                const int32_t lo = a < b ? a : b, hi = a < b ? b : a;
                his[(size_t)(start[lo] + __atomic_fetch_add(&fill[lo], 1, __ATOMIC_RELAXED))] = hi;
            });
    }
    std::vector<int64_t> ustart((size_t)nv + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int v = 0; v < nv; ++v) {
        int32_t* b = his.data() + start[v];
        int32_t* e = his.data() + start[v + 1];
        std::sort(b, e);
        ustart[v + 1] = std::unique(b, e) - b;
    }
    for (int v = 0; v < nv; ++v) ustart[v + 1] += ustart[v];
    edges.resize((size_t)ustart[nv] * 2);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int v = 0; v < nv; ++v) {
        const int32_t* b = his.data() + start[v];
        const int64_t n = ustart[v + 1] - ustart[v];
        int32_t* out = edges.data() + 2 * ustart[v];
        for (int64_t i = 0; i < n; ++i) { out[2 * i] = v; out[2 * i + 1] = b[i]; }
    }
}
}  // namespace

void ensure_edges(HostLevel& L) {
    if (L.have_edges) return;
    const int nen = L.dim + 1, nle = L.dim == 3 ? 6 : 3, dim = L.dim;
    const int32_t* el = L.elems.data();
    unique_edges(L.nv, L.ne, [=](int64_t e, auto emit) {
        const int32_t* v = el + (size_t)e * nen;
        for (int k = 0; k < nle; ++k) {
            const int* le = dim == 3 ? LE3[k] : LE2[k];
            emit(v[le[0]], v[le[1]]);
        }
    }, L.edges);
    L.have_edges = true;
}

// Edges of a regularly refined level from its parent instead of from its own 6 ne element edges: every fine edge is
//   (a) one half of a coarse edge k = (a, b):  (a, m_k), (b, m_k)                                 -- 2 per coarse edge, no duplicates
//   (b) an edge between two midpoints inside a coarse element: 2D the 3 sides of the inner triangle (child 3); 3D the 12 sides
//       of the inner octahedron + its diagonal = the edges of children 4..7 = {p, q, c_j, c_j+1} (refine_level): (p, q), (p, c_j),
//       (q, c_j), (c_j, c_j+1) -- 13 per tetrahedron, the octahedron sides on a coarse face are seen from both neighbours.
// 3.7x (3D) fewer candidates than the generic path, same sorted unique result (checked against it in the tests).
void edges_from_parent(const HostLevel& C, HostLevel& F) {
    const int dim = C.dim, nvc = C.nv;
    const int64_t nec = C.ne, nedc = C.nedges();
    const int32_t* ce = C.edges.data();
    const int32_t* fe = F.elems.data();
    unique_edges(F.nv, nec + nedc, [=](int64_t i, auto emit) {
        if (i >= nec) {
            const int64_t k = i - nec;
            const int32_t m = (int32_t)(nvc + k);
            emit(ce[2 * k], m);
            emit(ce[2 * k + 1], m);
        } else if (dim == 2) {
            const int32_t* c = fe + ((size_t)i * 4 + 3) * 3;          // child 3 = (mab, mbc, mca)
            emit(c[0], c[1]); emit(c[1], c[2]); emit(c[2], c[0]);
        } else {
            const int32_t* c4 = fe + ((size_t)i * 8 + 4) * 4;          // children 4..7 = (p, q, x, y), x/y possibly swapped
            const int32_t p = c4[0], q = c4[1];
            emit(p, q);
            for (int j = 0; j < 4; ++j) emit(c4[4 * j + 2], c4[4 * j + 3]);
            for (int j = 0; j < 4; j += 2)                             // c0, c1 from child 4; c2, c3 from child 6
                for (int t = 2; t < 4; ++t) { emit(p, c4[4 * j + t]); emit(q, c4[4 * j + t]); }
        }
    }, F.edges);
    F.have_edges = true;
}

namespace {
struct EdgeFinder {
    const int32_t* e;
    int64_t n;
    std::vector<int64_t> start;   // first edge with lo == v  (size nv+1)
    EdgeFinder(const HostLevel& L) : e(L.edges.data()), n(L.nedges()), start((size_t)L.nv + 1, 0) {
        for (int64_t i = 0; i < n; ++i) start[e[2 * i] + 1]++;
        for (int v = 0; v < L.nv; ++v) start[v + 1] += start[v];
    }
    inline int64_t find(int32_t a, int32_t b) const {
        int32_t lo = a < b ? a : b, hi = a < b ? b : a;
        int64_t l = start[lo], r = start[lo + 1];
        while (l < r) {
            int64_t m = (l + r) >> 1;
            if (e[2 * m + 1] < hi) l = m + 1; else r = m;
        }
        return l;   // caller guarantees existence
    }
};

inline double sqdist(const double* p, const double* q, int dim) {
    double d0 = p[0] - q[0];
    double acc = d0 * d0;
    for (int c = 1; c < dim; ++c) {
        double dc = p[c] - q[c];
        acc = acc + dc * dc;
    }
    return acc;
}
inline double det3(const double* x0, const double* x1, const double* x2, const double* x3) {
    double a[3], b[3], c[3];
    for (int k = 0; k < 3; ++k) { a[k] = x1[k] - x0[k]; b[k] = x2[k] - x0[k]; c[k] = x3[k] - x0[k]; }
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
struct SpEdge { int32_t lo, hi, sub; };
}  // namespace

void refine_level(const HostLevel& C, HostLevel& F) {
    const int dim = C.dim, nv = C.nv;
    const int64_t ned = C.nedges();
    EdgeFinder ef(C);
    F.dim = dim;
    F.nv_coarse = nv;
    F.nv = nv + (int)ned;
    F.xyz.resize((size_t)F.nv * dim);
    std::copy(C.xyz.begin(), C.xyz.end(), F.xyz.begin());
    F.pa.resize(ned);
    F.pb.resize(ned);
    const int vol_sub = C.esub.empty() ? 0 : C.esub[0];
    F.vsub.assign((size_t)F.nv, vol_sub);
    std::copy(C.vsub.begin(), C.vsub.end(), F.vsub.begin());
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ned; ++k) {
        int32_t a = C.edges[2 * k], b = C.edges[2 * k + 1];
        F.pa[k] = a;
        F.pb[k] = b;
        for (int c = 0; c < dim; ++c) F.xyz[(size_t)(nv + k) * dim + c] = 0.5 * (C.xyz[(size_t)a * dim + c] + C.xyz[(size_t)b * dim + c]);
    }
    const int64_t nse = (int64_t)C.sp_edges_sub.size();
    for (int64_t i = 0; i < nse; ++i) F.vsub[nv + ef.find(C.sp_edges[2 * i], C.sp_edges[2 * i + 1])] = C.sp_edges_sub[i];
    auto mid = [&](int32_t a, int32_t b) -> int32_t { return (int32_t)(nv + ef.find(a, b)); };

    const int nch = dim == 3 ? 8 : 4, nen = dim + 1;
    F.ne = C.ne * nch;
    F.elems.resize((size_t)F.ne * nen);
    F.esub.resize((size_t)F.ne);
    const double* X = F.xyz.data();
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < C.ne; ++e) {
        const int32_t* v = &C.elems[(size_t)e * nen];
        int32_t* out = &F.elems[(size_t)e * nch * nen];
        for (int c = 0; c < nch; ++c) F.esub[(size_t)e * nch + c] = C.esub[e];
        if (dim == 2) {
            int32_t a = v[0], b = v[1], c = v[2];
            int32_t mab = mid(a, b), mbc = mid(b, c), mca = mid(c, a);
            int32_t ch[4][3] = {{a, mab, mca}, {mab, b, mbc}, {mca, mbc, c}, {mab, mbc, mca}};
            std::memcpy(out, ch, sizeof(ch));
        } else {
            int32_t m01 = mid(v[0], v[1]), m02 = mid(v[0], v[2]), m03 = mid(v[0], v[3]);
            int32_t m12 = mid(v[1], v[2]), m13 = mid(v[1], v[3]), m23 = mid(v[2], v[3]);
            double d0 = sqdist(X + 3 * (size_t)m01, X + 3 * (size_t)m23, 3);
            double d1 = sqdist(X + 3 * (size_t)m02, X + 3 * (size_t)m13, 3);
            double d2 = sqdist(X + 3 * (size_t)m03, X + 3 * (size_t)m12, 3);
            int choice = 0;
            double best = d0;
            if (d1 < best) { choice = 1; best = d1; }
            if (d2 < best) { choice = 2; }
            int32_t p, q, c0, c1, c2, c3;
            if (choice == 0)      { p = m01; q = m23; c0 = m02; c1 = m03; c2 = m13; c3 = m12; }
            else if (choice == 1) { p = m02; q = m13; c0 = m01; c1 = m03; c2 = m23; c3 = m12; }
            else                  { p = m03; q = m12; c0 = m01; c1 = m02; c2 = m23; c3 = m13; }
            int32_t ch[8][4] = {{v[0], m01, m02, m03}, {m01, v[1], m12, m13}, {m02, m12, v[2], m23}, {m03, m13, m23, v[3]},
                                {p, q, c0, c1}, {p, q, c1, c2}, {p, q, c2, c3}, {p, q, c3, c0}};
            for (int c = 0; c < 8; ++c) {
                if (det3(X + 3 * (size_t)ch[c][0], X + 3 * (size_t)ch[c][1], X + 3 * (size_t)ch[c][2], X + 3 * (size_t)ch[c][3]) < 0)
                    std::swap(ch[c][2], ch[c][3]);
            }
            std::memcpy(out, ch, sizeof(ch));
        }
    }
    // special entities of the fine level
    std::vector<SpEdge> ne_;
    ne_.reserve((size_t)nse * 2 + C.sp_faces_sub.size() * 3);
    auto push = [&](int32_t a, int32_t b, int32_t s) { ne_.push_back({a < b ? a : b, a < b ? b : a, s}); };
    for (int64_t i = 0; i < nse; ++i) {
        int32_t a = C.sp_edges[2 * i], b = C.sp_edges[2 * i + 1], m = mid(a, b);
        push(a, m, C.sp_edges_sub[i]);
        push(b, m, C.sp_edges_sub[i]);
    }
    const int64_t nsf = (int64_t)C.sp_faces_sub.size();
    F.sp_faces.clear();
    F.sp_faces_sub.clear();
    if (dim == 3) {
        F.sp_faces.reserve((size_t)nsf * 12);
        for (int64_t i = 0; i < nsf; ++i) {
            int32_t a = C.sp_faces[3 * i], b = C.sp_faces[3 * i + 1], c = C.sp_faces[3 * i + 2], s = C.sp_faces_sub[i];
            int32_t mab = mid(a, b), mbc = mid(b, c), mca = mid(c, a);
            push(mab, mbc, s); push(mbc, mca, s); push(mca, mab, s);
            int32_t ch[4][3] = {{a, mab, mca}, {mab, b, mbc}, {mca, mbc, c}, {mab, mbc, mca}};
            for (auto& f : ch) {
                std::sort(f, f + 3);
                F.sp_faces.insert(F.sp_faces.end(), f, f + 3);
                F.sp_faces_sub.push_back(s);
            }
        }
    }
    std::sort(ne_.begin(), ne_.end(), [](const SpEdge& x, const SpEdge& y) { return x.lo != y.lo ? x.lo < y.lo : x.hi < y.hi; });
    F.sp_edges.resize(ne_.size() * 2);
    F.sp_edges_sub.resize(ne_.size());
    for (size_t i = 0; i < ne_.size(); ++i) {
        F.sp_edges[2 * i] = ne_[i].lo;
        F.sp_edges[2 * i + 1] = ne_[i].hi;
        F.sp_edges_sub[i] = ne_[i].sub;
    }
    F.have_edges = false;
    F.edges.clear();
    if (!getenv("ADMM_B200_GENERIC_EDGES")) edges_from_parent(C, F);
}

// Rows of the P1 vertex graph: [lower neighbours ascending | diagonal | upper neighbours ascending].  The edges are sorted by
// (lo, hi), so the upper part of row lo is a contiguous run of the edge list (copied in parallel); the lower part of row hi
// collects the edges (lo, hi) from all over the list -- scattered with an atomic cursor and then sorted per row by the edge
// index (= ascending lo), which makes the result independent of the scatter order.
void build_pattern(HostLevel& L, HostPattern& P) {
    ensure_edges(L);
    const int nv = L.nv;
    const int64_t ned = L.nedges();
    const int32_t* E = L.edges.data();
    std::vector<int32_t> nlow((size_t)nv, 0);
    std::vector<int64_t> estart((size_t)nv + 1, 0);       // first edge with lo == v
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ned; ++k) {
        __atomic_fetch_add(&nlow[E[2 * k + 1]], 1, __ATOMIC_RELAXED);
        if (k == 0 || E[2 * k] != E[2 * k - 2])
            estart[E[2 * k]] = k + 1;                      // marker (+1 so that 0 means "no edge starts here")
    }
    {   // vertices without an upper edge inherit the next start
        int64_t next = ned;
        for (int v = nv - 1; v >= 0; --v) {
            if (estart[v]) next = estart[v] - 1;
            estart[v] = next;
        }
        estart[nv] = ned;
    }
    P.rowptr.resize((size_t)nv + 1);
    P.rowptr[0] = 0;
    for (int i = 0; i < nv; ++i) P.rowptr[i + 1] = P.rowptr[i] + nlow[i] + (int32_t)(estart[i + 1] - estart[i]) + 1;
    const int64_t nnz = P.rowptr[nv];
    P.colidx.resize(nnz);
    P.mid.resize(nnz);
    P.diagpos.resize(nv);
    std::vector<int32_t> fill((size_t)nv, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nv; ++i) {
        const int32_t d = P.rowptr[i] + nlow[i];
        P.diagpos[i] = d;
        P.colidx[d] = i;
        P.mid[d] = i;
        int32_t o = d + 1;
        for (int64_t k = estart[i]; k < estart[i + 1]; ++k, ++o) { P.colidx[o] = E[2 * k + 1]; P.mid[o] = (int32_t)(nv + k); }
    }
    // lower parts: mid holds the edge index until the per-row sort
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ned; ++k) {
        const int32_t hi = E[2 * k + 1];
        P.mid[P.rowptr[hi] + __atomic_fetch_add(&fill[hi], 1, __ATOMIC_RELAXED)] = (int32_t)(nv + k);
    }
#pragma omp parallel for schedule(dynamic, 4096)
    for (int i = 0; i < nv; ++i) {
        int32_t* m = P.mid.data() + P.rowptr[i];
        std::sort(m, m + nlow[i]);
        for (int j = 0; j < nlow[i]; ++j) P.colidx[P.rowptr[i] + j] = E[2 * (int64_t)(m[j] - nv)];
    }
}

// vertex -> element incidence (CSR), elements ascending per vertex: the fixed summation order of the row-owner assembly.
// Deliberately serial: the element loop visits the vertices with good locality, and a scatter with atomic cursors plus a per-vertex
// sort was measured slower than this on 8 threads (level 5: 1.8 s serial, 2.2 s threaded).
void build_v2e(const HostLevel& L, std::vector<int32_t>& ptr, std::vector<int32_t>& idx) {
    const int N = L.dim + 1, nv = L.nv;
    const int64_t n = (int64_t)L.ne * N;
    ptr.assign((size_t)nv + 1, 0);
    idx.resize((size_t)n);
    for (int64_t k = 0; k < n; ++k) ptr[L.elems[k] + 1]++;
    for (int v = 0; v < nv; ++v) ptr[v + 1] += ptr[v];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < L.ne; ++e)
        for (int a = 0; a < N; ++a) idx[fill[L.elems[(size_t)e * N + a]]++] = (int32_t)e;
}

// ---------------------------------------------------------------------------------------------
// UGX reader (the subset of the format the shipped grids use)
// ---------------------------------------------------------------------------------------------
namespace {
bool tag_body(const std::string& s, size_t from, size_t to, const std::string& tag, std::string& body, size_t* end_pos = nullptr) {
    size_t p = s.find("<" + tag, from);
    while (p != std::string::npos && p < to) {
        char c = s[p + 1 + tag.size()];
        if (c == '>' || c == ' ') break;
        p = s.find("<" + tag, p + 1);
    }
    if (p == std::string::npos || p >= to) return false;
    size_t gt = s.find('>', p);
    size_t close = s.find("</" + tag + ">", gt);
    if (gt == std::string::npos || close == std::string::npos || close > to) return false;
    body = s.substr(gt + 1, close - gt - 1);
    if (end_pos) *end_pos = close + tag.size() + 3;
    return true;
}
void parse_ints(const std::string& b, std::vector<int64_t>& out) {
    const char* p = b.c_str();
    char* q;
    while (true) {
        long long v = strtoll(p, &q, 10);
        if (q == p) break;
        out.push_back(v);
        p = q;
    }
}
void parse_doubles(const std::string& b, std::vector<double>& out) {
    const char* p = b.c_str();
    char* q;
    while (true) {
        double v = strtod(p, &q);
        if (q == p) break;
        out.push_back(v);
        p = q;
    }
}
}  // namespace

bool load_ugx(const std::string& path, HostMesh& mesh, std::string& err) {
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    size_t sh = s.find("<subset_handler");
    if (sh == std::string::npos) { err = "no <subset_handler> in " + path; return false; }
    size_t vp = s.find("<vertices coords=\"");
    if (vp == std::string::npos || vp > sh) { err = "no <vertices coords=..> in " + path; return false; }
    int nc = s[vp + 18] - '0';
    std::string body;
    tag_body(s, vp, sh, "vertices", body);
    std::vector<double> xyz3;
    parse_doubles(body, xyz3);
    std::vector<int64_t> edges, tris, tets;
    if (tag_body(s, 0, sh, "edges", body)) parse_ints(body, edges);
    if (tag_body(s, 0, sh, "triangles", body)) parse_ints(body, tris);
    if (tag_body(s, 0, sh, "tetrahedrons", body)) parse_ints(body, tets);
    const int dim = tets.empty() ? 2 : 3;
    if (dim == 2 && tris.empty()) { err = "grid has neither triangles nor tetrahedrons"; return false; }
    HostLevel L;
    L.dim = dim;
    L.nv = (int)(xyz3.size() / nc);
    L.xyz.resize((size_t)L.nv * dim);
    for (int v = 0; v < L.nv; ++v)
        for (int c = 0; c < dim; ++c) L.xyz[(size_t)v * dim + c] = xyz3[(size_t)v * nc + c];
    const std::vector<int64_t>& el = dim == 3 ? tets : tris;
    L.ne = (int)(el.size() / (dim + 1));
    L.elems.assign(el.begin(), el.end());
    L.vsub.assign(L.nv, -1);
    L.esub.assign(L.ne, -1);
    std::vector<int32_t> esubs(edges.size() / 2, -1), fsubs(tris.size() / 3, -1);
    mesh.subset_names.clear();
    size_t pos = sh;
    int si = 0;
    while (true) {
        size_t p = s.find("<subset name=\"", pos);
        if (p == std::string::npos) break;
        size_t q = s.find('"', p + 14);
        mesh.subset_names.push_back(s.substr(p + 14, q - p - 14));
        size_t end = s.find("</subset>", q);
        std::vector<int64_t> ids;
        if (tag_body(s, q, end, "vertices", body)) { ids.clear(); parse_ints(body, ids); for (auto i : ids) L.vsub[i] = si; }
        if (tag_body(s, q, end, "edges", body)) { ids.clear(); parse_ints(body, ids); for (auto i : ids) esubs[i] = si; }
        if (tag_body(s, q, end, "faces", body)) {
            ids.clear(); parse_ints(body, ids);
            for (auto i : ids) { if (dim == 3) fsubs[i] = si; else L.esub[i] = si; }
        }
        if (tag_body(s, q, end, "volumes", body)) { ids.clear(); parse_ints(body, ids); for (auto i : ids) L.esub[i] = si; }
        pos = end + 9;
        ++si;
    }
    for (int v = 0; v < L.nv; ++v) if (L.vsub[v] < 0) { err = "vertex without subset"; return false; }
    for (int e = 0; e < L.ne; ++e) if (L.esub[e] != L.esub[0] || L.esub[e] < 0) { err = "exactly one element subset expected"; return false; }
    const int vol_sub = L.esub[0];
    std::vector<SpEdge> sp;
    for (size_t i = 0; i < esubs.size(); ++i)
        if (esubs[i] != vol_sub && esubs[i] >= 0) {
            int32_t a = (int32_t)edges[2 * i], b = (int32_t)edges[2 * i + 1];
            sp.push_back({a < b ? a : b, a < b ? b : a, esubs[i]});
        }
    std::sort(sp.begin(), sp.end(), [](const SpEdge& x, const SpEdge& y) { return x.lo != y.lo ? x.lo < y.lo : x.hi < y.hi; });
    for (auto& e : sp) { L.sp_edges.push_back(e.lo); L.sp_edges.push_back(e.hi); L.sp_edges_sub.push_back(e.sub); }
    if (dim == 3)
        for (size_t i = 0; i < fsubs.size(); ++i)
            if (fsubs[i] != vol_sub && fsubs[i] >= 0) {
                int32_t t[3] = {(int32_t)tris[3 * i], (int32_t)tris[3 * i + 1], (int32_t)tris[3 * i + 2]};
                std::sort(t, t + 3);
                L.sp_faces.insert(L.sp_faces.end(), t, t + 3);
                L.sp_faces_sub.push_back(fsubs[i]);
            }
    mesh.dim = dim;
    mesh.levels.clear();
    mesh.levels.push_back(std::move(L));
    return true;
}

}  // namespace ab
