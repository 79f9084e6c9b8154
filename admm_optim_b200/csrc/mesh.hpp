// Host-side grid hierarchy: .ugx reader, regular refinement, P1 sparsity pattern.
// Replaces (for the hot path) UG4 lib_grid + refinement_util: LoadDomain (3d_admm.lua:108-109),
// util.refinement.CreateRegularHierarchy (3d_admm.lua:186).  Conventions are documented in DESIGN.md
// ("Mesh conventions") and restated independently by oracle/mesh_np.py.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ab {

struct HostLevel {
    int dim = 0, nv = 0, ne = 0, nv_coarse = 0;
    std::vector<double> xyz;       // nv*dim, current coordinates
    std::vector<int32_t> elems;    // ne*(dim+1)
    std::vector<int32_t> vsub;     // nv
    std::vector<int32_t> esub;     // ne
    std::vector<int32_t> sp_edges, sp_edges_sub;   // special edges (subset != element subset): 2 ints each, sorted (lo,hi)
    std::vector<int32_t> sp_faces, sp_faces_sub;   // special faces (3D): 3 ints each, ascending
    std::vector<int32_t> pa, pb;   // parents of vertices nv_coarse..nv-1 (edge midpoints)
    std::vector<int32_t> edges;    // 2*nedges, sorted unique (lo,hi); filled by ensure_edges()
    bool have_edges = false;
    int64_t nedges() const { return (int64_t)edges.size() / 2; }
};

struct HostPattern {               // P1 vertex graph incl. diagonal, columns ascending
    std::vector<int32_t> rowptr, colidx, diagpos;
    std::vector<int32_t> mid;      // per entry: vertex id of the edge midpoint on the next level (diag: the vertex itself)
};

struct HostMesh {
    int dim = 0;
    std::vector<std::string> subset_names;
    std::vector<HostLevel> levels;
    int subset_index(const std::string& name) const;
};

void ensure_edges(HostLevel& L);
void edges_from_parent(const HostLevel& coarse_with_edges, HostLevel& fine);
void refine_level(const HostLevel& coarse_with_edges, HostLevel& fine);
void build_pattern(HostLevel& L, HostPattern& P);
void build_v2e(const HostLevel& L, std::vector<int32_t>& ptr, std::vector<int32_t>& idx);
bool load_ugx(const std::string& path, HostMesh& mesh, std::string& err);

}  // namespace ab
