"""Host-side domain decomposition for the multi-GPU path (one process per GPU).

Replaces, for the hot path, UG4's ParMETIS partitioning + hierarchical distribution (3d_admm.lua:124-186):
the level-0 elements are split by recursive coordinate bisection, every rank keeps the sub-grid of its own
elements and refines it locally (children inherit the parent's rank, `clusteredSiblings`, 3d_admm.lua:147), so
each rank owns a complete nested hierarchy of its subdomain.  Vertices shared between ranks are found per
level by exact coordinate matching of candidate lists (both sides run the same refinement arithmetic, so
shared vertices have bit-identical coordinates); the matched lists are sorted lexicographically, which gives
both sides the same canonical order without any further negotiation.

Hierarchical agglomeration (3d_admm.lua:151-183 keeps level 0 on one process and widens the process set level by
level): grid levels whose GLOBAL size is below a threshold are not decomposed at all -- rank 0 holds them as one global
hierarchy; the helpers at the end of this file compute the vertical-interface maps (local -> global vertex ids and
local -> global matrix-block positions on the gather level).  A problem whose top level is below the threshold is
not decomposed in the first place (ug4.Backend).

Pure NumPy + a tiny `gather` callable (torch.distributed.all_gather_object in production, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def rcb_partition(centroids: np.ndarray, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection along the longest extent; parts are proportional for any nparts >= 1."""
    part = np.zeros(len(centroids), np.int32)

    def rec(idx, first, count):
        if count == 1 or len(idx) == 0:
            part[idx] = first
            return
        left = count // 2
        ext = centroids[idx].max(axis=0) - centroids[idx].min(axis=0)
        ax = int(np.argmax(ext))
        order = idx[np.lexsort((idx, centroids[idx, ax]))]          # ties broken by element id: deterministic
        k = (len(idx) * left) // count
        rec(order[:k], first, left)
        rec(order[k:], first + left, count - left)

    rec(np.arange(len(centroids)), 0, nparts)
    return part


def vertex_rank_masks(elems: np.ndarray, part: np.ndarray, nv: int) -> np.ndarray:
    """uint64 bitmask per vertex: bit q set <=> some element of rank q contains the vertex (nparts <= 64)."""
    mask = np.zeros(nv, np.uint64)
    bits = (np.uint64(1) << part.astype(np.uint64))
    for a in range(elems.shape[1]):
        np.bitwise_or.at(mask, elems[:, a], bits)
    return mask


_LOCAL_EDGES = {2: [(0, 1), (1, 2), (0, 2)], 3: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]}
_LOCAL_FACES = [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]


def _rows_in(rows, table, nv):
    """boolean mask: which rows (sorted tuples) occur in table (sorted tuples)."""
    if len(rows) == 0:
        return np.zeros(0, bool)
    def key(a):
        k = a[:, 0].astype(np.int64)
        for c in range(1, a.shape[1]):
            k = k * nv + a[:, c]
        return k
    return np.isin(key(rows), key(table))


def extract_submesh(g: dict, part: np.ndarray, rank: int) -> dict:
    """Sub-grid of the elements owned by `rank` (local vertex numbering keeps the global order).
    `g`: dim, xyz, elems, vsub, esub, sp_edges(+_sub), sp_faces(+_sub) of the global level-0 grid."""
    dim = int(g["dim"])
    mine = np.flatnonzero(part == rank)
    elems_g = g["elems"][mine]
    l2g = np.unique(elems_g)
    g2l = -np.ones(len(g["xyz"]), np.int64)
    g2l[l2g] = np.arange(len(l2g))
    elems = g2l[elems_g].astype(np.int32)
    nv = len(g["xyz"])
    # boundary-subset edges / faces that are edges / faces of a local element
    le = np.concatenate([np.sort(np.stack([elems_g[:, i], elems_g[:, j]], 1), axis=1) for i, j in _LOCAL_EDGES[dim]])
    se, ses = g["sp_edges"], g["sp_edges_sub"]
    keep = _rows_in(se, le, nv)
    sp_edges, sp_edges_sub = g2l[se[keep]].astype(np.int32), ses[keep].astype(np.int32)
    if dim == 3 and len(g["sp_faces"]):
        lf = np.concatenate([np.sort(np.stack([elems_g[:, i], elems_g[:, j], elems_g[:, k]], 1), axis=1) for i, j, k in _LOCAL_FACES])
        sf, sfs = g["sp_faces"], g["sp_faces_sub"]
        keepf = _rows_in(sf, lf, nv)
        sp_faces, sp_faces_sub = g2l[sf[keepf]].astype(np.int32), sfs[keepf].astype(np.int32)
    else:
        sp_faces, sp_faces_sub = np.zeros((0, 3), np.int32), np.zeros(0, np.int32)
    return dict(dim=dim, xyz=np.ascontiguousarray(g["xyz"][l2g]), elems=np.ascontiguousarray(elems), vsub=g["vsub"][l2g].astype(np.int32),
                esub=g["esub"][mine].astype(np.int32), sp_edges=np.ascontiguousarray(sp_edges).reshape(-1, 2), sp_edges_sub=sp_edges_sub,
                sp_faces=np.ascontiguousarray(sp_faces).reshape(-1, 3), sp_faces_sub=sp_faces_sub, l2g=l2g.astype(np.int32))


def refine_masks(mask_coarse: np.ndarray, parent_a: np.ndarray, parent_b: np.ndarray) -> np.ndarray:
    """Candidate sharing masks of the next level: copies keep theirs, a midpoint can only be shared with ranks that
    share both parents (a superset of the truth; the coordinate matching removes the false positives)."""
    return np.concatenate([mask_coarse, mask_coarse[parent_a] & mask_coarse[parent_b]])


def _void_rows(x):
    x = np.ascontiguousarray(x)
    return x.view([("", x.dtype)] * x.shape[1]).ravel()


def match_level(xyz: np.ndarray, mask: np.ndarray, rank: int, nranks: int, gather):
    """Interfaces of one level. `gather(obj)` returns the list of every rank's obj (all_gather_object).
    Returns (neigh, offsets, idx, owned)."""
    cand = {}
    for q in range(nranks):
        if q == rank:
            continue
        c = np.flatnonzero(mask & (np.uint64(1) << np.uint64(q)))
        if len(c):
            cand[q] = c
    everyone = gather({q: xyz[c] for q, c in cand.items()})
    neigh, offsets, idx = [], [0], []
    owned = np.ones(len(xyz), np.uint8)
    for q in sorted(cand):
        theirs = everyone[q].get(rank)
        if theirs is None or len(theirs) == 0:
            continue
        mine = xyz[cand[q]]
        _, ia, _ = np.intersect1d(_void_rows(mine), _void_rows(theirs), return_indices=True)   # sorted by coordinates
        if len(ia) == 0:
            continue
        loc = cand[q][ia].astype(np.int32)
        neigh.append(q)
        idx.append(loc)
        offsets.append(offsets[-1] + len(loc))
        if q < rank:
            owned[loc] = 0
    idx = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32)
    return np.array(neigh, np.int32), np.array(offsets, np.int32), idx, owned


# ------------------------------------------------------------------------------------------------
# hierarchical agglomeration: which levels are gathered, and the vertical-interface maps
# ------------------------------------------------------------------------------------------------
def global_level_counts(g: dict, refs: int):
    """Vertices per level of the GLOBAL hierarchy from the regular-refinement recurrences V' = V+E, E' = 2E+3F(+T),
    F' = 4F(+8T), T' = 8T, seeded with the level-0 entity counts of the grid."""
    el = np.asarray(g["elems"])
    dim = int(g["dim"])
    V = len(g["xyz"])
    E = len(np.unique(np.sort(np.concatenate([el[:, [i, j]] for i, j in _LOCAL_EDGES[dim]]), axis=1), axis=0))
    if dim == 3:
        T = len(el)
        F = len(np.unique(np.sort(np.concatenate([el[:, list(f)] for f in _LOCAL_FACES]), axis=1), axis=0))
    else:
        T, F = 0, len(el)
    out = [V]
    for _ in range(refs):
        V, E, F, T = V + E, 2 * E + 3 * F + T, 4 * F + 8 * T, 8 * T
        out.append(V)
    return out


def gather_level(nv_levels, dim: int, max_dofs: int) -> int:
    """Highest level whose global P1 deformation space has at most max_dofs unknowns (level 0 always qualifies):
    levels 0..gather_level are held undivided by rank 0, the levels above are decomposed over all ranks."""
    lg = 0
    for l, nv in enumerate(nv_levels):
        if nv * dim <= max_dofs:
            lg = l
    return lg


def propagate_l2g(l2g_coarse, nvc_global, gpa, gpb, lpa, lpb):
    """local -> global vertex ids one level up.  Copies keep their ids (coarse vertices are a prefix of the fine ones on
    both sides); a local midpoint is the global midpoint of the global edge between the images of its parents."""
    l2g_coarse = np.asarray(l2g_coarse, np.int64)
    nvg = int(nvc_global)
    gkey = np.minimum(gpa, gpb).astype(np.int64) * nvg + np.maximum(gpa, gpb)
    order = np.argsort(gkey, kind="stable")
    a, b = l2g_coarse[lpa], l2g_coarse[lpb]
    lkey = np.minimum(a, b) * nvg + np.maximum(a, b)
    pos = np.searchsorted(gkey[order], lkey)
    if len(lkey) and (pos.max(initial=0) >= len(gkey) or not np.array_equal(gkey[order][pos], lkey)):
        raise ValueError("a local edge has no global counterpart")
    return np.concatenate([l2g_coarse, nvg + order[pos]]).astype(np.int64)


def pattern_keys(elems, nv: int):
    """Sorted keys i*nv + j of the P1 block pattern (vertex pairs sharing an element, diagonal included) = the BSR order of
    ab_domain_level_pattern: rows ascending, columns ascending."""
    el = np.asarray(elems, np.int64)
    n = el.shape[1]
    keys = np.concatenate([el[:, a] * nv + el[:, b] for a in range(n) for b in range(n)])
    return np.unique(keys)


def block_positions(local_elems, nv_local: int, l2g, global_keys, nv_global: int):
    """Position in the global BSR pattern of every block of the local pattern (local BSR order)."""
    lk = pattern_keys(local_elems, nv_local)
    i, j = lk // nv_local, lk % nv_local
    l2g = np.asarray(l2g, np.int64)
    gk = l2g[i] * nv_global + l2g[j]
    pos = np.searchsorted(global_keys, gk)
    if len(gk) and (pos.max(initial=0) >= len(global_keys) or not np.array_equal(global_keys[pos], gk)):
        raise ValueError("a local matrix block has no global counterpart")
    return pos.astype(np.int32)


def csr_keys(rowptr, colidx, nv: int):
    """The keys of pattern_keys from a BSR pattern (ab_domain_level_pattern): already sorted, no np.unique needed."""
    rowptr = np.asarray(rowptr, np.int64)
    rows = np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))
    return rows * nv + np.asarray(colidx, np.int64)


def block_positions_csr(rowptr_l, colidx_l, l2g, rowptr_g, colidx_g):
    """block_positions from the two BSR patterns the native side builds (same result, a fraction of the host time)."""
    l2g = np.asarray(l2g, np.int64)
    nvg = len(rowptr_g) - 1
    rowptr_l = np.asarray(rowptr_l, np.int64)
    rows = np.repeat(np.arange(len(rowptr_l) - 1, dtype=np.int64), np.diff(rowptr_l))
    gk = l2g[rows] * nvg + l2g[np.asarray(colidx_l, np.int64)]
    gkeys = csr_keys(rowptr_g, colidx_g, nvg)
    pos = np.searchsorted(gkeys, gk)
    if len(gk) and (pos.max(initial=0) >= len(gkeys) or not np.array_equal(gkeys[pos], gk)):
        raise ValueError("a local matrix block has no global counterpart")
    return pos.astype(np.int32)


# ------------------------------------------------------------------------------------------------
# shared matrix blocks of a decomposed level (exact Gershgorin bound of the additive operators)
# ------------------------------------------------------------------------------------------------
def match_blocks(rowptr, colidx, neigh, offsets, idx, rank: int, gather):
    """Matrix blocks (i, j) whose two vertices are shared with a neighbour rank AND that exist in both local patterns: their
    values are additive over the ranks, so sum_j |a_ij| of the global operator needs their sum before the absolute value.
    `neigh/offsets/idx`: the vertex interface of the level (match_level).  Every pair of neighbours derives the same ordered
    list from integers only: a block is keyed by the positions of its two vertices in the pair's common vertex order.
    Returns (offsets_b, slot_block, bpos, brow, mult): per neighbour (same order as `neigh`) the slots offsets_b[n]..offsets_b[n+1]
    into slot_block = compact ids of the shared blocks; bpos = local block position of every compact id, brow = its row vertex,
    mult = number of ranks holding it."""
    rowptr = np.asarray(rowptr, np.int64)
    colidx = np.asarray(colidx, np.int64)
    nv = len(rowptr) - 1
    cand = {}
    for n, q in enumerate(neigh):
        verts = np.asarray(idx[offsets[n]:offsets[n + 1]], np.int64)
        m = len(verts)
        pos = -np.ones(nv, np.int64)
        pos[verts] = np.arange(m)
        cnt = rowptr[verts + 1] - rowptr[verts]
        start = np.repeat(rowptr[verts], cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        b = start + within                                           # local block positions of the rows of the shared vertices
        pi = np.repeat(np.arange(m), cnt)
        pj = pos[colidx[b]]
        keep = pj >= 0
        key = pi[keep] * m + pj[keep]
        order = np.argsort(key)
        cand[int(q)] = (key[order], b[keep][order])
    everyone = gather({q: k for q, (k, _) in cand.items()})
    offsets_b, slots = [0], []
    for q in [int(x) for x in neigh]:
        mine_k, mine_b = cand[q]
        theirs = everyone[q].get(rank)
        if theirs is None:
            theirs = np.zeros(0, np.int64)
        _, ia, _ = np.intersect1d(mine_k, theirs, return_indices=True)    # ascending common keys: the same order on both sides
        slots.append(mine_b[ia])
        offsets_b.append(offsets_b[-1] + len(ia))
    allb = np.concatenate(slots) if slots else np.zeros(0, np.int64)
    bpos, inv, counts = np.unique(allb, return_inverse=True, return_counts=True)
    brow = np.searchsorted(rowptr, bpos, side="right") - 1
    return (np.array(offsets_b, np.int32), inv.astype(np.int32), bpos.astype(np.int32), brow.astype(np.int32), (counts + 1).astype(np.int32))
