// wcfe_partition.cu -- host-only part of the C ABI: cutting ONE substrate mesh into strips for a multi-GPU run and producing
// the tables wcfe_halo_connect wants (include/wcfe.h, "one mesh over several GPUs").  No device code, no CUDA calls: a C++ (or
// MPI, or Python) host calls wcfe_strip_partition on every rank with the same global mesh, uploads the rank's local strip with
// wcfe_upload_mesh, exchanges the wcfe_halo_info blobs and hands the lists below to wcfe_halo_connect.
//
// The reference has no such mode (its only parallelism is a job array of independent processes, s_sim_031021_1.sub).
// The partition is the one wholecellfe_b200/partition.py computes (strip_partition + connect_plan) -- tests/test_partition.py
// holds the two against each other:
//   * nodes sorted by x, ties by node id; cut into `world` runs of equal node count, or of equal total weight;
//   * a rank holds every element with at least one owned vertex, hence a one-ring of ghost nodes;
//   * local numbering keeps the global order (ascending global id), local elements keep ascending global element id, so every
//     nodal sum is formed in the same order as on a single GPU and the owned rows of all ranks tile the single-GPU result bit for bit.
#include "../../include/wcfe.h"
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

struct wcfe_strip {
    int rank = 0, world = 1;
    std::vector<int> local_nodes;      // global id of every local node, ascending
    std::vector<int> tri_local;        // [3*m] local vertex ids
    std::vector<int> elem_gid;         // [m] global element ids, ascending
    std::vector<int> owned_local;      // local ids of the owned nodes
    std::vector<int> ghost_local;      // local ids of the ghost nodes
    std::vector<int> peer_rank, my_slot_at_peer, rows_from_peer;
    std::vector<int> export_node, export_peer, export_row;
};

namespace {

struct Part {
    std::vector<int> local_nodes;                       // ascending global ids
    std::vector<int> ghosts;                            // ascending global ids
    std::vector<int> neighbours;                        // ranks this strip exchanges rows with, ascending
    int local_of(int g) const { return (int)(std::lower_bound(local_nodes.begin(), local_nodes.end(), g) - local_nodes.begin()); }
};

// the nodes of every rank's strip (owned + ghosts) and its neighbour list; element lists only for `keep`
void build_parts(int nn, int ne, const int *tri, const std::vector<int> &owner, int world, int keep, std::vector<Part> &parts,
                 std::vector<int> &keep_elems) {
    parts.assign((size_t)world, Part());
    std::vector<std::vector<int>> nodes((size_t)world);
    for (int e = 0; e < ne; e++) {
        const int a = tri[3 * (size_t)e], b = tri[3 * (size_t)e + 1], c = tri[3 * (size_t)e + 2];
        int rs[3] = {owner[a], owner[b], owner[c]};
        std::sort(rs, rs + 3);
        for (int q = 0; q < 3; q++) {
            if (q && rs[q] == rs[q - 1]) continue;
            std::vector<int> &v = nodes[(size_t)rs[q]];
            v.push_back(a); v.push_back(b); v.push_back(c);
            if (rs[q] == keep) keep_elems.push_back(e);
        }
    }
    for (int r = 0; r < world; r++) {
        std::vector<int> &v = nodes[(size_t)r];
        std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
        Part &p = parts[(size_t)r];
        p.local_nodes.swap(v);
        for (int g : p.local_nodes) if (owner[g] != r) p.ghosts.push_back(g);
    }
    // r and q are neighbours when one holds a ghost the other owns (the relation is symmetric: both hold the shared element)
    std::vector<std::vector<char>> nb((size_t)world, std::vector<char>((size_t)world, 0));
    for (int r = 0; r < world; r++) for (int g : parts[(size_t)r].ghosts) { nb[(size_t)r][(size_t)owner[g]] = 1; nb[(size_t)owner[g]][(size_t)r] = 1; }
    for (int r = 0; r < world; r++) for (int q = 0; q < world; q++) if (q != r && nb[(size_t)r][(size_t)q]) parts[(size_t)r].neighbours.push_back(q);
    (void)nn;
}

}   // namespace

extern "C" {

int wcfe_strip_partition(int nn, const double *xyz, int ne, const int *tri, const double *weights, int world, int rank, wcfe_strip **out) {
    if (!out) return WCFE_ERR_ARG;
    *out = nullptr;
    if (nn <= 0 || ne <= 0 || !xyz || !tri || world < 1 || rank < 0 || rank >= world) return WCFE_ERR_ARG;
    for (size_t q = 0; q < 3 * (size_t)ne; q++) if (tri[q] < 0 || tri[q] >= nn) return WCFE_ERR_ARG;
    // owner of every node: position in the (x, id) order against the cut points
    std::vector<int> order((size_t)nn);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return xyz[3 * (size_t)a] < xyz[3 * (size_t)b]; });
    std::vector<long long> bounds((size_t)world + 1, 0);
    bounds[(size_t)world] = nn;
    if (!weights) {
        for (int r = 0; r <= world; r++) bounds[(size_t)r] = ((long long)nn * r) / world;
    } else {
        std::vector<double> cw((size_t)nn);
        double acc = 0.0;
        for (int k = 0; k < nn; k++) { acc += weights[order[(size_t)k]]; cw[(size_t)k] = acc; }
        for (int r = 1; r < world; r++)       // first position whose running weight reaches the r-th share (numpy.searchsorted, side='left')
            bounds[(size_t)r] = (long long)(std::lower_bound(cw.begin(), cw.end(), cw.back() * r / world) - cw.begin());
    }
    std::vector<int> owner((size_t)nn, 0);
    for (int r = 0; r < world; r++) for (long long k = bounds[(size_t)r]; k < bounds[(size_t)r + 1]; k++) owner[(size_t)order[(size_t)k]] = r;

    std::vector<Part> parts;
    std::vector<int> elems;
    build_parts(nn, ne, tri, owner, world, rank, parts, elems);
    const Part &me = parts[(size_t)rank];
    wcfe_strip *s = new wcfe_strip;
    s->rank = rank; s->world = world;
    s->local_nodes = me.local_nodes;
    s->elem_gid = elems;                                 // ascending: elements were visited in order
    s->tri_local.resize(3 * elems.size());
    for (size_t m = 0; m < elems.size(); m++) for (int j = 0; j < 3; j++) s->tri_local[3 * m + j] = me.local_of(tri[3 * (size_t)elems[m] + j]);
    for (size_t l = 0; l < me.local_nodes.size(); l++) (owner[(size_t)me.local_nodes[l]] == rank ? s->owned_local : s->ghost_local).push_back((int)l);
    // halo tables, neighbour by neighbour in ascending rank order (= the peer-slot order of wcfe_halo_connect)
    s->peer_rank = me.neighbours;
    for (size_t slot = 0; slot < me.neighbours.size(); slot++) {
        const int q = me.neighbours[slot];
        const Part &pq = parts[(size_t)q];
        s->my_slot_at_peer.push_back((int)(std::find(pq.neighbours.begin(), pq.neighbours.end(), rank) - pq.neighbours.begin()));
        int rows = 0;
        for (int g : me.ghosts) if (owner[(size_t)g] == q) rows++;
        s->rows_from_peer.push_back(rows);
        for (int g : pq.ghosts) {                        // q's ghosts that this rank owns, in q's ascending order
            if (owner[(size_t)g] != rank) continue;
            s->export_node.push_back(me.local_of(g)); s->export_peer.push_back((int)slot); s->export_row.push_back(pq.local_of(g));
        }
    }
    *out = s;
    return WCFE_OK;
}

int wcfe_strip_sizes(const wcfe_strip *s, int *n_local, int *n_elem, int *n_owned, int *n_ghost, int *n_peers, int *n_export) {
    if (!s) return WCFE_ERR_ARG;
    if (n_local) *n_local = (int)s->local_nodes.size();
    if (n_elem) *n_elem = (int)s->elem_gid.size();
    if (n_owned) *n_owned = (int)s->owned_local.size();
    if (n_ghost) *n_ghost = (int)s->ghost_local.size();
    if (n_peers) *n_peers = (int)s->peer_rank.size();
    if (n_export) *n_export = (int)s->export_node.size();
    return WCFE_OK;
}

static void put(const std::vector<int> &v, int *dst) { if (dst && !v.empty()) memcpy(dst, v.data(), sizeof(int) * v.size()); }

int wcfe_strip_mesh(const wcfe_strip *s, int *local_nodes, int *tri_local, int *elem_gid, int *owned_local, int *ghost_local) {
    if (!s) return WCFE_ERR_ARG;
    put(s->local_nodes, local_nodes); put(s->tri_local, tri_local); put(s->elem_gid, elem_gid); put(s->owned_local, owned_local); put(s->ghost_local, ghost_local);
    return WCFE_OK;
}

int wcfe_strip_halo(const wcfe_strip *s, int *peer_rank, int *my_slot_at_peer, int *rows_from_peer, int *export_node, int *export_peer, int *export_row) {
    if (!s) return WCFE_ERR_ARG;
    put(s->peer_rank, peer_rank); put(s->my_slot_at_peer, my_slot_at_peer); put(s->rows_from_peer, rows_from_peer);
    put(s->export_node, export_node); put(s->export_peer, export_peer); put(s->export_row, export_row);
    return WCFE_OK;
}

int wcfe_strip_free(wcfe_strip *s) {
    if (!s) return WCFE_ERR_ARG;
    delete s;
    return WCFE_OK;
}

}   // extern "C"
