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
#include <cmath>
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
    std::vector<int> owner;            // [nn global] rank that owns every node of the uncut mesh
};

struct wcfe_cell_plan {
    std::vector<int> peer_rank, my_slot_at_peer, patch_list, patch_owned, export_node, export_peer, import_node;
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

// nodes sorted by key (ties by node id) and cut into `world` runs: key = x gives vertical strips, key = the angle about the cell's
// centroid gives sectors (partition.py sector_partition: every rank then holds a wedge of the substrate and of the cell above it)
int wcfe_key_partition(int nn, const double *key, int ne, const int *tri, const double *weights, int world, int rank, wcfe_strip **out) {
    if (!out) return WCFE_ERR_ARG;
    *out = nullptr;
    if (nn <= 0 || ne <= 0 || !key || !tri || world < 1 || rank < 0 || rank >= world) return WCFE_ERR_ARG;
    for (size_t q = 0; q < 3 * (size_t)ne; q++) if (tri[q] < 0 || tri[q] >= nn) return WCFE_ERR_ARG;
    // owner of every node: position in the (key, id) order against the cut points
    std::vector<int> order((size_t)nn);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[(size_t)a] < key[(size_t)b]; });
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
    s->owner = owner;
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

int wcfe_strip_partition(int nn, const double *xyz, int ne, const int *tri, const double *weights, int world, int rank, wcfe_strip **out) {
    if (!out) return WCFE_ERR_ARG;
    *out = nullptr;
    if (nn <= 0 || !xyz) return WCFE_ERR_ARG;
    std::vector<double> key((size_t)nn);
    for (int i = 0; i < nn; i++) key[(size_t)i] = xyz[3 * (size_t)i];
    return wcfe_key_partition(nn, key.data(), ne, tri, weights, world, rank, out);
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

// Cost weights per substrate node for a balanced cut (partition.py cell_weights + cell_cost_weights, measured on B200: DESIGN.md 6):
// 1, plus under_cell for a node inside the cell's padded bounding box (it also reads its epoch stamp and walks anchor buckets), plus --
// when the cell is partitioned too, cell_cost > 0 -- cell_cost for every cell node charged to one substrate node of the raster cell
// (size 1.5 h) under it.
int wcfe_partition_weights(int nn_subs, const double *subs_xyz, int nn_cell, const double *cell_xyz, double under_cell, double cell_cost, double *weights) {
    if (nn_subs <= 0 || !subs_xyz || nn_cell <= 0 || !cell_xyz || !weights) return WCFE_ERR_ARG;
    double lo[2] = {cell_xyz[0], cell_xyz[1]}, hi[2] = {cell_xyz[0], cell_xyz[1]};
    for (int i = 1; i < nn_cell; i++) for (int c = 0; c < 2; c++) { lo[c] = std::min(lo[c], cell_xyz[3 * (size_t)i + c]); hi[c] = std::max(hi[c], cell_xyz[3 * (size_t)i + c]); }
    const double pad = std::max(2.0, 0.25 * std::max(hi[0] - lo[0], hi[1] - lo[1]));
    for (int i = 0; i < nn_subs; i++) {
        const double x = subs_xyz[3 * (size_t)i], y = subs_xyz[3 * (size_t)i + 1];
        const bool inside = x >= lo[0] - pad && x <= hi[0] + pad && y >= lo[1] - pad && y <= hi[1] + pad;
        weights[i] = 1.0 + under_cell * (inside ? 1.0 : 0.0);
    }
    if (!(cell_cost > 0.0)) return WCFE_OK;
    double xmin = subs_xyz[0], xmax = subs_xyz[0], ymin = subs_xyz[1], ymax = subs_xyz[1];
    for (int i = 1; i < nn_subs; i++) {
        xmin = std::min(xmin, subs_xyz[3 * (size_t)i]); xmax = std::max(xmax, subs_xyz[3 * (size_t)i]);
        ymin = std::min(ymin, subs_xyz[3 * (size_t)i + 1]); ymax = std::max(ymax, subs_xyz[3 * (size_t)i + 1]);
    }
    const double h = std::sqrt((xmax - xmin) * (ymax - ymin) / (double)std::max(1, nn_subs)), g = 1.5 * h;
    if (!(g > 0.0)) return WCFE_ERR_ARG;
    const long long nx = (long long)((xmax - xmin) / g) + 1, ny = (long long)((ymax - ymin) / g) + 1;
    if (nx * ny > (1ll << 28)) return WCFE_ERR_ARG;
    std::vector<int> rep((size_t)(nx * ny), -1);
    for (int i = 0; i < nn_subs; i++)          // the last node filed in a raster cell represents it
        rep[(size_t)((long long)((subs_xyz[3 * (size_t)i + 1] - ymin) / g) * nx + (long long)((subs_xyz[3 * (size_t)i] - xmin) / g))] = i;
    std::vector<double> extra((size_t)nn_subs, 0.0);
    for (int i = 0; i < nn_cell; i++) {
        long long ix = (long long)((cell_xyz[3 * (size_t)i] - xmin) / g), iy = (long long)((cell_xyz[3 * (size_t)i + 1] - ymin) / g);
        ix = std::min(std::max(ix, 0ll), nx - 1); iy = std::min(std::max(iy, 0ll), ny - 1);
        const int t = rep[(size_t)(iy * nx + ix)];
        if (t >= 0) extra[(size_t)t] += cell_cost;
    }
    for (int i = 0; i < nn_subs; i++) weights[i] = weights[i] + extra[(size_t)i];
    return WCFE_OK;
}

// ---- node numbering.  The patch kernels own 128 consecutive nodes per CTA and stage their one-ring in shared memory: a mesh numbered
// without locality (a mesh generator's advancing front, COMSOL's export order) has patches whose halo does not fit, and wcfe_upload_mesh
// then selects the general kernels.  wcfe_morton_order renumbers nodes along a Morton (Z-order) curve over the 16-bit quantised (x, y)
// and sorts the triangles by the key of their centroid -- what meshgen.py does for the synthetic configurations.
namespace {
unsigned long long spread16(unsigned long long v) {
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
}   // namespace
// perm[new] = old node id; tri_out[3*ne]: the triangles in the new numbering and order (vertex order kept, hence orientation)
int wcfe_morton_order(int nn, const double *xyz, int ne, const int *tri, int *perm, int *tri_out) {
    if (nn <= 0 || !xyz || ne < 0 || (ne > 0 && !tri) || !perm || (ne > 0 && !tri_out)) return WCFE_ERR_ARG;
    for (size_t q = 0; q < 3 * (size_t)ne; q++) if (tri[q] < 0 || tri[q] >= nn) return WCFE_ERR_ARG;
    double lo[2] = {xyz[0], xyz[1]}, hi[2] = {xyz[0], xyz[1]};
    for (int i = 1; i < nn; i++) for (int c = 0; c < 2; c++) { lo[c] = std::min(lo[c], xyz[3 * (size_t)i + c]); hi[c] = std::max(hi[c], xyz[3 * (size_t)i + c]); }
    const double span[2] = {std::max(hi[0] - lo[0], 1e-300), std::max(hi[1] - lo[1], 1e-300)};
    auto key_of = [&](double x, double y) {
        const long long qx = std::min((long long)((x - lo[0]) / span[0] * 65535.0), 65535ll), qy = std::min((long long)((y - lo[1]) / span[1] * 65535.0), 65535ll);
        return spread16((unsigned long long)qx) | (spread16((unsigned long long)qy) << 1);
    };
    std::vector<unsigned long long> key((size_t)nn);
    for (int i = 0; i < nn; i++) key[(size_t)i] = key_of(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1]);
    std::iota(perm, perm + nn, 0);
    std::stable_sort(perm, perm + nn, [&](int a, int b) { return key[(size_t)a] < key[(size_t)b]; });
    std::vector<int> inv((size_t)nn);
    for (int n = 0; n < nn; n++) inv[(size_t)perm[n]] = n;
    std::vector<unsigned long long> tkey((size_t)ne);
    std::vector<int> torder((size_t)ne);
    for (int e = 0; e < ne; e++) {
        const double *a = xyz + 3 * (size_t)tri[3 * (size_t)e], *b = xyz + 3 * (size_t)tri[3 * (size_t)e + 1], *c = xyz + 3 * (size_t)tri[3 * (size_t)e + 2];
        tkey[(size_t)e] = key_of(((a[0] + b[0]) + c[0]) / 3.0, ((a[1] + b[1]) + c[1]) / 3.0);
    }
    std::iota(torder.begin(), torder.end(), 0);
    std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) { return tkey[(size_t)a] < tkey[(size_t)b]; });
    for (int n = 0; n < ne; n++) for (int j = 0; j < 3; j++) tri_out[3 * (size_t)n + j] = inv[(size_t)tri[3 * (size_t)torder[(size_t)n] + j]];
    return WCFE_OK;
}

// ---- ensembles (BASELINE configs[4]; the reference's mechanism is a job array of independent processes, s_sim_031021_1.sub): the
// parameter sweep and its shard over ranks, deterministic from the cell id so that every rank rebuilds the table without communication
// (ensemble.py sweep_grid / partition).  Cell `id` of n_cells: laws (WCFE_K1 / _K31 / _KMD ids) x n_rho log-spaced ligand densities
// rho_i_max in [rho_min, rho_max] (UL:1231), n_rho = round(sqrt(n_cells / n_laws)), law-major; the seed counts how often the grid
// has been passed through (from 1).
int wcfe_sweep_cell(int n_cells, int n_laws, const int *laws, double rho_min, double rho_max, int id, wcfe_params *params, unsigned int *seed) {
    if (n_cells <= 0 || n_laws <= 0 || !laws || id < 0 || id >= n_cells || !params || !(rho_min > 0.0) || !(rho_max >= rho_min)) return WCFE_ERR_ARG;
    const int n_rho = std::max(1, (int)std::nearbyint(std::sqrt((double)n_cells / (double)n_laws)));
    const int per_pass = n_laws * n_rho, k = id % per_pass, law = laws[k / n_rho], j = k % n_rho;
    int rc = wcfe_default_params(params, law);
    if (rc != WCFE_OK) return rc;
    double rho = rho_min;
    if (n_rho > 1) rho = j == 0 ? rho_min : (j == n_rho - 1 ? rho_max : rho_min * std::pow(rho_max / rho_min, (double)j / (double)(n_rho - 1)));
    params->rho_i_max = rho;
    if (seed) *seed = 1u + (unsigned int)(id / per_pass);
    return WCFE_OK;
}
// contiguous, balanced shard [lo, hi) of the cell ids for `rank` (sizes differ by at most one)
int wcfe_ensemble_shard(int n_cells, int world, int rank, int *lo, int *hi) {
    if (n_cells < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return WCFE_ERR_ARG;
    const int base = n_cells / world, extra = n_cells % world;
    *lo = rank * base + std::min(rank, extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
    return WCFE_OK;
}

int wcfe_strip_owner(const wcfe_strip *s, int *owner) {
    if (!s || !owner) return WCFE_ERR_ARG;
    put(s->owner, owner);
    return WCFE_OK;
}

// ---- the CELL of a partitioned run: which 128-node patches a rank computes, owns, pushes and receives, for ANY partition of the
// substrate given as a node -> rank map (strips: wcfe_strip_owner).  The lists wcfe_cell_halo_connect wants; the same rules, raster and
// arithmetic as wholecellfe_b200/partition.py cell_partition + cell_connect_plan (tests/test_partition.py holds the two against each
// other, list for list):
//   * an anchor lies within `margin` of its cell node and a rank's local substrate triangles reach one edge length beyond its owned
//     nodes, so a rank computes every patch with a node in a raster cell (size margin + 1.5 h, h the mean substrate spacing) that --
//     after dilation by one cell -- holds a substrate node the rank owns; a patch on a cut is computed by every rank it touches;
//   * a patch is OWNED by the lowest rank that owns substrate in the raster cell under the patch's centroid: that rank publishes the
//     patch's area / centroid partial sums and pushes its rows to the ranks that read them as halo rows without computing them;
//   * a rank with no patch at all takes the one nearest to its strip (every rank runs the step's kernels).
namespace {
// the summation order of numpy's float64 add (pairwise, blocks of 128 with eight accumulators): the centroids decide raster cells, and
// the lists are meant to equal partition.py's bit for bit
double pairwise_sum(const double *a, long long n) {
    if (n < 8) { double r = 0.0; for (long long i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; k++) r[k] = a[k];
        long long i = 8;
        for (; i < n - (n % 8); i += 8) for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    long long n2 = n / 2; n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}
constexpr int kCellPatch = 128;      // owned nodes per cell patch (wcfe_patch.cuh kPatch)
}   // namespace

int wcfe_cell_partition(int world, int rank, int nn_subs, const double *subs_xyz, const int *subs_owner,
                        int nn_cell, const double *cell_xyz, int ne_cell, const int *cell_tri, double margin, wcfe_cell_plan **out) {
    if (!out) return WCFE_ERR_ARG;
    *out = nullptr;
    if (world < 1 || world > 32 || rank < 0 || rank >= world || nn_subs <= 0 || !subs_xyz || !subs_owner || nn_cell <= 0 || !cell_xyz ||
        ne_cell <= 0 || !cell_tri || !(margin >= 0.0)) return WCFE_ERR_ARG;
    for (int i = 0; i < nn_subs; i++) if (subs_owner[i] < 0 || subs_owner[i] >= world) return WCFE_ERR_ARG;
    for (size_t q = 0; q < 3 * (size_t)ne_cell; q++) if (cell_tri[q] < 0 || cell_tri[q] >= nn_cell) return WCFE_ERR_ARG;
    double xmin = subs_xyz[0], xmax = subs_xyz[0], ymin = subs_xyz[1], ymax = subs_xyz[1];
    for (int i = 1; i < nn_subs; i++) {
        xmin = std::min(xmin, subs_xyz[3 * (size_t)i]); xmax = std::max(xmax, subs_xyz[3 * (size_t)i]);
        ymin = std::min(ymin, subs_xyz[3 * (size_t)i + 1]); ymax = std::max(ymax, subs_xyz[3 * (size_t)i + 1]);
    }
    const double h = std::sqrt((xmax - xmin) * (ymax - ymin) / (double)std::max(1, nn_subs));
    const double g = margin + 1.5 * h;
    if (!(g > 0.0)) return WCFE_ERR_ARG;
    const long long nx = (long long)((xmax - xmin) / g) + 1, ny = (long long)((ymax - ymin) / g) + 1;
    if (nx * ny > (1ll << 28)) return WCFE_ERR_ARG;
    std::vector<unsigned int> mask((size_t)(nx * ny), 0u), dil;
    for (int i = 0; i < nn_subs; i++) {
        const long long ix = (long long)((subs_xyz[3 * (size_t)i] - xmin) / g), iy = (long long)((subs_xyz[3 * (size_t)i + 1] - ymin) / g);
        mask[(size_t)(iy * nx + ix)] |= 1u << subs_owner[i];
    }
    dil = mask;
    for (long long y = 0; y < ny; y++) for (long long x = 0; x < nx; x++) {
        unsigned int m = 0;
        for (long long dy = -1; dy <= 1; dy++) for (long long dx = -1; dx <= 1; dx++) {
            const long long yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < ny && xx >= 0 && xx < nx) m |= mask[(size_t)(yy * nx + xx)];
        }
        dil[(size_t)(y * nx + x)] = m;
    }
    auto cell_of = [&](double px, double py) {
        long long ix = (long long)((px - xmin) / g), iy = (long long)((py - ymin) / g);
        ix = std::min(std::max(ix, 0ll), nx - 1); iy = std::min(std::max(iy, 0ll), ny - 1);
        return (size_t)(iy * nx + ix);
    };
    const int npatch = (nn_cell + kCellPatch - 1) / kCellPatch;
    std::vector<unsigned int> pbits((size_t)npatch, 0u);
    std::vector<double> cx((size_t)npatch), cy((size_t)npatch), col((size_t)kCellPatch);
    std::vector<int> powner((size_t)npatch, 0);
    for (int p = 0; p < npatch; p++) {
        const int b = p * kCellPatch, e = std::min(nn_cell, b + kCellPatch), n = e - b;
        for (int i = b; i < e; i++) pbits[(size_t)p] |= dil[cell_of(cell_xyz[3 * (size_t)i], cell_xyz[3 * (size_t)i + 1])];
        for (int c = 0; c < 2; c++) {
            for (int i = b; i < e; i++) col[(size_t)(i - b)] = cell_xyz[3 * (size_t)i + c];
            const double sum = n > 1 ? col[0] + pairwise_sum(col.data() + 1, n - 1) : col[0];      // numpy add.reduceat: first element, then the rest
            (c == 0 ? cx : cy)[(size_t)p] = sum / (double)n;
        }
        const size_t cc = cell_of(cx[(size_t)p], cy[(size_t)p]);
        unsigned int cb = mask[cc];
        if (!cb) cb = dil[cc];
        if (!cb) cb = 1u;
        int o = 0;
        while (!((cb >> o) & 1u)) o++;
        powner[(size_t)p] = o;
    }
    std::vector<std::vector<char>> listed((size_t)world, std::vector<char>((size_t)npatch, 0));
    for (int r = 0; r < world; r++) {
        bool any = false;
        for (int p = 0; p < npatch; p++) {
            listed[(size_t)r][(size_t)p] = (char)(((pbits[(size_t)p] >> r) & 1u) || powner[(size_t)p] == r);
            any = any || listed[(size_t)r][(size_t)p];
        }
        if (!any) {
            std::vector<double> ox, oy;
            for (int i = 0; i < nn_subs; i++) if (subs_owner[i] == r) { ox.push_back(subs_xyz[3 * (size_t)i]); oy.push_back(subs_xyz[3 * (size_t)i + 1]); }
            if (ox.empty()) return WCFE_ERR_ARG;          // a rank that owns no substrate at all
            const double px = (0.0 + pairwise_sum(ox.data(), (long long)ox.size())) / (double)ox.size();
            const double py = (0.0 + pairwise_sum(oy.data(), (long long)oy.size())) / (double)oy.size();
            int best = 0; double bd = 0.0;
            for (int p = 0; p < npatch; p++) {
                const double dx = cx[(size_t)p] - px, dy = cy[(size_t)p] - py, d2 = dx * dx + dy * dy;
                if (p == 0 || d2 < bd) { bd = d2; best = p; }
            }
            listed[(size_t)r][(size_t)best] = 1;
        }
    }
    // halo rows a rank needs and does not compute: node b of an element edge (a, b) with a in a listed patch and b in one that is not
    std::vector<std::vector<int>> imports((size_t)world);
    for (int e = 0; e < ne_cell; e++) {
        const int v[3] = {cell_tri[3 * (size_t)e], cell_tri[3 * (size_t)e + 1], cell_tri[3 * (size_t)e + 2]};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            if (i == j) continue;
            const int pa = v[i] / kCellPatch, pb = v[j] / kCellPatch;
            if (pa == pb) continue;
            for (int r = 0; r < world; r++) if (listed[(size_t)r][(size_t)pa] && !listed[(size_t)r][(size_t)pb]) imports[(size_t)r].push_back(v[j]);
        }
    }
    for (auto &v : imports) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    wcfe_cell_plan *pl = new wcfe_cell_plan;
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        const int slot = (int)pl->peer_rank.size();
        pl->peer_rank.push_back(q);
        pl->my_slot_at_peer.push_back(rank < q ? rank : rank - 1);
        for (int node : imports[(size_t)q]) if (powner[(size_t)(node / kCellPatch)] == rank) { pl->export_node.push_back(node); pl->export_peer.push_back(slot); }
    }
    for (int p = 0; p < npatch; p++) if (listed[(size_t)rank][(size_t)p]) { pl->patch_list.push_back(p); pl->patch_owned.push_back(powner[(size_t)p] == rank ? 1 : 0); }
    pl->import_node = imports[(size_t)rank];
    *out = pl;
    return WCFE_OK;
}

int wcfe_cell_plan_sizes(const wcfe_cell_plan *p, int *n_peers, int *n_patches, int *n_export, int *n_import) {
    if (!p) return WCFE_ERR_ARG;
    if (n_peers) *n_peers = (int)p->peer_rank.size();
    if (n_patches) *n_patches = (int)p->patch_list.size();
    if (n_export) *n_export = (int)p->export_node.size();
    if (n_import) *n_import = (int)p->import_node.size();
    return WCFE_OK;
}

int wcfe_cell_plan_lists(const wcfe_cell_plan *p, int *peer_rank, int *my_slot_at_peer, int *patch_list, int *patch_owned, int *export_node,
                         int *export_peer, int *import_node) {
    if (!p) return WCFE_ERR_ARG;
    put(p->peer_rank, peer_rank); put(p->my_slot_at_peer, my_slot_at_peer); put(p->patch_list, patch_list); put(p->patch_owned, patch_owned);
    put(p->export_node, export_node); put(p->export_peer, export_peer); put(p->import_node, import_node);
    return WCFE_OK;
}

int wcfe_cell_plan_free(wcfe_cell_plan *p) {
    if (!p) return WCFE_ERR_ARG;
    delete p;
    return WCFE_OK;
}

}   // extern "C"
