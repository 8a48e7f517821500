// wcfe_rows.h -- host-only helper of build_patches (wcfe_api.cu); a header of its own so that the CPU test build
// (tests/hostmath) can exercise it without a CUDA toolchain.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace wcfe {
// Row numbering of a substrate patch's node table (k_subs_ring) against shared-memory bank conflicts.  The table is gathered: at
// fan position e the lanes of a warp read the rows of THEIR e-th neighbours, 16 bytes per lane, a quarter-warp (8 lanes) per
// wavefront when the eight rows fall into eight different 16-byte bank groups (row mod 8) -- with rows numbered as the nodes are,
// the eight behave like random numbers and a gather costs 2.2 wavefronts instead of one (ncu: 11 of 29 M wavefronts per launch, on
// the unit that bounds the kernel).  The access pattern is static, so the numbering is chosen for it: nodes in order of how often
// they are read, each into the residue class (mod 8, of limited capacity) in which it meets the fewest rows it is ever read
// together with; inside the class, the row whose residue mod 16 does the same for the 8-byte reads of a half-warp.
// jk: (j, k) local ids per fan entry, rptr: CSR over the owned nodes; returns row_of[local id] (a permutation of [0, NL)).
inline std::vector<unsigned short> ring_table_rows(const std::vector<unsigned short> &jk, const std::vector<unsigned short> &rptr, int P, int NL) {
    std::vector<unsigned short> row_of((size_t)NL);
    for (int l = 0; l < NL; l++) row_of[l] = (unsigned short)l;
    int maxval = 0;
    for (int l = 0; l < P; l++) maxval = std::max(maxval, (int)rptr[l + 1] - (int)rptr[l]);
    if (!maxval) return row_of;
    // groups of local ids one load instruction of one quarter-warp touches (J only where the walk really loads it)
    std::vector<std::vector<int>> groups;
    for (int e = 0; e < maxval; e++)
        for (int q = 0; q < P / 8; q++)
            for (int slot = 0; slot < 2; slot++) {
                std::vector<int> g;
                for (int t = 8 * q; t < 8 * q + 8; t++) {
                    const int r0 = rptr[t], n = rptr[t + 1] - r0;
                    if (e >= n) continue;
                    if (slot == 0 && e > 0 && jk[2 * (size_t)(r0 + e)] == jk[2 * (size_t)(r0 + e - 1) + 1]) continue;     // J == previous K: kept in registers
                    g.push_back(jk[2 * (size_t)(r0 + e) + slot]);
                }
                if (g.size() > 1) groups.push_back(std::move(g));
            }
    // the table fill: eight consecutive threads store their own rows, eight consecutive halo positions theirs
    for (int q = 0; q < NL / 8; q++) { std::vector<int> g; for (int t = 8 * q; t < 8 * q + 8 && t < NL; t++) g.push_back(t); groups.push_back(std::move(g)); }
    std::vector<int> freq((size_t)NL, 0);
    std::vector<std::vector<int>> of_node((size_t)NL);
    for (size_t gi = 0; gi < groups.size(); gi++) for (int v : groups[gi]) { freq[v]++; of_node[v].push_back((int)gi); }
    std::vector<int> order;
    for (int l = 0; l < NL; l++) if (freq[l]) order.push_back(l);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return freq[a] > freq[b]; });
    std::vector<int> row((size_t)NL, -1);
    std::vector<char> taken((size_t)NL, 0);
    std::vector<int> left(8, 0);
    for (int r = 0; r < NL; r++) left[r % 8]++;
    for (int v : order) {
        int cost8[8] = {0}, cost16[16] = {0};
        for (int gi : of_node[v])
            for (int u : groups[gi]) if (u != v && row[u] >= 0) { cost8[row[u] % 8]++; cost16[row[u] % 16]++; }
        int best = -1;
        for (int c = 0; c < 8; c++) if (left[c] && (best < 0 || cost8[c] < cost8[best] || (cost8[c] == cost8[best] && left[c] > left[best]))) best = c;
        int pick = -1;
        for (int r = best; r < NL; r += 8) if (!taken[r] && (pick < 0 || cost16[r % 16] < cost16[pick % 16])) pick = r;
        row[v] = pick; taken[pick] = 1; left[best]--;
    }
    int next = 0;
    for (int l = 0; l < NL; l++) {          // nodes nobody gathers (none on a closed mesh) take what is left
        if (row[l] >= 0) continue;
        while (taken[next]) next++;
        row[l] = next; taken[next] = 1;
    }
    for (int l = 0; l < NL; l++) row_of[l] = (unsigned short)row[l];
    return row_of;
}


// what the numbering is judged by: wavefronts of the 16-byte gathers and of the table fill (max multiplicity of a residue mod 8 over
// the eight rows a quarter-warp touches at once), for a given row_of
inline long long ring_table_wavefronts(const std::vector<unsigned short> &jk, const std::vector<unsigned short> &rptr, int P, int NL,
                                       const std::vector<unsigned short> &row_of) {
    long long total = 0;
    int maxval = 0;
    for (int l = 0; l < P; l++) maxval = std::max(maxval, (int)rptr[l + 1] - (int)rptr[l]);
    auto cost = [&](const std::vector<int> &g) { int m[8] = {0}, mx = 0; for (int v : g) mx = std::max(mx, ++m[row_of[v] % 8]); return mx; };
    for (int e = 0; e < maxval; e++)
        for (int q = 0; q < P / 8; q++)
            for (int slot = 0; slot < 2; slot++) {
                std::vector<int> g;
                for (int t = 8 * q; t < 8 * q + 8; t++) {
                    const int r0 = rptr[t], n = rptr[t + 1] - r0;
                    if (e >= n) continue;
                    if (slot == 0 && e > 0 && jk[2 * (size_t)(r0 + e)] == jk[2 * (size_t)(r0 + e - 1) + 1]) continue;
                    g.push_back(jk[2 * (size_t)(r0 + e) + slot]);
                }
                if (!g.empty()) total += cost(g);
            }
    for (int q = 0; q < NL / 8; q++) { std::vector<int> g; for (int t = 8 * q; t < 8 * q + 8; t++) g.push_back(t); total += cost(g); }
    return total;
}
}  // namespace wcfe
