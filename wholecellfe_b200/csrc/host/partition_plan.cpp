// partition_plan -- what a C++ / MPI host does before a partitioned run, with nothing but the C ABI's host-only entry points
// (include/wcfe.h, csrc/wcfe_partition.cu; no GPU needed): cost weights, the substrate cut (strips, or sectors about the cell's
// centroid), the rank's local strip and its wcfe_halo_connect tables, the cell's patch lists and its wcfe_cell_halo_connect tables.
//
//   partition_plan <cell mesh (ecomsol)> <substrate mesh (ecomsol)> <world> [strips|sectors] [margin]
//
// prints one line per rank: the sizes a rank would allocate and upload.  In a real host every rank runs the same calls for its own
// `rank`, uploads the strip with wcfe_upload_mesh, all-gathers the wcfe_halo_info / wcfe_cell_halo_info blobs and connects
// (INTEGRATION.md 3).  tests/test_partition.py holds the numbers against wholecellfe_b200/partition.py.
#include "../../../include/wcfe.h"
#include "cellbp_io.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

int main(int argc, char **argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: partition_plan <cell mesh> <substrate mesh> <world> [strips|sectors] [margin]\n"); return 2; }
    try {
        std::vector<double> xc, xs; std::vector<int> tc, ts;
        cellbp::read_ecomsol(argv[1], xc, tc);
        cellbp::read_ecomsol(argv[2], xs, ts);
        const int world = std::atoi(argv[3]);
        const bool sectors = argc > 4 && std::string(argv[4]) == "sectors";
        const double margin = argc > 5 ? std::atof(argv[5]) : 0.15;
        const int nc = (int)xc.size() / 3, ns = (int)xs.size() / 3, mc = (int)tc.size() / 3, ms = (int)ts.size() / 3;
        std::vector<double> w((size_t)ns), key((size_t)ns);
        if (wcfe_partition_weights(ns, xs.data(), nc, xc.data(), 2.0, 4.0, w.data()) != WCFE_OK) throw std::runtime_error("wcfe_partition_weights");
        double cx = 0, cy = 0;
        for (int i = 0; i < nc; i++) { cx += xc[3 * (size_t)i]; cy += xc[3 * (size_t)i + 1]; }
        cx /= nc; cy /= nc;
        for (int i = 0; i < ns; i++) key[(size_t)i] = sectors ? std::atan2(xs[3 * (size_t)i + 1] - cy, xs[3 * (size_t)i] - cx) : xs[3 * (size_t)i];
        std::vector<int> owner((size_t)ns);
        for (int rank = 0; rank < world; rank++) {
            wcfe_strip *strip = nullptr;
            if (wcfe_key_partition(ns, key.data(), ms, ts.data(), w.data(), world, rank, &strip) != WCFE_OK) throw std::runtime_error("wcfe_key_partition");
            int n_local, n_elem, n_owned, n_ghost, n_peers, n_export;
            wcfe_strip_sizes(strip, &n_local, &n_elem, &n_owned, &n_ghost, &n_peers, &n_export);
            wcfe_strip_owner(strip, owner.data());
            wcfe_cell_plan *plan = nullptr;
            if (wcfe_cell_partition(world, rank, ns, xs.data(), owner.data(), nc, xc.data(), mc, tc.data(), margin, &plan) != WCFE_OK) throw std::runtime_error("wcfe_cell_partition");
            int c_peers, c_patches, c_export, c_import;
            wcfe_cell_plan_sizes(plan, &c_peers, &c_patches, &c_export, &c_import);
            std::vector<int> pl((size_t)c_patches), po((size_t)c_patches);
            wcfe_cell_plan_lists(plan, nullptr, nullptr, pl.data(), po.data(), nullptr, nullptr, nullptr);
            int owned_patches = 0;
            for (int v : po) owned_patches += v;
            std::printf("rank %d: substrate local %d owned %d ghost %d elements %d peers %d export rows %d | cell patches %d owned %d export %d import %d\n",
                        rank, n_local, n_owned, n_ghost, n_elem, n_peers, n_export, c_patches, owned_patches, c_export, c_import);
            wcfe_cell_plan_free(plan);
            wcfe_strip_free(strip);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "partition_plan: %s\n", e.what());
        return 1;
    }
    return 0;
}
