// C interface over the host shell's parser / reader / writer so that the CPU test-suite can exercise them (ctypes).
#include "cellbp_io.h"
#include "cellbp_script.h"
#include <cstring>
#include <string>
#include <vector>

static std::string g_err;
static std::vector<double> g_xyz; static std::vector<int> g_tri;

extern "C" {
const char *cellbp_last_error(void) { return g_err.c_str(); }
// out9: frame_dt sim_dt end_sim_t curr_t curr_t_specified relative perform_improvement allow_topology_changes collision_safety
int cellbp_read_script(const char *path, double *out9, char *out_path, int out_path_cap, char *mesh_path, int mesh_path_cap) {
    try {
        const cellbp::ScriptInfo s = cellbp::read_script(path);
        out9[0] = s.frame_dt; out9[1] = s.sim_dt; out9[2] = s.end_sim_t; out9[3] = s.curr_t; out9[4] = s.curr_t_specified;
        out9[5] = s.output_path_is_relative; out9[6] = s.perform_improvement; out9[7] = s.allow_topology_changes; out9[8] = s.collision_safety;
        strncpy(out_path, s.output_path.c_str(), out_path_cap - 1); out_path[out_path_cap - 1] = 0;
        strncpy(mesh_path, s.mesh_path.c_str(), mesh_path_cap - 1); mesh_path[mesh_path_cap - 1] = 0;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int cellbp_read_ecomsol(const char *path, int *nn, int *ne) {
    try { cellbp::read_ecomsol(path, g_xyz, g_tri); *nn = (int)g_xyz.size() / 3; *ne = (int)g_tri.size() / 3; return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
void cellbp_get_mesh(double *xyz, int *tri) { memcpy(xyz, g_xyz.data(), g_xyz.size() * 8); memcpy(tri, g_tri.data(), g_tri.size() * 4); }
int cellbp_write_vtk_cell(const char *path, double time, int nn, const double *xyz, int ne, const int *tri, const double *cA, const double *cC,
                          const double *fa_u, const double *ea, double kint) {
    try {
        cellbp::write_vtk_cell(path, time, std::vector<double>(xyz, xyz + 3 * nn), std::vector<int>(tri, tri + 3 * ne), std::vector<double>(cA, cA + nn),
                               std::vector<double>(cC, cC + nn), std::vector<double>(fa_u, fa_u + 3 * nn), std::vector<double>(ea, ea + 6 * nn), kint);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int cellbp_write_vtk_subs(const char *path, double time, int nn, const double *xyz, int ne, const int *tri, const double *ea, int nfa, const double *fa_u) {
    try {
        cellbp::write_vtk_subs(path, time, std::vector<double>(xyz, xyz + 3 * nn), std::vector<int>(tri, tri + 3 * ne), std::vector<double>(ea, ea + 6 * nn),
                               std::vector<double>(fa_u, fa_u + 3 * nfa));
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
}
