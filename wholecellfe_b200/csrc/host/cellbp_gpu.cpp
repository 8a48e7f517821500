// cellbp_gpu -- stand-alone host shell with the reference's command line, script format and .vtk frame output
// (cellBP/main.cpp:636-740, :427-489; scriptinit.cpp; iomesh.cpp), driving the UL step on the GPU through the C ABI.
//
//   cellbp_gpu <scriptfile_cell> <scriptfile_subs> [outputbasedirectory] [--variant k1|k31|kMD] [--device N] [--quiet] [--no-eltopo]
//
// El Topo is not linked: the scripts must switch remeshing and collision handling off (perform_improvement 0,
// allow_topology_changes 0, collision_safety 0; scriptinit.cpp:57-65) -- with those flags the reference's step is exactly
// the El-Topo-free path this shell runs.  Use the adapter to keep El Topo in the loop.
// --no-eltopo accepts scripts that leave those switches on (the shipped scripts/cell_example.txt does) and runs them WITHOUT mesh
// improvement and collision handling, saying so: the result is the reference's as long as El Topo would have left the mesh alone --
// which it does for k1 on the shipped example, all 100 frames byte-identical (tests/test_host_shell.py) -- and is not otherwise
// (k31 / kMD remesh from step ~300 on).
#include "../../../include/wcfe.h"
#include "cellbp_io.h"
#include "cellbp_script.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <iostream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace {

// framestepper.h:14-76
struct FrameStepper {
    double frame_length, current_time = 0; int frame_count = 0, step_count = 0;
    explicit FrameStepper(double len) : frame_length(len) {}
    double get_step_length(double max_step) const { return (current_time + max_step > frame_length) ? frame_length - current_time : max_step; }
    bool done_frame() const { return current_time >= frame_length - 1e-7; }
    void advance_step(double dt) { current_time += dt; ++step_count; }
    void next_frame() { current_time = 0; step_count = 0; ++frame_count; }
};

std::string next_run_dir(const std::string &base) {            // main.cpp:494-533 (with the sprintf overlap of :528 fixed)
    int max_dir = 0;
    if (DIR *dp = opendir(base.c_str())) {
        while (dirent *ep = readdir(dp)) { int i; if (sscanf(ep->d_name, "run%04d", &i) > 0 && i > max_dir) max_dir = i; }
        closedir(dp);
    }
    char buf[32]; snprintf(buf, sizeof buf, "/run%04d", max_dir + 1);
    return base + buf;
}
void mkdir_p(const std::string &path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0777); }
        if (i < path.size()) cur += path[i];
    }
}
void copy_file(const std::string &from, const std::string &to) {
    FILE *a = fopen(from.c_str(), "rb"), *b = fopen(to.c_str(), "wb");
    if (a && b) { char buf[4096]; size_t n; while ((n = fread(buf, 1, sizeof buf, a)) > 0) fwrite(buf, 1, n, b); }
    if (a) fclose(a);
    if (b) fclose(b);
}
void ck(wcfe_ctx *c, int rc, const char *what) {
    if (rc != WCFE_OK) { std::fprintf(stderr, "cellbp_gpu: %s failed (%d): %s\n", what, rc, wcfe_last_error(c)); std::exit(1); }
}

}   // namespace

int main(int argc, char **argv) {
    std::cout << "\nCellBP (GPU UL step): cell biophysics on B200\n--------------------------------------------------\n\n";
    std::vector<std::string> pos;
    int variant = WCFE_K1, device = 0; bool quiet = false, no_eltopo = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--variant" && i + 1 < argc) { const std::string v = argv[++i]; variant = v == "k31" ? WCFE_K31 : (v == "kMD" ? WCFE_KMD : WCFE_K1); }
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--quiet") quiet = true;
        else if (a == "--no-eltopo") no_eltopo = true;
        else pos.push_back(a);
    }
    if (pos.size() < 2) {
        std::cout << "Usage: <executable> <scriptfile_cell> <scriptfile_subs> <outputbasedirectory> [--variant k1|k31|kMD] [--device N]\n";
        return 0;
    }
    const std::string base = pos.size() > 2 ? pos[2] : "./";
    try {
        const cellbp::ScriptInfo sc = cellbp::read_script(pos[0]), ss = cellbp::read_script(pos[1]);
        for (const cellbp::ScriptInfo *s : {&sc, &ss})
            if (s->perform_improvement || s->allow_topology_changes || s->collision_safety) {
                if (!no_eltopo)
                    throw std::runtime_error("this shell does not link El Topo: set perform_improvement 0, allow_topology_changes 0, collision_safety 0 "
                                             "in surftrack_parameters of both scripts, pass --no-eltopo to run them without it, or use the adapter of INTEGRATION.md");
                std::fprintf(stderr, "cellbp_gpu: --no-eltopo: %s asks for mesh improvement / topology changes / collision handling; El Topo is not linked, "
                                     "they are SKIPPED (same frames as the reference only while El Topo would leave the mesh alone)\n", s == &sc ? pos[0].c_str() : pos[1].c_str());
            }
        std::vector<double> xc, xs; std::vector<int> tc, ts;
        cellbp::read_ecomsol(sc.mesh_path, xc, tc);
        cellbp::read_ecomsol(ss.mesh_path, xs, ts);
        std::string out = sc.output_path_is_relative ? base + "/" + sc.output_path : sc.output_path;      // main.cpp:557-563
        mkdir_p(out);
        out = next_run_dir(out);
        mkdir_p(out);
        std::cout << "Output path: " << out << std::endl;
        copy_file(pos[0], out + "/aaa-script_cell.txt"); copy_file(pos[1], out + "/aaa-script_subs.txt");   // main.cpp:687-707

        wcfe_params prm; wcfe_default_params(&prm, variant);
        wcfe_ctx *ctx = nullptr;
        int rc = wcfe_create(&ctx, device, &prm);
        if (rc != WCFE_OK) { std::fprintf(stderr, "cellbp_gpu: wcfe_create failed (%d): %s -- there is no CPU fallback\n", rc, wcfe_last_error(nullptr)); return 1; }
        const int nc = (int)xc.size() / 3, ns = (int)xs.size() / 3;
        ck(ctx, wcfe_upload_mesh(ctx, WCFE_CELL, nc, xc.data(), (int)tc.size() / 3, tc.data()), "upload_mesh(cell)");
        ck(ctx, wcfe_upload_mesh(ctx, WCFE_SUBS, ns, xs.data(), (int)ts.size() / 3, ts.data()), "upload_mesh(subs)");
        ck(ctx, wcfe_set_rng(ctx, WCFE_RNG_GLIBC_REPLAY, 1), "set_rng");                                    // main.cpp:709 srand(1)

        FrameStepper fs(sc.frame_dt);
        double t = sc.curr_t_specified ? sc.curr_t : 0.0;
        const double max_t = sc.end_sim_t, kint_out = prm.kint_bond;
        std::vector<double> cA(nc), cC(nc), fa(3 * (size_t)nc), eac(6 * (size_t)nc), eas(6 * (size_t)ns);
        auto write_frame = [&](int frame) {
            ck(ctx, wcfe_download_state(ctx, WCFE_CELL, WCFE_F_X, xc.data()), "download x");
            ck(ctx, wcfe_download_state(ctx, WCFE_CELL, WCFE_F_CA, cA.data()), "download cA");
            ck(ctx, wcfe_download_state(ctx, WCFE_CELL, WCFE_F_CC, cC.data()), "download cC");
            ck(ctx, wcfe_download_state(ctx, WCFE_CELL, WCFE_F_FAU, fa.data()), "download fa_u");
            ck(ctx, wcfe_download_state(ctx, WCFE_CELL, WCFE_F_EA, eac.data()), "download ea");
            ck(ctx, wcfe_download_state(ctx, WCFE_SUBS, WCFE_F_EA, eas.data()), "download ea(subs)");
            char name[64];
            snprintf(name, sizeof name, "/cell_frame%04d.vtk", frame);
            cellbp::write_vtk_cell(out + name, t, xc, tc, cA, cC, fa, eac, kint_out);                       // main.cpp:445-447
            snprintf(name, sizeof name, "/subs_frame%04d.vtk", frame);
            cellbp::write_vtk_subs(out + name, t, xs, ts, eas, fa);                                          // main.cpp:448-451
        };
        write_frame(0);                                                                                     // main.cpp:716-722
        long steps = 0;
        bool running = true;
        while (running) {                                                                                   // run_simulation, main.cpp:480-489
            if (!quiet) std::cout << "\n --------------- frame " << fs.frame_count << " --------------- \n" << std::endl;
            while (!fs.done_frame()) {                                                                      // advance_frame, main.cpp:440-452
                const double dt = fs.get_step_length(sc.sim_dt);
                // advance_sim (main.cpp:168-418): with El Topo off one pass takes the whole dt, unless max_t cuts it short
                double accum = 0;
                while (accum < 0.99 * dt && t + accum < max_t) {
                    double cur = dt - accum;
                    if (max_t - t - accum < cur) cur = max_t - t - accum;
                    double tt = t;                                          // t_tot handed to the physics is m_curr_t at the start of advance_sim
                    ck(ctx, wcfe_step_n(ctx, 1, cur, &tt), "step");
                    accum += cur; steps++;
                }
                t += accum;                                                 // main.cpp:405
                if (t > max_t - 1e-7) running = false;                      // simulation.h:72-75
                fs.advance_step(dt);
            }
            // the reference rewrites the frame's file after every step (quirk q8): what survives is the state at the end of the frame
            write_frame(fs.frame_count);
            if (!quiet) {
                double d[8]; wcfe_get_diagnostics(ctx, d);
                std::cout << "total area " << d[0] << "\ncentroid " << d[1] << ", " << d[2] << ", " << d[3] << "\n";
            }
            fs.next_frame();
        }
        std::cout << "done: " << steps << " steps, t = " << t << ", " << wcfe_kernel_launches(ctx) << " kernel launches\n";
        wcfe_destroy(ctx);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "cellbp_gpu: %s\n", e.what());
        return 1;
    }
    return 0;
}
