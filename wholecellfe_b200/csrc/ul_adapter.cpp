// ul_adapter.cpp -- the reference's UL entry points (the five main.cpp calls, and the three drivers it declares and leaves commented
// out) on top of the C ABI (include/wcfe.h).
//
// Compile this translation unit against the REFERENCE's own headers (common/vec.h, eltopo3d/nondestructivetrimesh.h,
// cellBP/Eigen) and link it, together with libwcfe.so, in place of cellBP/UL_growth_explicit.o: the unmodified
// cellBP/main.cpp then runs its explicit updated-Lagrangian step on the GPU.  The prototypes below are
// the ones main.cpp re-declares by hand (wholeCellFE_k1/cellBP/main.cpp:72-84) and calls from advance_sim()
// (main.cpp:224, :293, :329, :354, :360, :361).
//
// Ownership follows the reference: the caller (SurfTrack) owns every array; V_t, A_t, ea_t, c_C, c_A, fa_u are mutated in
// place; V_hlf is the caller's temporary.  Because El Topo may remesh between steps, the host-owned nodal state is
// re-uploaded every step and connectivity whenever it changed.  All physics runs on the device; this file only moves data.
//
//   -DWCFE_VARIANT=WCFE_K1 | WCFE_K31 | WCFE_KMD   selects which reference tree is being replaced (default WCFE_K1).
#include "wcfe.h"
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <Eigen/Dense>
#include "vec.h"
#include "nondestructivetrimesh.h"

#ifndef WCFE_VARIANT
#define WCFE_VARIANT WCFE_K1
#endif

namespace {

struct MeshSig { size_t nv = 0, nt = 0; unsigned long long hash = 0; };

struct Adapter {
    wcfe_ctx *ctx = nullptr;
    MeshSig sig[2];
    std::vector<double> buf;          // AoS scratch
    double dt_real = 0;
    bool cell_pending = false;        // updateFA seen, device step not yet run
    int ndrv = 0;                     // driver calls since the last substrate update (main.cpp: cell :224, substrate :293, cell again :329)
} g;

// WCFE_ADAPTER_PROFILE=1: wall time spent inside each of the five entry points, printed at exit (where the whole application's
// step goes: bench.py other_configs.whole_app)
struct Prof {
    double t[5] = {0, 0, 0, 0, 0}; long n[5] = {0, 0, 0, 0, 0}; bool on = false;
    Prof() { on = std::getenv("WCFE_ADAPTER_PROFILE") != nullptr; }
    ~Prof() {
        if (!on) return;
        const char *name[5] = {"driverVelocities_Explicit", "driverVelocities_Biaxial", "updateFA", "updateForces_Explicit(cell)", "updateForces_Explicit(subs)"};
        double tot = 0;
        for (int q = 0; q < 5; q++) tot += t[q];
        std::fprintf(stderr, "wcfe adapter profile: %.3f ms per step inside the adapter (%ld steps)\n", n[2] ? 1e3 * tot / n[2] : 0.0, n[2]);
        for (int q = 0; q < 5; q++) std::fprintf(stderr, "  %-30s %8ld calls %9.3f ms/call\n", name[q], n[q], n[q] ? 1e3 * t[q] / n[q] : 0.0);
    }
} g_prof;
struct ProfScope {
    int q; std::chrono::steady_clock::time_point t0;
    explicit ProfScope(int q_) : q(q_) { if (g_prof.on) t0 = std::chrono::steady_clock::now(); }
    ~ProfScope() { if (g_prof.on) { g_prof.t[q] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); g_prof.n[q]++; } }
};

void die(const char *where, int rc) {
    std::fprintf(stderr, "wcfe adapter: %s failed (%d): %s\n", where, rc, wcfe_last_error(g.ctx));
    std::abort();                     // the reference is fail-stop as well (assert / exit)
}
#define OK(call) do { int rc_ = (call); if (rc_ != WCFE_OK) die(#call, rc_); } while (0)

void ensure_context() {
    if (g.ctx) return;
    wcfe_params p;
    OK(wcfe_default_params(&p, WCFE_VARIANT));
    int dev = 0;
    if (const char *e = std::getenv("WCFE_DEVICE")) dev = std::atoi(e);
    int rc = wcfe_create(&g.ctx, dev, &p);
    if (rc != WCFE_OK) { std::fprintf(stderr, "wcfe adapter: wcfe_create failed (%d): %s\n", rc, wcfe_last_error(nullptr)); std::abort(); }
    unsigned int seed = 1;            // main.cpp:709 srand(1); the per-step srand(time(NULL)) of UL:1460 is dropped
    if (const char *e = std::getenv("WCFE_SEED")) seed = (unsigned int)std::strtoul(e, nullptr, 10);
    OK(wcfe_set_rng(g.ctx, WCFE_RNG_GLIBC_REPLAY, seed));
}

const double *flat(const std::vector<Vec3d> &v) { return v.empty() ? nullptr : &v[0][0]; }   // Vec3d == 3 contiguous doubles
double *flat(std::vector<Vec3d> &v) { return v.empty() ? nullptr : &v[0][0]; }

// (re)upload connectivity when the mesh changed (remeshing: SurfTrack::improve_mesh / defrag_mesh).  with_coords: the signature
// also covers the coordinates -- the SUBSTRATE's device copy (positions, point-location grid, patch tables) is refreshed only
// here, and El Topo's smoothing / collision resolution may move its vertices without touching the connectivity; the cell's
// positions travel every step anyway.
void sync_mesh(int surface, const std::vector<Vec3d> &x, const NonDestructiveTriMesh &mesh, bool with_coords) {
    MeshSig s;
    s.nv = x.size(); s.nt = mesh.num_triangles();
    unsigned long long h = 1469598103934665603ull;
    for (size_t t = 0; t < s.nt; t++) {
        const Vec3st &tri = mesh.get_triangle(t);
        for (int j = 0; j < 3; j++) { h ^= (unsigned long long)tri[j]; h *= 1099511628211ull; }
    }
    if (with_coords) {
        for (size_t v = 0; v < s.nv; v++) for (int j = 0; j < 3; j++) {
            unsigned long long bits; const double d = x[v][j]; std::memcpy(&bits, &d, 8);
            h ^= bits; h *= 1099511628211ull;
        }
    }
    s.hash = h;
    MeshSig &old = g.sig[surface];
    if (old.nv == s.nv && old.nt == s.nt && old.hash == s.hash) return;
    std::vector<int> tris(3 * s.nt);
    for (size_t t = 0; t < s.nt; t++) { const Vec3st &tri = mesh.get_triangle(t); for (int j = 0; j < 3; j++) tris[3 * t + j] = (int)tri[j]; }
    OK(wcfe_upload_mesh(g.ctx, surface, (int)s.nv, flat(x), (int)s.nt, tris.data()));
    old = s;
}

// the reference keeps the strain as one heap-allocated Eigen 6-vector per node: flatten / scatter it around the batched transfers
const double *flat_ea(const std::vector<Eigen::VectorXd> &ea) {
    g.buf.resize(6 * ea.size() + 1);
    for (size_t i = 0; i < ea.size(); i++) for (int q = 0; q < 6; q++) g.buf[6 * i + q] = ea[i](q);
    return g.buf.data();
}
void scatter_ea(std::vector<Eigen::VectorXd> &ea) {
    for (size_t i = 0; i < ea.size(); i++) for (int q = 0; q < 6; q++) ea[i](q) = g.buf[6 * i + q];
}

}   // namespace

// main.cpp:224, :329
void driverVelocities_Explicit(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt) {
    ProfScope ps(0);
    ensure_context();
    g.ndrv++;
    OK(wcfe_driver_velocities(g.ctx, WCFE_CELL, (int)X_t.size(), flat(X_t), flat(V_t), flat(A_t), dt, 0.0, flat(V_t_hlf)));
}

// main.cpp:293
void driverVelocities_Biaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot) {
    ProfScope ps(1);
    ensure_context();
    g.ndrv++;
    OK(wcfe_driver_velocities(g.ctx, WCFE_SUBS, (int)X_t.size(), flat(X_t), flat(V_t), flat(A_t), dt, t_tot, flat(V_t_hlf)));
}

// The drivers main.cpp declares (:73, :74, :76) and keeps commented out at its two call sites (:217-220, :291-295).  Their signature
// does not say which surface they serve; a maintainer who enables one puts it at one of the two call sites, so the position in
// main.cpp's call order decides: the second driver call of a step is the substrate's (:293), the first and a third (:329) the cell's.
// On the substrate the next update uses this V_hlf (UL:499); on the cell the update refuses (its drag force reads V_hlf, UL:571-573).
namespace {
void driver_of_kind(int kind, const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot) {
    ensure_context();
    const int surface = (g.ndrv == 1) ? WCFE_SUBS : WCFE_CELL;
    ProfScope ps(surface);
    g.ndrv++;
    OK(wcfe_driver_velocities_kind(g.ctx, surface, kind, (int)X_t.size(), flat(X_t), flat(V_t), flat(A_t), dt, t_tot, flat(V_t_hlf)));
}
}   // namespace
void driverVelocities_Uniaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot) {
    driver_of_kind(WCFE_DRV_UNIAXIAL, X_t, V_t, V_t_hlf, A_t, dt, t_tot);
}
void driverVelocities_OffBiaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot) {
    driver_of_kind(WCFE_DRV_OFFBIAXIAL, X_t, V_t, V_t_hlf, A_t, dt, t_tot);
}
void driverVelocities_BulgeTest(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot) {
    driver_of_kind(WCFE_DRV_BULGETEST, X_t, V_t, V_t_hlf, A_t, dt, t_tot);
}

// main.cpp:354.  The device step needs V, A, ea too, which the reference only hands to updateForces_Explicit; main.cpp does
// not look at c_C / fa_u between the two calls (it passes the same vectors on), so positions and bond state are uploaded
// here and the step itself runs inside updateForces_Explicit(cell), which then writes c_C, c_A, fa_u back as well.
void updateFA(std::vector<double> &c_A, std::vector<double> &c_C, std::vector<Vec3d> &fa_u, double dt_real,
              const std::vector<Vec3d> &X_t_cell, const std::vector<Vec3d> &x_t_cell, const NonDestructiveTriMesh &mesh_cell,
              const std::vector<Vec3d> &X_t_subs, const std::vector<Vec3d> &x_t_subs, const NonDestructiveTriMesh &mesh_subs, double kint) {
    (void)kint;                       // kint is fixed by the variant (main.cpp:350-354)
    ProfScope ps(2);
    ensure_context();
    // the device step assumes a FROZEN substrate (the literal 0* of main.cpp:299: x_t_subs == X_t_subs, F == I, subs_u == 0).  A
    // host integrator that moved a substrate vertex within this step (collision resolution) has left that regime: say so instead
    // of silently computing with stale positions
    if (x_t_subs.size() != X_t_subs.size()) { std::fprintf(stderr, "wcfe adapter: substrate changed size inside a step\n"); std::abort(); }
    for (size_t v = 0; v < x_t_subs.size(); v++)
        if (x_t_subs[v][0] != X_t_subs[v][0] || x_t_subs[v][1] != X_t_subs[v][1] || x_t_subs[v][2] != X_t_subs[v][2]) {
            std::fprintf(stderr, "wcfe adapter: substrate vertex %zu moved within the step (x_t != X_t): a moving substrate is not supported by the frozen-substrate device path (WCFE_ERR_UNSUPPORTED)\n", v);
            std::abort();
        }
    sync_mesh(WCFE_SUBS, X_t_subs, mesh_subs, true);     // vertices moved BETWEEN steps (smoothing): fresh positions, grid and tables
    sync_mesh(WCFE_CELL, X_t_cell, mesh_cell, false);
    {   // one pinned transfer for the five fields (X_prev after x: an upload of x alone would also set X_prev := x)
        const int f[5] = {WCFE_F_X, WCFE_F_XPREV, WCFE_F_CA, WCFE_F_CC, WCFE_F_FAU};
        const double *h[5] = {flat(x_t_cell), flat(X_t_cell), c_A.data(), c_C.data(), flat(fa_u)};
        OK(wcfe_upload_states(g.ctx, WCFE_CELL, 5, f, h));
    }
    OK(wcfe_mark_positions_committed(g.ctx));
    g.dt_real = dt_real;
    g.cell_pending = true;
}

// main.cpp:360
void updateForces_Explicit(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf,
                           std::vector<Vec3d> &A_t, std::vector<Eigen::VectorXd> &ea_t, std::vector<double> &c_A, std::vector<double> &c_B,
                           std::vector<double> &c_C, std::vector<Vec3d> &fa_u, double dt, double dt_real, const NonDestructiveTriMesh &mesh, double t_tot) {
    // V_t_hlf is not uploaded: the device forms it from V_t, A_t and the dt of the LAST driverVelocities_Explicit call, which
    // wcfe_driver_velocities recorded -- the caller's vector bit for bit, also when El Topo shortened the step and main.cpp
    // re-ran the driver with the smaller dt (main.cpp:329) while `dt` here is still ori_dt (main.cpp:360)
    (void)X_t; (void)x_t; (void)V_t_hlf; (void)c_B; (void)mesh;
    ProfScope ps(3);
    if (!g.cell_pending) { std::fprintf(stderr, "wcfe adapter: updateForces_Explicit(cell) without a preceding updateFA\n"); std::abort(); }
    {
        const int f[3] = {WCFE_F_V, WCFE_F_A, WCFE_F_EA};
        const double *h[3] = {flat(V_t), flat(A_t), flat_ea(ea_t)};
        OK(wcfe_upload_states(g.ctx, WCFE_CELL, 3, f, h));
    }
    OK(wcfe_update_cell(g.ctx, dt, dt_real, t_tot));
    {
        const int f[6] = {WCFE_F_V, WCFE_F_A, WCFE_F_EA, WCFE_F_CA, WCFE_F_CC, WCFE_F_FAU};
        double *h[6] = {flat(V_t), flat(A_t), g.buf.data(), c_A.data(), c_C.data(), flat(fa_u)};
        OK(wcfe_download_states(g.ctx, WCFE_CELL, 6, f, h));
        scatter_ea(ea_t);
    }
    g.cell_pending = false;
}

// main.cpp:361
void updateForces_Explicit(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf,
                           std::vector<Vec3d> &A_t, std::vector<Eigen::VectorXd> &ea_t, double dt, double dt_real, const NonDestructiveTriMesh &mesh,
                           double t_tot, std::vector<Vec3d> &fa_u, const std::vector<double> &c_C, const std::vector<Vec3d> &x_cell,
                           const NonDestructiveTriMesh &mesh_cell) {
    (void)X_t; (void)x_t; (void)V_t_hlf; (void)mesh; (void)fa_u; (void)c_C; (void)x_cell; (void)mesh_cell;   // already on the device
    ProfScope ps(4);
    {
        const int f[3] = {WCFE_F_V, WCFE_F_A, WCFE_F_EA};
        const double *h[3] = {flat(V_t), flat(A_t), flat_ea(ea_t)};
        OK(wcfe_upload_states(g.ctx, WCFE_SUBS, 3, f, h));
    }
    OK(wcfe_update_subs(g.ctx, dt, dt_real, t_tot));
    g.ndrv = 0;
    {
        const int f[3] = {WCFE_F_V, WCFE_F_A, WCFE_F_EA};
        double *h[3] = {flat(V_t), flat(A_t), g.buf.data()};
        OK(wcfe_download_states(g.ctx, WCFE_SUBS, 3, f, h));
        scatter_ea(ea_t);
    }
}

// test hooks (used by the adapter parity test; harmless in production)
extern "C" void wcfe_adapter_reset(void) {
    if (g.ctx) wcfe_destroy(g.ctx);
    g = Adapter();
}
extern "C" void wcfe_adapter_seed(unsigned int seed) { ensure_context(); OK(wcfe_set_rng(g.ctx, WCFE_RNG_GLIBC_REPLAY, seed)); }
