// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A flat C interface over the UNMODIFIED reference translation units
//   <variant>/cellBP/UL_growth_explicit.cpp  and  <variant>/eltopo3d/nondestructivetrimesh.cpp
// compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libulref_{k1,k31,kMD}.so.  Nothing from the reference is copied here: this file
// only re-declares the reference's free-function prototypes (as the reference's own main.cpp does by
// hand, cellBP/main.cpp:69-87) and calls them.
//
// Sequencing of ulref_step() follows cellBP/main.cpp:168-418 (advance_sim) with El Topo reduced to
// what it does under `perform_improvement 0`, `collision_safety 0` (eltopo3d/dynamicsurface.cpp:
// integrate() == positions <- newpositions, actual_dt = dt), i.e. the El-Topo-free fast path of
// SURVEY.md Appendix A.
//
// Determinism: the Makefile links with -Wl,--wrap=srand, which turns the per-step `srand(time(NULL))`
// (UL_growth_explicit.cpp:1460) into a call of the no-op __wrap_srand below; the real libc srand() is
// only reachable through ulref_srand() -> __real_srand().

#include <vector>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <streambuf>
#include <Eigen/Dense>
#include "vec.h"
#include "nondestructivetrimesh.h"

#ifndef ULREF_KINT
#error "compile with -DULREF_KINT=<1000|31000|-1> (cellBP/main.cpp:350-354 of the variant)"
#endif

// ---- prototypes of the reference's free functions (cellBP/main.cpp:69-87; UL_growth_explicit.cpp:20-35)
void driverVelocities_Explicit(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt);
void driverVelocities_Biaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot);
void driverVelocities_Uniaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot);
void driverVelocities_OffBiaxial(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot);
void driverVelocities_BulgeTest(const std::vector<Vec3d> &X_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, double dt, double t_tot);
void updateFA(std::vector<double> &c_A, std::vector<double> &c_C, std::vector<Vec3d> &fa_u, double dt_real, const std::vector<Vec3d> &X_t_cell, const std::vector<Vec3d> &x_t_cell, const NonDestructiveTriMesh &mesh_cell, const std::vector<Vec3d> &X_t_subs, const std::vector<Vec3d> &x_t_subs, const NonDestructiveTriMesh &mesh_subs, double kint);
void updateForces_Explicit(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, std::vector<Eigen::VectorXd> &ea_t,
                           std::vector<double> &c_A, std::vector<double> &c_B, std::vector<double> &c_C, std::vector<Vec3d> &fa_u, double dt, double dt_real, const NonDestructiveTriMesh &mesh, double t_tot);
void updateForces_Explicit(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Vec3d> &V_t, std::vector<Vec3d> &V_t_hlf, std::vector<Vec3d> &A_t, std::vector<Eigen::VectorXd> &ea_t, double dt, double dt_real, const NonDestructiveTriMesh &mesh, double t_tot, std::vector<Vec3d> &fa_u, const std::vector<double> &c_C, const std::vector<Vec3d> &x_cell, const NonDestructiveTriMesh &mesh_cell);
void calculateRes(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Eigen::VectorXd> &ea_t, const std::vector<double> &c_A, const std::vector<double> &c_C, const NonDestructiveTriMesh &mesh, double dt_real, Eigen::VectorXd &Res, Eigen::VectorXd &Res_ea, Eigen::VectorXd &MM, Eigen::VectorXd &Res_Adif, Eigen::VectorXd &Res_Cdif, int mat, Eigen::Vector3d &centroid, double time);
void calculateRes(const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, std::vector<Eigen::VectorXd> &ea_t, const NonDestructiveTriMesh &mesh, double dt_real, Eigen::VectorXd &Res, Eigen::VectorXd &Res_ea, Eigen::VectorXd &MM, int mat, double time);
void evalElementRe(Eigen::Vector3d &node1_X, Eigen::Vector3d &node2_X, Eigen::Vector3d &node3_X, Eigen::Vector3d &node1_x, Eigen::Vector3d &node2_x, Eigen::Vector3d &node3_x, Eigen::VectorXd &node1_ea, Eigen::VectorXd &node2_ea, Eigen::VectorXd &node3_ea, std::vector<double> &node_c_A, std::vector<double> &node_c_C, double dt_real, Eigen::VectorXd &Re, Eigen::VectorXd &Re_ea, Eigen::VectorXd &Re_Adif, Eigen::VectorXd &Re_Cdif, Eigen::Vector3d &MMe, int mat, Eigen::Vector3d &centroid, double time);
void evalElementRe(Eigen::Vector3d &node1_X, Eigen::Vector3d &node2_X, Eigen::Vector3d &node3_X, Eigen::Vector3d &node1_x, Eigen::Vector3d &node2_x, Eigen::Vector3d &node3_x, Eigen::VectorXd &node1_ea, Eigen::VectorXd &node2_ea, Eigen::VectorXd &node3_ea, double dt_real, Eigen::VectorXd &Re, Eigen::VectorXd &Re_ea, Eigen::Vector3d &MMe, int mat, double time);
void calculateForce_FA(Eigen::VectorXd &F_constraint, const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, const std::vector<double> &c_C, std::vector<Vec3d> &fa_u, const NonDestructiveTriMesh &mesh, double kint);
void calculateForce_FA_subs(Eigen::VectorXd &F_constraint, const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, const NonDestructiveTriMesh &mesh, const std::vector<double> &c_C, const std::vector<Vec3d> &x_cell, std::vector<Vec3d> &fa_u, const NonDestructiveTriMesh &mesh_cell, double kint);
void calculateForce_AP(Eigen::VectorXd &F_constraint, const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, const std::vector<double> &c_A, const NonDestructiveTriMesh &mesh, double dt_real);
void calculateBendingForce2D(Eigen::VectorXd &F_bending, const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, const NonDestructiveTriMesh &mesh);
void calculateConstraintForce(Eigen::VectorXd &F_constraint, const std::vector<Vec3d> &x_t);
void calculatePressureForce(Eigen::VectorXd &F_constraint, const std::vector<Vec3d> &x_t, const Eigen::VectorXd &MM, const NonDestructiveTriMesh &mesh, double t_tot);
Eigen::Vector4d calculateAreaTot(const std::vector<Vec3d> &x_t, const NonDestructiveTriMesh &mesh);
void calculateAreaConstraintForce(Eigen::VectorXd &F_areaConstraint, const std::vector<Vec3d> &X_t, const std::vector<Vec3d> &x_t, const NonDestructiveTriMesh &mesh, double tot_area);

namespace {

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

// the reference prints ~30 lines per step to std::cout; silence it while inside the wrapper
struct Quiet {
    std::streambuf *old; NullBuf nb;
    Quiet() : old(std::cout.rdbuf(&nb)) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

struct Surface {
    NonDestructiveTriMesh mesh;
    std::vector<Vec3d> x, V, A, fa_u;
    std::vector<Eigen::VectorXd> ea;
    std::vector<double> cA, cB, cC;
    void init(int nn, const double *xyz, int ne, const int *tri) {
        x.resize(nn); V.assign(nn, Vec3d(0, 0, 0)); A.assign(nn, Vec3d(0, 0, 0)); fa_u.assign(nn, Vec3d(0, 0, 0));
        for (int i = 0; i < nn; i++) x[i] = Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        ea.resize(nn);
        for (int i = 0; i < nn; i++) { ea[i] = Eigen::VectorXd(6); ea[i].setZero(); }   // scriptinit.cpp:567-580
        cA.assign(nn, 0.0); cB.assign(nn, 0.0); cC.assign(nn, 0.0);
        mesh.set_num_vertices(nn);
        for (int e = 0; e < ne; e++) mesh.add_triangle(Vec3st(tri[3 * e], tri[3 * e + 1], tri[3 * e + 2]));   // iomesh.cpp:548
    }
};

void put3(const std::vector<Vec3d> &v, double *out) { for (size_t i = 0; i < v.size(); i++) for (int c = 0; c < 3; c++) out[3 * i + c] = v[i][c]; }
void get3(std::vector<Vec3d> &v, const double *in) { for (size_t i = 0; i < v.size(); i++) for (int c = 0; c < 3; c++) v[i][c] = in[3 * i + c]; }

}  // namespace

struct ulref {
    Surface cell, subs;
    // last-step temporaries kept for inspection
    std::vector<Vec3d> Xc_prev, Xs_prev, Vh_c, Vh_s;
};

extern "C" {

#ifdef ULREF_ADAPTER_ONLY
// Adapter build: the very same driver, linked against wholecellfe_b200/csrc/ul_adapter.cpp + libwcfe.so instead of the
// reference's UL_growth_explicit.o -- i.e. the five symbols below resolve to the GPU drop-in.
void wcfe_adapter_seed(unsigned int);
void wcfe_adapter_reset(void);
void ulref_srand(unsigned int seed) { wcfe_adapter_seed(seed); }
int ulref_rand(void) { return 0; }
#else
// target of -Wl,--wrap=srand for every srand() call in this .so: a no-op
void __wrap_srand(unsigned int) {}
void __real_srand(unsigned int);

void ulref_srand(unsigned int seed) { __real_srand(seed); }   // main.cpp:709 does srand(1)
int ulref_rand(void) { return rand(); }
#endif
double ulref_default_kint(void) { return (double)(ULREF_KINT); }

ulref *ulref_create(int nn_c, const double *xyz_c, int ne_c, const int *tri_c,
                    int nn_s, const double *xyz_s, int ne_s, const int *tri_s) {
#ifdef ULREF_ADAPTER_ONLY
    wcfe_adapter_reset();
#endif
    ulref *h = new ulref;
    h->cell.init(nn_c, xyz_c, ne_c, tri_c);
    h->subs.init(nn_s, xyz_s, ne_s, tri_s);
    for (int i = 0; i < nn_c; i++) { h->cell.cA[i] = 100.0; h->cell.cC[i] = 0.0; }   // main.cpp:671-674
    return h;
}
void ulref_destroy(ulref *h) { delete h; }

// field ids: 0 x, 1 V, 2 A, 3 ea(6), 4 cA, 5 cB, 6 cC, 7 fa_u, 8 X_prev (positions at the start of the
// last step), 9 V_hlf of the last step
int ulref_get(ulref *h, int surface, int field, double *out) {
    Surface &s = surface == 0 ? h->cell : h->subs;
    switch (field) {
    case 0: put3(s.x, out); break;
    case 1: put3(s.V, out); break;
    case 2: put3(s.A, out); break;
    case 3: for (size_t i = 0; i < s.ea.size(); i++) for (int q = 0; q < 6; q++) out[6 * i + q] = s.ea[i](q); break;
    case 4: memcpy(out, s.cA.data(), s.cA.size() * 8); break;
    case 5: memcpy(out, s.cB.data(), s.cB.size() * 8); break;
    case 6: memcpy(out, s.cC.data(), s.cC.size() * 8); break;
    case 7: put3(s.fa_u, out); break;
    case 8: put3(surface == 0 ? h->Xc_prev : h->Xs_prev, out); break;
    case 9: put3(surface == 0 ? h->Vh_c : h->Vh_s, out); break;
    default: return -1;
    }
    return 0;
}
int ulref_set(ulref *h, int surface, int field, const double *in) {
    Surface &s = surface == 0 ? h->cell : h->subs;
    switch (field) {
    case 0: get3(s.x, in); break;
    case 1: get3(s.V, in); break;
    case 2: get3(s.A, in); break;
    case 3: for (size_t i = 0; i < s.ea.size(); i++) for (int q = 0; q < 6; q++) s.ea[i](q) = in[6 * i + q]; break;
    case 4: memcpy(s.cA.data(), in, s.cA.size() * 8); break;
    case 5: memcpy(s.cB.data(), in, s.cB.size() * 8); break;
    case 6: memcpy(s.cC.data(), in, s.cC.size() * 8); break;
    case 7: get3(s.fa_u, in); break;
    default: return -1;
    }
    return 0;
}

// vertex->triangle incidence and boundary flags as the reference's mesh class builds them
int ulref_onering(ulref *h, int surface, int v, int *out, int cap) {
    Surface &s = surface == 0 ? h->cell : h->subs;
    const std::vector<size_t> &r = s.mesh.m_vertex_to_triangle_map[v];
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = (int)r[i];
    return (int)r.size();
}
void ulref_boundary_flags(ulref *h, int surface, int *out) {
    Surface &s = surface == 0 ? h->cell : h->subs;
    for (size_t i = 0; i < s.x.size(); i++) out[i] = s.mesh.m_is_boundary_vertex[i] ? 1 : 0;
}

// One UL step, El Topo off.  main.cpp:168-418; dt = requested = actual.  subs_scale is the literal
// `0*` factor of main.cpp:299 (pass 0.0 for the shipped drivers).
void ulref_step(ulref *h, double dt, double t_curr, double kint, double subs_scale) {
    Quiet q;
    Surface &c = h->cell, &s = h->subs;
    const int nc = (int)c.x.size(), ns = (int)s.x.size();
    h->Vh_c.assign(nc, Vec3d(0, 0, 0)); h->Vh_s.assign(ns, Vec3d(0, 0, 0));
    driverVelocities_Explicit(c.x, c.V, h->Vh_c, c.A, dt);                       // main.cpp:224
    std::vector<Vec3d> newp(nc);
    for (int i = 0; i < nc; i++) newp[i] = c.x[i] + dt * h->Vh_c[i];           // main.cpp:227-231
    h->Xc_prev = c.x;                                                            // main.cpp:242
    c.x = newp;                                                                  // integrate(), collision_safety 0
    driverVelocities_Biaxial(s.x, s.V, h->Vh_s, s.A, dt, t_curr);                // main.cpp:293
    std::vector<Vec3d> newps(ns);
    for (int i = 0; i < ns; i++) newps[i] = s.x[i] + subs_scale * dt * h->Vh_s[i];   // main.cpp:298-302
    h->Xs_prev = s.x;                                                            // main.cpp:311
    s.x = newps;
    updateFA(c.cA, c.cC, c.fa_u, dt, h->Xc_prev, c.x, c.mesh, h->Xs_prev, s.x, s.mesh, kint);   // main.cpp:354
    updateForces_Explicit(h->Xc_prev, c.x, c.V, h->Vh_c, c.A, c.ea, c.cA, c.cB, c.cC, c.fa_u, dt, dt, c.mesh, t_curr);   // :360
    updateForces_Explicit(h->Xs_prev, s.x, s.V, h->Vh_s, s.A, s.ea, dt, dt, s.mesh, t_curr, c.fa_u, c.cC, c.x, c.mesh);   // :361
}

// The step of ulref_step (subs_scale 0) with the SUBSTRATE's driver swapped for one of the alternatives main.cpp keeps commented out
// at that call site (main.cpp:291-295): kind 1 Biaxial (== ulref_step), 2 Uniaxial, 3 OffBiaxial, 4 BulgeTest.
static void substrate_driver(int kind, Surface &s, std::vector<Vec3d> &Vh, double dt, double t_curr) {
    switch (kind) {
    case 2: driverVelocities_Uniaxial(s.x, s.V, Vh, s.A, dt, t_curr); break;
    case 3: driverVelocities_OffBiaxial(s.x, s.V, Vh, s.A, dt, t_curr); break;
    case 4: driverVelocities_BulgeTest(s.x, s.V, Vh, s.A, dt, t_curr); break;
    default: driverVelocities_Biaxial(s.x, s.V, Vh, s.A, dt, t_curr); break;
    }
}
void ulref_step_driver(ulref *h, double dt, double t_curr, double kint, int subs_kind) {
    Quiet q;
    Surface &c = h->cell, &s = h->subs;
    const int nc = (int)c.x.size(), ns = (int)s.x.size();
    h->Vh_c.assign(nc, Vec3d(0, 0, 0)); h->Vh_s.assign(ns, Vec3d(0, 0, 0));
    driverVelocities_Explicit(c.x, c.V, h->Vh_c, c.A, dt);                       // main.cpp:224
    std::vector<Vec3d> newp(nc);
    for (int i = 0; i < nc; i++) newp[i] = c.x[i] + dt * h->Vh_c[i];           // main.cpp:227-231
    h->Xc_prev = c.x;
    c.x = newp;
    substrate_driver(subs_kind, s, h->Vh_s, dt, t_curr);                         // main.cpp:291-295
    h->Xs_prev = s.x;                                                            // main.cpp:299: x + 0*dt*V_hlf
    updateFA(c.cA, c.cC, c.fa_u, dt, h->Xc_prev, c.x, c.mesh, h->Xs_prev, s.x, s.mesh, kint);   // main.cpp:354
    updateForces_Explicit(h->Xc_prev, c.x, c.V, h->Vh_c, c.A, c.ea, c.cA, c.cB, c.cC, c.fa_u, dt, dt, c.mesh, t_curr);   // :360
    updateForces_Explicit(h->Xs_prev, s.x, s.V, h->Vh_s, s.A, s.ea, dt, dt, s.mesh, t_curr, c.fa_u, c.cC, c.x, c.mesh);   // :361
}

// The same step when the host integrator accepts LESS than the requested dt (El Topo's integrate() with collision_safety on,
// dynamicsurface.cpp:462-637, returns actual_dt <= dt and leaves the vertices part of the way): main.cpp:244-335 literally, with
// integrate() modelled as "move the fraction frac of the way, actual_dt = frac * dt".  frac_subs < 1 also takes the redo branch
// of main.cpp:322-335 (the cell is advanced a second time with the substrate's smaller dt -- as the reference does).
void ulref_step_short(ulref *h, double dt, double t_curr, double kint, double frac_cell, double frac_subs) {
    Quiet q;
    Surface &c = h->cell, &s = h->subs;
    const int nc = (int)c.x.size(), ns = (int)s.x.size();
    auto integrate = [](std::vector<Vec3d> &pos, const std::vector<Vec3d> &newpos, double frac, double dt_in, double &actual) {
        for (size_t i = 0; i < pos.size(); i++) pos[i] = frac == 1.0 ? newpos[i] : pos[i] + frac * (newpos[i] - pos[i]);
        actual = frac * dt_in;
    };
    double curr_dt = dt, actual_dt;
    h->Vh_c.assign(nc, Vec3d(0, 0, 0)); h->Vh_s.assign(ns, Vec3d(0, 0, 0));
    driverVelocities_Explicit(c.x, c.V, h->Vh_c, c.A, curr_dt);                  // main.cpp:224
    std::vector<Vec3d> newp(nc);
    for (int i = 0; i < nc; i++) newp[i] = c.x[i] + curr_dt * h->Vh_c[i];      // main.cpp:227-231
    h->Xc_prev = c.x;                                                            // initial_positions, main.cpp:242
    integrate(c.x, newp, frac_cell, curr_dt, actual_dt);                         // main.cpp:248
    const double ori_dt = curr_dt;                                               // main.cpp:251
    curr_dt = actual_dt;
    driverVelocities_Biaxial(s.x, s.V, h->Vh_s, s.A, curr_dt, t_curr);           // main.cpp:293
    std::vector<Vec3d> newps(ns);
    for (int i = 0; i < ns; i++) newps[i] = s.x[i] + 0 * curr_dt * h->Vh_s[i];  // main.cpp:298-302
    h->Xs_prev = s.x;                                                            // main.cpp:311
    integrate(s.x, newps, frac_subs, curr_dt, actual_dt);                        // main.cpp:316
    const double ori_dt2 = curr_dt;
    curr_dt = actual_dt;
    if (actual_dt < ori_dt2) {                                                   // main.cpp:322-335
        driverVelocities_Explicit(c.x, c.V, h->Vh_c, c.A, curr_dt);
        for (int i = 0; i < nc; i++) newp[i] = c.x[i] + curr_dt * h->Vh_c[i];
        integrate(c.x, newp, 1.0, curr_dt, actual_dt);
        curr_dt = actual_dt;
    }
    updateFA(c.cA, c.cC, c.fa_u, actual_dt, h->Xc_prev, c.x, c.mesh, h->Xs_prev, s.x, s.mesh, kint);   // main.cpp:354
    updateForces_Explicit(h->Xc_prev, c.x, c.V, h->Vh_c, c.A, c.ea, c.cA, c.cB, c.cC, c.fa_u, ori_dt, actual_dt, c.mesh, t_curr);   // :360
    updateForces_Explicit(h->Xs_prev, s.x, s.V, h->Vh_s, s.A, s.ea, ori_dt, actual_dt, s.mesh, t_curr, c.fa_u, c.cC, c.x, c.mesh);   // :361
}

// every driverVelocities_* the reference defines (UL:103-345; main.cpp:69-76 declares all five, :217-224 / :291-295 call two),
// as pure functions of flat AoS arrays.  kind: 0 Explicit, 1 Biaxial, 2 Uniaxial, 3 OffBiaxial, 4 BulgeTest.  In the adapter build the
// five symbols resolve to ul_adapter.cpp (the device).
int ulref_driver(int kind, int nn, const double *X, const double *V, const double *A, double dt, double t_tot, double *Vh) {
    Quiet q;
    std::vector<Vec3d> x(nn), v(nn), a(nn), vh(nn, Vec3d(0, 0, 0));
    get3(x, X); get3(v, V); get3(a, A);
    switch (kind) {
    case 0: driverVelocities_Explicit(x, v, vh, a, dt); break;
    case 1: driverVelocities_Biaxial(x, v, vh, a, dt, t_tot); break;
    case 2: driverVelocities_Uniaxial(x, v, vh, a, dt, t_tot); break;
    case 3: driverVelocities_OffBiaxial(x, v, vh, a, dt, t_tot); break;
    case 4: driverVelocities_BulgeTest(x, v, vh, a, dt, t_tot); break;
    default: return -1;
    }
    put3(vh, Vh);
    return 0;
}

#ifndef ULREF_ADAPTER_ONLY
// ---- fine-grained entry points for known-answer tests -------------------------------------------

void ulref_eval_element(int mat, const double *X9, const double *x9, const double *ea18, const double *cA3, const double *cC3,
                        double dt, const double *centroid3, double time, double *Re9, double *Re_ea18, double *MMe3) {
    Quiet q;
    Eigen::Vector3d X[3], x[3]; Eigen::VectorXd ea[3];
    for (int i = 0; i < 3; i++) {
        X[i] << X9[3 * i], X9[3 * i + 1], X9[3 * i + 2];
        x[i] << x9[3 * i], x9[3 * i + 1], x9[3 * i + 2];
        ea[i] = Eigen::VectorXd(6);
        for (int k = 0; k < 6; k++) ea[i](k) = ea18[6 * i + k];
    }
    Eigen::VectorXd Re(9), Re_ea(18), RA(3), RC(3); Eigen::Vector3d MMe;
    Re.setZero(); Re_ea.setZero(); MMe.setZero(); RA.setZero(); RC.setZero();
    if (mat == 1) {
        std::vector<double> cA(cA3, cA3 + 3), cC(cC3, cC3 + 3);
        Eigen::Vector3d cen; cen << centroid3[0], centroid3[1], centroid3[2];
        evalElementRe(X[0], X[1], X[2], x[0], x[1], x[2], ea[0], ea[1], ea[2], cA, cC, dt, Re, Re_ea, RA, RC, MMe, mat, cen, time);
    } else {
        evalElementRe(X[0], X[1], X[2], x[0], x[1], x[2], ea[0], ea[1], ea[2], dt, Re, Re_ea, MMe, mat, time);
    }
    for (int i = 0; i < 9; i++) Re9[i] = Re(i);
    for (int i = 0; i < 18; i++) Re_ea18[i] = Re_ea(i);
    for (int i = 0; i < 3; i++) MMe3[i] = MMe(i);
}

// Individual nodal force fields on the CURRENT state of the handle (X = Xc_prev of the last step, or x
// itself if no step was taken).  which: 0 Res(cell) 1 F_FA(cell) 2 F_AP 3 F_bending 4 F_area
// 5 Res(subs) 6 F_FA_subs.  out has 3*Nn doubles; for which==0/5 also returns MM in mm (Nn) and
// Res_ea in res_ea (6*Nn) when non-null.  aux4 receives (area, cx, cy, cz) for the cell.
int ulref_force_field(ulref *h, int which, double dt, double t_curr, double kint, double *out, double *mm, double *res_ea, double *aux4) {
    Quiet q;
    Surface &c = h->cell, &s = h->subs;
    const std::vector<Vec3d> &Xc = h->Xc_prev.size() == c.x.size() ? h->Xc_prev : c.x;
    const std::vector<Vec3d> &Xs = h->Xs_prev.size() == s.x.size() ? h->Xs_prev : s.x;
    Eigen::Vector4d area = calculateAreaTot(c.x, c.mesh);
    if (aux4) for (int i = 0; i < 4; i++) aux4[i] = area(i);
    const int nc = (int)c.x.size(), ns = (int)s.x.size();
    Eigen::VectorXd F;
    if (which == 0) {
        Eigen::VectorXd Res(nc * 3), Rea(nc * 6), MM(nc), RA(nc), RC(nc);
        Eigen::Vector3d cen; cen << area(1), area(2), area(3);
        calculateRes(Xc, c.x, c.ea, c.cA, c.cC, c.mesh, dt, Res, Rea, MM, RA, RC, 1, cen, t_curr);
        for (int i = 0; i < nc * 3; i++) out[i] = Res(i);
        if (mm) for (int i = 0; i < nc; i++) mm[i] = MM(i);
        if (res_ea) for (int i = 0; i < nc * 6; i++) res_ea[i] = Rea(i);
        return 0;
    }
    if (which >= 7 && which <= 9) {          // dormant: the diffusion residuals calculateRes(cell) returns and nobody reads (UL:617-618); pressure (UL:577)
        Eigen::VectorXd Res(nc * 3), Rea(nc * 6), MM(nc), RA(nc), RC(nc);
        Eigen::Vector3d cen; cen << area(1), area(2), area(3);
        calculateRes(Xc, c.x, c.ea, c.cA, c.cC, c.mesh, dt, Res, Rea, MM, RA, RC, 1, cen, t_curr);
        if (which == 9) {
            Eigen::VectorXd Fp(nc * 3); Fp.setZero();
            calculatePressureForce(Fp, c.x, MM, c.mesh, t_curr);
            for (int i = 0; i < nc * 3; i++) out[i] = Fp(i);
        } else for (int i = 0; i < nc; i++) out[i] = which == 7 ? RA(i) : RC(i);
        return 0;
    }
    if (which == 10) {                       // dormant: sphere constraint (UL:551)
        Eigen::VectorXd Fc(nc * 3); Fc.setZero();
        calculateConstraintForce(Fc, c.x);
        for (int i = 0; i < nc * 3; i++) out[i] = Fc(i);
        return 0;
    }
    if (which == 5) {
        Eigen::VectorXd Res(ns * 3), Rea(ns * 6), MM(ns);
        calculateRes(Xs, s.x, s.ea, s.mesh, dt, Res, Rea, MM, 2, t_curr);
        for (int i = 0; i < ns * 3; i++) out[i] = Res(i);
        if (mm) for (int i = 0; i < ns; i++) mm[i] = MM(i);
        if (res_ea) for (int i = 0; i < ns * 6; i++) res_ea[i] = Rea(i);
        return 0;
    }
    if (which == 6) {
        F = Eigen::VectorXd(ns * 3); F.setZero();
        calculateForce_FA_subs(F, Xs, s.x, s.mesh, c.cC, c.x, c.fa_u, c.mesh, kint);
        for (int i = 0; i < ns * 3; i++) out[i] = F(i);
        return 0;
    }
    F = Eigen::VectorXd(nc * 3); F.setZero();
    switch (which) {
    case 1: calculateForce_FA(F, Xc, c.x, c.cC, c.fa_u, c.mesh, kint); break;
    case 2: calculateForce_AP(F, Xc, c.x, c.cA, c.mesh, dt); break;
    case 3: calculateBendingForce2D(F, Xc, c.x, c.mesh); break;
    case 4: calculateAreaConstraintForce(F, Xc, c.x, c.mesh, area(0)); break;
    default: return -1;
    }
    for (int i = 0; i < nc * 3; i++) out[i] = F(i);
    return 0;
}

// ---- the reference's own .vtk writers and mesh reader (cellBP/iomesh.cpp), for byte-level golden files ------------
}  // extern "C"
#include <iomesh.h>
extern "C" {
int ulref_write_vtk_cell(ulref *h, double time, const char *path) {
    Surface &c = h->cell;
#if ULREF_KINT < 0
    write_vtk(c.mesh, c.x, c.cA, c.cB, c.cC, c.fa_u, c.ea, time, path, (double)(ULREF_KINT));   // kMD main.cpp:448-450
#else
    write_vtk(c.mesh, c.x, c.cA, c.cB, c.cC, c.fa_u, c.ea, time, path);                         // main.cpp:446-447
#endif
    return 0;
}
// fa_source: the reference hands the substrate writer the CELL's fa_u at frame 0 (main.cpp:722) -- pass 0 for that; 1 writes
// a zero vector of the substrate's size (what g_surf_subs->get_fa_u() would hold if it were sized correctly, quirk q6)
int ulref_write_vtk_subs(ulref *h, double time, const char *path, int fa_source) {
    Surface &s = h->subs;
    std::vector<Vec3d> fa = fa_source == 0 ? h->cell.fa_u : std::vector<Vec3d>(s.x.size(), Vec3d(0, 0, 0));
    if (fa.size() < s.x.size()) fa.resize(s.x.size(), Vec3d(0, 0, 0));
    write_vtk(s.mesh, s.x, s.ea, fa, time, path);                                               // main.cpp:451
    return 0;
}
#endif   // ULREF_ADAPTER_ONLY

}  // extern "C"
