/* oracle/ul_oracle.h -- TEST INFRASTRUCTURE ONLY (see ul_oracle.c). */
#ifndef UL_ORACLE_H
#define UL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ulo_mesh {
    int nn, ne;
    int *tri;               /* [3*ne] vertex ids, as read (iomesh.cpp:548) */
    int *ring_ptr;          /* [nn+1] CSR into ring */
    int *ring;              /* incident triangles in ascending triangle index
                               (nondestructivetrimesh.cpp:138-194 push_back order) */
    unsigned char *is_bnd;  /* m_is_boundary_vertex (nondestructivetrimesh.cpp:527-543) */
} ulo_mesh;

/* Everything the three reference variants hard-code, gathered in one struct.
   ulo_default_params() fills the reference's literal values. */
typedef struct ulo_params {
    int law;             /* 0 linear spring (k1,k31), 1 piecewise-linear MD table (kMD) */
    double kint_force;   /* kint hard-coded inside updateForces_Explicit: 1000 (k1 :458,:559) / 31000 (k31,kMD) */
    double kint_bond;    /* kint main.cpp hands to updateFA: 1000 / 31000 / -1 (main.cpp:350-354) */
    double subs_scale;   /* the literal `0*` of main.cpp:299 */
    double mu_cell, mu_subs;       /* UL:721-728 */
    double t_act_rate, t_act_max, t_act_t1;   /* UL:742-756: t<2 ? 200 t : 400 */
    double thick, kappa;           /* UL:767, :871 */
    double damp, rho, drag;        /* UL:596-597, :569 */
    double rho_i_max;              /* UL:1231 */
    double k_bend;                 /* UL:1632 */
    double k_area, area_0;         /* UL:1755-1756 */
    double ap_rate, ap_quantum, ap_density;   /* UL:1528-1539: 10, 5, 100 */
    double k_on, ka, kb, Fa, Fb;   /* UL:2079-2083 */
    double pin_limit, t_i, t_f;    /* UL:298-301: 49, 3, 8 */
} ulo_params;

typedef struct ulo_rng { uint32_t r[31]; int f, b; } ulo_rng;

typedef struct ulo_surface {
    ulo_mesh m;
    double *x, *X, *V, *Vh, *A, *ea, *cA, *cB, *cC, *fa_u;   /* AoS like the reference: [3n] / [6n] / [n] */
} ulo_surface;

typedef struct ulo_sim {
    ulo_surface cell, subs;
    ulo_params p;
    ulo_rng rng;
    int use_grid;          /* 0: literal O(Nn_cell*Ne_subs) loops; 1: uniform-grid candidates, same predicate and order */
    /* discrete state of the last step, for bit-exact checks */
    int *loc_bond;         /* [nn_cell] last containing substrate triangle in updateFA, -1 none */
    int *loc_force_cnt;    /* [nn_cell] number of containing triangles in calculateForce_FA_subs */
    int *loc_force_first;  /* [nn_cell] first containing triangle, -1 none */
    unsigned char *ap_fire;   /* [nn_cell] Pr < rate*dt*exp(-rate*dt) on boundary vertices */
    unsigned char *fa_reset;  /* [nn_cell] else-branch of UL:2107 taken */
    unsigned char *cc_clamp;  /* [nn_cell] c_C<0 clamp taken */
    double diag[8];        /* area, cx, cy, cz, |Res_cell|, |Res_subs| */
    /* force fields of the last step as they were inside updateForces_Explicit (for the 1e-10 force gate):
       0 Res(cell) 1 F_FA 2 F_AP 3 F_bending 4 F_area 5 Res(subs) 6 F_FA_subs; mm / res_ea per surface */
    double *sf[7], *s_mm[2], *s_rea[2];
    void *grid;
} ulo_sim;

void ulo_default_params(ulo_params *p, int variant /*0 k1, 1 k31, 2 kMD*/);

void ulo_srand(ulo_rng *g, unsigned int seed);
int  ulo_rand(ulo_rng *g);

int  ulo_mesh_build(ulo_mesh *m, int nn, int ne, const int *tri);
void ulo_mesh_free(ulo_mesh *m);

ulo_sim *ulo_create(int variant, int nn_c, const double *xyz_c, int ne_c, const int *tri_c,
                    int nn_s, const double *xyz_s, int ne_s, const int *tri_s);
void ulo_destroy(ulo_sim *s);
ulo_params *ulo_params_of(ulo_sim *s);
void ulo_set_use_grid(ulo_sim *s, int on);
void ulo_seed(ulo_sim *s, unsigned int seed);
void ulo_rng_discard(ulo_sim *s, long n);   /* skip n draws (resume a replay mid-run) */
/* field ids as oracle/ref_wrap.cpp: 0 x 1 V 2 A 3 ea 4 cA 5 cB 6 cC 7 fa_u 8 X_prev 9 V_hlf */
int  ulo_get(ulo_sim *s, int surface, int field, double *out);
int  ulo_set(ulo_sim *s, int surface, int field, const double *in);
/* discrete state: 0 loc_bond 1 loc_force_cnt 2 loc_force_first 3 ap_fire 4 fa_reset 5 cc_clamp */
int  ulo_get_discrete(ulo_sim *s, int which, int *out);
void ulo_get_diag(ulo_sim *s, double *out8);
void ulo_step(ulo_sim *s, double dt, double t_curr);

/* ---- the reference's functions, one by one -------------------------------------------------- */
void ulo_eval_element(const ulo_params *p, int mat, const double *X9, const double *x9, const double *ea18,
                      double dt, const double *centroid3, double time, double *Re9, double *Re_ea18, double *MMe3);
void ulo_calculate_res(const ulo_params *p, int mat, const ulo_mesh *m, const double *X, const double *x, const double *ea,
                       double dt, const double *centroid3, double time, double *Res, double *Res_ea, double *MM);
void ulo_area_tot(const ulo_mesh *m, const double *x, double *out4);
void ulo_force_FA(const ulo_params *p, const ulo_mesh *m, const double *x, const double *cC, const double *fa_u, double *F);
void ulo_force_FA_subs(ulo_sim *s, double *F);
void ulo_force_AP(ulo_sim *s, const double *x, double dt, double *F);
void ulo_force_bending(const ulo_mesh *m, const double *x, double k_bend, double *F);
void ulo_force_area(const ulo_params *p, const ulo_mesh *m, const double *x, double area_tot, double *F);
void ulo_update_FA(ulo_sim *s, double dt);
void ulo_driver_explicit(int nn, const double *V, const double *A, double dt, double *Vh);
void ulo_driver_biaxial(const ulo_params *p, int nn, const double *X, const double *V, const double *A, double dt, double t, double *Vh);
/* dormant physics: defined by the reference, calls commented out (UL:617-618, :577, :551) */
void ulo_diffusion_res(const ulo_params *p, const ulo_mesh *m, const double *X, const double *x, const double *cA, const double *cC,
                       double dt, double *Res_Adif, double *Res_Cdif);
void ulo_force_pressure(const ulo_mesh *m, const double *x, const double *MM, double t_tot, double *F);
void ulo_force_sphere(int nn, const double *x, double *F);
int  ulo_boundary_neighbours(ulo_sim *s, int ni, double *j1, double *j2);
/* every driverVelocities_* of UL:103-345 by kind: 0 Explicit, 1 Biaxial, 2 Uniaxial, 3 OffBiaxial, 4 BulgeTest; -1 on an unknown kind */
int  ulo_driver_kind(const ulo_params *p, int kind, int nn, const double *X, const double *V, const double *A, double dt, double t, double *Vh);
/* force field by id on the current state, same ids as ulref_force_field (7 Res_Adif, 8 Res_Cdif: out[nn]; 9 pressure, 10 sphere
 * constraint: the dormant ones) */
int  ulo_force_field(ulo_sim *s, int which, double dt, double t_curr, double *out, double *mm, double *res_ea, double *aux4);
/* the same fields as they were inside the last ulo_step() */
int  ulo_step_force_field(ulo_sim *s, int which, double *out, double *mm, double *res_ea);

#ifdef __cplusplus
}
#endif
#endif
