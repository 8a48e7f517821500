// wcfe_patch.cuh -- planar fast path: patch-tiled fused kernels that compute every element ONCE.
//
// The mesh is cut into patches of kPatch consecutive nodes (callers number nodes along a space-filling curve,
// meshgen.py does).  One CTA owns one patch:
//   load    nodal state of the owned nodes (coalesced) and of the halo nodes (gather, L2-resident: they are owned by
//           neighbouring patches) into shared memory;
//   element one thread per element incident to an owned node: closed-form planar element (wcfe_math.cuh) from shared
//           memory, result record (8-9 doubles) back into shared memory -- nothing is staged through HBM;
//   node    one thread per owned node: gathers its one-ring from the shared-memory records in ascending element order
//           (the reference's own summation order, UL_growth_explicit.cpp:692-705), adds the nodal forces, applies the
//           lumped-mass update and writes the new state (coalesced).
// Halo elements (incident to nodes of two or three patches) are recomputed by each of those patches: 15-30 % extra
// FLOPs instead of the 3x of the node-centric kernels, no atomics, bit-reproducible run to run.
//
// HBM traffic per step is the compulsory planar state (x,V,A in-plane + 3 strain components) plus the patch tables
// (~20 B/element): below the 3-D "algorithmic bytes" of SURVEY.md 8d because the z words of a planar state are dead.
#pragma once
#include "wcfe_kernels.cuh"

namespace wcfe {

constexpr int kPatch = 256;          // owned nodes per patch == threads per CTA

struct PatchView {
    int npatch;
    int max_halo, max_elem;          // shared-memory strides
    const int *elem_ptr;             // [npatch+1] into elem_gid / elem_lv
    const int *elem_gid;             // global element id of every patch-local element (ascending inside a patch)
    const unsigned short *elem_lv;   // [3*total] local vertex ids: < kPatch owned (node - first), >= kPatch halo
    const int *halo_ptr;             // [npatch+1] into halo_ids
    const int *halo_ids;             // global node ids of the halo nodes
    const unsigned short *ring16;    // parallel to SurfView::ring: (local element << 2) | slot
    const unsigned short *cand_l;    // parallel to BoundaryView::cand: local node ids (cell only)
};

__host__ __device__ inline size_t patch_smem_bytes(int n_node_arrays, int n_rec, int max_halo, int max_elem) {
    const size_t NL = (size_t)kPatch + max_halo, EM = (size_t)max_elem;
    return sizeof(double) * (n_node_arrays * NL + n_rec * EM);
}

// ---- substrate: updateForces_Explicit (UL:429-504) + driverVelocities_Biaxial (UL:275-345) + calculateForce_FA_subs ---
// (UL:1297-1454) for a frozen substrate (main.cpp:299): X == x, so only x is staged.
template <bool DBG>
__global__ void __launch_bounds__(kPatch) k_subs_patch(SurfView s, PatchView pv, SurfView c, BucketView bk, Consts k, double dt, double dt_real,
                                                       double t_tot, double bx0, double bx1, double by0, double by1, DebugView dbg) {
    extern __shared__ double sm[];
    const int NL = kPatch + pv.max_halo, EM = pv.max_elem;
    double *sx0 = sm, *sx1 = sx0 + NL, *se0 = sx1 + NL, *se1 = se0 + NL, *se3 = se1 + NL;
    double *rec = se3 + NL;                       // [8][EM]: r1x r1y r2x r2y es0 es1 es3 A0
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n0 = p * kPatch;
    const int nown = min(kPatch, s.nn - n0);
    const int h0 = pv.halo_ptr[p], nh = pv.halo_ptr[p + 1] - h0;
    const int e0 = pv.elem_ptr[p], nel = pv.elem_ptr[p + 1] - e0;
    const int np = s.np;
    // ---- load
    for (int l = tid; l < kPatch + nh; l += kPatch) {
        int g;
        if (l < kPatch) { if (l >= nown) continue; g = n0 + l; } else g = pv.halo_ids[h0 + l - kPatch];
        sx0[l] = s.x[g]; sx1[l] = s.x[np + g];
        se0[l] = s.ea_in[g]; se1[l] = s.ea_in[np + g]; se3[l] = s.ea_in[3 * np + g];
    }
    __syncthreads();
    // ---- elements
    for (int el = tid; el < nel; el += kPatch) {
        const unsigned short *lv = pv.elem_lv + (size_t)3 * (e0 + el);
        const int a = lv[0], b = lv[1], d = lv[2];
        const double X[3][2] = {{sx0[a], sx1[a]}, {sx0[b], sx1[b]}, {sx0[d], sx1[d]}};
        const double e3[3] = {1.0 / 3.0 * se0[a] + 1.0 / 3.0 * se0[b] + 1.0 / 3.0 * se0[d], 1.0 / 3.0 * se1[a] + 1.0 / 3.0 * se1[b] + 1.0 / 3.0 * se1[d],
                              1.0 / 3.0 * se3[a] + 1.0 / 3.0 * se3[b] + 1.0 / 3.0 * se3[d]};
        ElemOut2 o;
        elem_eval_planar(X, X, e3, k.mu_subs, 0.0, k.thick, o);
        rec[el] = o.r1[0]; rec[EM + el] = o.r1[1]; rec[2 * EM + el] = o.r2[0]; rec[3 * EM + el] = o.r2[1];
        rec[4 * EM + el] = o.eas[0]; rec[5 * EM + el] = o.eas[1]; rec[6 * EM + el] = o.eas[2]; rec[7 * EM + el] = o.A0;
    }
    __syncthreads();
    // ---- nodes
    if (tid >= nown) return;
    const int i = n0 + tid;
    const double ea0 = se0[tid], ea1 = se1[tid], ea3 = se3[tid];
    double rx = 0.0, ry = 0.0, q0 = 0.0, q1 = 0.0, q3 = 0.0, mm = 0.0;
    const int rb = s.ring_ptr[i], re = s.ring_ptr[i + 1];
    for (int r = rb; r < re; r++) {
        const int code = pv.ring16[r];
        const int el = code >> 2, slot = code & 3;
        const double r1x = rec[el], r1y = rec[EM + el], r2x = rec[2 * EM + el], r2y = rec[3 * EM + el];
        rx += slot == 0 ? -(r1x + r2x) : (slot == 1 ? r1x : r2x);
        ry += slot == 0 ? -(r1y + r2y) : (slot == 1 ? r1y : r2y);
        const double A0 = rec[7 * EM + el];
        q0 += (1. / 3.) * (rec[4 * EM + el] - ea0) * A0 * k.thick;       // UL:1170
        q1 += (1. / 3.) * (rec[5 * EM + el] - ea1) * A0 * k.thick;
        q3 += (1. / 3.) * (rec[6 * EM + el] - ea3) * A0 * k.thick;
        mm += (1. / 3.) * A0 * k.thick;                                  // UL:1172
    }
    const double Xi0 = sx0[tid], Xi1 = sx1[tid];
    double Fx = 0.0, Fy = 0.0;
    if (Xi0 >= bx0 && Xi0 <= bx1 && Xi1 >= by0 && Xi1 <= by1) {
        for (int r = rb; r < re; r++) {
            const int code = pv.ring16[r];
            const int el = code >> 2, slot = code & 3;
            const int e = pv.elem_gid[e0 + el];
            const unsigned long long h = bk.head[e];
            if ((unsigned int)(h >> 32) != bk.epoch) continue;
            int cnt = (int)(unsigned int)h;
            if (cnt > bk.cap) cnt = bk.cap;
            const unsigned short *lv = pv.elem_lv + (size_t)3 * (e0 + el);
            const double n1[3] = {sx0[lv[0]], sx1[lv[0]], 0.0}, n2[3] = {sx0[lv[1]], sx1[lv[1]], 0.0}, n3[3] = {sx0[lv[2]], sx1[lv[2]], 0.0};
            int prev = -1;
            for (int m = 0; m < cnt; m++) {          // ascending cell-node order: deterministic whatever the insertion order was
                int ni = 0x7fffffff;
                for (int q = 0; q < cnt; q++) { const int v = bk.items[(size_t)e * bk.cap + q]; if (v > prev && v < ni) ni = v; }
                prev = ni;
                const double u[3] = {c.fa[ni], c.fa[c.np + ni], 0.0};
                const double pt[3] = {c.x[ni] - u[0], c.x[c.np + ni] - u[1], 0.0};     // UL:1303
                double Ar[3], At;
                point_in_triangle(n1, n2, n3, pt, Ar, At);
                double fi[3];
                integrin_force(k, u, fi);
                const double w = Ar[slot] / At, cC = c.cC[ni], As = c.asum[ni];
                Fx += cC * fi[0] * w * As * k.rho_i_max;                                 // UL:1414-1422
                Fy += cC * fi[1] * w * As * k.rho_i_max;
            }
        }
    }
    double V0 = s.V[i], V1 = s.V[np + i], A0v = s.A[i], A1v = s.A[np + i];
    double Vh0 = V0 + (dt / 2.0) * A0v, Vh1 = V1 + (dt / 2.0) * A1v;
    if (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f) {            // UL:298-341, Vbiax == 0
        if (Xi0 < -k.pin_limit || Xi0 > k.pin_limit) Vh0 = 0.0;
        if (Xi1 > k.pin_limit || Xi1 < -k.pin_limit) Vh1 = 0.0;
    }
    node_update(Fx - rx, mm, k.damp, k.rho, dt, dt_real, Vh0, V0, A0v);                  // UL:471-503
    node_update(Fy - ry, mm, k.damp, k.rho, dt, dt_real, Vh1, V1, A1v);
    s.V[i] = V0; s.V[np + i] = V1; s.A[i] = A0v; s.A[np + i] = A1v;
    s.ea_out[i] = ea0 + q0 / mm; s.ea_out[np + i] = ea1 + q1 / mm; s.ea_out[3 * np + i] = ea3 + q3 / mm;
    if (DBG) {
        dbg.res[i] = rx; dbg.res[np + i] = ry; dbg.f_fa[i] = Fx; dbg.f_fa[np + i] = Fy; dbg.mm[i] = mm;
        dbg.res_ea[i] = q0; dbg.res_ea[np + i] = q1; dbg.res_ea[3 * np + i] = q3;
    }
}

}   // namespace wcfe
