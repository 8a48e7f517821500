// wcfe_patch_alt.cuh -- k_subs_patch: the element-once substrate kernel (one thread per patch-local element, records in shared memory,
// one-ring gather in ascending element order, buckets walked per node).  NOT on the production path: k_subs_ring (wcfe_patch.cuh)
// is 1.9x faster on B200.  Kept selectable (WCFE_SUBS_KERNEL=0) as an independent implementation of the same physics for the parity
// tests (tests/test_gpu_parity.py::test_alternative_substrate_kernels_agree, ::test_strained_substrate_against_oracle).
#pragma once
#include "wcfe_patch.cuh"

namespace wcfe {

// ---- substrate: updateForces_Explicit (UL:429-504) + driverVelocities_Biaxial (UL:275-345) + calculateForce_FA_subs ---
// (UL:1297-1454) for a frozen substrate (main.cpp:299): X == x, F == I, closed form elem_eval_frozen.
// Shared memory: node table packed as (x,y) / (e_xx,e_yy) double2 + e_xy, element records as five double2 per element
// (force on local node 0, 1, 2; (e_s.xx,e_s.yy); (e_s.xy, A0)) so that gathers are 128-bit and branch-free.
template <bool DBG>
__global__ void __launch_bounds__(kThreads, kMinBlk) k_subs_patch(SurfView s, PatchView pv, SurfView c, BucketView bk, Consts k, double dt, double dt_real,
                                                                  double t_tot, double bx0, double bx1, double by0, double by1, DebugView dbg, double dt_drv) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int NL = kPatch + pv.max_halo, EM = pv.max_elem;
    unsigned long long *bar = (unsigned long long *)smraw;
    unsigned char *sblob = smraw + 16;
    double2 *sxy = (double2 *)(sblob + pv.stride);       // [NL]
    double2 *se01 = sxy + NL;                             // [NL]
    double *se3 = (double *)(se01 + NL);                  // [NL] (NL is a multiple of 4)
    double2 *rec = (double2 *)(se3 + NL);                 // [5][EM]; first used as the SoA landing zone of the bulk copies
    double *stage = (double *)rec;                        // [5][kPatch]: x0 x1 e0 e1 e3 of the owned nodes
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n0 = p * kPatch;
    const int nown = min(kPatch, s.nn - n0);
    const int ncopy = min(kPatch, s.np - n0);     // padded count: bulk copies move multiples of 16 B
    const int np = s.np;
    if (tid == 0) {
        mbar_init(bar, 1);
        const unsigned int nb = (unsigned int)ncopy * 8u;
        mbar_expect_tx(bar, (unsigned int)pv.stride + 5u * nb);
        bulk_g2s(sblob, pv.blob + (size_t)p * pv.stride, (unsigned int)pv.stride, bar);
        bulk_g2s(stage, s.x + n0, nb, bar); bulk_g2s(stage + kPatch, s.x + np + n0, nb, bar);
        bulk_g2s(stage + 2 * kPatch, s.ea_in + n0, nb, bar); bulk_g2s(stage + 3 * kPatch, s.ea_in + np + n0, nb, bar);
        bulk_g2s(stage + 4 * kPatch, s.ea_in + 3 * (size_t)np + n0, nb, bar);
    }
    // own V, A straight to registers while the copies fly
    const int i = n0 + tid;
    const bool own = tid < nown;
    double V0 = 0, V1 = 0, A0v = 0, A1v = 0;
    if (own) { V0 = s.V[i]; V1 = s.V[np + i]; A0v = s.A[i]; A1v = s.A[np + i]; }
    __syncthreads();                              // barrier init visible to all waiters
    mbar_wait(bar, 0);
    const int *hdr = (const int *)sblob;
    const int nh = hdr[0], nel = hdr[1];
    const int *halo_ids = (const int *)(sblob + pv.off_halo);
    const unsigned short *elem_lv = (const unsigned short *)(sblob + pv.off_lv);
    const unsigned short *rptr = (const unsigned short *)(sblob + pv.off_rptr);
    const unsigned short *ring16 = (const unsigned short *)(sblob + pv.off_ring);
    // repack the owned nodes, gather the halo nodes (L2-resident: owned by neighbouring patches)
    double2 myxy = make_double2(0, 0), mye01 = make_double2(0, 0); double mye3 = 0;
    if (tid < kPatch) {
        myxy = make_double2(stage[tid], stage[kPatch + tid]); mye01 = make_double2(stage[2 * kPatch + tid], stage[3 * kPatch + tid]); mye3 = stage[4 * kPatch + tid];
        sxy[tid] = myxy; se01[tid] = mye01; se3[tid] = mye3;
    }
    for (int l = tid; l < nh; l += kThreads) {
        const int g = halo_ids[l];
        sxy[kPatch + l] = make_double2(s.x[g], s.x[np + g]);
        se01[kPatch + l] = make_double2(s.ea_in[g], s.ea_in[np + g]);
        se3[kPatch + l] = s.ea_in[3 * (size_t)np + g];
    }
    __syncthreads();                              // node table complete; the landing zone may now be overwritten by records
    // ---- elements
    for (int el = tid; el < nel; el += kThreads) {
        const unsigned short *lv = elem_lv + 3 * el;
        const int a = lv[0], b = lv[1], d = lv[2];
        const double2 pa = sxy[a], pb = sxy[b], pd = sxy[d], ea = se01[a], eb = se01[b], ed = se01[d];
        const double X[3][2] = {{pa.x, pa.y}, {pb.x, pb.y}, {pd.x, pd.y}};
        const double e3[3] = {1.0 / 3.0 * ea.x + 1.0 / 3.0 * eb.x + 1.0 / 3.0 * ed.x, 1.0 / 3.0 * ea.y + 1.0 / 3.0 * eb.y + 1.0 / 3.0 * ed.y,
                              1.0 / 3.0 * se3[a] + 1.0 / 3.0 * se3[b] + 1.0 / 3.0 * se3[d]};          // UL:1055
        ElemOutFrozen o;
        elem_eval_frozen(X, e3, k.mu_subs, k.thick, o);
        rec[el] = make_double2(-(o.r1[0] + o.r2[0]), -(o.r1[1] + o.r2[1]));   // force on local node 0, 1, 2: the node phase
        rec[EM + el] = make_double2(o.r1[0], o.r1[1]);                       // reads only the one of its slot
        rec[2 * EM + el] = make_double2(o.r2[0], o.r2[1]);
        rec[3 * EM + el] = make_double2(e3[0], e3[1]);           // e_s == averaged e for F == I
        rec[4 * EM + el] = make_double2(e3[2], o.A0);
    }
    __syncthreads();
    // ---- nodes
    if (!own) return;
    const double ea0 = mye01.x, ea1 = mye01.y, ea3 = mye3;
    double rx = 0.0, ry = 0.0, q0 = 0.0, q1 = 0.0, q3 = 0.0, mm = 0.0;
    const int rb = rptr[tid], re = rptr[tid + 1];
    // The one-ring gather is a chain of dependent shared-memory loads (code -> record); process the ring in chunks of
    // kRingChunk entries with every load of a chunk issued before the first use, padding short chunks with zero-weight
    // entries (adding +0 leaves the ascending-element summation order and the result unchanged).
    for (int base = rb; base < re; base += kRingChunk) {
        int code[kRingChunk];
#pragma unroll
        for (int j = 0; j < kRingChunk; j++) code[j] = base + j < re ? (int)ring16[base + j] : -1;
        double2 f[kRingChunk], es[kRingChunk], ez[kRingChunk];
#pragma unroll
        for (int j = 0; j < kRingChunk; j++) {
            const int el = code[j] < 0 ? 0 : code[j] >> 2, slot = code[j] & 3;
            f[j] = rec[slot * EM + el]; es[j] = rec[3 * EM + el]; ez[j] = rec[4 * EM + el];   // 48 B per (node, element)
        }
#pragma unroll
        for (int j = 0; j < kRingChunk; j++) {
            const bool on = code[j] >= 0;
            const double w = on ? (1. / 3.) * ez[j].y * k.thick : 0.0;   // UL:1172 == the weight of UL:1170
            rx += on ? f[j].x : 0.0; ry += on ? f[j].y : 0.0;
            q0 += (es[j].x - ea0) * w; q1 += (es[j].y - ea1) * w; q3 += (ez[j].x - ea3) * w;
            mm += w;
        }
    }
    const double Xi0 = myxy.x, Xi1 = myxy.y;
    double Fx = 0.0, Fy = 0.0;
    if (Xi0 >= bx0 && Xi0 <= bx1 && Xi1 >= by0 && Xi1 <= by1) {
        const int e0 = pv.elem_ptr[p];
        for (int r = rb; r < re; r++) {
            const int code = ring16[r];
            const int el = code >> 2, slot = code & 3;
            const int e = pv.elem_gid[e0 + el];
            const unsigned long long h = bk.head[e];
            if ((unsigned int)(h >> 32) != bk.epoch) continue;
            int cnt = (int)(unsigned int)h;
            if (cnt > bk.cap) cnt = bk.cap;
            const unsigned short *lv = elem_lv + 3 * el;
            const double2 p1 = sxy[lv[0]], p2 = sxy[lv[1]], p3 = sxy[lv[2]];
            const double n1[3] = {p1.x, p1.y, 0.0}, n2[3] = {p2.x, p2.y, 0.0}, n3[3] = {p3.x, p3.y, 0.0};
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
    double Vh0 = WCFE_ADD(V0, WCFE_MUL(dt_drv / 2.0, A0v)), Vh1 = WCFE_ADD(V1, WCFE_MUL(dt_drv / 2.0, A1v));   // UL:295, with the dt the driver was called with
    if (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f) {            // UL:298-341, Vbiax == 0
        if (Xi0 < -k.pin_limit || Xi0 > k.pin_limit) Vh0 = 0.0;
        if (Xi1 > k.pin_limit || Xi1 < -k.pin_limit) Vh1 = 0.0;
    }
    const double imm = div_rn(1.0, mm);                                                   // one reciprocal for the five quotients
    node_update_r(Fx - rx, mm, imm, k.damp, k.rho, dt, dt_real, Vh0, V0, A0v);            // UL:471-503
    node_update_r(Fy - ry, mm, imm, k.damp, k.rho, dt, dt_real, Vh1, V1, A1v);
    s.V[i] = V0; s.V[np + i] = V1; s.A[i] = A0v; s.A[np + i] = A1v;
    s.ea_out[i] = ea0 + q0 * imm; s.ea_out[np + i] = ea1 + q1 * imm; s.ea_out[3 * np + i] = ea3 + q3 * imm;
    if (DBG) {
        dbg.res[i] = rx; dbg.res[np + i] = ry; dbg.f_fa[i] = Fx; dbg.f_fa[np + i] = Fy; dbg.mm[i] = mm;
        dbg.res_ea[i] = q0; dbg.res_ea[np + i] = q1; dbg.res_ea[3 * np + i] = q3;
    }
}

}   // namespace wcfe
