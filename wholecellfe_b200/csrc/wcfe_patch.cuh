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

#ifndef WCFE_PATCH
#define WCFE_PATCH 128
#endif
#ifndef WCFE_MINBLK
#define WCFE_MINBLK 5
#endif
#ifndef WCFE_THREADS
#define WCFE_THREADS 128
#endif
constexpr int kPatch = WCFE_PATCH;     // owned nodes per patch
constexpr int kThreads = WCFE_THREADS; // threads per CTA: a patch has ~2.35 elements per owned node, so 1.25 x kPatch threads
                                       // finish the element phase in two full rounds instead of two and a tail
constexpr int kMinBlk = WCFE_MINBLK;   // resident CTAs per SM the register allocation is tuned for
#ifndef WCFE_RING_CHUNK
#define WCFE_RING_CHUNK 3
#endif
constexpr int kRingChunk = WCFE_RING_CHUNK;   // one-ring entries gathered per batch of independent shared-memory loads

// Per-patch descriptor ("blob"): everything a CTA needs to know about its patch, contiguous and of one fixed stride for
// all patches of a surface, so that ONE TMA bulk copy (cp.async.bulk -> UBLKCP) stages it without any dependent lookup:
//   int32  hdr[4]           nh (halo nodes), nel (patch-local elements), 0, 0
//   int32  halo_ids[Hp]     global node ids of the halo nodes
//   uint16 elem_lv[3*Ep]    local vertex ids per element: < kPatch owned (node - first), >= kPatch halo
//   uint16 ring_ptr[kPatch+8]  CSR offsets of the owned nodes into ring16 (patch-relative)
//   uint16 ring16[Rp]       (local element << 2) | slot, ascending global element id per node
// (every section padded to 16 B).  The owned nodes' state arrays are contiguous too and ride the same mbarrier.
struct PatchView {
    int npatch;
    int max_halo, max_elem;          // shared-memory strides of the node / record arrays
    int stride;                      // blob bytes per patch
    int off_halo, off_lv, off_rptr, off_ring;   // byte offsets of the sections inside a blob
    const unsigned char *blob;       // [npatch * stride]
    const int *elem_ptr;             // [npatch+1] into elem_gid
    const int *elem_gid;             // global element id of every patch-local element (substrate: bucket lookup)
    const unsigned short *cand_l;    // parallel to BoundaryView::cand: local node ids (cell only)
    // node-centric substrate kernel (k_subs_ring): hdr | halo ids | ring_ptr | (j,k) local vertex pairs per ring entry
    int ring_stride, roff_halo, roff_rptr, roff_jk, roff_row;      // roff_row: table row of every local node (uint16; wcfe_api.cu ring_table_rows)
    const unsigned char *ring_blob;
    const int *ring_cyc;             // (element << 2 | slot) per ring entry in the blob's cyclic order (same offsets as SurfView::ring)
};

__host__ __device__ inline size_t ring_smem_bytes(int max_halo, int ring_stride) {
    return 16 + (size_t)ring_stride + sizeof(double) * 5 * ((size_t)kPatch + max_halo) + sizeof(double) * 5 * kPatch + 64;
}
__host__ __device__ inline size_t patch_smem_bytes(int n_node_arrays, int n_rec, int max_halo, int max_elem, int stride) {
    const size_t NL = (size_t)kPatch + max_halo, EM = (size_t)max_elem;
    size_t recs = sizeof(double) * n_rec * EM;
    const size_t staging = sizeof(double) * 8 * (size_t)kPatch;      // owned-node SoA landing zone (aliases the record area)
    if (recs < staging) recs = staging;
    return 16 + (size_t)stride + sizeof(double) * n_node_arrays * NL + recs + 64;
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ----------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int phase) {
    unsigned int ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}

// ---- substrate, node-centric variant -----------------------------------------------------------------------------------------
// The frozen-substrate element is so cheap (one division) that recomputing it at each of its three nodes costs fewer
// instructions AND less shared-memory traffic than computing it once and handing an 80-byte record to three gatherers
// (measured: k_subs_patch runs at 80 % of the shared-memory wavefront peak, 44 % of those wavefronts bank conflicts).  Here
// a thread owns a node and walks its one-ring: for every incident element it reads the two other vertices (x, y, e) from
// the packed node table and evaluates, with itself as local vertex 0,
//     S = mu (b_t - det(I-2e) I),   f_me = -(r1 + r2) = -(t/2) sign(W) S (G2y-G1y, G1x-G2x),   A0 = |W|/2,
// i.e. no element phase, no records, one barrier, 17 KB of shared memory per CTA.  Each node's fan is walked in cyclic order so
// that consecutive elements share a vertex already in registers (half the shared-memory reads); the nodal sums are therefore
// formed in fan order, not in the reference's ascending element order (UL:977-982): same terms, rounding-level difference,
// still fixed and bit-reproducible.
// Its own descriptor blob: hdr | halo ids | valence[kPatch] (uint8) | (j, k)[step][node] (ushort2: local ids of the element's
// other two vertices in cyclic order; step-major so that a warp's index loads are conflict-free).
#ifndef WCFE_RING_MINBLK
#define WCFE_RING_MINBLK 5
#endif

// ---- partitioned substrate (BASELINE configs[3], SURVEY 8e second mode): halo exchange fused into the kernel -------------------
// A rank holds one strip of the substrate plus a one-ring of GHOST nodes owned by its neighbours.  The only per-step data a
// neighbour needs is the new strain of the strip's border nodes (positions are frozen, V and A are read by their owner alone).
// HALO = true: the thread that owns a border node stores the node's new strain into its own ea_out AND, through peer memory
// (NVLink: the peer's buffer mapped with cudaIpcOpenMemHandle, or a plain pointer when both strips live in one process),
// straight into the ghost row of every neighbour that mirrors the node: plain stores, no fence (a system-scope release per
// border thread was measured at +0.03 ms per step: MEMBAR.SYS waits for the SM's whole store backlog).  The arrival is
// signalled once per neighbour by ONE thread of the next kernel in the stream, the following step's cell kernel: the kernel
// boundary (griddepcontrol.wait / stream order) makes all of this kernel's stores visible to that thread, which then bumps the
// neighbour's arrival counter by the number of rows with a system-scope release.  A CTA of the next step's substrate kernel
// whose node table contains ghost rows spins (acquire, system scope) until every neighbour's counter has reached steps x rows
// before it reads them.  No host involvement, no separate exchange kernel, no staging buffer: the transfer rides on the stores
// the kernel performs anyway, the signal on a kernel that runs anyway.
//   Safety of writing into the buffer the peer may still be reading (ea is double-buffered, out(s) == in(s-1)): every CTA that
// pushes owns a border node, i.e. a node next to a ghost, so its own node table holds ghost rows and it WAITS first -- for the
// neighbours' signal of step s-1, which they send only after their substrate kernel of step s-1 has completed, all of its ghost
// reads included.  Deadlock-free: substrate(s) waits for the neighbour's cell(s), which waits for nobody.
struct HaloPeer {
    double *ea[2];                   // the peer's strain buffer this rank writes at even / odd steps since the connect
    long long np;                    // its pitch
    unsigned long long *flag;        // this rank's arrival counter in the peer's memory
    unsigned long long nrecv;        // rows the peer pushes to this rank per step
    unsigned long long nsend;        // rows this rank pushes to the peer per step
};
struct HaloView {
                                     // (per-patch role bits live in the ring descriptor's header word 2: bit 0 the patch's node table
                                     //  holds ghost rows -- wait first --, bit 1 the patch owns ghost or border nodes)
    const unsigned char *node_role;  // [np] 1 ghost (computed by its owner: no stores here), 2 border (mirrored on a peer)
    const int *exp_ptr;              // [npatch+1] into exp_ent
    const int *exp_ent;              // triples (local thread, peer slot, row at the peer), grouped by patch
    const HaloPeer *peer;            // [npeer]
    const unsigned long long *flags; // [npeer] arrival counters in MY memory, bumped by the peers
    int npeer;
    unsigned int step;               // steps since the connect
    int *err;
};
enum : int { ERR_HALO_TIMEOUT = 16 };
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one thread of the kernel that FOLLOWS a halo substrate kernel in the stream, after its griddepcontrol.wait: tell every
// neighbour that the rows of the step just finished are in its ghost rows
__device__ __noinline__ void halo_signal(const HaloPeer *peer, int npeer) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    for (int q = 0; q < npeer; q++) if (peer[q].nsend) red_release_sys_add(peer[q].flag, peer[q].nsend);
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// one thread: wait until every neighbour's rows of the previous step have arrived.  Bounded: after 10 s (a peer died) the error
// flag is raised and this and every later wait returns at once -- a wrong answer that is reported, never a hung GPU.
// (scalars by value: a reference to the kernel parameter would make every thread copy the struct to its local stack)
__device__ __noinline__ void halo_wait(const unsigned long long *flags, const HaloPeer *peer, int npeer, unsigned int step, int *err) {
    if (step == 0) return;
    const unsigned long long t0 = global_ns();
    for (int q = 0; q < npeer; q++) {
        const unsigned long long want = (unsigned long long)step * peer[q].nrecv;
        while (ld_acquire_sys(flags + q) < want) {
            if (*(volatile int *)err & ERR_HALO_TIMEOUT) return;
            if (global_ns() - t0 > 10000000000ull) { atomicOr(err, ERR_HALO_TIMEOUT); return; }
            __nanosleep(64);
        }
    }
}

// ---- what one owned node does once its patch's node table is complete ----------------------------------------------------------
// The table holds, per local node, (x, y), (e_xx, e_yy) and e_xy as stored in global memory; a node's row is where the descriptor
// says (row_of, chosen on the host against bank conflicts in the gathers below); (j, k) of a fan entry are rows.
struct RingTable { const double2 *sxy, *se01; const double *se3; };

template <bool DBG, bool HALO>
__device__ __forceinline__ void ring_node(const SurfView &s, const PatchView &pv, const SurfView &c, const BucketView &bk, const Consts &k,
                                          double dt, double dt_real, double t_tot, double bx0, double bx1, double by0, double by1,
                                          const DebugView &dbg, const HaloView &hv, double dt_drv, const RingTable T,
                                          const unsigned char *valence, const ushort2 *rjk, int p, int tid, int i, int nrole, bool zero_list,
                                          double V0, double V1, double A0v, double A1v, int myrow) {
    const int np = s.np;
    const double2 *sxy = T.sxy, *se01 = T.se01;
    const double *se3 = T.se3;
    const double2 myxy = sxy[myrow], mye01 = se01[myrow];
    const double ea0 = mye01.x, ea1 = mye01.y, ea3 = se3[myrow];
    double rx = 0.0, ry = 0.0, q0 = 0.0, q1 = 0.0, q3 = 0.0, mm = 0.0;
    const int nring = valence[tid];
    // one incident element, with this node as local vertex 0 and (j, k) the other two in cyclic order.
    // Frozen substrate, F == I (UL:988-1174 specialised, wcfe_math.cuh elem_eval_frozen), with a = I - 2 e_avg, b01 = -a01:
    //   S = mu (a^-1 - det(a) I),  Re_0 = -(t/2) sign(W) S (d x, d y),  Re_ea = (e_avg - ea_0) A0 t / 3,  MM = A0 t / 3
    // written so that every scale factor rides on the accumulating multiply-add.  I - 2 e_avg = 1 - (2/3)(e_i + e_j + e_k): the
    // own node's share is formed once, the neighbours' enter through one addition and one multiply-add per component.
    struct Nb { double2 p, t; double t3; };
    auto load_nb = [&](int l) { Nb n; n.p = sxy[l]; n.t = se01[l]; n.t3 = se3[l]; return n; };
    const double two3 = 2.0 / 3.0;
    // every product and sum below is contracted BY HAND (__dmul_rn / fma): this routine is instantiated in several kernels, and
    // a strip of a partitioned run must equal the whole mesh on one GPU bit for bit whatever each instantiation's compiler pass
    // would have fused
    const double c0 = __dsub_rn(1.0, __dmul_rn(two3, ea0)), c1 = __dsub_rn(1.0, __dmul_rn(two3, ea1)), c3 = __dmul_rn(two3, ea3);   // own node's share of a00, a11, b01
    const double A0c = __dsub_rn(1.0, __dmul_rn(2.0, ea0)), A1c = __dsub_rn(1.0, __dmul_rn(2.0, ea1)), B3c = __dmul_rn(2.0, ea3);   // a at e_avg == ea_own: Re_ea vanishes there
    const double mh = __dmul_rn(__dmul_rn(0.5, k.mu_subs), k.thick), cw = div_rn(k.thick, 6.0), cwh = div_rn(k.thick, 12.0);
    struct Contrib { double f1, f2, m, a00, a11, b01, aw; };
    auto entry = [&](const Nb &J, const Nb &K) {
        Contrib c;
        c.a00 = fma(-two3, __dadd_rn(J.t.x, K.t.x), c0); c.a11 = fma(-two3, __dadd_rn(J.t.y, K.t.y), c1); c.b01 = fma(two3, __dadd_rn(J.t3, K.t3), c3);
        const double G1x = __dsub_rn(J.p.x, myxy.x), G1y = __dsub_rn(J.p.y, myxy.y), G2x = __dsub_rn(K.p.x, myxy.x), G2y = __dsub_rn(K.p.y, myxy.y);
        const double Wz = fma(G1x, G2y, -__dmul_rn(G1y, G2x));
        const double det2 = fma(c.a00, c.a11, -__dmul_rn(c.b01, c.b01));
        const double idet = rcp_newton(det2);
        const double dx = __dsub_rn(G2y, G1y), dy = __dsub_rn(G1x, G2x);
        c.f1 = fma(idet, fma(c.a11, dx, __dmul_rn(c.b01, dy)), -__dmul_rn(det2, dx));                // (S / mu) (dx, dy)
        c.f2 = fma(idet, fma(c.b01, dx, __dmul_rn(c.a00, dy)), -__dmul_rn(det2, dy));
        c.m = Wz > 0 ? -mh : mh;                                           // -(t/2) mu sign(W)
        c.aw = fabs(Wz);                                                   // 2 A0
        return c;
    };
    auto add = [&](const Contrib &c) {
        rx = fma(c.m, c.f1, rx); ry = fma(c.m, c.f2, ry);
        mm = fma(c.aw, cw, mm);                                            // UL:1172
        const double wh = __dmul_rn(c.aw, cwh);                            // A0 t / 6: (e_avg - ea) = (a_own - a) / 2
        q0 = fma(__dsub_rn(A0c, c.a00), wh, q0); q1 = fma(__dsub_rn(A1c, c.a11), wh, q1); q3 = fma(__dsub_rn(c.b01, B3c), wh, q3);   // UL:1170 with e_s == e for F == I
    };
    // The fan is walked two elements at a time: their dependency chains (loads -> strain -> determinant -> reciprocal -> stress)
    // are independent, which doubles the instruction-level parallelism of a loop that is otherwise latency bound.
    // (keeping a closed fan's first neighbour in registers for its last element was measured: 0.188 -> 0.224 ms, the selects and
    // the ten extra live registers cost more than the saved table read.)
    int lastk = -1;
    Nb prev = {make_double2(0, 0), make_double2(0, 0), 0.0};
    int r = 0;
    for (; r + 1 < nring; r += 2) {
        const ushort2 jkA = rjk[r * kPatch + tid], jkB = rjk[(r + 1) * kPatch + tid];
        const Nb JA = (int)jkA.x == lastk ? prev : load_nb(jkA.x);
        const Nb KA = load_nb(jkA.y);
        const Nb JB = jkB.x == jkA.y ? KA : load_nb(jkB.x);
        const Nb KB = load_nb(jkB.y);
        const Contrib cA = entry(JA, KA), cB = entry(JB, KB);
        add(cA); add(cB);
        prev = KB; lastk = jkB.y;
    }
    if (r < nring) {
        const ushort2 jk = rjk[r * kPatch + tid];
        const Nb J = (int)jk.x == lastk ? prev : load_nb(jk.x);
        const Nb K = load_nb(jk.y);
        add(entry(J, K));
    }
    const double Xi0 = myxy.x, Xi1 = myxy.y;
    double Fx = 0.0, Fy = 0.0;
    // one thread always waits: this kernel's completion then implies that of the boundary finish, which the next cell kernel
    // (chained behind this kernel only) depends on; it also re-zeroes the triangle list k_bucket_reduce has consumed
    if (zero_list) {
        pdl_wait();
        if (bk.big_count && !bk.persist) *bk.big_count = 0;
        // partitioned cell: this step's cell kernel, boundary finish and reduction are complete on this rank -- every row it pushes is
        // out, every pushed row it reads has been read
        if (bk.cell_peer) cell_part_signal(bk.cell_peer, bk.cell_npeer, 0);
    }
    if (HALO && nrole == 1) return;          // ghost node (incomplete fan): its owner computes it and stores the result here
    unsigned int stamp = bk.epoch - 1u;
    if (Xi0 >= bx0 && Xi0 <= bx1 && Xi1 >= by0 && Xi1 <= by1) {
        // first touch of what the cell kernel produced (stamps, buckets, anchor records); all above overlaps its tail
        if (bk.wait_cell) {
            if (ld_acquire_gpu(bk.cell_done) < bk.cell_target) {        // all CTAs of this step's cell kernel are resident or done: bounded anyway
                const unsigned long long t0 = global_ns();
                while (ld_acquire_gpu(bk.cell_done) < bk.cell_target) {
                    if (global_ns() - t0 > 2000000000ull) { atomicOr(bk.chain_err, ERR_CHAIN); break; }
                    __nanosleep(32);
                }
            }
        } else pdl_wait();
        stamp = bk.node_epoch[i];
    }
    if (stamp == bk.epoch) {      // an incident triangle received an anchor in this step (two thirds of the nodes in the box skip this)
        // calculateForce_FA_subs (UL:1297-1454) as a gather: k_bucket_reduce has formed, for every triangle that holds anchors, the
        // six sums  sum_anchors H (A_j / A_tri)  (anchors in ascending cell-node order) and parked them in the triangle's own item
        // storage; this node adds the one of its slot, triangles in fan order.  element id -> bucket head -> sums are dependent
        // global loads: fetched for four incident triangles at a time.
        const int gb = s.ring_ptr[i];
        for (int r0 = 0; r0 < nring; r0 += 4) {
            int code4[4]; unsigned long long h4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) code4[u] = r0 + u < nring ? pv.ring_cyc[gb + r0 + u] : -1;
#pragma unroll
            for (int u = 0; u < 4; u++) h4[u] = code4[u] >= 0 ? bk.head[code4[u] >> 2] : 0ull;
            double2 f4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool on = code4[u] >= 0 && (unsigned int)(h4[u] >> 32) == bk.epoch;
                const double2 *sums = bk.persist ? reinterpret_cast<const double2 *>(bk.sums + (size_t)6 * (code4[u] >> 2))
                                                 : reinterpret_cast<const double2 *>(bk.items + (size_t)(code4[u] >> 2) * bk.cap);
                f4[u] = on ? sums[code4[u] & 3] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { Fx = __dadd_rn(Fx, f4[u].x); Fy = __dadd_rn(Fy, f4[u].y); }      // (+0 for the triangles without anchors)
        }
    }
    double Vh0 = WCFE_ADD(V0, WCFE_MUL(dt_drv / 2.0, A0v)), Vh1 = WCFE_ADD(V1, WCFE_MUL(dt_drv / 2.0, A1v));   // UL:295, with the dt the driver was called with
    if (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f) {            // UL:298-341, Vbiax == 0
        if (Xi0 < -k.pin_limit || Xi0 > k.pin_limit) Vh0 = 0.0;
        if (Xi1 > k.pin_limit || Xi1 < -k.pin_limit) Vh1 = 0.0;
    }
    // UL:471-503 with one reciprocal of the lumped mass for the five quotients, hand-contracted like the walk
    const double imm = div_rn(1.0, mm), ir = div_rn(imm, k.rho), dmr = __dmul_rn(__dmul_rn(k.damp, mm), k.rho);
    const double hdt = dt / 2.0, ddt = __dsub_rn(dt_real, hdt);
    A0v = __dmul_rn(fma(-dmr, V0, __dsub_rn(Fx, rx)), ir);
    A1v = __dmul_rn(fma(-dmr, V1, __dsub_rn(Fy, ry)), ir);
    V0 = dt_real > hdt ? fma(ddt, A0v, Vh0) : fma(dt_real, A0v, V0);
    V1 = dt_real > hdt ? fma(ddt, A1v, Vh1) : fma(dt_real, A1v, V1);
    s.V[i] = V0; s.V[np + i] = V1; s.A[i] = A0v; s.A[np + i] = A1v;
    const double en0 = fma(q0, imm, ea0), en1 = fma(q1, imm, ea1), en3 = fma(q3, imm, ea3);
    s.ea_out[i] = en0; s.ea_out[np + i] = en1; s.ea_out[3 * np + i] = en3;
    if (HALO && nrole == 2) {
        // border node: the same three words go straight into the ghost row of every neighbour that mirrors it (signalled by the
        // next kernel in the stream, halo_signal)
        for (int q = hv.exp_ptr[p]; q < hv.exp_ptr[p + 1]; q++) {
            if (hv.exp_ent[3 * q] != tid) continue;
            const HaloPeer *pr = hv.peer + hv.exp_ent[3 * q + 1];
            const long long row = hv.exp_ent[3 * q + 2], rnp = pr->np;
            double *dst = pr->ea[hv.step & 1];
            dst[row] = en0; dst[rnp + row] = en1; dst[3 * rnp + row] = en3;
        }
    }
    if (DBG) {
        dbg.res[i] = rx; dbg.res[np + i] = ry; dbg.f_fa[i] = Fx; dbg.f_fa[np + i] = Fy; dbg.mm[i] = mm;
        dbg.res_ea[i] = q0; dbg.res_ea[np + i] = q1; dbg.res_ea[3 * np + i] = q3;
    }
}

// ---- one CTA per patch.  (A persistent, software-pipelined form -- every CTA with three patches in flight: descriptor + owned
//      rows of patch j+2 by cp.async.bulk, halo rows of patch j+1 by cp.async/LDGSTS into a second node table, fans of patch j --
//      was built and measured in round 2: 0.198 ms against 0.185 ms for this form; the three barriers per patch and the table fill
//      on every iteration's critical path cost more than the hidden latencies return.  profiles/r02_notes.md.) -----------------
template <bool DBG, bool BATCH = false, bool HALO = false>
__global__ void __launch_bounds__(kPatch, WCFE_RING_MINBLK) k_subs_ring(SurfView s, PatchView pv, SurfView c, BucketView bk, Consts k_, double dt, double dt_real,
                                                         double t_tot, double bx0, double bx1, double by0, double by1, DebugView dbg, BatchView bt, HaloView hv, double dt_drv) {
    extern __shared__ __align__(128) unsigned char smraw[];
    pdl_launch_dependents();      // the next step's cell kernel may stage its patches behind this kernel's last wave
    const Consts &k = BATCH ? bt.k[blockIdx.y] : k_;
    if (BATCH) { const size_t r = blockIdx.y; batch_shift(s, r); batch_shift(c, r); batch_shift(bk, bt, r); }
    const int NL = kPatch + pv.max_halo;
    unsigned long long *bar = (unsigned long long *)smraw;
    unsigned char *sblob = smraw + 16;
    double2 *sxy = (double2 *)(sblob + pv.ring_stride);  // [NL]
    double2 *se01 = sxy + NL;                             // [NL]
    double *se3 = (double *)(se01 + NL);                  // [NL]
    double *stage = se3 + NL;                             // [5][kPatch] landing zone of the bulk copies
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n0 = p * kPatch;
    const int nown = min(kPatch, s.nn - n0);
    const int ncopy = min(kPatch, s.np - n0);
    const int np = s.np;
    if (tid == 0) {
        mbar_init(bar, 1);
        const unsigned int nb = (unsigned int)ncopy * 8u;
        mbar_expect_tx(bar, (unsigned int)pv.ring_stride + 5u * nb);
        bulk_g2s(sblob, pv.ring_blob + (size_t)p * pv.ring_stride, (unsigned int)pv.ring_stride, bar);
        bulk_g2s(stage, s.x + n0, nb, bar); bulk_g2s(stage + kPatch, s.x + np + n0, nb, bar);
        bulk_g2s(stage + 2 * kPatch, s.ea_in + n0, nb, bar); bulk_g2s(stage + 3 * kPatch, s.ea_in + np + n0, nb, bar);
        bulk_g2s(stage + 4 * kPatch, s.ea_in + 3 * (size_t)np + n0, nb, bar);
    }
    const int i = n0 + tid;
    const bool own = tid < nown;
    double V0 = 0, V1 = 0, A0v = 0, A1v = 0;
    if (own) { V0 = s.V[i]; V1 = s.V[np + i]; A0v = s.A[i]; A1v = s.A[np + i]; }
    __syncthreads();
    mbar_wait(bar, 0);
    const int nh = ((const int *)sblob)[0];
    // partitioned run: the patch's role rides in the descriptor's header (bit 0: its node table holds ghost rows, bit 1: it owns
    // ghost or border nodes), so the staging above is issued without any dependent load and interior patches pay nothing
    const int prole = HALO ? ((const int *)sblob)[2] : 0;
    int nrole = 0;
    if (HALO && prole) {
        if (own && (prole & 2)) nrole = (int)hv.node_role[i];
        if (prole & 1) {
            if (tid == 0) halo_wait(hv.flags, hv.peer, hv.npeer, hv.step, hv.err);   // the neighbours' rows of the previous step have landed
            __syncthreads();
            // ghost nodes inside the patch's own range were staged before the wait: read them again, past the (incoherent) L1
            if (nrole == 1) { stage[2 * kPatch + tid] = __ldcg(s.ea_in + i); stage[3 * kPatch + tid] = __ldcg(s.ea_in + np + i); stage[4 * kPatch + tid] = __ldcg(s.ea_in + 3 * (size_t)np + i); }
        }
    }
    const int *halo_ids = (const int *)(sblob + pv.roff_halo);
    const unsigned char *valence = sblob + pv.roff_rptr;
    const ushort2 *rjk = (const ushort2 *)(sblob + pv.roff_jk);          // [step][node]: conflict-free across a warp
    const unsigned short *row_of = (const unsigned short *)(sblob + pv.roff_row);
    const int myrow = row_of[tid];
    sxy[myrow] = make_double2(stage[tid], stage[kPatch + tid]); se01[myrow] = make_double2(stage[2 * kPatch + tid], stage[3 * kPatch + tid]); se3[myrow] = stage[4 * kPatch + tid];
    if (HALO && (prole & 1)) {
        // ghost rows were written by a peer GPU into L2 while this kernel was already running: read them past the (incoherent) L1
        for (int l = tid; l < nh; l += kPatch) {
            const int g = halo_ids[l], r = row_of[kPatch + l];
            sxy[r] = make_double2(s.x[g], s.x[np + g]);
            se01[r] = make_double2(__ldcg(s.ea_in + g), __ldcg(s.ea_in + np + g));
            se3[r] = __ldcg(s.ea_in + 3 * (size_t)np + g);
        }
    } else {
        for (int l = tid; l < nh; l += kPatch) {
            const int g = halo_ids[l], r = row_of[kPatch + l];
            sxy[r] = make_double2(s.x[g], s.x[np + g]);
            se01[r] = make_double2(s.ea_in[g], s.ea_in[np + g]);
            se3[r] = s.ea_in[3 * (size_t)np + g];
        }
    }
    __syncthreads();
    if (!own) return;
    ring_node<DBG, HALO>(s, pv, c, bk, k, dt, dt_real, t_tot, bx0, bx1, by0, by1, dbg, hv, dt_drv, RingTable{sxy, se01, se3}, valence, rjk, p, tid, i, nrole,
                         blockIdx.x == 0 && tid == 0, V0, V1, A0v, A1v, myrow);
}

// ---- integrin force on the substrate, per triangle ---------------------------------------------------------------------------------
// calculateForce_FA_subs (UL:1297-1454) adds, for every anchor inside a substrate triangle, H * A_j/A_tri to the triangle's three
// nodes, anchors in ascending cell-node order.  The cell kernel lists every triangle that received its first anchor of the step;
// here each listed triangle evaluates the containment sub-areas of its anchors ONCE (a per-node gather would do it three times,
// after three dependent lookups) and leaves the six sums in the head of the bucket's own item storage (6 doubles = 12 ints <=
// cap), where k_subs_ring picks up the one of its slot.
//   WARP = false  one thread per triangle (fine substrates: a couple of anchors per triangle, ~ne_cell/2 triangles listed)
//   WARP = true   one warp per triangle (coarse substrate under a fine cell: 88 anchors in one triangle of the example meshes):
//                 the lanes rank the items (all pairs, 1/32 each), evaluate them in parallel, park the terms in shared memory
//                 by rank, and six lanes form the six sums in the reference's order.
constexpr int kReduceWarps = 8, kReduceWindow = 128;
// (the work of block `bid` of `nblk`; the caller has passed griddepcontrol.wait.  term: [kReduceWarps][kReduceWindow][6] doubles of shared
// memory for WARP, unused otherwise)
template <bool BATCH, bool WARP>
__device__ __forceinline__ void bucket_reduce_body(int bid, int nblk, double (*term)[kReduceWindow][6], SurfView s, SurfView c, BucketView bk, const BatchView &bt) {
    if (BATCH) { const size_t r = blockIdx.y; batch_shift(s, r); batch_shift(c, r); batch_shift(bk, bt, r); }
    const int nbig = (int)min(*bk.big_count, (unsigned int)bk.big_cap);
    const double2 *anc2 = reinterpret_cast<const double2 *>(c.anc);
    if (!WARP) {
        for (int b = bid * blockDim.x + threadIdx.x; b < nbig; b += nblk * blockDim.x) {
            const int e = bk.big_list[b];
            const unsigned long long hw = bk.head[e];
            int cnt = (int)((unsigned int)hw & (bk.persist ? kHeadCount : 0xffffffffu));
            if (bk.persist && cnt == 0) continue;    // a listed triangle nobody is in at the moment (it stays listed: an anchor that comes back does not list it twice)
            if (cnt > bk.cap) cnt = bk.cap;
            int *items = bk.items + (size_t)e * bk.cap;
            const int i0 = s.tri[e], i1 = s.tri[s.nep + e], i2 = s.tri[2 * s.nep + e];
            const double n1x = s.x[i0], n1y = s.x[s.np + i0], n2x = s.x[i1], n2y = s.x[s.np + i1], n3x = s.x[i2], n3y = s.x[s.np + i2];
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
            int nlive = 0, last = -1; bool dirty = false, sorted = true;
            for (int q = 0; q < cnt; q++) {
                const int v = items[q];
                if (v < 0) { dirty = true; continue; }
                nlive++; dirty |= (v & kItemVolatile) != 0;
                sorted &= (v & kItemId) > last; last = v & kItemId;
            }
            auto add_anchor = [&](int ni) {
                const double2 ap = anc2[2 * (size_t)ni], ah = anc2[2 * (size_t)ni + 1];
                double Ar[3], At;                                                        // UL:1303-1340; planar form, bit-identical
                point_in_triangle_2d(n1x, n1y, n2x, n2y, n3x, n3y, ap.x, ap.y, Ar, At);
                const Recip rt = recip_of(At);
                const double w0 = quot(Ar[0], rt), w1 = quot(Ar[1], rt), w2 = quot(Ar[2], rt);
                a0 += ah.x * w0; a1 += ah.y * w0; a2 += ah.x * w1; a3 += ah.y * w1; a4 += ah.x * w2; a5 += ah.y * w2;   // UL:1414-1422
            };
            // ascending cell-node order: deterministic whatever the insertion order was.  A persistent bucket is left in that order
            // (below), so from the second step on the entries are simply walked; otherwise the next larger id is searched each time.
            if (sorted) {
                for (int q = 0; q < cnt; q++) { const int v = items[q]; if (v >= 0) add_anchor(v & kItemId); }
            } else {
                int prev = -1;
                for (int m = 0; m < nlive; m++) {
                    int ni = 0x7fffffff;
                    for (int q = 0; q < cnt; q++) { const int v = items[q]; const int id = v & kItemId; if (v >= 0 && id > prev && id < ni) ni = id; }
                    prev = ni;
                    add_anchor(ni);
                }
            }
            // the triangle's three nodes have something to gather in this step
            bk.node_epoch[i0] = bk.epoch; bk.node_epoch[i1] = bk.epoch; bk.node_epoch[i2] = bk.epoch;
            if (bk.persist) {
                double2 *o = reinterpret_cast<double2 *>(bk.sums + (size_t)6 * e);
                o[0] = make_double2(a0, a1); o[1] = make_double2(a2, a3); o[2] = make_double2(a4, a5);
                int w = cnt;
                if (dirty) { w = 0; for (int q = 0; q < cnt; q++) { const int v = items[q]; if (v >= 0 && !(v & kItemVolatile)) items[w++] = v; } }   // this thread alone owns the bucket now
                if (!sorted)          // (membership changed, or the first step: once; insertion sort of a handful of ids)
                    for (int a = 1; a < w; a++) { const int key = items[a]; int q = a - 1; while (q >= 0 && items[q] > key) { items[q + 1] = items[q]; q--; } items[q + 1] = key; }
                bk.head[e] = ((unsigned long long)bk.epoch << 32) | kHeadListed | (unsigned int)w;
            } else {
                double2 *o = reinterpret_cast<double2 *>(items);      // this thread alone reads and writes the bucket; its count starts the next step at zero
                o[0] = make_double2(a0, a1); o[1] = make_double2(a2, a3); o[2] = make_double2(a4, a5);
                bk.head[e] = (unsigned long long)bk.epoch << 32;
            }
        }
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double (*tw)[6] = term[WARP ? warp : 0];
    for (int b = bid * kReduceWarps + warp; b < nbig; b += nblk * kReduceWarps) {
        const int e = bk.big_list[b];
        int cnt = (int)((unsigned int)bk.head[e] & (bk.persist ? kHeadCount : 0xffffffffu));
        if (bk.persist && cnt == 0) continue;        // a listed triangle nobody is in at the moment (uniform across the warp)
        if (cnt > bk.cap) cnt = bk.cap;
        int *items = bk.items + (size_t)e * bk.cap;
        const int i0 = s.tri[e], i1 = s.tri[s.nep + e], i2 = s.tri[2 * s.nep + e];
        const double n1[3] = {s.x[i0], s.x[s.np + i0], 0.0}, n2[3] = {s.x[i1], s.x[s.np + i1], 0.0}, n3[3] = {s.x[i2], s.x[s.np + i2], 0.0};
        int nlive = 0, dirty = 0, ordered = 1;
        for (int m = lane; m < cnt; m += 32) {
            const int v = items[m];
            if (v < 0) { dirty = 1; continue; }
            nlive++; dirty |= (v & kItemVolatile) != 0;
            if (m) { const int u = items[m - 1]; ordered &= (u >= 0 && (u & kItemId) < (v & kItemId)); }
        }
        nlive = __reduce_add_sync(0xffffffffu, nlive); dirty = __any_sync(0xffffffffu, dirty);
        const bool sorted = !dirty && __all_sync(0xffffffffu, ordered);      // a persistent bucket is left in ascending order (below): rank == position
        double acc = 0.0;                                   // lanes 0..5: (vertex, component) = (lane >> 1, lane & 1)
        for (int w0 = 0; w0 < nlive; w0 += kReduceWindow) {   // one window unless a bucket holds more than 128 anchors
            for (int m = lane; m < cnt; m += 32) {
                const int vv = items[m];
                if (vv < 0) continue;
                const int v = vv & kItemId;
                int rank = m;
                if (!sorted) { rank = 0; for (int q = 0; q < cnt; q++) { const int u = items[q]; rank += (u >= 0 && (u & kItemId) < v); } }      // node ids inside a bucket are distinct
                if (rank < w0 || rank >= w0 + kReduceWindow) continue;
                const double2 ap = anc2[2 * (size_t)v], ah = anc2[2 * (size_t)v + 1];
                double Ar[3], At;
                point_in_triangle_2d(n1[0], n1[1], n2[0], n2[1], n3[0], n3[1], ap.x, ap.y, Ar, At);     // planar form, bit-identical
                double *t6 = tw[rank - w0];
                const Recip rt = recip_of(At);
#pragma unroll
                for (int j = 0; j < 3; j++) { const double w = quot(Ar[j], rt); t6[2 * j] = ah.x * w; t6[2 * j + 1] = ah.y * w; }
            }
            __syncwarp();
            if (lane < 6) {
                const int n = min(kReduceWindow, nlive - w0);
#pragma unroll 8
                for (int r = 0; r < n; r++) acc += tw[r][lane];      // the loads are independent of the running sum: unrolled, they pipeline
            }
            __syncwarp();
        }
        if (bk.persist) {
            if (lane < 6) bk.sums[(size_t)6 * e + lane] = acc;
            int w = cnt;
            if (!sorted) {
                // membership changed (or the first step): leave the survivors in ascending order -- every lane ranks its entries
                // among them, the order is staged in this warp's shared scratch (all lanes are past its last use) and copied back
                int npers = 0;
                for (int m = lane; m < cnt; m += 32) { const int v = items[m]; npers += (v >= 0 && !(v & kItemVolatile)); }
                npers = __reduce_add_sync(0xffffffffu, npers);
                int *stg = reinterpret_cast<int *>(&tw[0][0]);
                if (npers <= kReduceWindow * 6 * 2) {
                    for (int m = lane; m < cnt; m += 32) {
                        const int v = items[m];
                        if (v < 0 || (v & kItemVolatile)) continue;
                        int rank = 0;
                        for (int q = 0; q < cnt; q++) { const int u = items[q]; rank += (u >= 0 && !(u & kItemVolatile) && u < v); }
                        stg[rank] = v;
                    }
                    __syncwarp();
                    for (int m = lane; m < npers; m += 32) items[m] = stg[m];
                    __syncwarp();
                    w = npers;
                } else if (dirty) {      // (a bucket too large for the scratch: compacted by one lane, order left as it is)
                    if (lane == 6) { w = 0; for (int q = 0; q < cnt; q++) { const int v = items[q]; if (v >= 0 && !(v & kItemVolatile)) items[w++] = v; } }
                    w = __shfl_sync(0xffffffffu, w, 6);
                }
            }
            if (lane == 6) {
                bk.node_epoch[i0] = bk.epoch; bk.node_epoch[i1] = bk.epoch; bk.node_epoch[i2] = bk.epoch;
                bk.head[e] = ((unsigned long long)bk.epoch << 32) | kHeadListed | (unsigned int)w;
            }
        } else {
            if (lane < 6) reinterpret_cast<double *>(items)[lane] = acc;     // every lane has finished reading the items (syncwarp above)
            if (lane == 6) { bk.node_epoch[i0] = bk.epoch; bk.node_epoch[i1] = bk.epoch; bk.node_epoch[i2] = bk.epoch; bk.head[e] = (unsigned long long)bk.epoch << 32; }
        }
    }
}

template <bool BATCH, bool WARP>
__global__ void __launch_bounds__(kReduceWarps * 32) k_bucket_reduce(SurfView s, SurfView c, BucketView bk, BatchView bt) {
    __shared__ double term[WARP ? kReduceWarps : 1][WARP ? kReduceWindow : 1][6];
    pdl_launch_dependents();
    pdl_wait();                   // the list and the buckets are the cell kernel's output
    bucket_reduce_body<BATCH, WARP>(blockIdx.x, gridDim.x, reinterpret_cast<double (*)[kReduceWindow][6]>(term), s, c, bk, bt);
}

// ---- boundary finish and bucket reduction in ONE launch (production chain) ------------------------------------------------------
// Both follow the cell kernel, neither needs the other, both are a handful of latency-bound blocks: as separate links of the
// programmatic-dependent-launch chain the reduction could not start before the LAST block of the finish kernel had retired
// (griddepcontrol.wait waits for the whole prerequisite grid) -- 13-18 us per step in which the GPU idles.  Blocks [0, nfb) finish
// the boundary, the others reduce; the substrate kernel's wait then covers both.
template <bool BATCH, bool PART, bool WARP>
__global__ void __launch_bounds__(kReduceWarps * 32) k_finish_reduce(int nfb, SurfView cs, BoundaryView b, Consts k_, int npart, const double *__restrict__ partials,
                                                                     const double *__restrict__ V_in, const double *__restrict__ A_in, double dt, double dt_real,
                                                                     double *__restrict__ diag, int *__restrict__ err, BatchView bt, double dt_drv, CellPartView cp,
                                                                     SurfView ss, BucketView bk) {
    __shared__ double smem_d[WARP ? kReduceWarps * kReduceWindow * 6 : 4 * 32 + 4];
    static_assert(kReduceWarps * kReduceWindow * 6 >= 4 * 32 + 4, "the finish blocks' scratch fits the reduction's");
    pdl_launch_dependents();
    pdl_wait();                   // everything below reads what the cell kernel wrote
    if ((int)blockIdx.x < nfb)
        boundary_finish_body<false, BATCH, PART>(blockIdx.x, nfb, smem_d, smem_d + 4 * 32, cs, b, k_, npart, partials, V_in, A_in, dt, dt_real, diag, err, DebugView{}, bt, dt_drv, cp);
    else
        bucket_reduce_body<BATCH, WARP>((int)blockIdx.x - nfb, (int)gridDim.x - nfb, reinterpret_cast<double (*)[kReduceWindow][6]>(smem_d), ss, cs, bk, bt);
}

// ---- cell: the whole cell side of one UL step in ONE kernel ---------------------------------------------------------
//   PREDICT: driverVelocities_Explicit + inline predictor (UL:103-124, main.cpp:224-242) for owned AND halo nodes, read
//            from the current position buffer and written to the other one (ping-pong: the old buffer becomes X_t).
//            Because a patch predicts its halo nodes from THEIR V, A, no patch may overwrite V, A in place: x, V, A and ea
//            are all double-buffered (read buffer / write buffer, swapped by the host after the launch);
//   updateFA (UL:2005-2114) and the anchor location / bucket insertion for calculateForce_FA_subs (UL:1301-1340);
//   updateForces_Explicit, cell overload (UL:508-630): residual, integrin, drag, bending, actin, update.
// Boundary vertices wait for the total area exactly as in k_cell_node (finished by k_boundary_finish).
struct CellStepArgs {
    double dt, dt_real, t_tot, ap_threshold;
    double dt_drv;                  // the dt driverVelocities_Explicit formed V_hlf with (main.cpp:224, :329): differs from dt when El Topo shortened a step
    double ix0, ix1, iy0, iy1;
    int rng_mode; unsigned int seed, step;
    const unsigned int *tape;
    double *partials;
    int *d_flags, *d_loc_bond, *d_loc_cnt, *d_loc_first, *d_ap_fire, *err;
    // per cell node, where its anchor was last LOCATED: [2][np] double2 = (anchor x, anchor y) at that moment and (c^2, triangle as
    // the bits of a 64-bit integer), c the radius around that point inside which every point lies in the triangle with all three
    // sub-areas above the margin of the previous-triangle shortcut (anchor_clearance2).  While the anchor stays inside the circle --
    // thousands of steps at the rates anchors creep -- one coalesced 32-byte read and two multiply-adds replace the containment test
    // and the gathers loc_first -> tri -> x behind it.  triangle < 0 or c^2 <= 0 (also NaN: the 0xff fill): nothing usable, locate.
    // A geometric fact about the frozen substrate: invalidated when substrate positions are uploaded, by nothing else.
    double2 *loc_cache;
    int debug;
    const HaloPeer *halo_peer;      // partitioned run, not the first step: signal the previous step's halo rows (halo_signal)
    int halo_npeer;
    CellPartView part;              // partitioned cell (PART instantiation): patch list, pushes, counters
};

// Shared memory: the seven node arrays in SoA (TMA destinations), element records as array-of-structs, kRec doubles per
// element at a compile-time stride (one address computation per one-ring entry, immediate offsets for the nine loads).
constexpr int kRec = 9;           // r1x r1y r2x r2y e_s.xx e_s.yy e_s.xy w a3:  w = A0 t / 3 (UL:911-913), a3 = a_cur / 3 (UL:1251)
__host__ __device__ inline size_t cell_smem_bytes(int max_halo, int max_elem, int stride) {
    const size_t NL = (size_t)kPatch + max_halo;
    return 16 + (size_t)stride + sizeof(double) * 7 * NL + sizeof(double) * kRec * (size_t)max_elem + 64;
}
// append a cell node to the bucket of substrate triangle t; returns the slot.
//   With the per-triangle reduction behind this kernel (bk.big_count != nullptr, the production chain): ONE 64-bit ATOM.ADD on the
//   head word.  Its low half is the anchor count of the step; k_finish_reduce / k_bucket_reduce, which visits every triangle that
//   received an anchor, leaves it at zero again (and the current epoch in the high half, for the substrate kernel) -- no reset pass,
//   no epoch test, no second atomic here, and the result (the slot) may stay unconsumed for as long as the caller likes.
//   Without it (k_subs_patch walks the buckets itself): heads are (epoch << 32 | count) and never reset; a fire-and-forget RED.MAX
//   brings the head to the current epoch first (same-address atomics of one thread are performed in program order).
// Measured (timing experiment, bench cell): anchor location + insertion cost 33 of k_cell_patch's 95 us -- 2.5 M scattered atomics
// and partial-sector stores per launch; this form and the epoch stamps moved into the reduction leave 1 M.
// margin of the previous-triangle shortcut (relative sub-area) and the squared radius around p inside which it holds: a sub-area is
// linear in p and changes by at most |edge| |dp| / 2, so |dp| < 2 (min A_k - margin A_tri) / max |edge| keeps all three above the
// margin; 0.999 absorbs the rounding of this very computation (the margin itself is seven orders above the predicate's).
constexpr double kAnchorMargin = 1e-9;
__device__ __forceinline__ double anchor_clearance2(const double *Ar, double At, double n1x, double n1y, double n2x, double n2y, double n3x, double n3y) {
    const double amin = fmin(Ar[0], fmin(Ar[1], Ar[2])) - kAnchorMargin * At;
    if (!(amin > 0.0)) return 0.0;
    const double e1 = (n2x - n1x) * (n2x - n1x) + (n2y - n1y) * (n2y - n1y), e2 = (n3x - n1x) * (n3x - n1x) + (n3y - n1y) * (n3y - n1y),
                 e3 = (n3x - n2x) * (n3x - n2x) + (n3y - n2y) * (n3y - n2y);
    const double l2 = fmax(e1, fmax(e2, e3));
    return 0.999 * (4.0 * amin * amin) / l2;
}
__device__ __forceinline__ unsigned int bucket_take_slot(const BucketView &bk, int t) {
    if (!bk.big_count) atomicMax(bk.head + t, (unsigned long long)bk.epoch << 32);
    return (unsigned int)atomicAdd(bk.head + t, 1ull);
}

template <bool PREDICT, bool DBG, bool BATCH = false, bool PART = false>
__global__ void __launch_bounds__(kThreads, kMinBlk) k_cell_patch(SurfView s, PatchView pv, SurfView sub, GridView g, BucketView bk, BoundaryView b, Consts k_,
                                                       CellStepArgs a, const double *__restrict__ x_in, double *__restrict__ x_out,
                                                       double *__restrict__ V_out, double *__restrict__ A_out, DebugView dbg, BatchView bt) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ double shred[3 * 8];               // per-warp partial sums of (area, area x, area y)
    __shared__ int s_big[2 * kThreads];           // triangles that received their first anchor from this CTA, published with one global atomic
    __shared__ int s_nbig, s_bigbase;
    const Consts &k = BATCH ? bt.k[blockIdx.y] : k_;
    if (BATCH) {
        const size_t r = blockIdx.y, n3 = 3 * (size_t)s.np * r;
        batch_shift(s, r); batch_shift(sub, r); batch_shift(bk, bt, r);
        b.buf += bt.bnd * r; a.partials += bt.partials * r;
        a.d_flags += bt.flags * r; a.d_loc_bond += bt.flags * r; a.d_loc_cnt += bt.flags * r; a.d_loc_first += bt.flags * r; a.d_ap_fire += bt.flags * r; a.err += r;
        a.loc_cache += 2 * (size_t)s.np * r;
        if (bt.seed) a.seed = bt.seed[r];
        if (PREDICT) { x_in += n3; x_out += n3; }
        V_out += n3; A_out += n3;
    }
    const int NL = kPatch + pv.max_halo;
    unsigned long long *bar = (unsigned long long *)smraw;
    unsigned char *sblob = smraw + 16;
    double *sX0 = (double *)(sblob + pv.stride), *sX1 = sX0 + NL, *sx0 = sX1 + NL, *sx1 = sx0 + NL, *se0 = sx1 + NL, *se1 = se0 + NL, *se3 = se1 + NL;
    double *rec = se3 + NL;                       // [max_elem][kRec]
    const int p = PART ? a.part.patch_list[blockIdx.x] : blockIdx.x, tid = threadIdx.x;      // (the list is static: read before the wait)
    const int n0 = p * kPatch;
    const int nown = min(kPatch, s.nn - n0);
    const int ncopy = min(kPatch, s.np - n0);
    const int np = s.np;
    // ---- stage: TMA bulk copies of the patch descriptor and of the owned nodes' state
    if (tid == 0) {
        s_nbig = 0;
        mbar_init(bar, 1);
        const unsigned int nb = (unsigned int)ncopy * 8u;
        mbar_expect_tx(bar, (unsigned int)pv.stride + (PREDICT ? 5u : 7u) * nb);
        bulk_g2s(sblob, pv.blob + (size_t)p * pv.stride, (unsigned int)pv.stride, bar);      // static descriptor: no kernel writes it
    }
    // Programmatic dependent launch, the rule that keeps the chain sound: before its wait a kernel touches only data that no
    // kernel of the chain writes, and it releases its own dependents only AFTER its wait.  By induction, once all CTAs of this
    // kernel have been released, the previous step's substrate kernel (our prerequisite) and everything before it are complete
    // -- which is what the substrate kernel of THIS step relies on when it reads its own previous output before waiting.
    pdl_wait();
    pdl_launch_dependents();
    if (!BATCH && a.halo_peer && blockIdx.x == 0 && tid == kThreads - 1) halo_signal(a.halo_peer, a.halo_npeer);   // the previous substrate kernel is complete
    if (tid == 0) {
        const unsigned int nb = (unsigned int)ncopy * 8u;
        if (PREDICT) { bulk_g2s(sX0, x_in + n0, nb, bar); bulk_g2s(sX1, x_in + np + n0, nb, bar); }
        else { bulk_g2s(sX0, s.X + n0, nb, bar); bulk_g2s(sX1, s.X + np + n0, nb, bar); bulk_g2s(sx0, s.x + n0, nb, bar); bulk_g2s(sx1, s.x + np + n0, nb, bar); }
        bulk_g2s(se0, s.ea_in + n0, nb, bar); bulk_g2s(se1, s.ea_in + np + n0, nb, bar); bulk_g2s(se3, s.ea_in + 3 * (size_t)np + n0, nb, bar);
    }
    const bool own = tid < nown;
    const int i = n0 + tid;
    double V0 = 0, V1 = 0, Av0 = 0, Av1 = 0;      // own node's V, A (kept for the update)
    double fa0 = 0, fa1 = 0, cC = 0;
    int bi = -1;                                   // boundary ordinal of the own node
    // the triangle the anchor was in at the previous step and its (frozen) vertices, from the per-node cache: one coalesced load,
    // in flight behind the staging wait
    int tprev = -1;
    double lc_x = 0, lc_y = 0, lc_c2 = 0;
    if (own) {
        V0 = s.V[i]; V1 = s.V[np + i]; Av0 = s.A[i]; Av1 = s.A[np + i]; fa0 = s.fa[i]; fa1 = s.fa[np + i]; cC = s.cC[i];
        bi = b.index[i];
        const double2 c0 = a.loc_cache[i], c1 = a.loc_cache[(size_t)np + i];
        lc_x = c0.x; lc_y = c0.y; lc_c2 = c1.x;
        const long long tl = __double_as_longlong(c1.y);
        tprev = (tl >= 0 && tl < (long long)sub.ne) ? (int)tl : -1;
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int *hdr = (const int *)sblob;
    const int nh = hdr[0], nel = hdr[1];
    const int *halo_ids = (const int *)(sblob + pv.off_halo);
    const unsigned short *elem_lv = (const unsigned short *)(sblob + pv.off_lv);
    const unsigned short *rptr = (const unsigned short *)(sblob + pv.off_rptr);
    const unsigned short *ring16 = (const unsigned short *)(sblob + pv.off_ring);
    // partitioned cell: header word 2 of the descriptor carries the patch's role on THIS rank (bit 0: it reads rows other ranks push,
    // or pushes rows itself -- wait for every peer's step counter first; bit 1: it pushes; bit 2: this rank owns its partial sums)
    const int prole = PART ? hdr[2] : 0;
    if (PART && (prole & 1)) {
        if (tid == 0) cell_part_wait(a.part.flags, 0, a.part.npeer, (unsigned long long)a.part.step, a.part.err);
        __syncthreads();
    }
    // ---- halo gather first (its global loads fly while the own node's bond kinetics run) + the same predictor: every patch that
    //      sees a node computes the same x_new
    for (int l = tid; l < nh; l += kThreads) {
        const int gi = halo_ids[l];
        const int d = kPatch + l;
        if (PART && (prole & 1)) {
            // rows a peer GPU stored into L2 while this kernel may already have been running: read them past the (incoherent) L1
            const double xo0 = __ldcg(x_in + gi), xo1 = __ldcg(x_in + np + gi);
            const double v0 = __ldcg(s.V + gi), v1 = __ldcg(s.V + np + gi), a0 = __ldcg(s.A + gi), a1 = __ldcg(s.A + np + gi);
            const double vh0 = WCFE_ADD(v0, WCFE_MUL(a.dt / 2.0, a0)), vh1 = WCFE_ADD(v1, WCFE_MUL(a.dt / 2.0, a1));
            sX0[d] = xo0; sX1[d] = xo1; sx0[d] = WCFE_ADD(xo0, WCFE_MUL(a.dt, vh0)); sx1[d] = WCFE_ADD(xo1, WCFE_MUL(a.dt, vh1));
            se0[d] = __ldcg(s.ea_in + gi); se1[d] = __ldcg(s.ea_in + np + gi); se3[d] = __ldcg(s.ea_in + 3 * (size_t)np + gi);
            continue;
        }
        if (PREDICT) {
            const double xo0 = x_in[gi], xo1 = x_in[np + gi];
            const double v0 = s.V[gi], v1 = s.V[np + gi], a0 = s.A[gi], a1 = s.A[np + gi];
            const double vh0 = WCFE_ADD(v0, WCFE_MUL(a.dt / 2.0, a0)), vh1 = WCFE_ADD(v1, WCFE_MUL(a.dt / 2.0, a1));
            sX0[d] = xo0; sX1[d] = xo1; sx0[d] = WCFE_ADD(xo0, WCFE_MUL(a.dt, vh0)); sx1[d] = WCFE_ADD(xo1, WCFE_MUL(a.dt, vh1));
        } else {
            sX0[d] = s.X[gi]; sX1[d] = s.X[np + gi]; sx0[d] = s.x[gi]; sx1[d] = s.x[np + gi];
        }
        se0[d] = s.ea_in[gi]; se1[d] = s.ea_in[np + gi]; se3[d] = s.ea_in[3 * (size_t)np + gi];
    }
    // ---- predictor for the own node (UL:118, main.cpp:228; uncontracted so that x_new is bit-identical to the x86 reference)
    if (PREDICT && own) {
        const double vh0 = WCFE_ADD(V0, WCFE_MUL(a.dt / 2.0, Av0)), vh1 = WCFE_ADD(V1, WCFE_MUL(a.dt / 2.0, Av1));
        const double xn0 = WCFE_ADD(sX0[tid], WCFE_MUL(a.dt, vh0)), xn1 = WCFE_ADD(sX1[tid], WCFE_MUL(a.dt, vh1));
        sx0[tid] = xn0; sx1[tid] = xn1;
        x_out[i] = xn0; x_out[np + i] = xn1;
    }
    // ---- bond kinetics + anchor location for the own node (needs only its own X, x)
    double anc_px = 0.0, anc_py = 0.0, myx0 = 0.0, myx1 = 0.0;
    unsigned int slot_prev = 0xffffffffu;          // deferred: the slot taken in the previous step's triangle (result of an atomic in flight)
    int cnt = 0, first = -1;
    // a triangle received an anchor: record the node, stamp the triangle's three nodes and list the triangle when it was the first
    // (persistent membership, BucketView: the slot word's bit 31 says the triangle is listed already; item = node id | flags)
    auto finish_insert = [&](int t, unsigned int slot_raw, int v0, int v1, int v2, int item) {
        const unsigned int slot = bk.persist ? (slot_raw & kHeadCount) : slot_raw;
        if (slot < (unsigned int)bk.cap) bk.items[(size_t)t * bk.cap + slot] = item; else atomicOr(a.err, ERR_BUCKET_OVERFLOW);
        if (slot_raw == 0) {             // first anchor of the step in this triangle (persistent: of an empty, unlisted triangle)
            if (!bk.big_count) { bk.node_epoch[v0] = bk.epoch; bk.node_epoch[v1] = bk.epoch; bk.node_epoch[v2] = bk.epoch; }   // (else: stamped by the reduction)
            if (bk.big_count) {          // the reduction has a triangle to reduce
                if (bk.persist) atomicOr(bk.head + t, (unsigned long long)kHeadListed);
                const int q = atomicAdd(&s_nbig, 1);
                if (q < 2 * kThreads) s_big[q] = t;
                else { const unsigned int idx = atomicAdd(bk.big_count, 1u); if (idx < (unsigned int)bk.big_cap) bk.big_list[idx] = t; else atomicOr(a.err, ERR_BUCKET_OVERFLOW); }
            }
        }
    };
    if (own) {
        myx0 = sx0[tid]; myx1 = sx1[tid];
        const double Xi[3] = {sX0[tid], sX1[tid], 0.0}, xi[3] = {myx0, myx1, 0.0};
        double fa[3] = {fa0, fa1, 0.0};
        if (a.debug) {
            const double pt0[3] = {Xi[0] - fa[0], Xi[1] - fa[1], 0.0};                  // UL:2020
            int last = -1;
            locate(g, sub, pt0, [&](int t, const double *, double) { last = t; });
            a.d_loc_bond[i] = last;
        }
        const double cell_u[3] = {xi[0] - Xi[0], xi[1] - Xi[1], 0.0}, subs_u[3] = {0.0, 0.0, 0.0};
        a.d_flags[i] = bond_update(k, a.dt_real, cell_u, subs_u, fa, cC);
        fa0 = fa[0]; fa1 = fa[1];
        s.fa[i] = fa0; s.fa[np + i] = fa1; s.cC[i] = cC;
        const double pt[3] = {xi[0] - fa0, xi[1] - fa1, 0.0};                           // UL:1303
        anc_px = pt[0]; anc_py = pt[1];
        if (!(pt[0] >= a.ix0 && pt[0] <= a.ix1 && pt[1] >= a.iy0 && pt[1] <= a.iy1)) atomicOr(a.err, ERR_REGION);
        if (PART && !(fa0 * fa0 + fa1 * fa1 <= a.part.fa_max2)) atomicOr(a.err, ERR_REGION);      // the anchor may sit under a strip whose rank does not compute this patch
        // Anchors rarely change triangle between steps: try the triangle of the previous step first.  A point STRICTLY inside
        // it (all three sub-areas > 0) lies in no other triangle of a conforming substrate mesh, so the grid search -- which
        // finds every containing triangle, as the reference's exhaustive loop does (UL:1309-1452) -- is only needed otherwise.
        bool done = false;
        // (no circle -- the anchor sits within the margin of an edge --: nothing to remember, the next step locates again)
        auto remember = [&](int t, const double *Ar, double At, double n1x, double n1y, double n2x, double n2y, double n3x, double n3y) {
            const double c2 = anchor_clearance2(Ar, At, n1x, n1y, n2x, n2y, n3x, n3y);
            a.loc_cache[i] = make_double2(pt[0], pt[1]);
            a.loc_cache[(size_t)np + i] = make_double2(c2, __longlong_as_double(c2 > 0.0 ? (long long)t : -1ll));
            return c2 > 0.0;
        };
        if (tprev >= 0) {
            const double dx = pt[0] - lc_x, dy = pt[1] - lc_y;
            if (fma(dx, dx, dy * dy) < lc_c2) {              // still inside the circle that was inside the triangle, margin included
                if (!bk.persist) slot_prev = bucket_take_slot(bk, tprev);      // consumed after the element phase: the round trip hides behind it
                cnt = 1; first = tprev; done = true;                           // (persistent membership: nothing to do at all)
            } else {
                // out of the circle (or no circle: the anchor sits within the margin of an edge): the containment test itself.
                // margin: within 1e-9 of an edge (relative sub-area) the reference's rounded predicate may ALSO accept the neighbour
                // (A1 = A_tri - A2 - A3 is formed in floating point, UL:1338): such anchors take the grid search, which reports every
                // accepting triangle; the margin is seven orders above the rounding it guards against
                const int v0 = sub.tri[tprev], v1 = sub.tri[sub.nep + tprev], v2 = sub.tri[2 * sub.nep + tprev];
                const double n1x = sub.x[v0], n1y = sub.x[sub.np + v0], n2x = sub.x[v1], n2y = sub.x[sub.np + v1], n3x = sub.x[v2], n3y = sub.x[sub.np + v2];
                double Ar[3], At;      // planar path: z == 0 everywhere (checked at upload), the 2-D predicate is bit-identical to the 3-D one
                const bool in = point_in_triangle_2d(n1x, n1y, n2x, n2y, n3x, n3y, pt[0], pt[1], Ar, At);
                const double mrg = kAnchorMargin * At;
                if (in && Ar[0] > mrg && Ar[1] > mrg && Ar[2] > mrg && remember(tprev, Ar, At, n1x, n1y, n2x, n2y, n3x, n3y)) {
                    if (!bk.persist) slot_prev = bucket_take_slot(bk, tprev);
                    cnt = 1; first = tprev; done = true;
                } else if (bk.persist) {
                    // the anchor leaves the triangle it is a member of (or comes within the margin of its edges): its entry dies
                    const int n = min((int)((unsigned int)bk.head[tprev] & kHeadCount), bk.cap);
                    int *it = bk.items + (size_t)tprev * bk.cap;
                    for (int q = 0; q < n; q++) if (it[q] == i) { it[q] = kItemDead; break; }
                }
            }
        }
        if (!done) {
            bool stay = false;           // strictly inside one triangle, margin included: a persistent member
            locate(g, sub, pt, [&](int t, const double *Ar, double At) {
                const int v0 = sub.tri[t], v1 = sub.tri[sub.nep + t], v2 = sub.tri[2 * sub.nep + t];
                if (cnt == 0) {          // the anchor changed triangle (or had none): remember the new one
                    first = t;
                    stay = remember(t, Ar, At, sub.x[v0], sub.x[sub.np + v0], sub.x[v1], sub.x[sub.np + v1], sub.x[v2], sub.x[sub.np + v2]);
                }
                cnt++;
                finish_insert(t, bucket_take_slot(bk, t), v0, v1, v2, (bk.persist && (!stay || cnt > 1)) ? (i | kItemVolatile) : i);
            });
            if (cnt == 0 && tprev >= 0) a.loc_cache[(size_t)np + i] = make_double2(0.0, __longlong_as_double(-1ll));
        }
        a.d_loc_cnt[i] = cnt; a.d_loc_first[i] = first;
    }
    __syncthreads();                              // node table complete (halo rows, predicted own rows)
    // ---- elements: closed-form planar element from shared memory, one record per patch-local element; the thread that
    //      evaluates an element whose local vertex 0 this patch owns also counts it for calculateAreaTot (UL:1655-1661)
    const double tk = t_act_of(k, a.t_tot) * k.kappa;
    const Recip r3 = recip_of(3.0);
    double red0 = 0.0, red1 = 0.0, red2 = 0.0;
    for (int el = tid; el < nel; el += kThreads) {
        const unsigned short *lv = elem_lv + 3 * el;
        const int n1 = lv[0], n2 = lv[1], n3 = lv[2];
        const double X[3][2] = {{sX0[n1], sX1[n1]}, {sX0[n2], sX1[n2]}, {sX0[n3], sX1[n3]}};
        const double x[3][2] = {{sx0[n1], sx1[n1]}, {sx0[n2], sx1[n2]}, {sx0[n3], sx1[n3]}};
        const double e3[3] = {1.0 / 3.0 * se0[n1] + 1.0 / 3.0 * se0[n2] + 1.0 / 3.0 * se0[n3], 1.0 / 3.0 * se1[n1] + 1.0 / 3.0 * se1[n2] + 1.0 / 3.0 * se1[n3],
                              1.0 / 3.0 * se3[n1] + 1.0 / 3.0 * se3[n2] + 1.0 / 3.0 * se3[n3]};
        ElemOut2 o;
        elem_eval_planar(X, x, e3, k.mu_cell, tk, k.thick, o);
        double *R = rec + kRec * el;
        // what every incident node needs of A0 and a_cur is formed here once per element instead of once per (node, element)
        R[0] = o.r1[0]; R[1] = o.r1[1]; R[2] = o.r2[0]; R[3] = o.r2[1]; R[4] = o.eas[0]; R[5] = o.eas[1]; R[6] = o.eas[2];
        R[7] = __dmul_rn(__dmul_rn(1. / 3., o.A0), k.thick); R[8] = quot(o.a_cur, r3);
        if (n1 < kPatch) {
            red0 += o.a_cur;
            red1 += (1. / 3.) * ((x[0][0] + x[1][0]) + x[2][0]) * o.a_cur;
            red2 += (1. / 3.) * ((x[0][1] + x[1][1]) + x[2][1]) * o.a_cur;
        }
    }
    {   // fixed-order partial sums: lanes (shuffle tree), then warps (thread 0, after the barrier below)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { red0 += __shfl_down_sync(0xffffffffu, red0, off); red1 += __shfl_down_sync(0xffffffffu, red1, off); red2 += __shfl_down_sync(0xffffffffu, red2, off); }
        if ((tid & 31) == 0) { shred[tid >> 5] = red0; shred[8 + (tid >> 5)] = red1; shred[16 + (tid >> 5)] = red2; }
    }
    if (own && slot_prev != 0xffffffffu) {      // the deferred insert
        int tv0 = 0, tv1 = 0, tv2 = 0;
        if (!bk.big_count) { tv0 = sub.tri[tprev]; tv1 = sub.tri[sub.nep + tprev]; tv2 = sub.tri[2 * sub.nep + tprev]; }   // (the stamps of the protocol without reduction)
        finish_insert(tprev, slot_prev, tv0, tv1, tv2, i);
    }
    __syncthreads();                              // records complete; every insert of this CTA listed
    if (tid == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < kThreads / 32; w++) { t0 += shred[w]; t1 += shred[8 + w]; t2 += shred[16 + w]; }
        double *o = a.partials + (size_t)4 * p;
        o[0] = t0; o[1] = t1; o[2] = t2; o[3] = 0.0;
        if (PART && (prole & 4))             // this rank owns the patch: its sums go into everybody's array (a patch on a cut is computed
            for (int q = 0; q < a.part.npeer; q++) {      // twice, bit-identically, and published once)
                double *po = a.part.peer[q].partials[a.part.step & 1] + (size_t)4 * p;
                po[0] = t0; po[1] = t1; po[2] = t2; po[3] = 0.0;
            }
        const int nbig = min(s_nbig, 2 * kThreads);
        if (bk.big_count && nbig) s_bigbase = (int)atomicAdd(bk.big_count, (unsigned int)nbig);
    }
    // ---- nodes
    if (own) {
        const double ea0 = se0[tid], ea1 = se1[tid], ea3 = se3[tid];
        double rx = 0.0, ry = 0.0, q0 = 0.0, q1 = 0.0, q3 = 0.0, mm = 0.0, asum = 0.0;
        const int rb = rptr[tid], re = rptr[tid + 1];
        // the one-ring gather is a chain of dependent shared-memory loads (code -> record): two entries per trip, both codes and
        // both records' loads issued before the first use; a missing second entry adds +0 (the ascending-element summation order
        // of UL:692-705 and the result are unchanged)
        for (int r = rb; r < re; r += 2) {
            const bool two = r + 1 < re;
            const int cA = ring16[r], cB = two ? (int)ring16[r + 1] : 0;
            const double *RA = rec + kRec * (cA >> 2), *RB = rec + kRec * (cB >> 2);
            const int sA = cA & 3, sB = cB & 3;
            const double a0 = RA[0], a1 = RA[1], a2 = RA[2], a3 = RA[3], a4 = RA[4], a5 = RA[5], a6 = RA[6], a7 = RA[7], a8 = RA[8];
            const double b0 = RB[0], b1 = RB[1], b2 = RB[2], b3 = RB[3], b4 = RB[4], b5 = RB[5], b6 = RB[6], b7 = two ? RB[7] : 0.0, b8 = two ? RB[8] : 0.0;
            // (explicit fma: the contraction must not depend on the template instantiation, batched and single runs compare bitwise)
            rx += sA == 0 ? -(a0 + a2) : (sA == 1 ? a0 : a2);
            ry += sA == 0 ? -(a1 + a3) : (sA == 1 ? a1 : a3);
            q0 = fma(a4 - ea0, a7, q0);                                // UL:911
            q1 = fma(a5 - ea1, a7, q1);
            q3 = fma(a6 - ea3, a7, q3);
            mm += a7;                                                  // UL:913
            asum += a8;                                                // UL:1251
            const double fxB = sB == 0 ? -(b0 + b2) : (sB == 1 ? b0 : b2), fyB = sB == 0 ? -(b1 + b3) : (sB == 1 ? b1 : b3);
            rx += two ? fxB : 0.0; ry += two ? fyB : 0.0;
            q0 = fma(b4 - ea0, b7, q0);                                // b7 == 0 without a second entry: +0
            q1 = fma(b5 - ea1, b7, q1);
            q3 = fma(b6 - ea3, b7, q3);
            mm += b7;
            asum += b8;
        }
        double Vh0 = V0 + (a.dt_drv / 2.0) * Av0, Vh1 = V1 + (a.dt_drv / 2.0) * Av1;         // UL:118
        const double u[3] = {fa0, fa1, 0.0};
        double fi[3];
        integrin_force(k, u, fi);
        const double Ffx = -cC * fi[0] * asum * k.rho_i_max, Ffy = -cC * fi[1] * asum * k.rho_i_max;   // UL:1259
        const double Fvx = -k.drag * Vh0, Fvy = -k.drag * Vh1;                       // UL:571
        s.asum[i] = asum;
        reinterpret_cast<double2 *>(s.anc)[2 * (size_t)i] = make_double2(anc_px, anc_py);
        reinterpret_cast<double2 *>(s.anc)[2 * (size_t)i + 1] = make_double2(-Ffx, -Ffy);
        const Recip rmm = recip_of(mm), rmr = recip_of(mm * k.rho);      // shared by the three strain and the two acceleration quotients
        s.ea_out[i] = ea0 + quot(q0, rmm); s.ea_out[np + i] = ea1 + quot(q1, rmm); s.ea_out[3 * np + i] = ea3 + quot(q3, rmm);   // UL:607-612
        if (DBG) {
            dbg.res[i] = rx; dbg.res[np + i] = ry; dbg.f_fa[i] = Ffx; dbg.f_fa[np + i] = Ffy; dbg.mm[i] = mm;
            dbg.res_ea[i] = q0; dbg.res_ea[np + i] = q1; dbg.res_ea[3 * np + i] = q3;
            dbg.f_ap[i] = 0.0; dbg.f_ap[np + i] = 0.0; dbg.f_bend[i] = 0.0; dbg.f_bend[np + i] = 0.0; dbg.f_area[i] = 0.0; dbg.f_area[np + i] = 0.0;
        }
        if (bi < 0) {
            node_update_q(Fvx + Ffx - rx, mm, rmr, k.damp, k.rho, a.dt, a.dt_real, Vh0, V0, Av0);    // UL:592-628
            node_update_q(Fvy + Ffy - ry, mm, rmr, k.damp, k.rho, a.dt, a.dt_real, Vh1, V1, Av1);
            // V, A go to the OTHER buffer: neighbouring patches still read this node's old V, A for their halo predictor
            V_out[i] = V0; V_out[np + i] = V1; A_out[i] = Av0; A_out[np + i] = Av1;
        } else {
            const double xi[3] = {myx0, myx1, 0.0};
            double j1[3] = {0, 0, 0}, j2[3] = {0, 0, 0};
            for (int q = b.cand_ptr[bi]; q < b.cand_ptr[bi + 1]; q++) {                                // UL:1479-1507
                const int lj = pv.cand_l[2 * q], lint = pv.cand_l[2 * q + 1];
                const double xb[3] = {sx0[lj], sx1[lj], 0.0}, xin[3] = {sx0[lint], sx1[lint], 0.0};
                if (bnd_is_first(xi, xb, xin)) { j1[0] = xb[0]; j1[1] = xb[1]; } else { j2[0] = xb[0]; j2[1] = xb[1]; }
            }
            double nrm[3], tl;
            bnd_normal(xi, j1, j2, nrm, tl);
            const double kap = bnd_curvature(xi, j1, j2, tl);
            double Pr;
            if (a.rng_mode == 0) Pr = (double)a.tape[bi] / 2147483648.0;                               // UL:1525
            else Pr = (double)philox_u32(a.seed, a.step, (unsigned int)bi) * (1.0 / 4294967296.0);
            const bool fire = Pr < a.ap_threshold;
            const double factin = fire ? k.ap_quantum * ap_trips(k.ap_density, tl) : 0.0;
            a.d_ap_fire[i] = fire ? 1 : 0;
            double *o = b.buf + (size_t)16 * bi;
            const double Fbx = k.k_bend * kap * nrm[0], Fby = k.k_bend * kap * nrm[1], Fax = factin * nrm[0], Fay = factin * nrm[1];
            o[0] = ((Fvx + Fbx) + Fax) + Ffx; o[1] = ((Fvy + Fby) + Fay) + Ffy; o[2] = 0.0;
            o[3] = rx; o[4] = ry; o[5] = 0.0; o[6] = mm; o[7] = tl; o[8] = nrm[0]; o[9] = nrm[1]; o[10] = 0.0;
            if (DBG) { dbg.f_bend[i] = Fbx; dbg.f_bend[np + i] = Fby; dbg.f_ap[i] = Fax; dbg.f_ap[np + i] = Fay; }
        }
    }
    if (PART && PREDICT && (prole & 2) && own) {
        // rows other ranks read as halo rows of THEIR patches: the new state goes to the same index of their write buffers
        // (positions and strains from here for every such node, V and A only for interior nodes: a boundary vertex gets them
        // from the boundary finish)
        const int par = (int)(a.part.step & 1u);
        for (int q = a.part.exp_ptr[p]; q < a.part.exp_ptr[p + 1]; q++) {
            if (a.part.exp_ent[2 * q] != tid) continue;
            const CellPeer *pr = a.part.peer + a.part.exp_ent[2 * q + 1];
            pr->x[par][i] = myx0; pr->x[par][np + i] = myx1;
            pr->ea[par][i] = s.ea_out[i]; pr->ea[par][np + i] = s.ea_out[np + i]; pr->ea[par][3 * (size_t)np + i] = s.ea_out[3 * (size_t)np + i];
            if (bi < 0) { pr->V[par][i] = V0; pr->V[par][np + i] = V1; pr->A[par][i] = Av0; pr->A[par][np + i] = Av1; }
        }
    }
    if (bk.big_count) {                       // uniform: publish this CTA's share of the triangle list
        __syncthreads();                      // s_bigbase
        const int nbig = min(s_nbig, 2 * kThreads);
        for (int q = tid; q < nbig; q += kThreads) {
            const int idx = s_bigbase + q;
            if (idx < bk.big_cap) bk.big_list[idx] = s_big[q]; else atomicOr(a.err, ERR_BUCKET_OVERFLOW);
        }
    }
    if (bk.cell_done) {
        // every thread's global writes (state, buckets, stamps, anchor records) precede this barrier: one release by thread 0
        // publishes them all to whoever acquires the counter (k_subs_ring)
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(bk.cell_done, 1ull); }
    }
}

}   // namespace wcfe
