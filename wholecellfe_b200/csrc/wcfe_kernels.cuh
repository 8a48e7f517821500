// wcfe_kernels.cuh -- CUDA kernels of the UL step (sm_100a), general 3-D path ("v1": node-centric gather).
//
// Data layout in HBM: struct-of-arrays FP64.  A field of width w lives in one allocation of w*np doubles,
// component c at [c*np + node] (np = node count padded to 32, so every component starts 256-B aligned and
// a warp's loads of one component are one or two 128-B lines).  Connectivity is int32 SoA tri[c*nep + e].
//
// Deterministic assembly without atomics: every nodal sum (Res, Res_ea, MM, node area) is a gather over the
// vertex->triangle CSR in ascending element order -- the reference's own summation order at a node
// (serial scatter-add in element order, UL_growth_explicit.cpp:692-705) -- evaluated by the thread that owns
// the node.  Elements are recomputed by each of their three nodes (the per-element closed form is ~0.4 kFLOP,
// wcfe_math.cuh); nothing is staged through HBM.
#pragma once
#include "wcfe_math.cuh"

namespace wcfe {

constexpr int kBlock = 128;
constexpr int kMaxLoc = 8;   // containing triangles recorded per anchor (a point on a mesh vertex has ~6)

struct SurfView {            // device pointers of one surface
    int nn, np, ne, nep;
    const int *tri;          // [3*nep]
    const int *ring_ptr;     // [nn+1]
    const int *ring;         // (element<<2)|slot, ascending element
    double *x, *X;           // [3*np] current / start-of-step positions (substrate: X == x, frozen)
    double *V, *A;           // [3*np]
    const double *ea_in;     // [6*np]
    double *ea_out;          // [6*np]
    double *cC, *fa, *asum;  // cell only: [np], [3*np], [np]
    double *anc;             // cell only, planar path: [4*np] anchor records (p.x, p.y, H.x, H.y) written by k_cell_patch:
                             //   p = x - fa_u (UL:1303), H = c_C f(fa_u) Asum rho = minus the cell node's own F_FA (UL:1259, :1414)
};

struct GridView {            // uniform grid over substrate triangles (candidate filter for point location)
    int nx, ny;
    double x0, y0, inv_h;
    const int *ptr, *items;  // CSR cell -> triangles, ascending triangle index inside a cell
};

struct CellPeer;
struct BucketView {          // triangle -> anchored cell nodes, rebuilt every step, no reset pass:
    unsigned long long *head;   // [ne_subs] (epoch<<32 | count); a stale epoch reads as count 0
    int *items;                 // [ne_subs*cap]
    int cap;
    unsigned int epoch;
    // planar path: every triangle that receives its first anchor of the step is listed here by the inserting thread;
    // k_bucket_reduce turns each listed bucket into the six nodal sums of calculateForce_FA_subs (nullptr: no list, the
    // substrate kernel walks the buckets itself -- k_subs_patch, k_subs_node)
    unsigned int *big_count;        // entries of big_list; zeroed on the stream before every cell kernel
    int *big_list;                  // [big_cap] triangle ids
    int big_cap;
    // planar path: epoch stamp per SUBSTRATE NODE, written by the cell kernel on the three vertices of every triangle it drops an
    // anchor into.  A substrate node under the cell whose stamp is stale has nothing to gather and skips the bucket walk (two
    // thirds of the nodes inside the padded bounding box of a round cell).
    unsigned int *node_epoch;       // [np_subs], or nullptr
    // completion counter of the planar cell kernel: every CTA of k_cell_patch adds one when all of its global writes are out (it
    // never resets; the host keeps the running target).  The substrate kernel's under-the-cell threads wait for THIS instead of
    // griddepcontrol.wait, whose prerequisite is the boundary-finish kernel they have no use for (it sits between the two in the
    // stream): 20-30 us that only show once a strip's substrate kernel is as short as the finish kernel (partitioned runs).
    unsigned long long *cell_done;  // or nullptr
    unsigned long long cell_target;
    int wait_cell;                  // substrate side: 1 = spin on the counter, 0 = griddepcontrol.wait (k_bucket_reduce sits in between)
    int *chain_err;
    const CellPeer *cell_peer;      // partitioned cell: the substrate kernel's first thread bumps every peer's step counter
    int cell_npeer;
    // PERSISTENT membership (production chain with the reduction, persist != 0): an anchor changes triangle about once in a thousand
    // steps, so buckets and the triangle list are kept from step to step and the cell kernel touches them only on a change --
    //   head word, low half: bit 31 = the triangle is listed, bits 0..30 = entries in use (live, tombstones, this step's appends);
    //   items: node id; kItemDead = a member that left (its own thread overwrites its entry); | kItemVolatile = an anchor within
    //   the margin of an edge, located and inserted anew every step (possibly into several triangles), dropped by the reduction;
    //   the reduction compacts a bucket when it finds either and puts the six sums into sums[6 * triangle] instead of the head of the
    //   item storage, which now has to survive.  A triangle stays listed once it has held an anchor (an empty one costs the reduction
    //   one load): the list grows with the area the anchors have visited since the last reset, never with the number of crossings.
    // Valid together with CellStepArgs::loc_cache ("the node is a member of the triangle its cache row names"): the host resets both
    // whenever one of them is (path change, substrate upload, mesh upload, partition connect).
    double *sums;                   // [6 * ne_subs], or nullptr
    int persist;
};
constexpr int kItemDead = -1, kItemVolatile = 0x40000000, kItemId = 0x3fffffff;
constexpr unsigned int kHeadListed = 0x80000000u, kHeadCount = 0x7fffffffu;

// ---- partitioned CELL (BASELINE configs[3]): every rank keeps the whole cell's arrays, with the single-GPU numbering, but computes
// only the patches that can put an anchor into its substrate strip (a patch on a cut is computed by both neighbours).  What a rank
// does not compute goes stale there -- except the nodes its own patches read as halo rows: those are pushed, every step, by the
// rank that owns their patch, through peer memory, to the same index of the reader's arrays (replicated storage: no row
// translation).  Two epoch counters per ordered pair of ranks pace the exchange, both bumped by one per step:
//   step counter     by the substrate kernel's first thread after its griddepcontrol.wait (cell kernel, boundary finish and the
//                    reduction of this step are complete: every pushed row is out, every halo row has been read).  A patch that
//                    reads pushed rows OR pushes rows waits until every peer's counter has reached the step it is about to run:
//                    the rows of the previous step have landed, and the buffers it is about to push into are no longer read.
//   partials counter by the boundary-finish kernel's first thread after its wait (this rank's cell kernel is complete: the area /
//                    centroid partial sums of the patches it OWNS have been stored into every peer's partials array).  Every
//                    rank then adds ALL patches' partials in the single-GPU order (UL:1641-1667): the same total, bit for bit.
struct CellPeer {
    double *x[2], *V[2], *A[2], *ea[2];   // the peer's write buffers of the first / second step after the connect (they swap every step)
    double *partials[2];                  // its partials arrays, by step parity
    unsigned long long *flag_step, *flag_part;   // this rank's counters in the peer's memory
};
struct CellPartView {
    const int *patch_list;               // [gridDim.x] patches this rank computes, ascending
    const int *exp_ptr, *exp_ent;        // CSR over ALL patches: pairs (local node, peer slot) of the nodes a patch pushes
    const CellPeer *peer; int npeer;
    const unsigned long long *flags;     // mine: [0 .. 7] step counters, [8 .. 15] partials counters, one per peer slot
    unsigned int step;                   // steps since the connect
    const unsigned char *bnd_local;      // [nb] the boundary vertex sits in a patch this rank computes
    const int *bexp_ptr, *bexp_peer;     // CSR over boundary ordinals: peer slots its V, A go to (written by the boundary finish)
    int *err;
    double fa_max2;                      // (anchor margin)^2 the patch lists were built for: |fa_u| beyond it is an error
};
enum : int { ERR_CELL_HALO_TIMEOUT = 32 };

struct BoundaryView {        // cell boundary vertices in ascending vertex order (the RNG draw order, UL:1463-1465)
    int nb;
    const int *node;         // [nb]
    const int *index;        // [nn] ordinal or -1
    const int *cand_ptr;     // [nb+1]
    const int *cand;         // pairs (nj, nint) in the reference's one-ring scan order (UL:1479-1507)
    double *buf;             // [16*nb] per boundary vertex: f_pre[3], Res[3], MM, tl, normal[2], ...
};

struct DebugView {           // optional force-field capture (parity gate on nodal forces)
    double *res, *f_fa, *f_ap, *f_bend, *f_area, *mm, *res_ea;   // cell or substrate set
};

// Ensembles: R independent replicas of the same two meshes advance in lockstep in ONE launch per kernel
// (gridDim.y = R).  Connectivity, patch descriptors and boundary tables are shared; every state array holds R consecutive
// images, and a CTA shifts its views to replica blockIdx.y before it starts.  Per-replica constants (spring law, ligand
// density, ...) come from BatchView::k instead of the by-value kernel parameter.
// Programmatic dependent launch (griddepcontrol, sm_90+): the step's three kernels are chained cell -> finish -> substrate.  A
// kernel lets its dependent start as soon as all of its own CTAs have been dispatched (pdl_launch_dependents, first instruction),
// and the dependent blocks (pdl_wait) only where it first touches what the predecessor produced.  Both are no-ops in a launch
// without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct BatchView {
    const Consts *k;             // [R]
    const unsigned int *seed;    // [R] Philox keys, or nullptr: every replica draws the same numbers (common random numbers)
    long long flags, partials, bnd, head, items, snodes;   // per-replica element counts: int flag arrays, area partials, boundary buffer, bucket heads / items, substrate node stamps
};
__device__ __forceinline__ void batch_shift(SurfView &s, size_t r) {
    const size_t n = (size_t)s.np * r;
    s.x += 3 * n; s.X += 3 * n; s.V += 3 * n; s.A += 3 * n; s.ea_in += 6 * n; s.ea_out += 6 * n;
    if (s.cC) { s.cC += n; s.fa += 3 * n; s.asum += n; }
    if (s.anc) s.anc += 4 * n;
}
__device__ __forceinline__ void batch_shift(BucketView &bk, const BatchView &bt, size_t r) {
    bk.head += bt.head * r; bk.items += bt.items * r;
    if (bk.big_count) { bk.big_count += r; bk.big_list += (size_t)bk.big_cap * r; }
    if (bk.node_epoch) bk.node_epoch += bt.snodes * r;
    if (bk.sums) bk.sums += 6 * bt.head * r;
}

enum : int { ERR_BUCKET_OVERFLOW = 1, ERR_LOC_OVERFLOW = 2, ERR_NAN = 4, ERR_REGION = 8, ERR_CHAIN = 32 };   // 16: ERR_HALO_TIMEOUT (wcfe_patch.cuh)

// ---- predictor: driverVelocities_Explicit + inline predictor (UL:103-124, main.cpp:224-242) ---------------
// (aos != nullptr: the predicted positions also as the array of Vec3d the host integrator receives -- the round trip's hand-over
// without a transpose pass of its own)
__global__ void __launch_bounds__(kBlock) k_predict_cell(int nn, int np, double *__restrict__ x, double *__restrict__ X,
                                                         const double *__restrict__ V, const double *__restrict__ A, double dt, double *__restrict__ aos = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // uncontracted (two roundings per multiply-add) so that x_new, and with it cell_u = x - X in updateFA, is
        // bit-identical to the x86 reference given the same V, A
        const double xo = x[c * np + i];
        double vh = WCFE_ADD(V[c * np + i], WCFE_MUL(dt / 2.0, A[c * np + i]));
        if (c == 2) vh = 0.0;
        X[c * np + i] = xo;
        const double xn = WCFE_ADD(xo, WCFE_MUL(dt, vh));
        x[c * np + i] = xn;
        if (aos) aos[3 * (size_t)i + c] = xn;
    }
}

// ---- host images: the reference's arrays are AoS (std::vector<Vec3d>, 6-vector strains); the device state is SoA --------------
__global__ void __launch_bounds__(kBlock) k_aos_to_soa(int nn, int np, int w, const double *__restrict__ aos, double *__restrict__ soa) {
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= (long long)np * w) return;
    const int i = (int)(idx / w), q = (int)(idx - (long long)i * w);
    soa[(size_t)q * np + i] = i < nn ? aos[idx] : 0.0;
}
__global__ void __launch_bounds__(kBlock) k_soa_to_aos(int nn, int np, int w, const double *__restrict__ soa, double *__restrict__ aos) {
    const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= (long long)nn * w) return;
    const int i = (int)(idx / w), q = (int)(idx - (long long)i * w);
    aos[idx] = soa[(size_t)q * np + i];
}

// ---- stand-alone drivers for the adapter: host-owned X, V, A in, V_hlf out.  kind 0 driverVelocities_Explicit UL:103-124,
//      1 _Biaxial UL:275-345 (the two main.cpp calls), 2 _Uniaxial UL:125-168, 3 _OffBiaxial UL:214-273, 4 _BulgeTest UL:170-212
//      (defined by the reference, commented out at main.cpp:217-220, :291-295).  Thresholds are the reference's literals; the
//      moving thresholds are formed uncontracted like the reference's x86 code.
__global__ void __launch_bounds__(kBlock) k_driver(int kind, int nn, const double *__restrict__ X, const double *__restrict__ V,
                                                   const double *__restrict__ A, double *__restrict__ Vh, Consts k, double dt, double t_tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = WCFE_ADD(V[3 * i + c], WCFE_MUL(dt / 2.0, A[3 * i + c]));
    const double X0 = X[3 * i], X1 = X[3 * i + 1], X2 = X[3 * i + 2];
    if (kind != 4) v[2] = 0.0;
    if (kind == 1 && (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f)) {
        if (X0 < -k.pin_limit || X0 > k.pin_limit) v[0] = 0.0;
        if (X1 > k.pin_limit || X1 < -k.pin_limit) v[1] = 0.0;
    } else if (kind == 2) {
        if (X0 < -0.0499) { v[0] = 0.0; v[1] = 0.0; }
        if (X0 > WCFE_ADD(0.9499, WCFE_MUL(0.1, t_tot)) && t_tot <= 2.0) { v[0] = 0.1; v[1] = 0.0; }
        if (X0 > 0.9499 + 2.00 * 0.1 && t_tot >= 2.0) { v[0] = 0.0; v[1] = 0.0; }
    } else if (kind == 3) {
        if (X0 < -0.04) { v[0] = 0.0; v[1] = 0.0; }
        if (X0 > WCFE_ADD(0.94, WCFE_MUL(0.1, t_tot)) && t_tot < 1.0) { v[0] = 0.1; v[1] = 0.0; }
        if (X0 > 0.94 + 0.99 * 0.1 && t_tot >= 1.0) { v[0] = 0.0; v[1] = 0.0; }
        if (X1 > 0.99 || X1 < 0.01) v[1] = 0.0;
    } else if (kind == 4) {
        const bool mid = X2 < 0.001 && X2 > -0.001;
        if (mid && (X0 < 0.001 || X0 > 0.999 || X1 < 0.001 || X1 > 0.999)) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) Vh[3 * i + c] = v[c];
}

// the substrate's nodal update after a driver other than _Biaxial: V = V_hlf + (dt_real - dt/2) A (UL:499) with the V_hlf that driver
// returned (AoS, kept on the device) and the A the substrate kernel has just written; launched only when dt_real > dt/2 (UL:498)
__global__ void __launch_bounds__(kBlock) k_subs_vhalf(int nn, int np, double *__restrict__ V, const double *__restrict__ A,
                                                       const double *__restrict__ Vh, double dt, double dt_real) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double ddt = dt_real - dt / 2.0;
#pragma unroll
    for (int c = 0; c < 3; c++) V[c * np + i] = fma(ddt, A[c * np + i], Vh[3 * i + c]);
}

// ---- dormant physics of the cell: fields the reference defines and whose calls / consumers are commented out.
//      One thread per node, one-ring gather in ascending element order (the order the reference's serial scatter-add produces),
//      on the CURRENT device state (x, X = start-of-step positions, c_A, c_C).  Diagnostic kernels, not part of the step.
//      which 0 / 1: diffusion residuals Res_Adif / Res_Cdif (evalElementRe UL:833-842, :915-925; assembled calculateRes UL:702-704;
//                   consumers commented out UL:617-618), out[np]
//      which 2:     calculatePressureForce UL:1765-1816 (call commented out UL:577): -pres (MM / 0.1) n, MM the lumped mass of
//                   calculateRes (one third of the one-ring reference areas times thick), out[3*np] SoA
//      which 3:     calculateConstraintForce UL:1178-1194 (call commented out UL:551), out[3*np] SoA
__global__ void __launch_bounds__(kBlock) k_dormant(SurfView s, const double *__restrict__ cA, const double *__restrict__ cC, int which,
                                                    double dt, double pres, double thick, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nn) return;
    const int np = s.np;
    if (which == 3) {
        const double x0 = s.x[i], x1 = s.x[np + i], x2 = s.x[2 * np + i];
        const double dist = sqrt(x0 * x0 + x1 * x1 + x2 * x2), f = 4000.0 * (350.0 - dist);
        out[i] = f * (x0 / dist); out[np + i] = f * (x1 / dist); out[2 * np + i] = f * (x2 / dist);
        return;
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, asum = 0.0, mm = 0.0;
    const double *conc = which == 0 ? cA : cC;
    const double D = which == 0 ? 1.0 : 1e-10;                                   // UL:919-920
    for (int r = s.ring_ptr[i]; r < s.ring_ptr[i + 1]; r++) {
        const int e = s.ring[r] >> 2, slot = s.ring[r] & 3;
        const int n0 = s.tri[e], n1 = s.tri[s.nep + e], n2 = s.tri[2 * s.nep + e];
        double G1[3], G2[3], g1[3], g2[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            G1[c] = s.X[c * np + n1] - s.X[c * np + n0]; G2[c] = s.X[c * np + n2] - s.X[c * np + n0];
            g1[c] = s.x[c * np + n1] - s.x[c * np + n0]; g2[c] = s.x[c * np + n2] - s.x[c * np + n0];
        }
        const double Gx[3] = {G1[1] * G2[2] - G1[2] * G2[1], G1[2] * G2[0] - G1[0] * G2[2], G1[0] * G2[1] - G1[1] * G2[0]};
        const double gx[3] = {g1[1] * g2[2] - g1[2] * g2[1], g1[2] * g2[0] - g1[0] * g2[2], g1[0] * g2[1] - g1[1] * g2[0]};
        const double nG = sqrt(Gx[0] * Gx[0] + Gx[1] * Gx[1] + Gx[2] * Gx[2]), ng = sqrt(gx[0] * gx[0] + gx[1] * gx[1] + gx[2] * gx[2]);
        if (which == 2) {
            acc0 += 0.5 * gx[0]; acc1 += 0.5 * gx[1]; acc2 += 0.5 * gx[2];       // area_ti * unit normal_ti (UL:1799-1803)
            asum += 0.5 * ng;
            mm += (1. / 3.) * nG / 2.0 * thick;                                   // UL:913
            continue;
        }
        // rows 0 and 1 of [g1 g2 n]^-1 (UL:834-836): the reciprocal basis g2 x n / det, n x g1 / det
        const double n[3] = {gx[0] / ng, gx[1] / ng, gx[2] / ng};
        const double c0[3] = {g2[1] * n[2] - g2[2] * n[1], g2[2] * n[0] - g2[0] * n[2], g2[0] * n[1] - g2[1] * n[0]};
        const double c1[3] = {n[1] * g1[2] - n[2] * g1[1], n[2] * g1[0] - n[0] * g1[2], n[0] * g1[1] - n[1] * g1[0]};
        const double det = g1[0] * c0[0] + g1[1] * c0[1] + g1[2] * c0[2];
        const double a0 = conc[n0], a1 = conc[n1], a2 = conc[n2];
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double r0 = c0[c] / det, r1 = c1[c] / det;
            const double grad = a0 * r0 + a1 * r1 + a2 * (-r0 - r1);              // UL:839-840
            const double gN = slot == 0 ? r0 : (slot == 1 ? r1 : -r0 - r1);         // UL:917
            dot += grad * gN;
        }
        const double mean = (1. / 3.) * (a0 + a1 + a2);                           // UL:841-842
        acc0 += (mean * (1. / 3.) - dt * D * dot) * nG / 2.0 * thick;             // UL:924-925
    }
    if (which == 2) {
        // the reference normalises the area-weighted sum twice (UL:1805-1806): the direction is that of the sum
        const double l = sqrt(acc0 * acc0 + acc1 * acc1 + acc2 * acc2), f = -pres * (mm / 0.1);
        (void)asum;
        out[i] = f * (acc0 / l); out[np + i] = f * (acc1 / l); out[2 * np + i] = f * (acc2 / l);
    } else out[i] = acc0;
}

// ---- point location of one anchor in the substrate -------------------------------------------------------
// visits candidates in ascending triangle index; calls f(tri, A[3], A_tri) for every containing triangle
template <class F>
__device__ __forceinline__ void locate_at(const GridView &g, const SurfView &s, const double *__restrict__ pos, const double pt[3], F &&f) {
    const double fx = (pt[0] - g.x0) * g.inv_h, fy = (pt[1] - g.y0) * g.inv_h;
    if (!(fx >= 0.0) || !(fy >= 0.0) || fx >= (double)g.nx || fy >= (double)g.ny) return;
    const int cidx = (int)fy * g.nx + (int)fx;
    const int b = g.ptr[cidx], e = g.ptr[cidx + 1];
    for (int k = b; k < e; k++) {
        const int t = g.items[k];
        const int i0 = s.tri[t], i1 = s.tri[s.nep + t], i2 = s.tri[2 * s.nep + t];
        const double n1[3] = {pos[i0], pos[s.np + i0], pos[2 * s.np + i0]};
        const double n2[3] = {pos[i1], pos[s.np + i1], pos[2 * s.np + i1]};
        const double n3[3] = {pos[i2], pos[s.np + i2], pos[2 * s.np + i2]};
        double Ar[3], At;
        if (point_in_triangle(n1, n2, n3, pt, Ar, At)) f(t, Ar, At);
    }
}
template <class F>
__device__ __forceinline__ void locate(const GridView &g, const SurfView &s, const double pt[3], F &&f) { locate_at(g, s, s.x, pt, f); }

// ---- moving substrate (the reference's `0*` of main.cpp:299 replaced by wcfe_params::subs_scale): driverVelocities_Biaxial
//      (UL:275-345) + predictor x_new = x + scale * dt * V_hlf (main.cpp:298-302), general 3-D path only.  Also watches how far the
//      nodes have drifted from the positions the point-location grid was built on (grid_ref): beyond half the grid's padding the host
//      rebuilds the grid before the next step (flag).
__global__ void __launch_bounds__(kBlock) k_predict_subs(int nn, int np, double *__restrict__ x, double *__restrict__ X, const double *__restrict__ V,
                                                         const double *__restrict__ A, double dt, double t_tot, double scale, Consts k,
                                                         const double *__restrict__ grid_ref, double half_pad2, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double xo[3] = {x[i], x[np + i], x[2 * np + i]};
    double vh[3];
#pragma unroll
    for (int c = 0; c < 3; c++) vh[c] = WCFE_ADD(V[c * np + i], WCFE_MUL(dt / 2.0, A[c * np + i]));
    vh[2] = 0.0;
    if (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f) {            // UL:298-341, Vbiax == 0
        if (xo[0] < -k.pin_limit || xo[0] > k.pin_limit) vh[0] = 0.0;
        if (xo[1] > k.pin_limit || xo[1] < -k.pin_limit) vh[1] = 0.0;
    }
    double d2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double xn = WCFE_ADD(xo[c], WCFE_MUL(WCFE_MUL(scale, dt), vh[c]));         // main.cpp:299: x + scale*dt*V_hlf, uncontracted
        X[c * np + i] = xo[c];
        x[c * np + i] = xn;
        const double d = xn - grid_ref[c * np + i];
        d2 += d * d;
    }
    if (d2 > half_pad2) *flag = 1;
}

// ---- updateFA (UL:2005-2114) + anchor location for calculateForce_FA_subs (UL:1301-1340) ---------------------
// One thread per cell node.  Substrate frozen (main.cpp:299 `0*`): subs_u == 0 whatever triangle holds the anchor,
// so the reference's first point location only matters for the discrete-state record (debug).
__global__ void __launch_bounds__(kBlock) k_bond(SurfView c, SurfView s, GridView g, BucketView bk, Consts k, double dt,
                                                 double ix0, double ix1, double iy0, double iy1,
                                                 int *__restrict__ d_flags, int *__restrict__ d_loc_bond,
                                                 int *__restrict__ d_loc_cnt, int *__restrict__ d_loc_first, int debug, int *__restrict__ err, int moving) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.nn) return;
    const double xi[3] = {c.x[i], c.x[c.np + i], c.x[2 * c.np + i]};
    const double Xi[3] = {c.X[i], c.X[c.np + i], c.X[2 * c.np + i]};
    double fa[3] = {c.fa[i], c.fa[c.np + i], c.fa[2 * c.np + i]};
    double cC = c.cC[i];
    double subs_u[3] = {0.0, 0.0, 0.0};          // (no containing triangle: the reference leaves it uninitialised, UL:2018 -- defined 0)
    if (debug || moving) {
        // the anchor in the substrate's coordinates at the START of the step (s.X; == s.x for the frozen substrate), UL:2020-2057; with a
        // moving substrate its displacement over the step is interpolated there: last match wins (UL:2057-2064)
        const double pt0[3] = {Xi[0] - fa[0], Xi[1] - fa[1], Xi[2] - fa[2]};
        int last = -1;
        const int np = s.np;
        locate_at(g, s, s.X, pt0, [&](int t, const double *Ar, double At) {
            last = t;
            if (moving) {
                const int i0 = s.tri[t], i1 = s.tri[s.nep + t], i2 = s.tri[2 * s.nep + t];
                const double w1 = Ar[0] / At, w2 = Ar[1] / At, w3 = Ar[2] / At;
#pragma unroll
                for (int q = 0; q < 3; q++)      // UL:2061, left to right, uncontracted
                    subs_u[q] = WCFE_ADD(WCFE_ADD(WCFE_MUL(w1, WCFE_SUB(s.x[q * np + i0], s.X[q * np + i0])), WCFE_MUL(w2, WCFE_SUB(s.x[q * np + i1], s.X[q * np + i1]))),
                                         WCFE_MUL(w3, WCFE_SUB(s.x[q * np + i2], s.X[q * np + i2])));
            }
        });
        if (debug) d_loc_bond[i] = last;
    }
    const double cell_u[3] = {xi[0] - Xi[0], xi[1] - Xi[1], xi[2] - Xi[2]};      // UL:2016
    const int fl = bond_update(k, dt, cell_u, subs_u, fa, cC);
    c.fa[i] = fa[0]; c.fa[c.np + i] = fa[1]; c.fa[2 * c.np + i] = fa[2];
    c.cC[i] = cC;
    d_flags[i] = fl;
    // anchor on the substrate for the force exchange: x_cell - fa_u (UL:1303), current substrate coordinates
    const double pt[3] = {xi[0] - fa[0], xi[1] - fa[1], xi[2] - fa[2]};
    // the substrate kernel only looks for anchors near substrate nodes inside the padded region sized on the host
    if (!(pt[0] >= ix0 && pt[0] <= ix1 && pt[1] >= iy0 && pt[1] <= iy1)) atomicOr(err, ERR_REGION);
    int cnt = 0, first = -1;
    locate(g, s, pt, [&](int t, const double *, double) {
        if (cnt == 0) first = t;
        cnt++;
        // append node i to triangle t's bucket (epoch-stamped counter: no reset pass between steps)
        unsigned long long *h = bk.head + t;
        unsigned long long old = *h, want;
        int slot;
        do {
            const unsigned long long seen = old;
            if ((unsigned int)(seen >> 32) != bk.epoch) { want = ((unsigned long long)bk.epoch << 32) | 1ull; slot = 0; }
            else { want = seen + 1ull; slot = (int)(unsigned int)seen; }
            old = atomicCAS(h, seen, want);
            if (old == seen) break;
        } while (true);
        if (slot < bk.cap) bk.items[(size_t)t * bk.cap + slot] = i; else atomicOr(err, ERR_BUCKET_OVERFLOW);
    });
    d_loc_cnt[i] = cnt; d_loc_first[i] = first;
}

// ---- gather over the one-ring: shared by cell and substrate ------------------------------------------------------
struct NodeSums { double res[3], rea[6], mm, asum, area, cen[3]; };

__device__ __forceinline__ void gather_ring(const SurfView &s, int i, double mu, double tk, double theta_g, double thick, NodeSums &a) {
    for (int k = 0; k < 3; k++) { a.res[k] = 0.0; a.cen[k] = 0.0; }
    for (int k = 0; k < 6; k++) a.rea[k] = 0.0;
    a.mm = 0.0; a.asum = 0.0; a.area = 0.0;
    double eai[6];
#pragma unroll
    for (int q = 0; q < 6; q++) eai[q] = s.ea_in[q * s.np + i];
    const int rb = s.ring_ptr[i], re = s.ring_ptr[i + 1];
    for (int r = rb; r < re; r++) {
        const int code = s.ring[r];
        const int e = code >> 2, slot = code & 3;
        int id[3] = {s.tri[e], s.tri[s.nep + e], s.tri[2 * s.nep + e]};
        double X[3][3], x[3][3], e6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 3; j++) {
#pragma unroll
            for (int c = 0; c < 3; c++) { X[j][c] = s.X[c * s.np + id[j]]; x[j][c] = s.x[c * s.np + id[j]]; }
        }
        // ea at the integration point: 1/3 e1 + 1/3 e2 + 1/3 e3 (UL:781)
#pragma unroll
        for (int q = 0; q < 6; q++)
            e6[q] = 1.0 / 3.0 * s.ea_in[q * s.np + id[0]] + 1.0 / 3.0 * s.ea_in[q * s.np + id[1]] + 1.0 / 3.0 * s.ea_in[q * s.np + id[2]];
        ElemOut o;
        elem_eval(X, x, e6, mu, tk, theta_g, thick, o);
#pragma unroll
        for (int c = 0; c < 3; c++) a.res[c] += slot == 0 ? -(o.r1[c] + o.r2[c]) : (slot == 1 ? o.r1[c] : o.r2[c]);
#pragma unroll
        for (int q = 0; q < 6; q++) a.rea[q] += (1. / 3.) * (o.eas[q] - eai[q]) * o.A0 * thick;       // UL:911
        a.mm += (1. / 3.) * o.A0 * thick;                                                               // UL:913
        a.asum += o.a_cur / 3;                                                                          // UL:1251
        if (slot == 0) {                                                                                // each element once: UL:1655-1661
            a.area += o.a_cur;
#pragma unroll
            for (int c = 0; c < 3; c++) a.cen[c] += (1. / 3.) * ((x[0][c] + x[1][c]) + x[2][c]) * o.a_cur;
        }
    }
}

// deterministic block reduction (fixed tree) of 4 doubles; result valid in thread 0
__device__ __forceinline__ void block_reduce4(double v[4], double *sh /*[4*32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], off);
    if (lane == 0) for (int q = 0; q < 4; q++) sh[q * 32 + wid] = v[q];
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double t = lane < nw ? sh[q * 32 + lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
            v[q] = t;
        }
    }
}

// ---- updateForces_Explicit, cell overload (UL:508-630), fused: residual gather + all nodal forces + update ----------
// Boundary vertices need the total area (UL:589), a grid-wide sum of this very kernel's element areas: they park
// their partial force in BoundaryView::buf and are finished by k_boundary_finish.
template <bool DBG>
__global__ void __launch_bounds__(kBlock) k_cell_node(SurfView s, BoundaryView b, Consts k, double dt, double dt_real, double t_tot,
                                                      int rng_mode, const unsigned int *__restrict__ tape, unsigned int seed, unsigned int step,
                                                      double ap_threshold, double *__restrict__ partials, int *__restrict__ d_ap_fire, DebugView dbg, double dt_drv) {
    __shared__ double sh[4 * 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < s.nn) {
        const double theta_g = 1.0 + k.thetag_dot * dt_real;
        NodeSums a;
        gather_ring(s, i, k.mu_cell, t_act_of(k, t_tot) * k.kappa, theta_g, k.thick, a);
        red[0] = a.area; red[1] = a.cen[0]; red[2] = a.cen[1]; red[3] = a.cen[2];
        const int np = s.np;
        double V[3] = {s.V[i], s.V[np + i], s.V[2 * np + i]};
        double A[3] = {s.A[i], s.A[np + i], s.A[2 * np + i]};
        double Vh[3];
#pragma unroll
        for (int c = 0; c < 3; c++) Vh[c] = V[c] + (dt_drv / 2.0) * A[c];                     // UL:118 (dt_drv: what the driver was called with)
        Vh[2] = 0.0;
        // integrin force on the cell node (UL:1258-1261; kMD UL:1246-1269)
        const double u[3] = {s.fa[i], s.fa[np + i], s.fa[2 * np + i]};
        const double cC = s.cC[i];
        double fi[3], F_FA[3], F_vd[3];
        integrin_force(k, u, fi);
#pragma unroll
        for (int c = 0; c < 3; c++) { F_FA[c] = -cC * fi[c] * a.asum * k.rho_i_max; F_vd[c] = -k.drag * Vh[c]; }   // UL:1259, :571
        s.asum[i] = a.asum;
        // strain update (UL:607-612)
#pragma unroll
        for (int q = 0; q < 6; q++) s.ea_out[q * np + i] = s.ea_in[q * np + i] + a.rea[q] / a.mm;
        if (DBG) {
#pragma unroll
            for (int c = 0; c < 3; c++) { dbg.res[c * np + i] = a.res[c]; dbg.f_fa[c * np + i] = F_FA[c]; dbg.f_ap[c * np + i] = 0.0; dbg.f_bend[c * np + i] = 0.0; dbg.f_area[c * np + i] = 0.0; }
            dbg.mm[i] = a.mm;
#pragma unroll
            for (int q = 0; q < 6; q++) dbg.res_ea[q * np + i] = a.rea[q];
        }
        const int bi = b.index[i];
        if (bi < 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double f = F_vd[c] + F_FA[c] - a.res[c];                            // UL:592 with the boundary-only terms == 0
                node_update(f, a.mm, k.damp, k.rho, dt, dt_real, Vh[c], V[c], A[c]);
                s.V[c * np + i] = V[c]; s.A[c * np + i] = A[c];
            }
        } else {
            // boundary vertex: neighbours along the boundary curve (UL:1479-1507), later matches overwrite
            const double xi[3] = {s.x[i], s.x[np + i], s.x[2 * np + i]};
            double j1[3] = {0, 0, 0}, j2[3] = {0, 0, 0};
            for (int q = b.cand_ptr[bi]; q < b.cand_ptr[bi + 1]; q++) {
                const int nj = b.cand[2 * q], nint = b.cand[2 * q + 1];
                const double xb[3] = {s.x[nj], s.x[np + nj], s.x[2 * np + nj]};
                const double xin[3] = {s.x[nint], s.x[np + nint], s.x[2 * np + nint]};
                if (bnd_is_first(xi, xb, xin)) { j1[0] = xb[0]; j1[1] = xb[1]; j1[2] = xb[2]; }
                else { j2[0] = xb[0]; j2[1] = xb[1]; j2[2] = xb[2]; }
            }
            double nrm[3], tl;
            bnd_normal(xi, j1, j2, nrm, tl);
            const double kap = bnd_curvature(xi, j1, j2, tl);
            // actin polymerisation: one draw per boundary vertex per step (UL:1525-1541)
            double Pr;
            if (rng_mode == 0) Pr = (double)tape[bi] / 2147483648.0;
            else Pr = (double)philox_u32(seed, step, (unsigned int)bi) * (1.0 / 4294967296.0);
            const bool fire = Pr < ap_threshold;
            const double factin = fire ? k.ap_quantum * ap_trips(k.ap_density, tl) : 0.0;
            d_ap_fire[i] = fire ? 1 : 0;
            double *o = b.buf + (size_t)16 * bi;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double F_b = k.k_bend * kap * nrm[c], F_AP = factin * nrm[c];       // UL:1633, :1542
                o[c] = ((F_vd[c] + F_b) + F_AP) + F_FA[c];                                // UL:592 up to F_area
                o[3 + c] = a.res[c];
                if (DBG) { dbg.f_bend[c * np + i] = F_b; dbg.f_ap[c * np + i] = F_AP; }
            }
            o[6] = a.mm; o[7] = tl; o[8] = nrm[0]; o[9] = nrm[1]; o[10] = nrm[2];
        }
    }
    block_reduce4(red, sh);
    if (threadIdx.x == 0) { for (int q = 0; q < 4; q++) partials[(size_t)4 * blockIdx.x + q] = red[q]; }
}

// ---- total area / centroid (UL:1641-1667) + area-constraint force and update of the boundary vertices (UL:1670-1763) --
// every block reduces the partial sums itself, in the same fixed order (bit-reproducible run to run and identical across
// blocks), then finishes its share of the boundary vertices; block 0 publishes the diagnostics
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// one thread: bump this rank's counter `which` (0 step, 1 partials) on every peer -- after a fence that orders the pushes of the
// kernels that completed before this thread's griddepcontrol.wait returned
__device__ __noinline__ void cell_part_signal(const CellPeer *peer, int npeer, int which) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    for (int q = 0; q < npeer; q++)
        asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(which ? peer[q].flag_part : peer[q].flag_step), "l"(1ull) : "memory");
}
// one thread: wait until every peer's counter (base 0 step, 8 partials) has reached `want`.  Bounded like halo_wait: 10 s, then the
// error flag is raised and every later wait returns at once -- a wrong answer that is reported, never a hung GPU.
__device__ __noinline__ void cell_part_wait(const unsigned long long *flags, int base, int npeer, unsigned long long want, int *err) {
    if (want == 0) return;
    const unsigned long long t0 = global_timer_ns();
    for (int q = 0; q < npeer; q++)
        while (ld_acquire_sys_u64(flags + base + q) < want) {
            if (*(volatile int *)err & ERR_CELL_HALO_TIMEOUT) return;
            if (global_timer_ns() - t0 > 10000000000ull) { atomicOr(err, ERR_CELL_HALO_TIMEOUT); return; }
            __nanosleep(64);
        }
}

// (the work of one block of `nblk`; the caller has passed griddepcontrol.wait.  sh: 4 * 32 doubles, tot: 4 doubles of shared memory)
template <bool DBG, bool BATCH, bool PART>
__device__ __forceinline__ void boundary_finish_body(int bid, int nblk, double *sh, double *tot, SurfView s, BoundaryView b, const Consts &k_, int npart,
                                                     const double *__restrict__ partials, const double *__restrict__ V_in, const double *__restrict__ A_in,
                                                     double dt, double dt_real, double *__restrict__ diag, int *__restrict__ err, const DebugView &dbg,
                                                     const BatchView &bt, double dt_drv, const CellPartView &cp) {
    if (PART) {
        // this rank's cell kernel is complete: tell everybody that the partial sums of its patches are in their arrays, then wait
        // for everybody else's
        if (bid == 0 && threadIdx.x == 0) cell_part_signal(cp.peer, cp.npeer, 1);
        if (threadIdx.x == 0) cell_part_wait(cp.flags, 8, cp.npeer, (unsigned long long)cp.step + 1ull, cp.err);
        __syncthreads();
    }
    const Consts &k = BATCH ? bt.k[blockIdx.y] : k_;
    if (BATCH) {
        const size_t r = blockIdx.y;
        batch_shift(s, r);
        b.buf += bt.bnd * r; partials += bt.partials * r; V_in += 3 * (size_t)s.np * r; A_in += 3 * (size_t)s.np * r; diag += 8 * r; err += r;
    }
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    // strided accumulation in a fixed order, then a fixed tree
    for (int p = threadIdx.x; p < npart; p += blockDim.x)
        for (int q = 0; q < 4; q++) v[q] += PART ? __ldcg(partials + (size_t)4 * p + q) : partials[(size_t)4 * p + q];      // (peer-written rows: past the incoherent L1)
    block_reduce4(v, sh);
    if (threadIdx.x == 0) {
        tot[0] = v[0]; tot[1] = v[1] / v[0]; tot[2] = v[2] / v[0]; tot[3] = v[3] / v[0];
        if (bid == 0) {
            for (int q = 0; q < 4; q++) diag[q] = tot[q];
            // NaN/Inf sentinel: the total area and the centroid see every cell node's position once per step
            if (!(isfinite(tot[0]) && isfinite(tot[1]) && isfinite(tot[2]))) atomicOr(err, ERR_NAN);
        }
    }
    __syncthreads();
    const double area_tot = tot[0];
    const int np = s.np;
    for (int bi = bid * blockDim.x + threadIdx.x; bi < b.nb; bi += nblk * blockDim.x) {
        if (PART && !cp.bnd_local[bi]) continue;      // its patch is another rank's: the record below is stale here
        const int i = b.node[bi];
        const double *o = b.buf + (size_t)16 * bi;
        const double mm = o[6], tl = o[7];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double F_area = -k.k_area * (area_tot - k.area_0) / 2 * tl * o[8 + c];   // UL:1757
            const double f = (o[c] + F_area) - o[3 + c];
            double V = V_in[c * np + i], A = A_in[c * np + i];       // start-of-step values (the planar path double-buffers V, A)
            double Vh = V + (dt_drv / 2.0) * A;
            if (c == 2) Vh = 0.0;
            node_update(f, mm, k.damp, k.rho, dt, dt_real, Vh, V, A);
            s.V[c * np + i] = V; s.A[c * np + i] = A;
            if (PART && c < 2)      // a boundary vertex other ranks read as a halo row: its V, A are final only here
                for (int q = cp.bexp_ptr[bi]; q < cp.bexp_ptr[bi + 1]; q++) {
                    const CellPeer *pr = cp.peer + cp.bexp_peer[q];
                    pr->V[cp.step & 1][c * np + i] = V; pr->A[cp.step & 1][c * np + i] = A;
                }
            if (DBG) dbg.f_area[c * np + i] = F_area;
        }
    }
}

template <bool DBG, bool BATCH = false, bool PART = false>
__global__ void __launch_bounds__(1024) k_boundary_finish(SurfView s, BoundaryView b, Consts k_, int npart, const double *__restrict__ partials,
                                                          const double *__restrict__ V_in, const double *__restrict__ A_in,
                                                          double dt, double dt_real, double *__restrict__ diag, int *__restrict__ err, DebugView dbg,
                                                          BatchView bt, double dt_drv, CellPartView cp) {
    __shared__ double sh[4 * 32];
    __shared__ double tot[4];
    pdl_launch_dependents();      // the substrate kernel may start its cell-independent part now
    pdl_wait();                   // everything below reads what the cell kernel wrote
    boundary_finish_body<DBG, BATCH, PART>(blockIdx.x, gridDim.x, sh, tot, s, b, k_, npart, partials, V_in, A_in, dt, dt_real, diag, err, dbg, bt, dt_drv, cp);
}

// ---- updateForces_Explicit, substrate overload (UL:429-504) incl. driverVelocities_Biaxial (UL:275-345) and
//      calculateForce_FA_subs (UL:1297-1454) as a gather over the per-triangle buckets filled by k_bond -----------------
template <bool DBG>
__global__ void __launch_bounds__(kBlock) k_subs_node(SurfView s, SurfView c, BucketView bk, Consts k, double dt, double dt_real, double t_tot,
                                                      double bx0, double bx1, double by0, double by1, DebugView dbg, double dt_drv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nn) return;
    const double theta_g = 1.0 + k.thetag_dot * dt_real;
    NodeSums a;
    gather_ring(s, i, k.mu_subs, 0.0, theta_g, k.thick, a);
    const int np = s.np;
    const double Xi[3] = {s.X[i], s.X[np + i], s.X[2 * np + i]};
    double F_FA[3] = {0.0, 0.0, 0.0};
    // force from the cell: only substrate nodes inside the (padded) bounding box of the anchors can receive any
    if (Xi[0] >= bx0 && Xi[0] <= bx1 && Xi[1] >= by0 && Xi[1] <= by1) {
        const int rb = s.ring_ptr[i], re = s.ring_ptr[i + 1];
        for (int r = rb; r < re; r++) {
            const int code = s.ring[r];
            const int e = code >> 2, slot = code & 3;
            const unsigned long long h = bk.head[e];
            if ((unsigned int)(h >> 32) != bk.epoch) continue;
            int cnt = (int)(unsigned int)h;
            if (cnt > bk.cap) cnt = bk.cap;
            const int i0 = s.tri[e], i1 = s.tri[s.nep + e], i2 = s.tri[2 * s.nep + e];
            const double n1[3] = {s.x[i0], s.x[np + i0], s.x[2 * np + i0]};
            const double n2[3] = {s.x[i1], s.x[np + i1], s.x[2 * np + i1]};
            const double n3[3] = {s.x[i2], s.x[np + i2], s.x[2 * np + i2]};
            // ascending cell-node order inside the bucket (insertion order is not deterministic)
            int prev = -1;
            for (int m = 0; m < cnt; m++) {
                int ni = 0x7fffffff;
                for (int q = 0; q < cnt; q++) { const int v = bk.items[(size_t)e * bk.cap + q]; if (v > prev && v < ni) ni = v; }
                prev = ni;
                const double u[3] = {c.fa[ni], c.fa[c.np + ni], c.fa[2 * c.np + ni]};
                const double pt[3] = {c.x[ni] - u[0], c.x[c.np + ni] - u[1], c.x[2 * c.np + ni] - u[2]};   // UL:1303
                double Ar[3], At;
                point_in_triangle(n1, n2, n3, pt, Ar, At);
                double fi[3];
                integrin_force(k, u, fi);
                const double w = Ar[slot] / At, cC = c.cC[ni], As = c.asum[ni];
#pragma unroll
                for (int d = 0; d < 3; d++) F_FA[d] += cC * fi[d] * w * As * k.rho_i_max;                  // UL:1414-1422
            }
        }
    }
    double V[3] = {s.V[i], s.V[np + i], s.V[2 * np + i]};
    double A[3] = {s.A[i], s.A[np + i], s.A[2 * np + i]};
    double Vh[3];
#pragma unroll
    for (int d = 0; d < 3; d++) Vh[d] = V[d] + (dt_drv / 2.0) * A[d];
    Vh[2] = 0.0;
    // pinned edges, UL:298-341 with Vbiax == 0: all three time windows zero the component; nothing fires at t == t_i, t_f
    if (t_tot < k.t_i || (t_tot > k.t_i && t_tot < k.t_f) || t_tot > k.t_f) {
        if (Xi[0] < -k.pin_limit || Xi[0] > k.pin_limit) Vh[0] = 0.0;
        if (Xi[1] > k.pin_limit || Xi[1] < -k.pin_limit) Vh[1] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double f = F_FA[d] - a.res[d];                                              // UL:471
        node_update(f, a.mm, k.damp, k.rho, dt, dt_real, Vh[d], V[d], A[d]);
        s.V[d * np + i] = V[d]; s.A[d * np + i] = A[d];
    }
#pragma unroll
    for (int q = 0; q < 6; q++) s.ea_out[q * np + i] = s.ea_in[q * np + i] + a.rea[q] / a.mm;   // UL:486-491
    if (DBG) {
#pragma unroll
        for (int d = 0; d < 3; d++) { dbg.res[d * np + i] = a.res[d]; dbg.f_fa[d * np + i] = F_FA[d]; }
        dbg.mm[i] = a.mm;
#pragma unroll
        for (int q = 0; q < 6; q++) dbg.res_ea[q * np + i] = a.rea[q];
    }
}

}   // namespace wcfe
