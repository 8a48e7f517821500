/* wcfe.h -- C ABI of libwcfe.so: the explicit updated-Lagrangian (UL) step of wholeCellFE's cellBP on B200.
 *
 * Drop-in boundary.  The reference has no plugin loader: the boundary is the
 * link-time free-function interface between cellBP/main.cpp and cellBP/UL_growth_explicit.cpp, whose
 * prototypes main.cpp re-declares by hand (wholeCellFE_k1/cellBP/main.cpp:69-87) and calls from
 * advance_sim() (main.cpp:168-418).  This header is the flat C surface a binding would target (same
 * style as El Topo's own C API, eltopo3d/eltopo.h:150-268: flattened `const double* xyz[3N]`,
 * `const int* tri[3M]`, caller-owned buffers, int status codes, no exceptions).  The C++ adapter
 * wholecellfe_b200/csrc/ul_adapter.cpp defines the reference's C++ symbols (the five main.cpp calls and the three drivers it keeps
 * commented out) on top of it so that
 * the unmodified reference main.cpp links against libwcfe instead of UL_growth_explicit.o
 * (INTEGRATION.md).
 *
 * Conventions: all host arrays are array-of-structs exactly as the reference stores them
 * (std::vector<Vec3d> == 3 contiguous doubles per node, common/vec.h:22-25; strains 6 per node in the
 * order xx,yy,zz,xy,xz,yz, UL_growth_explicit.cpp:784-786,886).  Device state is struct-of-arrays FP64
 * and never leaves the GPU unless asked for.  Every entry point returns WCFE_OK (0) or a negative
 * error code; wcfe_last_error() gives the text.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with WCFE_ERR_CUDA.
 */
#ifndef WCFE_H
#define WCFE_H

#ifdef __cplusplus
extern "C" {
#endif

#define WCFE_OK 0
#define WCFE_ERR_ARG (-1)          /* bad argument / state (e.g. no mesh uploaded) */
#define WCFE_ERR_CUDA (-2)         /* CUDA runtime error, or no device */
#define WCFE_ERR_UNSUPPORTED (-3)  /* a parameter combination the GPU path does not implement */
#define WCFE_ERR_CAPACITY (-4)     /* a device-side table overflowed (see wcfe_last_error) */
#define WCFE_ERR_NUMERIC (-5)      /* NaN/Inf sentinel tripped */
#define WCFE_ERR_PEER (-6)         /* partitioned run: a neighbouring rank's halo rows did not arrive */

#define WCFE_CELL 0
#define WCFE_SUBS 1

/* variants == which reference source tree: wholeCellFE_k1 / _k31 / _kMD */
#define WCFE_K1 0
#define WCFE_K31 1
#define WCFE_KMD 2

/* state fields (widths in doubles per node) */
#define WCFE_F_X 0      /* 3  positions                DynamicSurface::pm_positions  */
#define WCFE_F_V 1      /* 3  velocity                 m_Vt   (dynamicsurface.h:234-264) */
#define WCFE_F_A 2      /* 3  acceleration             m_At   */
#define WCFE_F_EA 3     /* 6  Euler-Almansi strain     m_eat  */
#define WCFE_F_CA 4     /* 1  free integrin conc.      m_cA   (cell only) */
#define WCFE_F_CB 5     /* 1                           m_cB   (cell only; never read by the UL path) */
#define WCFE_F_CC 6     /* 1  bound integrin fraction  m_cC   (cell only) */
#define WCFE_F_FAU 7    /* 3  integrin anchor vector   m_fau  (cell only) */
#define WCFE_F_XPREV 8  /* 3  positions at the start of the step, X_t (cell; uploadable in adapter mode, after WCFE_F_X) */

/* discrete state of the last step, one int per cell node (parity gates of SURVEY.md 8d) */
#define WCFE_D_LOC_BOND 0         /* containing substrate triangle in updateFA (last match), -1 none; debug mode only */
#define WCFE_D_LOC_FORCE_CNT 1    /* number of containing triangles in calculateForce_FA_subs */
#define WCFE_D_LOC_FORCE_FIRST 2  /* lowest-index containing triangle, -1 none */
#define WCFE_D_AP_FIRE 3          /* actin-polymerisation Poisson event fired (boundary vertices) */
#define WCFE_D_FA_RESET 4         /* anchor reset branch of UL_growth_explicit.cpp:2107-2111 */
#define WCFE_D_CC_CLAMP 5         /* c_C < 0 clamp (UL:2095) */

/* RNG modes for the actin-polymerisation draw (UL_growth_explicit.cpp:1525) */
#define WCFE_RNG_GLIBC_REPLAY 0   /* replays glibc rand() after srand(seed): bit-identical to the (patched) reference */
#define WCFE_RNG_PHILOX 1         /* counter-based Philox4x32-10 keyed by (seed, step, boundary ordinal) */

/* Everything the reference hard-codes in source, with its literal values as defaults
 * (wcfe_default_params).  File:line are in wholeCellFE_k1/cellBP/UL_growth_explicit.cpp. */
typedef struct wcfe_params {
    int law;             /* 0 linear integrin spring, 1 piecewise-linear MD table (kMD :1204-1214) */
    double kint_force;   /* kint inside updateForces_Explicit (:458,:559): 1000 / 31000 */
    double kint_bond;    /* kint main.cpp hands to updateFA (main.cpp:350-354): 1000 / 31000 / -1 */
    double subs_scale;   /* the literal `0*` of main.cpp:299 (0: frozen substrate, the planar fast path; != 0: moving substrate, general kernels, wcfe_step_n only) */
    double mu_cell, mu_subs;                  /* :724,:727 */
    double t_act_rate, t_act_max, t_act_t1;   /* :744-756  t<2 ? 200 t : 400 */
    double thick, kappa;                      /* :767,:871  (kappa must be 0.5 on the GPU: the a0 term vanishes) */
    double thetag_dot;                        /* :714 growth rate; theta_g = 1 + thetag_dot*dt */
    double damp, rho, drag;                   /* :596,:597,:569 */
    double rho_i_max;                         /* :1231 */
    double k_bend;                            /* :1632 */
    double k_area, area_0;                    /* :1755-1756 */
    double ap_rate, ap_quantum, ap_density;   /* :1528,:1539,:1530 */
    double k_on, ka, kb, Fa, Fb;              /* :2079-2083 */
    double pin_limit, t_i, t_f;               /* :298-301 */
    double md_px[6], md_py[6], md_k[6];       /* kMD :1208,:1211,:1214 */
    double c_A_init;                          /* main.cpp:672 */
} wcfe_params;

typedef struct wcfe_ctx wcfe_ctx;

int wcfe_default_params(wcfe_params *p, int variant);

/* lifetime; `device` is the CUDA ordinal (one context per process per GPU) */
int wcfe_create(wcfe_ctx **out, int device, const wcfe_params *p);
int wcfe_destroy(wcfe_ctx *c);
/* Ensembles (the reference runs parameter sweeps as independent processes, s_sim_031021_1.sub): ONE context that
 * holds `replicas` independent copies of the same cell + substrate pair, params[r] the parameters of replica r (spring law, ligand
 * density, ... ; ap_rate must be common).  wcfe_upload_mesh initialises every replica alike; wcfe_step_n advances them all in
 * lockstep with one launch per kernel (gridDim.y = replicas) on the same random-number tape (common random numbers); state
 * up/downloads, wcfe_set_params, wcfe_get_discrete and wcfe_get_diagnostics address the replica chosen by wcfe_select_replica
 * (default 0).  The piecewise entry points (predict / commit / update*, round trip, debug capture) are single-replica only and
 * return WCFE_ERR_UNSUPPORTED on a batched context; so does a non-planar mesh. */
int wcfe_create_batch(wcfe_ctx **out, int device, const wcfe_params *params, int replicas);
int wcfe_select_replica(wcfe_ctx *c, int replica);
int wcfe_replicas(wcfe_ctx *c);
/* per-replica Philox key (WCFE_RNG_PHILOX only); wcfe_set_rng resets every replica to its common seed */
int wcfe_set_replica_seed(wcfe_ctx *c, int replica, unsigned int seed);
int wcfe_set_params(wcfe_ctx *c, const wcfe_params *p);
/* Host-only: the parameter sweep of an ensemble (config 5: spring law x ligand density x seed) and its shard over ranks, deterministic
 * from the cell id (every rank rebuilds its part without communication; same table as wholecellfe_b200/ensemble.py).  wcfe_sweep_cell:
 * parameters (wcfe_default_params of the law with rho_i_max set) and RNG seed of cell `id` out of n_cells -- laws[n_laws] (WCFE_K1 ...) x
 * round(sqrt(n_cells / n_laws)) log-spaced rho_i_max in [rho_min, rho_max], law-major, seed = 1 + number of completed passes through
 * that grid.  wcfe_ensemble_shard: the contiguous balanced range [lo, hi) of cell ids of `rank`. */
int wcfe_sweep_cell(int n_cells, int n_laws, const int *laws, double rho_min, double rho_max, int id, wcfe_params *params, unsigned int *seed);
int wcfe_ensemble_shard(int n_cells, int world, int rank, int *lo, int *hi);
/* run on an externally owned cudaStream_t (e.g. torch's current stream); NULL = the context's own non-blocking stream;
 * pass cudaStreamLegacy ((cudaStream_t)0x1) to name the legacy default stream explicitly */
int wcfe_set_stream(wcfe_ctx *c, void *cuda_stream);
const char *wcfe_last_error(wcfe_ctx *c);

/* Mesh + state.  upload_mesh replaces connectivity, rebuilds the vertex->triangle CSR, the boundary
 * tables and (substrate) the point-location grid, and resets the surface's state the way
 * scriptinit.cpp:552-580 + main.cpp:671-674 do (V = A = ea = fa_u = 0, c_A = c_A_init, c_C = 0).
 * Replaces: ScriptInit/SurfTrack construction (main.cpp:597-609) and, after remeshing, the re-upload. */
int wcfe_upload_mesh(wcfe_ctx *c, int surface, int nn, const double *xyz, int ne, const int *tri);
int wcfe_upload_state(wcfe_ctx *c, int surface, int field, const double *host);
int wcfe_download_state(wcfe_ctx *c, int surface, int field, double *host);
/* The same for several fields of one surface at once: one pinned transfer and one stream synchronisation per call instead of one
 * per field (fields[q] is a WCFE_F_* id, host[q] its image laid out as for the single-field calls; processed in order). */
int wcfe_upload_states(wcfe_ctx *c, int surface, int n_fields, const int *fields, const double *const *host);
int wcfe_download_states(wcfe_ctx *c, int surface, int n_fields, const int *fields, double *const *host);
int wcfe_mesh_sizes(wcfe_ctx *c, int surface, int *nn, int *ne);

int wcfe_set_rng(wcfe_ctx *c, int mode, unsigned int seed);
/* resume a glibc replay mid-run (checkpoint/restart): discard n_draws values (one per boundary vertex per step) */
int wcfe_rng_skip(wcfe_ctx *c, long long n_draws);
int wcfe_set_debug(wcfe_ctx *c, int on);   /* also record WCFE_D_LOC_BOND (extra point location per step) and force fields */
/* per-kernel CUDA-event timing of wcfe_step_n (events on the launching stream; switches the PDL chain off while on); slots: 0 predictor
 * (k_predict_cell; fused into the cell kernel on the planar path) 1 k_bond 2 cell kernel (k_cell_patch / k_cell_node) 3 boundary finish
 * 4 substrate kernel (k_subs_ring / k_subs_patch / k_subs_node) 5 k_bucket_reduce.  ms8 = accumulated ms, count8 = launches timed. */
int wcfe_set_profile(wcfe_ctx *c, int on);
/* test hook: run the general 3-D (node-centric) kernels even when the state is planar */
int wcfe_set_force_general(wcfe_ctx *c, int on);
/* out8: planar fast path active (0/1), patch-local elements / elements for cell and substrate (recompute overhead),
 * max halo nodes / max elements per cell patch, max halo nodes per substrate patch, substrate kernel in use (2 k_subs_ring: node-centric,
 * the default; 0 k_subs_patch: element-once), triangle-bucket capacity */
int wcfe_get_path_info(wcfe_ctx *c, double *out8);
int wcfe_get_kernel_ms(wcfe_ctx *c, double *ms8, long long *count8);

/* ---- the reference's call sequence, one entry point per call site of advance_sim() --------------- */

/* driverVelocities_Explicit / driverVelocities_Biaxial + the inline predictor (main.cpp:224-231,
 * :293-302): V_hlf = V + dt/2 A (z = 0, substrate pins), x_new = x + dt V_hlf (substrate: + 0*).
 * This is the "mesh-driver interface": positions in, predicted positions out.  xyz_new_out may be NULL
 * (keep the prediction on the device). */
int wcfe_predict(wcfe_ctx *c, int surface, double dt, double t_curr, double *xyz_new_out);
/* positions the host integrator actually accepted (El Topo integrate(), main.cpp:248,:316);
 * NULL = accept the prediction as is (collision_safety 0) */
int wcfe_commit(wcfe_ctx *c, int surface, const double *xyz_final);
/* updateFA + updateForces_Explicit(cell) + updateForces_Explicit(substrate) (main.cpp:354,:360,:361) */
int wcfe_update(wcfe_ctx *c, double dt_req, double dt_real, double t_curr);
/* the same in two halves, for the adapter: the reference hands the cell's V, A, ea to updateForces_Explicit(cell)
 * (main.cpp:360) and the substrate's only to the substrate overload (main.cpp:361).  _cell = updateFA + cell overload
 * (+ anchor buckets), _subs = substrate overload; _subs must follow the _cell of the same step. */
int wcfe_update_cell(wcfe_ctx *c, double dt_req, double dt_real, double t_curr);
int wcfe_update_subs(wcfe_ctx *c, double dt_req, double dt_real, double t_curr);
/* driverVelocities_Explicit (surface == WCFE_CELL, UL:103-124) / driverVelocities_Biaxial (WCFE_SUBS, UL:275-345) as a pure
 * function of host-owned arrays (AoS [3*nn]): V_hlf = V + dt/2 A, z = 0, substrate edge pins.  Computed on the device. */
int wcfe_driver_velocities(wcfe_ctx *c, int surface, int nn, const double *X, const double *V, const double *A, double dt, double t_curr, double *Vh_out);
/* Every driverVelocities_* of UL:103-345 by kind (main.cpp:69-76 declares them, :217-224 / :291-295 call the first two and keep the others
 * commented out): V_hlf = V + dt/2 A with the driver's boundary conditions, computed on the device as a pure function of host-owned AoS
 * arrays.  wcfe_driver_velocities(c, surface, ...) is kind == surface.  `surface` says whose V_hlf this is for the update that follows:
 *  - WCFE_SUBS with any kind: the substrate's nodal update V = V_hlf + (dt_real - dt/2) A (UL:499, the only reader of V_hlf in that
 *    overload) uses this driver's V_hlf (one context, one GPU, frozen substrate; WCFE_ERR_UNSUPPORTED otherwise);
 *  - WCFE_CELL with a kind other than EXPLICIT: the function value is returned, but the cell's update refuses (WCFE_ERR_UNSUPPORTED):
 *    its drag force reads V_hlf (UL:571-573) and the kernels form it as driverVelocities_Explicit does.
 * The choice lasts for one update, as one driver call per step does in main.cpp. */
#define WCFE_DRV_EXPLICIT 0     /* UL:103-124 */
#define WCFE_DRV_BIAXIAL 1      /* UL:275-345 */
#define WCFE_DRV_UNIAXIAL 2     /* UL:125-168 */
#define WCFE_DRV_OFFBIAXIAL 3   /* UL:214-273 */
#define WCFE_DRV_BULGETEST 4    /* UL:170-212 (keeps the normal component) */
int wcfe_driver_velocities_kind(wcfe_ctx *c, int surface, int kind, int nn, const double *X, const double *V, const double *A, double dt, double t_curr, double *Vh_out);
/* adapter mode: x (WCFE_F_X) and X_prev (WCFE_F_XPREV) were uploaded by the host integrator instead of predict/commit */
int wcfe_mark_positions_committed(wcfe_ctx *c);

/* El-Topo-free fast path: n x (predict, commit(NULL), update) without leaving the device; time advances
 * as sim->m_curr_t += accum_dt does (main.cpp:405).  *t_io: in = time at the start, out = time after. */
int wcfe_step_n(wcfe_ctx *c, int n, double dt, double *t_io);

/* the same without waiting: returns once the n steps are queued on the context's stream (ensembles of many small
 * independent cells per GPU: one host thread keeps every context busy).  wcfe_sync() waits and reports device errors. */
int wcfe_step_n_async(wcfe_ctx *c, int n, double dt, double *t_io);

/* host round trip per step, as a host integrator in the loop needs it (El Topo integrate(), main.cpp:248,:316): predict -> D2H of
 * the predicted positions into pinned buffers (arrays of Vec3d) -> H2D of the same buffers as "accepted" positions -> update ->
 * D2H of the diagnostics.  Used for the end-to-end measurement; returns the bytes moved per step.
 * Mode 0 (default): only the CELL's positions travel -- the substrate is frozen (main.cpp:299), its predicted positions are bit
 * for bit the ones the host holds and wcfe_commit ignores what comes back for it -- and the hand-over of step q+1 overlaps the
 * substrate half of step q on a second stream.  Mode 1: both surfaces travel, everything serial (what ul_adapter.cpp does, where
 * the reference application owns every array).  Both give the same state as wcfe_step_n, bit for bit. */
int wcfe_set_roundtrip_mode(wcfe_ctx *c, int mode);
int wcfe_step_n_roundtrip(wcfe_ctx *c, int n, double dt, double *t_io, long long *h2d_bytes_per_step, long long *d2h_bytes_per_step);

/* out8: total area, centroid x,y,z (UL:531-535), |Res_cell|, |Res_subs| (UL:593,:472), step counter, time */
int wcfe_get_diagnostics(wcfe_ctx *c, double *out8);
int wcfe_get_discrete(wcfe_ctx *c, int which, int *out);
/* mesh incidence as built for the device: out arrays may be NULL */
int wcfe_get_boundary_flags(wcfe_ctx *c, int surface, int *out);
/* number of kernels launched since creation, and wall-free device time of the last step_n in ms */
long long wcfe_kernel_launches(wcfe_ctx *c);
double wcfe_last_step_ms(wcfe_ctx *c);
/* individual nodal force fields on the current device state, same ids as the oracle's force_field:
 * 0 Res(cell) 1 F_FA 2 F_AP 3 F_bending 4 F_area 5 Res(subs) 6 F_FA_subs; out[3*nn]; mm[nn] / res_ea[6*nn]
 * may be NULL.  Diagnostic entry point for the force-level parity gate (1e-10). */
int wcfe_force_field(wcfe_ctx *c, int which, double dt, double t_curr, double *out, double *mm, double *res_ea);
/* Dormant physics of the cell: fields the reference defines and whose calls / consumers are commented out, computed on the current
 * device state (x, start-of-step X, c_A, c_C); diagnostic entry point outside the step, general (3-D) arithmetic.  For a context that
 * holds the whole cell (single GPU, or a partitioned substrate with the cell replicated); with a partitioned CELL a rank's rows outside
 * its own patches are not current and neither are these fields there.  WCFE_ERR_UNSUPPORTED on a batched context.
 *   WCFE_DORMANT_RES_ADIF / _RES_CDIF  diffusion residuals of free / bound integrins (evalElementRe UL:833-842, :915-925, assembled in
 *                                      calculateRes UL:702-704; consumers commented out UL:617-618); out[nn]
 *   WCFE_DORMANT_PRESSURE              calculatePressureForce UL:1765-1816 (bulge-test follower load, call commented out UL:577); out[3*nn]
 *   WCFE_DORMANT_SPHERE                calculateConstraintForce UL:1178-1194 (call commented out UL:551); out[3*nn]
 * Not built: calculateBendingRes UL:1818-1885, a one-sided finite difference (delta 1e-6) of an acos of unit normals -- on a flat
 * sheet its value is the rounding noise of acos near 1, there is nothing to pin. */
#define WCFE_DORMANT_RES_ADIF 0
#define WCFE_DORMANT_RES_CDIF 1
#define WCFE_DORMANT_PRESSURE 2
#define WCFE_DORMANT_SPHERE 3
int wcfe_dormant_field(wcfe_ctx *c, int which, double dt, double t_curr, double *out);
/* raw device pointer of a state field (SoA: component q of node i at ptr[q*pitch + i]); x, V, A, ea are double-buffered, so
 * ask again after every step.  Zero-copy interop with torch.distributed for the halo exchange of partitioned meshes. */
int wcfe_device_ptr(wcfe_ctx *c, int surface, int field, void **ptr, long long *pitch);
int wcfe_sync(wcfe_ctx *c);
/* Device self test: n pseudo-random operand pairs (zero numerators, both signs, x / 3 included) through the division-free correctly
 * rounded quotients the kernels use (wcfe_math.cuh recip_of / quot) and through the compiler's IEEE `/`; *mismatch = results that
 * differ in any bit (the sign of a zero quotient excepted).  Must be 0. */
int wcfe_selftest_quotient(int device, long long n, unsigned int seed, long long *mismatch);
/* moving substrate (wcfe_params::subs_scale != 0): how often the point-location grid has been rebuilt because a substrate node had
 * drifted beyond half the grid's padding from where the grid was built (checked after every step of wcfe_step_n) */
long long wcfe_grid_refreshes(wcfe_ctx *c);

/* ---- one mesh over several GPUs (BASELINE configs[3]: "element-partitioned over 2/4/8 B200 with ... halo exchange"; the reference
 * itself has no such mode).  Each rank's context holds ONE STRIP of the substrate -- its owned nodes plus a one-ring of
 * ghost nodes owned by neighbouring ranks, uploaded as an ordinary mesh -- and the whole cell.  The substrate is frozen
 * (main.cpp:299), so the only per-step data a neighbour needs is the new strain ea of the strip's border nodes.  After
 * wcfe_halo_connect the substrate kernel itself stores those rows into the neighbours' ghost rows through peer memory (NVLink)
 * and waits, in-kernel, for its own ghost rows: wcfe_step_n / _async then advance all strips with no host-side exchange, and the
 * owned nodes of all ranks equal the single-GPU run bit for bit.  All ranks must execute the same number of steps.
 *   wcfe_halo_info_get   what a neighbour needs to reach this context's strain buffers and arrival counters: raw device addresses
 *                        (same process) and CUDA IPC handles (other processes of the same node).  Ship the struct as bytes.
 *   wcfe_halo_connect    peers[q]: the info of neighbour q; my_slot_at_peer[q]: this rank's index in q's own peer list;
 *                        rows_from_peer[q]: ghost rows q sends here per step; export_*: for every (border node, neighbour) pair the
 *                        local node, the peer slot and the node's local index AT the peer; ghost_node: local ids of the ghosts.
 * No step may run between info_get and connect; a host barrier across ranks must follow connect before the first step.
 * Teardown is collective too: the border rows of the last step are peer stores that nobody has awaited, so before ANY rank calls
 * wcfe_halo_disconnect, wcfe_destroy or wcfe_upload_mesh(WCFE_SUBS) every rank must have returned from wcfe_sync and met the
 * others at a host barrier (a rank that unmaps its buffers earlier faults a slower neighbour still storing into them). */
#define WCFE_HALO_MAX_PEERS 8
typedef struct wcfe_halo_info {
    long long pid;                     /* owner process */
    int device, ea_cur;                /* CUDA ordinal; which of the two strain buffers is current */
    long long np;                      /* pitch (doubles) of the strain table: component q of node i at [q*np + i] */
    unsigned long long ptr[3];         /* device addresses in the owner: ea buffer 0, ea buffer 1, arrival counters */
    unsigned long long offset[3];      /* of each inside its cudaMalloc allocation (an IPC handle names the allocation) */
    unsigned char ipc[3][64];          /* cudaIpcMemHandle_t of the three allocations */
} wcfe_halo_info;
int wcfe_halo_info_get(wcfe_ctx *c, wcfe_halo_info *out);
int wcfe_halo_connect(wcfe_ctx *c, int n_peers, const wcfe_halo_info *peers, const int *my_slot_at_peer, const int *rows_from_peer,
                      int n_export, const int *export_node, const int *export_peer, const int *export_row, int n_ghost, const int *ghost_node);
int wcfe_halo_disconnect(wcfe_ctx *c);

/* ---- the CELL of the same partitioned run.  Every rank uploads the WHOLE cell (same numbering, hence the same 128-node patches as on
 * one GPU) but computes only a subset of the patches: those that can drop an anchor into its substrate strip; a patch on a cut is
 * computed by both neighbours, bit-identically.  What a rank does not compute goes stale in its arrays -- except the nodes its own
 * patches read as halo rows, which the rank that OWNS their patch stores, every step, into the same index of the reader's arrays
 * through peer memory (from the cell kernel; V and A of boundary vertices from the boundary-finish kernel).  The area / centroid
 * partial sums of every patch go from its owner into every rank's partials array, so that each rank forms the total area in the
 * single-GPU order (UL:1641-1667, consumed at UL:589, :1757).  Two epoch counters per ordered pair of ranks pace all of it, in
 * kernel; no host work per step.  Results equal the single-GPU run bit for bit on every row a rank computes.
 *   wcfe_cell_halo_info_get   addresses + CUDA IPC handles of this context's cell buffers (ship as bytes to every other rank)
 *   wcfe_cell_halo_connect    peers[q]: EVERY other rank; my_slot_at_peer[q]: this rank's index in q's peer list;
 *                             patch_list: ascending patch ids this rank computes (at least one), patch_owned[k]: 1 where this rank
 *                             publishes the patch's partial sums (every patch has exactly one owner across the ranks);
 *                             export_node / export_peer: (node, peer slot) pairs -- nodes of OWNED patches the peer reads as halo
 *                             rows; import_node: nodes this rank receives (halo rows of its patches that it does not compute);
 *                             anchor_margin: the distance |fa_u| the patch lists allow an anchor to be from its node (an anchor
 *                             beyond it raises WCFE_ERR_CAPACITY: it may lie under a strip whose rank does not compute the patch).
 * After the connect the context advances with wcfe_step_n / wcfe_step_n_async only, all ranks by the same number of steps.  The same
 * barriers as for the substrate apply: one after every rank's connect, one (after wcfe_sync) before any rank disconnects, destroys
 * or re-uploads. */
#define WCFE_CELL_HALO_NPTR 11
typedef struct wcfe_cell_halo_info {
    long long pid;
    int device, reserved;
    long long np;                                      /* pitch (doubles) of the cell's node arrays */
    unsigned long long ptr[WCFE_CELL_HALO_NPTR];       /* x, V, A, ea as written by the next step and by the one after (8), partials of even /
                                                          odd steps (2), counters (1) */
    unsigned long long offset[WCFE_CELL_HALO_NPTR];
    unsigned char ipc[WCFE_CELL_HALO_NPTR][64];
} wcfe_cell_halo_info;
int wcfe_cell_halo_info_get(wcfe_ctx *c, wcfe_cell_halo_info *out);
int wcfe_cell_halo_connect(wcfe_ctx *c, int n_peers, const wcfe_cell_halo_info *peers, const int *my_slot_at_peer,
                           int n_patches, const int *patch_list, const int *patch_owned,
                           int n_export, const int *export_node, const int *export_peer, int n_import, const int *import_node,
                           double anchor_margin);
int wcfe_cell_halo_disconnect(wcfe_ctx *c);

/* Host-side helper (no GPU work, usable from any host language): cut ONE mesh into `world` vertical strips and produce, for
 * `rank`, the local strip to upload and the tables wcfe_halo_connect takes.  Nodes are sorted by x (ties by id) and cut into runs
 * of equal node count, or -- with one weight per node -- of equal total weight (nodes under the cell cost more: see
 * wholecellfe_b200/partition.py cell_weights).  A rank holds every element with at least one owned vertex; local nodes and
 * elements keep ascending global ids, so nodal sums are formed in the single-GPU order.  Call it with the same arguments on every
 * rank.  Sizes first, then fetch the arrays into caller-owned buffers (any pointer may be NULL):
 *   mesh: local_nodes[n_local] global id of every local node; tri_local[3*n_elem]; elem_gid[n_elem]; owned_local[n_owned];
 *         ghost_local[n_ghost] (= ghost_node of wcfe_halo_connect)
 *   halo: peer_rank[n_peers] ascending = the peer-slot order; my_slot_at_peer, rows_from_peer [n_peers]; export_* [n_export] */
typedef struct wcfe_strip wcfe_strip;
int wcfe_strip_partition(int nn, const double *xyz, int ne, const int *tri, const double *weights, int world, int rank, wcfe_strip **out);
int wcfe_strip_sizes(const wcfe_strip *s, int *n_local, int *n_elem, int *n_owned, int *n_ghost, int *n_peers, int *n_export);
int wcfe_strip_mesh(const wcfe_strip *s, int *local_nodes, int *tri_local, int *elem_gid, int *owned_local, int *ghost_local);
int wcfe_strip_halo(const wcfe_strip *s, int *peer_rank, int *my_slot_at_peer, int *rows_from_peer, int *export_node, int *export_peer, int *export_row);
int wcfe_strip_free(wcfe_strip *s);
/* The same cut along any per-node key instead of x (ties by node id): key = atan2(y - cy, x - cx) about the cell's centroid gives the
 * sectors of partition.py sector_partition -- every rank then holds a wedge of the substrate AND of the cell above it (world - 1 <=
 * WCFE_HALO_MAX_PEERS: all sectors meet at the centre).  Same handle and accessors as wcfe_strip_partition. */
int wcfe_key_partition(int nn, const double *key, int ne, const int *tri, const double *weights, int world, int rank, wcfe_strip **out);
/* Host-only: renumber a mesh for the patch kernels (128 consecutive nodes per CTA, one-ring staged in shared memory: without locality
 * in the numbering wcfe_upload_mesh falls back to the general kernels).  Nodes along a Morton (Z-order) curve over the 16-bit quantised
 * (x, y), ties by old id; triangles by the key of their centroid, vertex order (orientation) kept.  perm[new] = old id; apply it to
 * every nodal array before the upload and its inverse to what comes back.  What meshgen.py does for the synthetic configurations. */
int wcfe_morton_order(int nn, const double *xyz, int ne, const int *tri, int *perm, int *tri_out);
/* weights[nn_subs] for a cost-balanced cut: 1 + under_cell for a substrate node inside the cell's padded bounding box (measured ~3x a
 * plain node: under_cell 2) + cell_cost per cell node above it when the cell is partitioned as well (measured ~4; 0 = replicated cell).
 * partition.py cell_weights + cell_cost_weights. */
int wcfe_partition_weights(int nn_subs, const double *subs_xyz, int nn_cell, const double *cell_xyz, double under_cell, double cell_cost, double *weights);
/* owner[nn]: the rank that owns every node of the uncut mesh (the strips' node -> rank map) */
int wcfe_strip_owner(const wcfe_strip *s, int *owner);

/* Host-only as well: the arguments of wcfe_cell_halo_connect for `rank` -- which 128-node cell patches it computes (every patch that
 * can drop an anchor, at most `margin` from its node, into one of the rank's substrate triangles), which of them it owns (the lowest
 * rank owning substrate under the patch's centroid), the rows it pushes to and receives from the other ranks -- for ANY partition of
 * the substrate given as subs_owner[nn_subs] (strips: wcfe_strip_owner).  world <= 32.  Same lists as wholecellfe_b200/partition.py
 * cell_partition + cell_connect_plan.
 *   sizes: n_peers (= world - 1), n_patches, n_export, n_import
 *   lists: peer_rank, my_slot_at_peer [n_peers]; patch_list, patch_owned [n_patches]; export_node, export_peer [n_export];
 *          import_node [n_import] */
typedef struct wcfe_cell_plan wcfe_cell_plan;
int wcfe_cell_partition(int world, int rank, int nn_subs, const double *subs_xyz, const int *subs_owner,
                        int nn_cell, const double *cell_xyz, int ne_cell, const int *cell_tri, double margin, wcfe_cell_plan **out);
int wcfe_cell_plan_sizes(const wcfe_cell_plan *p, int *n_peers, int *n_patches, int *n_export, int *n_import);
int wcfe_cell_plan_lists(const wcfe_cell_plan *p, int *peer_rank, int *my_slot_at_peer, int *patch_list, int *patch_owned, int *export_node,
                         int *export_peer, int *import_node);
int wcfe_cell_plan_free(wcfe_cell_plan *p);

#ifdef __cplusplus
}
#endif
#endif
