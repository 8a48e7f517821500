"""Element-partitioned substrate with a per-step halo exchange (BASELINE configs[3]; SURVEY.md 8e, second mode).

Round-1 scope: the SUBSTRATE (8 of the 9 million elements of config 4, and frozen in space, main.cpp:299) is cut into
vertical strips, one per rank; the cell (the remaining ninth) is replicated on every rank, where it evolves identically
because every kernel is deterministic -- so the only data that has to cross between GPUs each step is the strain `ea` of the
substrate nodes on the strip borders (3 doubles per node; positions never change, V and A are only read by their owner).
Ghost nodes are recomputed with incomplete one-rings and overwritten by the exchange before anything reads them.

Two ways to move the border strains:
  halo="peer" (default)  fused into the substrate kernel: the thread that owns a border node stores its new strain straight into
                         the neighbours' ghost rows through peer memory (NVLink; CUDA IPC mappings between the ranks' processes)
                         and bumps an arrival counter; the next step's border CTAs wait for the counters in-kernel.  No host work
                         per step, the programmatic-dependent-launch chain of the step stays intact (wcfe_halo_connect).
  halo="dist"            a host-driven exchange after every step through torch.distributed (NCCL on device tensors that alias
                         libwcfe's own buffers; gloo on CPU tensors in the tests).  Kept as the baseline the fused path is measured
                         against, and as the only path between nodes.
Host logic here: strip partition, local meshes, send/receive index lists, the connect plan.
"""
import numpy as np


class StripPart:
    """What one rank holds of the partitioned substrate."""

    def __init__(self, rank, world, owned, ghosts, local_nodes, tri_local, elem_gid, send, recv):
        self.rank, self.world = rank, world
        self.owned = owned                  # global ids of owned nodes (ascending)
        self.ghosts = ghosts                # global ids of ghost nodes (ascending)
        self.local_nodes = local_nodes      # global id of every local node, ascending (owned and ghosts interleaved)
        self.tri_local = tri_local          # [m,3] local vertex ids, elements in ascending global element id
        self.elem_gid = elem_gid            # global element ids of the local elements
        self.send = send                    # {neighbour rank: local ids of owned nodes that are ghosts there}
        self.recv = recv                    # {neighbour rank: local ids of ghost nodes owned there}
        self.owned_local = np.searchsorted(local_nodes, owned)


def cell_weights(subs_xyz, cell_xyz, under_cell=2.0):
    """Per-node cost weights for strip_partition: a substrate node inside the cell's bounding box (padded like libwcfe's anchor
    region) also reads its epoch stamp and, if an incident triangle holds anchors, walks their buckets; measured on B200 the nodes
    under the cell cost ~3x a plain node on average (profiles/r01f_ncu_summary.txt, tools/strip_probe.py)."""
    xs = np.asarray(subs_xyz); xc = np.asarray(cell_xyz)
    lo, hi = xc[:, :2].min(axis=0), xc[:, :2].max(axis=0)
    pad = max(2.0, 0.25 * float((hi - lo).max()))
    inside = np.all((xs[:, :2] >= lo - pad) & (xs[:, :2] <= hi + pad), axis=1)
    return 1.0 + under_cell * inside


def _cut_by_key(key, world, weights):
    """owner rank per node: nodes sorted by `key` (ties by id) and cut into `world` runs of equal count / equal total weight"""
    nn = len(key)
    order = np.lexsort((np.arange(nn), key))
    owner = np.empty(nn, dtype=np.int32)
    if weights is None:
        bounds = [(nn * r) // world for r in range(world + 1)]
    else:
        cw = np.cumsum(np.asarray(weights, dtype=np.float64)[order])
        bounds = [0] + [int(np.searchsorted(cw, cw[-1] * r / world)) for r in range(1, world)] + [nn]
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def partition_from_owner(xyz, tri, owner, world):
    """The StripPart of every rank from a node -> owner-rank map: a rank holds every element with at least one owned vertex, hence
    a one-ring of ghost nodes.  Local numbering keeps the global (space-filling-curve) order, and local elements keep ascending
    global ids, so every nodal sum is formed in the same order as on a single GPU."""
    tri = np.asarray(tri)
    tri_owner = owner[tri]                                   # [m,3]
    parts = []
    for r in range(world):
        emask = (tri_owner == r).any(axis=1)
        elem_gid = np.nonzero(emask)[0]
        local_nodes = np.unique(tri[elem_gid])
        is_owned = owner[local_nodes] == r
        owned = local_nodes[is_owned]
        ghosts = local_nodes[~is_owned]
        # nodes owned by r that no element touches cannot exist in a valid mesh; owned nodes all appear in local_nodes
        tri_local = np.searchsorted(local_nodes, tri[elem_gid]).astype(np.int32)
        parts.append([r, owned, ghosts, local_nodes, tri_local, elem_gid])
    out = []
    for r in range(world):
        _, owned, ghosts, local_nodes, tri_local, elem_gid = parts[r]
        recv = {}
        for q in np.unique(owner[ghosts]):
            g = ghosts[owner[ghosts] == q]
            recv[int(q)] = np.searchsorted(local_nodes, g)
        send = {}
        for q in range(world):
            if q == r:
                continue
            gq = parts[q][2]
            mine = gq[owner[gq] == r]                        # ghosts of q that r owns, in q's ascending order
            if len(mine):
                send[q] = np.searchsorted(local_nodes, mine)
        out.append(StripPart(r, world, owned, ghosts, local_nodes, tri_local, elem_gid, send, recv))
    return out


def strip_partition(xyz, tri, world, weights=None):
    """Cut the mesh into `world` vertical strips of (almost) equal node count -- or, with `weights` (one per node), of equal
    total weight, so that the strips under the cell hold fewer nodes.  Returns the list of StripPart.
    A node belongs to the strip its x coordinate falls in (ties by node id)."""
    xyz = np.asarray(xyz)
    return partition_from_owner(xyz, tri, _cut_by_key(xyz[:, 0], world, weights), world)


def sector_partition(xyz, tri, world, centre, weights=None):
    """Cut the mesh into `world` sectors around `centre` (the cell's centroid) of equal node count / equal total weight: every rank
    then holds a wedge of the substrate AND the wedge of the cell above it, so that a partitioned cell's work spreads over all
    ranks instead of the few whose vertical strip happens to pass under the cell.  All sectors meet at the centre: a rank may have
    every other rank as a neighbour (world - 1 <= WCFE_HALO_MAX_PEERS)."""
    xyz = np.asarray(xyz)
    ang = np.arctan2(xyz[:, 1] - centre[1], xyz[:, 0] - centre[0])
    return partition_from_owner(xyz, tri, _cut_by_key(ang, world, weights), world)


PATCH = 128      # libwcfe's owned nodes per patch (wcfe_patch.cuh kPatch)


class CellPart:
    """What one rank computes of the partitioned cell (replicated storage, partitioned patches: include/wcfe.h)."""

    def __init__(self, rank, world, patch_list, patch_owned, imports, exports):
        self.rank, self.world = rank, world
        self.patch_list = patch_list        # ascending patch ids this rank computes
        self.patch_owned = patch_owned      # 1 where it publishes the patch's partial sums / pushes its rows
        self.imports = imports              # node ids it receives (halo rows of its patches that it does not compute)
        self.exports = exports              # {peer rank: node ids of owned patches the peer reads}
        own = patch_list[patch_owned.astype(bool)]
        self.owned_patches = own

    def owned_nodes(self, nn):
        """node ids of the patches this rank owns (its authoritative rows)"""
        if len(self.owned_patches) == 0:
            return np.zeros(0, dtype=np.int64)
        ids = (self.owned_patches[:, None] * PATCH + np.arange(PATCH)[None, :]).ravel()
        return ids[ids < nn]


def _owner_raster(parts, subs_xyz, cell_size):
    """bit mask of the ranks that OWN a substrate node in every cell of a uniform raster over the substrate, and one node per cell"""
    xs = np.asarray(subs_xyz)
    x0, y0 = xs[:, 0].min(), xs[:, 1].min()
    nx = int((xs[:, 0].max() - x0) / cell_size) + 1; ny = int((xs[:, 1].max() - y0) / cell_size) + 1
    mask = np.zeros((ny, nx), dtype=np.uint32)
    rep = np.full((ny, nx), -1, dtype=np.int64)
    for r, p in enumerate(parts):
        ix = ((xs[p.owned, 0] - x0) / cell_size).astype(np.int64); iy = ((xs[p.owned, 1] - y0) / cell_size).astype(np.int64)
        np.bitwise_or.at(mask, (iy, ix), np.uint32(1 << r))
        rep[iy, ix] = p.owned
    return mask, rep, (x0, y0, nx, ny)


def _dilate(mask):
    out = mask.copy()
    ny, nx = mask.shape
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            ys = slice(max(0, dy), ny + min(0, dy)); yd = slice(max(0, -dy), ny + min(0, -dy))
            xs_ = slice(max(0, dx), nx + min(0, dx)); xd = slice(max(0, -dx), nx + min(0, -dx))
            out[yd, xd] |= mask[ys, xs_]
    return out


def cell_partition(parts, cell_xyz, cell_tri, subs_xyz, margin=0.15):
    """Which cell patches every rank computes, owns, pushes and receives -- for ANY partition of the substrate (strips, sectors).

    A rank computes every patch that can drop an anchor into one of its local substrate triangles.  An anchor lies within `margin`
    of its node (libwcfe raises an error when one leaves it), a local triangle has an owned vertex, i.e. reaches at most one edge
    length beyond the owned nodes: the ranks that own a substrate node within margin + 1.5 h of a cell node are found on a raster of
    that cell size, dilated by one cell (conservative).  A patch on a cut is therefore computed by every rank it touches; the
    rank that owns the substrate under the patch's centroid owns it: it alone publishes the patch's area / centroid partial sums
    and pushes its rows to the ranks that read them as halo rows of their own patches without computing them."""
    xyz = np.asarray(cell_xyz); tri = np.asarray(cell_tri); xs = np.asarray(subs_xyz)
    world = len(parts)
    if world > 32:
        raise ValueError("cell_partition: at most 32 ranks (rank sets are 32-bit masks)")
    nn = len(xyz); npatch = (nn + PATCH - 1) // PATCH
    h = np.sqrt((xs[:, 0].max() - xs[:, 0].min()) * (xs[:, 1].max() - xs[:, 1].min()) / max(1, len(xs)))
    g = margin + 1.5 * h
    mask, rep, (x0, y0, nx, ny) = _owner_raster(parts, xs, g)
    dil = _dilate(mask)
    ix = np.clip(((xyz[:, 0] - x0) / g).astype(np.int64), 0, nx - 1); iy = np.clip(((xyz[:, 1] - y0) / g).astype(np.int64), 0, ny - 1)
    starts = np.arange(0, nn, PATCH)
    pbits = np.bitwise_or.reduceat(dil[iy, ix], starts)
    cx = np.add.reduceat(xyz[:, 0], starts) / np.diff(np.append(starts, nn)); cy = np.add.reduceat(xyz[:, 1], starts) / np.diff(np.append(starts, nn))
    cix = np.clip(((cx - x0) / g).astype(np.int64), 0, nx - 1); ciy = np.clip(((cy - y0) / g).astype(np.int64), 0, ny - 1)
    cbits = mask[ciy, cix]
    cbits = np.where(cbits != 0, cbits, dil[ciy, cix])
    cbits = np.where(cbits != 0, cbits, np.uint32(1))
    owner = np.zeros(npatch, dtype=np.int32)
    for r in range(world - 1, -1, -1):                                     # lowest rank among the owners of the centroid's raster cell
        owner[(cbits >> np.uint32(r)) & np.uint32(1) == 1] = r
    listed = np.zeros((world, npatch), dtype=bool)
    for r in range(world):
        listed[r] = ((pbits >> np.uint32(r)) & np.uint32(1)) == 1
        listed[r, owner == r] = True
        if not listed[r].any():              # a rank far from the cell still runs the step's kernels: give it the nearest patch
            px = xs[parts[r].owned, 0].mean(); py = xs[parts[r].owned, 1].mean()
            listed[r, int(np.argmin((cx - px) ** 2 + (cy - py) ** 2))] = True
    pid = np.arange(nn) // PATCH
    a = tri[:, [0, 0, 1, 1, 2, 2]].ravel(); b = tri[:, [1, 2, 0, 2, 0, 1]].ravel()
    pa, pb = pid[a], pid[b]
    cross = pa != pb                                                       # b is a halo row of patch pa
    pa, b, pb = pa[cross], b[cross], pb[cross]
    imports = []
    for r in range(world):
        sel = listed[r][pa] & ~listed[r][pb]
        imports.append(np.unique(b[sel]))
    out = []
    for r in range(world):
        exports = {}
        for q in range(world):
            if q == r or len(imports[q]) == 0:
                continue
            mine = imports[q][owner[pid[imports[q]]] == r]
            if len(mine):
                exports[q] = mine
        pl = np.nonzero(listed[r])[0].astype(np.int32)
        out.append(CellPart(r, world, pl, (owner[pl] == r).astype(np.int32), imports[r].astype(np.int32), exports))
    return out


def cell_connect_plan(cparts, rank):
    """Arguments of wcfe_cell_halo_connect for `rank`: (peer ranks, my_slot_at_peer, patch_list, patch_owned, export_node,
    export_peer, import_node).  Every other rank is a peer (the partial sums go to all of them)."""
    me = cparts[rank]
    peers = [q for q in range(me.world) if q != rank]
    my_slot = [rank if rank < q else rank - 1 for q in peers]
    en, ep = [], []
    for slot, q in enumerate(peers):
        mine = me.exports.get(q)
        if mine is not None:
            en.append(mine); ep.append(np.full(len(mine), slot))
    cat = lambda v: np.concatenate(v).astype(np.int32) if v else np.zeros(0, np.int32)
    return peers, my_slot, me.patch_list, me.patch_owned, cat(en), cat(ep), me.imports


def cell_cost_weights(subs_xyz, cell_xyz, cell_cost=4.0):
    """Extra per-substrate-node weights for a partition whose cell is partitioned too: the cost of every cell node is charged to a
    substrate node under it (a cell node costs ~4 substrate nodes: 0.088 ms / 0.5 M against 0.185 ms / 4 M on B200)."""
    xs = np.asarray(subs_xyz); xc = np.asarray(cell_xyz)
    h = np.sqrt((xs[:, 0].max() - xs[:, 0].min()) * (xs[:, 1].max() - xs[:, 1].min()) / max(1, len(xs)))
    g = 1.5 * h
    x0, y0 = xs[:, 0].min(), xs[:, 1].min()
    nx = int((xs[:, 0].max() - x0) / g) + 1; ny = int((xs[:, 1].max() - y0) / g) + 1
    rep = np.full((ny, nx), -1, dtype=np.int64)
    rep[((xs[:, 1] - y0) / g).astype(np.int64), ((xs[:, 0] - x0) / g).astype(np.int64)] = np.arange(len(xs))
    ix = np.clip(((xc[:, 0] - x0) / g).astype(np.int64), 0, nx - 1); iy = np.clip(((xc[:, 1] - y0) / g).astype(np.int64), 0, ny - 1)
    tgt = rep[iy, ix]
    w = np.zeros(len(xs))
    np.add.at(w, tgt[tgt >= 0], cell_cost)
    return w


def neighbours(part):
    """ranks this strip exchanges rows with, ascending: the peer-slot order of wcfe_halo_connect"""
    return sorted(set(part.send) | set(part.recv))


def connect_plan(parts, rank):
    """Arguments of wcfe_halo_connect for `rank`, from the partition every rank computes identically: (neighbour ranks,
    my_slot_at_peer, rows_from_peer, export_node, export_peer, export_row, ghost_node)."""
    me = parts[rank]
    nb = neighbours(me)
    my_slot = [neighbours(parts[q]).index(rank) for q in nb]
    rows_from = [len(me.recv.get(q, ())) for q in nb]
    en, ep, er = [], [], []
    for slot, q in enumerate(nb):
        mine = me.send.get(q)
        if mine is None:
            continue
        theirs = parts[q].recv[rank]                        # the same nodes, as q numbers them (same order: test_partition_invariants)
        en.append(mine); ep.append(np.full(len(mine), slot)); er.append(theirs)
    cat = lambda v: np.concatenate(v).astype(np.int32) if v else np.zeros(0, np.int32)
    ghost_local = np.searchsorted(me.local_nodes, me.ghosts).astype(np.int32)
    return nb, my_slot, rows_from, cat(en), cat(ep), cat(er), ghost_local


def native_strip(xyz, tri, world, rank, weights=None, key=None):
    """The same partition and connect plan from the C ABI's host-side helper (wcfe_strip_partition, csrc/wcfe_partition.cu):
    what a C++ / MPI host would call.  key: cut along this per-node key instead of x (wcfe_key_partition; the angle about the cell's
    centroid gives sector_partition).  Returns a dict of numpy arrays named like the C arguments."""
    import ctypes as C
    from . import _lib as B
    L = B.load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); tri = np.ascontiguousarray(tri, dtype=np.int32)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    h = C.c_void_p()
    if key is None:
        rc = L.wcfe_strip_partition(len(xyz), xyz.ctypes.data_as(B._dp), len(tri), tri.ctypes.data_as(B._ip),
                                    None if w is None else w.ctypes.data_as(B._dp), world, rank, C.byref(h))
    else:
        key = np.ascontiguousarray(key, dtype=np.float64)
        rc = L.wcfe_key_partition(len(xyz), key.ctypes.data_as(B._dp), len(tri), tri.ctypes.data_as(B._ip),
                                  None if w is None else w.ctypes.data_as(B._dp), world, rank, C.byref(h))
    if rc:
        raise ValueError("wcfe_strip_partition failed (%d)" % rc)
    n = [C.c_int() for _ in range(6)]
    L.wcfe_strip_sizes(h, *[C.byref(v) for v in n])
    n_local, n_elem, n_owned, n_ghost, n_peers, n_export = [v.value for v in n]
    ia = lambda k: np.zeros(k, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(B._ip)
    out = {"local_nodes": ia(n_local), "tri_local": ia(3 * n_elem), "elem_gid": ia(n_elem), "owned_local": ia(n_owned), "ghost_local": ia(n_ghost),
           "peer_rank": ia(n_peers), "my_slot_at_peer": ia(n_peers), "rows_from_peer": ia(n_peers),
           "export_node": ia(n_export), "export_peer": ia(n_export), "export_row": ia(n_export)}
    L.wcfe_strip_mesh(h, ip(out["local_nodes"]), ip(out["tri_local"]), ip(out["elem_gid"]), ip(out["owned_local"]), ip(out["ghost_local"]))
    L.wcfe_strip_halo(h, ip(out["peer_rank"]), ip(out["my_slot_at_peer"]), ip(out["rows_from_peer"]), ip(out["export_node"]), ip(out["export_peer"]), ip(out["export_row"]))
    L.wcfe_strip_free(h)
    out["tri_local"] = out["tri_local"].reshape(-1, 3)
    return out


def native_cell_plan(subs_xyz, subs_owner, cell_xyz, cell_tri, world, rank, margin=0.15):
    """cell_partition + cell_connect_plan from the C ABI's host-side helper (wcfe_cell_partition, csrc/wcfe_partition.cu): what a C++ /
    MPI host would call.  subs_owner: rank owning every substrate node.  Returns a dict of numpy arrays named like the C arguments."""
    import ctypes as C
    from . import _lib as B
    L = B.load()
    xs = np.ascontiguousarray(subs_xyz, dtype=np.float64); ow = np.ascontiguousarray(subs_owner, dtype=np.int32)
    xc = np.ascontiguousarray(cell_xyz, dtype=np.float64); tc = np.ascontiguousarray(cell_tri, dtype=np.int32)
    h = C.c_void_p()
    ip = lambda a: a.ctypes.data_as(B._ip)
    rc = L.wcfe_cell_partition(world, rank, len(xs), xs.ctypes.data_as(B._dp), ip(ow), len(xc), xc.ctypes.data_as(B._dp), len(tc), ip(tc), margin, C.byref(h))
    if rc:
        raise ValueError("wcfe_cell_partition failed (%d)" % rc)
    n = [C.c_int() for _ in range(4)]
    L.wcfe_cell_plan_sizes(h, *[C.byref(v) for v in n])
    n_peers, n_patches, n_export, n_import = [v.value for v in n]
    ia = lambda k: np.zeros(k, dtype=np.int32)
    out = {"peer_rank": ia(n_peers), "my_slot_at_peer": ia(n_peers), "patch_list": ia(n_patches), "patch_owned": ia(n_patches),
           "export_node": ia(n_export), "export_peer": ia(n_export), "import_node": ia(n_import)}
    L.wcfe_cell_plan_lists(h, *[ip(out[k]) for k in ("peer_rank", "my_slot_at_peer", "patch_list", "patch_owned", "export_node", "export_peer", "import_node")])
    L.wcfe_cell_plan_free(h)
    return out


def make_partition(xs, ts, world, cell_xyz, balance=True, cell="replicated", shape="strips", cell_cost=4.0):
    """The substrate partition of a run: pieces of equal COST -- nodes under the cell also gather integrin forces, and with the cell
    partitioned too they carry the cost of the cell nodes above them (balance=False: equal node counts) -- as vertical strips, or
    as sectors around the cell's centroid (every rank then gets its share of the cell as well)."""
    w = None
    if balance:
        w = cell_weights(xs, cell_xyz)
        if cell == "partitioned":
            w = w + cell_cost_weights(xs, cell_xyz, cell_cost)
    if shape == "sectors":
        return sector_partition(xs, ts, world, np.asarray(cell_xyz)[:, :2].mean(axis=0), w)
    if shape != "strips":
        raise ValueError("shape must be 'strips' or 'sectors'")
    return strip_partition(xs, ts, world, w)


class InProcessStrips:
    """All strips of a partitioned run driven by ONE process: one context per strip on the given devices (several strips may share
    a GPU), connected through plain device pointers (peer access between GPUs).  Strips advance in lockstep, one step queued on
    every strip before the next: a strip's step s+1 waits in-kernel for its neighbours' step s, which is already in flight."""

    def __init__(self, variant, cell_mesh, subs_mesh, devices, rng=(1, 1), params=None, balance=True, cell="replicated", margin=0.15, shape="strips"):
        from . import ulstep
        from ._lib import CELL, SUBS
        xs, ts = subs_mesh
        self.parts = make_partition(xs, ts, len(devices), cell_mesh[0], balance, cell, shape)
        self.cparts = cell_partition(self.parts, cell_mesh[0], cell_mesh[1], xs, margin) if cell == "partitioned" else None
        self.ncell = len(cell_mesh[0])
        self.margin = margin
        self.CELL = CELL
        self.sims = []
        for part, dev in zip(self.parts, devices):
            sim = ulstep.ULSim(variant, device=dev, params=params)
            sim.upload_mesh(CELL, cell_mesh[0], cell_mesh[1])
            sim.upload_mesh(SUBS, xs[part.local_nodes], part.tri_local)
            sim.set_rng(*rng)
            self.sims.append(sim)
        self.SUBS = SUBS
        self.t = 0.0
        self.connected = False

    def set_subs_field(self, field, global_values):
        for part, sim in zip(self.parts, self.sims):
            sim.set(self.SUBS, field, np.asarray(global_values)[part.local_nodes])

    def connect(self):
        infos = [sim.halo_info() for sim in self.sims]
        for r, sim in enumerate(self.sims):
            nb, slot, rows, en, ep, er, gh = connect_plan(self.parts, r)
            sim.halo_connect([infos[q] for q in nb], slot, rows, en, ep, er, gh)
        if self.cparts is not None:
            cinfos = [sim.cell_halo_info() for sim in self.sims]
            for r, sim in enumerate(self.sims):
                peers, slot, pl, po, en, ep, im = cell_connect_plan(self.cparts, r)
                sim.cell_halo_connect([cinfos[q] for q in peers], slot, pl, po, en, ep, im, self.margin)
        self.connected = True

    def owned_cell_field(self, r, field):
        """(node ids, values) of the cell rows rank r owns (every row when the cell is replicated)"""
        ids = self.cparts[r].owned_nodes(self.ncell) if self.cparts is not None else np.arange(self.ncell)
        return ids, self.sims[r].get(self.CELL, field)[ids]

    def step_n(self, n, dt):
        if not self.connected:
            self.connect()
        t = self.t
        for _ in range(n):
            for sim in self.sims:
                t1 = sim.step_n_async(1, dt, t)
            t = t1
        for sim in self.sims:
            sim.sync()
        self.t = t

    def owned_field(self, r, field):
        part = self.parts[r]
        return part.owned, self.sims[r].get(self.SUBS, field)[part.owned_local]

    def close(self):
        for sim in self.sims:
            sim.close()


class _CudaView:
    """__cuda_array_interface__ holder: lets torch alias a libwcfe device buffer without copying."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def exchange_rows(dist, table, part, comps, device):
    """Refresh the ghost rows of `table` ([ncomp, pitch] tensor, node index last) from their owners.  Pairwise send/recv with
    every neighbour, batched (NCCL groups them into one launch).  With the gloo backend (CPU tests, or several ranks sharing
    one GPU) the packed rows are staged through host memory, since gloo cannot send device tensors."""
    import torch
    staged = dist.get_backend() == "gloo" and table.is_cuda
    comm = torch.device("cpu") if staged else device
    cache = part.__dict__.setdefault("_idx_cache", {})
    key = (str(device), tuple(comps))
    if key not in cache:                                    # index tensors live on the device for the whole run
        cache[key] = (torch.as_tensor(comps, device=device),
                      {q: torch.as_tensor(ids, device=device) for q, ids in part.send.items()},
                      {q: torch.as_tensor(ids, device=device) for q, ids in part.recv.items()})
    comps_t, send_idx, recv_idx = cache[key]
    sub = table.index_select(0, comps_t)                    # [len(comps), pitch]
    ops, keep = [], []
    for q in sorted(send_idx):
        buf = sub.index_select(1, send_idx[q]).contiguous().to(comm)
        ops.append(dist.P2POp(dist.isend, buf, q))
        keep.append(buf)
    rbufs = []
    for q in sorted(recv_idx):
        buf = torch.empty((len(comps), len(recv_idx[q])), dtype=table.dtype, device=comm)
        ops.append(dist.P2POp(dist.irecv, buf, q))
        rbufs.append((q, buf))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for q, buf in rbufs:
        buf = buf.to(device)
        for k, cmp in enumerate(comps):
            table[cmp].index_copy_(0, recv_idx[q], buf[k])


class PartitionedSim:
    """One rank of a substrate-partitioned run: full cell + this rank's substrate strip on one GPU."""

    def __init__(self, variant, cell_mesh, subs_mesh, dist, device_index, rng=(1, 1), params=None, halo="peer", balance=True,
                 cell="replicated", margin=0.15, shape="strips", cell_cost=4.0):
        import torch
        from . import ulstep
        from ._lib import CELL, SUBS
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device("cuda", device_index)
        self.halo = halo
        self.connected = False
        xs, ts = subs_mesh
        if cell == "partitioned" and halo != "peer":
            raise ValueError("a partitioned cell needs the fused peer-memory exchange (halo='peer')")
        self.parts = make_partition(xs, ts, self.world, cell_mesh[0], balance, cell, shape, cell_cost)
        self.part = self.parts[self.rank]
        self.cparts = cell_partition(self.parts, cell_mesh[0], cell_mesh[1], xs, margin) if cell == "partitioned" else None
        self.cpart = self.cparts[self.rank] if self.cparts is not None else None
        self.margin = margin
        self.ncell = len(cell_mesh[0])
        self.CELL = CELL
        self.sim = ulstep.ULSim(variant, device=device_index, params=params, stream=torch.cuda.current_stream().cuda_stream)
        self.sim.upload_mesh(CELL, cell_mesh[0], cell_mesh[1])
        self.sim.upload_mesh(SUBS, xs[self.part.local_nodes], self.part.tri_local)
        self.sim.set_rng(*rng)
        self.SUBS = SUBS
        self.n_owned_elements = int((np.isin(ts, self.part.owned).any(axis=1)).sum())   # == local elements
        self.t = 0.0

    def _ea_table(self):
        ptr, pitch = self.sim.device_ptr(self.SUBS, "ea")
        return self.torch.as_tensor(_CudaView(ptr, (6, pitch)), device=self.device)

    def connect(self):
        """halo="peer": exchange the wcfe_halo_info of every rank (addresses + CUDA IPC handles) and wire this strip to its
        neighbours.  Collective; called by the first step_n, after the last state upload."""
        infos = [None] * self.world
        self.dist.all_gather_object(infos, self.sim.halo_info())
        nb, slot, rows, en, ep, er, gh = connect_plan(self.parts, self.rank)
        self.sim.halo_connect([infos[q] for q in nb], slot, rows, en, ep, er, gh)
        if self.cparts is not None:
            cinfos = [None] * self.world
            self.dist.all_gather_object(cinfos, self.sim.cell_halo_info())
            peers, slot, pl, po, en, ep, im = cell_connect_plan(self.cparts, self.rank)
            self.sim.cell_halo_connect([cinfos[q] for q in peers], slot, pl, po, en, ep, im, self.margin)
        self.dist.barrier()                 # nobody pushes before everybody has zeroed counters and mapped buffers
        self.connected = True

    def step(self, dt):
        """one UL step on every rank, then the halo exchange of the substrate strain (xx, yy, xy: the planar components)"""
        self.t = self.sim.step_n_async(1, dt, self.t)
        exchange_rows(self.dist, self._ea_table(), self.part, [0, 1, 3], self.device)

    def step_n(self, n, dt):
        if self.halo == "peer":
            if not self.connected:
                self.connect()
            self.t = self.sim.step_n_async(n, dt, self.t)      # the exchange is inside the substrate kernel
        else:
            for _ in range(n):
                self.step(dt)
        self.sim.sync()

    def set_subs_field(self, field, global_values):
        """upload a substrate field given for the WHOLE substrate ([Nn_global, w]); every rank keeps its local rows"""
        self.sim.set(self.SUBS, field, np.asarray(global_values)[self.part.local_nodes])

    def owned_field(self, field):
        """(global node ids, values) of this rank's owned substrate nodes"""
        return self.part.owned, self.sim.get(self.SUBS, field)[self.part.owned_local]

    def owned_cell_field(self, field):
        """(node ids, values) of the cell rows this rank owns (every row when the cell is replicated)"""
        ids = self.cpart.owned_nodes(self.ncell) if self.cpart is not None else np.arange(self.ncell)
        return ids, self.sim.get(self.CELL, field)[ids]

    def close(self):
        """Collective.  The border rows of the LAST step are plain peer stores nobody has awaited yet: a rank that unmapped or
        freed its strain buffers while a slower neighbour's final substrate kernel was still storing into them would fault the
        neighbour.  Drain the own stream, meet everybody, only then disconnect and free (include/wcfe.h, wcfe_halo_disconnect)."""
        if self.sim is None:
            return
        self.sim.sync()
        self.dist.barrier()
        self.sim.close()
        self.sim = None
