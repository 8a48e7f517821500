"""Substrate strip partition + halo exchange (BASELINE configs[3] multi-GPU mode).
CPU: partition invariants and the exchange itself over gloo with two ranks.
GPU: two ranks (NCCL on two GPUs when present, else both ranks on cuda:0 over gloo) reproduce the single-GPU run bit for bit."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from wholecellfe_b200.meshgen import square_mesh, disk_mesh
from wholecellfe_b200.partition import (strip_partition, sector_partition, make_partition, cell_partition, cell_connect_plan, exchange_rows,
                                        connect_plan, neighbours, cell_weights, cell_cost_weights, native_strip, native_cell_plan, PATCH)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("balanced", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_partition_invariants(world, balanced):
    xs, ts = square_mesh(48)
    w = None
    if balanced:                                                            # cost weights: nodes under a (small, off-centre) cell count 3x
        xc = np.array([[-20.0, -5.0, 0.0], [0.0, 12.0, 0.0], [5.0, -8.0, 0.0]])
        w = cell_weights(xs, xc)
        assert w.min() == 1.0 and w.max() == 3.0
    parts = strip_partition(xs, ts, world, w)
    if balanced and world > 1:
        tot = [w[p.owned].sum() for p in parts]
        assert max(tot) - min(tot) <= 0.1 * np.mean(tot) + 7 * 48          # equal weight up to one column of nodes
    owned = np.sort(np.concatenate([p.owned for p in parts]))
    assert np.array_equal(owned, np.arange(len(xs)))                       # every node owned exactly once
    covered = np.zeros(len(ts), dtype=int)
    for p in parts:
        covered[p.elem_gid] += 1
        assert np.array_equal(p.local_nodes[p.tri_local], ts[p.elem_gid])  # local connectivity maps back
        assert np.all(np.diff(p.elem_gid) > 0) and np.all(np.diff(p.local_nodes) > 0)
        ring = np.unique(ts[p.elem_gid])
        assert np.array_equal(np.sort(np.concatenate([p.owned, p.ghosts])), ring)
        for q, ids in p.recv.items():                                       # what r receives from q is what q sends to r
            assert np.array_equal(p.local_nodes[ids], parts[q].local_nodes[parts[q].send[p.rank]])
    assert covered.min() >= 1
    # an owned node's complete one-ring is local
    inc = [set() for _ in range(len(xs))]
    for e, t in enumerate(ts):
        for v in t:
            inc[v].add(e)
    for p in parts:
        have = set(p.elem_gid.tolist())
        for v in p.owned[:: max(1, len(p.owned) // 50)]:
            assert inc[v] <= have


@pytest.mark.parametrize("shape", ["strips", "sectors"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_cell_partition_invariants(world, shape):
    """The partitioned cell's plan (host logic, no GPU): every patch has exactly one owner, who computes it; every halo row of a
    computed patch is either computed locally or imported; every import is exported by the owner of its patch, to the right peer
    slot; sectors tile the substrate like strips do; anchors within the margin of a computed patch's nodes can only fall into
    triangles local to that rank (checked on random points)."""
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    parts = make_partition(xs, ts, world, xc, True, "partitioned", shape)
    owned = np.sort(np.concatenate([p.owned for p in parts]))
    assert np.array_equal(owned, np.arange(len(xs)))
    margin = 0.15
    cps = cell_partition(parts, xc, tc, xs, margin)
    npatch = (len(xc) + PATCH - 1) // PATCH
    own = np.zeros(npatch, dtype=int)
    for c in cps:
        own[c.owned_patches] += 1
        assert set(c.owned_patches.tolist()) <= set(c.patch_list.tolist()) and len(c.patch_list) >= 1
        assert np.all(np.diff(c.patch_list) > 0)
    assert np.all(own == 1)
    pid = np.arange(len(xc)) // PATCH
    owner = np.zeros(npatch, dtype=int)
    for r, c in enumerate(cps):
        owner[c.owned_patches] = r
    for r, c in enumerate(cps):
        listed = np.zeros(npatch, dtype=bool); listed[c.patch_list] = True
        halo = set()
        for t in tc[np.any(listed[pid[tc]], axis=1)]:
            ps = pid[t]
            for v, pv in zip(t, ps):
                if not listed[pv] and np.any(listed[ps]):
                    halo.add(int(v))
        assert halo == set(c.imports.tolist())                               # exactly the rows its patches read and it does not compute
        plan = cell_connect_plan(cps, r)
        peers, slot = plan[0], plan[1]
        assert peers == [q for q in range(world) if q != r]
        for k, q in enumerate(peers):
            assert cell_connect_plan(cps, q)[0][slot[k]] == r
        for q in range(world):                                                # what r imports from q's patches, q exports to r
            if q == r:
                continue
            need = c.imports[owner[pid[c.imports]] == q]
            got = cps[q].exports.get(r, np.zeros(0, int))
            assert np.array_equal(np.sort(need), np.sort(got))
    # anchors: a point within the margin of a node of a patch NOT computed by rank r never lies in a triangle local to r
    rng = np.random.default_rng(3)
    tri_rank = [set(p.elem_gid.tolist()) for p in parts]
    cen = xs[ts].mean(axis=1)[:, :2]
    for _ in range(300):
        n = int(rng.integers(len(xc)))
        ang = rng.uniform(0, 2 * np.pi); pt = xc[n, :2] + margin * rng.uniform(0, 1) * np.array([np.cos(ang), np.sin(ang)])
        near = np.argsort(((cen - pt) ** 2).sum(axis=1))[:12]                 # the containing triangle is among the nearest centroids
        for e in near:
            a, b, c3 = xs[ts[e], :2]
            d = (b[0] - a[0]) * (c3[1] - a[1]) - (b[1] - a[1]) * (c3[0] - a[0])
            w1 = ((b[0] - pt[0]) * (c3[1] - pt[1]) - (b[1] - pt[1]) * (c3[0] - pt[0])) / d
            w2 = ((c3[0] - pt[0]) * (a[1] - pt[1]) - (c3[1] - pt[1]) * (a[0] - pt[0])) / d
            if w1 >= 0 and w2 >= 0 and 1 - w1 - w2 >= 0:
                for r in range(world):
                    if int(e) in tri_rank[r]:
                        assert pid[n] in set(cps[r].patch_list.tolist()), (r, n, e)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_connect_plan_is_consistent_across_ranks(world):
    """what wcfe_halo_connect is told on one rank agrees with what its neighbours are told: every exported (node, row at the
    peer) names the same global node on both sides, the row counts match, the slots are each other's inverse"""
    xs, ts = square_mesh(40)
    parts = strip_partition(xs, ts, world)
    plans = [connect_plan(parts, r) for r in range(world)]
    for r, (nb, slot, rows, en, ep, er, gh) in enumerate(plans):
        assert nb == neighbours(parts[r]) and len(nb) <= 8
        assert np.array_equal(parts[r].local_nodes[gh], parts[r].ghosts)
        assert not np.intersect1d(en, gh).size                               # a ghost is never exported
        for k, q in enumerate(nb):
            assert plans[q][0][slot[k]] == r                                 # my slot at q really is me
            sel = ep == k
            assert np.array_equal(parts[r].local_nodes[en[sel]], parts[q].local_nodes[er[sel]])
            back = plans[q][0].index(r)
            assert plans[q][2][back] == int(sel.sum())                       # q expects exactly the rows I push
        # every ghost row is pushed by exactly one neighbour
        got = np.concatenate([plans[q][5][plans[q][4] == plans[q][0].index(r)] for q in nb]) if nb else np.zeros(0, int)
        assert np.array_equal(np.sort(got), np.sort(gh))


@pytest.mark.parametrize("balanced", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_c_abi_strip_partition_equals_python(world, balanced):
    """wcfe_strip_partition (host-only C++ in libwcfe.so: what a C++ / MPI host calls) produces the same strips and the same
    wcfe_halo_connect tables as partition.py, on a structured and on an unstructured-numbered mesh."""
    from wholecellfe_b200.build import build
    build()
    for xs, ts in (square_mesh(37), disk_mesh(1500)):
        w = cell_weights(xs, np.array([[-20.0, -5.0, 0.0], [0.0, 12.0, 0.0], [5.0, -8.0, 0.0]])) if balanced else None
        parts = strip_partition(xs, ts, world, w)
        for r in range(world):
            nat = native_strip(xs, ts, world, r, w)
            p = parts[r]
            assert np.array_equal(nat["local_nodes"], p.local_nodes) and np.array_equal(nat["elem_gid"], p.elem_gid)
            assert np.array_equal(nat["tri_local"], p.tri_local)
            assert np.array_equal(nat["owned_local"], p.owned_local)
            nb, slot, rows, en, ep, er, gh = connect_plan(parts, r)
            assert np.array_equal(nat["ghost_local"], gh)
            assert list(nat["peer_rank"]) == nb and list(nat["my_slot_at_peer"]) == slot and list(nat["rows_from_peer"]) == rows
            assert np.array_equal(nat["export_node"], en) and np.array_equal(nat["export_peer"], ep) and np.array_equal(nat["export_row"], er)


def test_c_abi_partition_weights_equal_python():
    """wcfe_partition_weights = cell_weights (+ cell_cost_weights when the cell is partitioned too), value for value"""
    import ctypes as C
    from wholecellfe_b200 import _lib as B
    from wholecellfe_b200.build import build
    build()
    L = B.load()
    for (nc, ns, shift) in ((6000, 120, (0.0, 0.0)), (3001, 64, (17.3, -9.1)), (500, 33, (-60.0, 0.0))):       # the last cell hangs over the substrate's edge
        xc, _ = disk_mesh(nc); xs, _ = square_mesh(ns)
        xc = np.ascontiguousarray(xc + np.array([shift[0], shift[1], 0.0])); xs = np.ascontiguousarray(xs)
        for cost in (0.0, 4.0):
            w = np.zeros(len(xs))
            assert L.wcfe_partition_weights(len(xs), xs.ctypes.data_as(B._dp), len(xc), xc.ctypes.data_as(B._dp), 2.0, cost, w.ctypes.data_as(B._dp)) == 0
            want = cell_weights(xs, xc) + (cell_cost_weights(xs, xc, cost) if cost else 0.0)
            assert np.array_equal(w, want), (nc, ns, cost)
            assert w.min() >= 1.0 and (cost == 0.0 or w.sum() > len(xs) + 2.0 * 0)
    assert L.wcfe_partition_weights(0, None, 1, None, 2.0, 0.0, None) == B.ERR_ARG


@pytest.mark.parametrize("world", [2, 3, 8])
def test_c_abi_key_partition_gives_the_sectors_of_python(world):
    """wcfe_key_partition cut along the angle about a centre = sector_partition + connect_plan of partition.py (every rank may have
    every other as a neighbour); along x it is wcfe_strip_partition."""
    from wholecellfe_b200.build import build
    build()
    xs, ts = square_mesh(41)
    centre = (3.7, -2.2)
    ang = np.arctan2(xs[:, 1] - centre[1], xs[:, 0] - centre[0])
    w = cell_weights(xs, np.array([[-20.0, -5.0, 0.0], [0.0, 12.0, 0.0], [5.0, -8.0, 0.0]]))
    parts = sector_partition(xs, ts, world, centre, w)
    for r in range(world):
        nat = native_strip(xs, ts, world, r, w, key=ang)
        p = parts[r]
        assert np.array_equal(nat["local_nodes"], p.local_nodes) and np.array_equal(nat["elem_gid"], p.elem_gid) and np.array_equal(nat["tri_local"], p.tri_local)
        nb, slot, rows, en, ep, er, gh = connect_plan(parts, r)
        assert np.array_equal(nat["ghost_local"], gh) and list(nat["peer_rank"]) == nb and list(nat["my_slot_at_peer"]) == slot and list(nat["rows_from_peer"]) == rows
        assert np.array_equal(nat["export_node"], en) and np.array_equal(nat["export_peer"], ep) and np.array_equal(nat["export_row"], er)
        a = native_strip(xs, ts, world, r, w); b = native_strip(xs, ts, world, r, w, key=xs[:, 0])
        assert all(np.array_equal(a[k], b[k]) for k in a)


@pytest.mark.parametrize("shape", ["strips", "sectors"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_c_abi_cell_partition_equals_python(world, shape):
    """wcfe_cell_partition (host-only C++ in libwcfe.so) returns, list for list, what partition.py cell_partition +
    cell_connect_plan compute -- for strips (owner map from wcfe_strip_owner) and for sectors (owner map handed in), on a cell that
    covers part of the substrate (ranks far from it get their nearest patch), with and without cost balancing."""
    import ctypes as C
    from wholecellfe_b200 import _lib as B
    from wholecellfe_b200.build import build
    build()
    for (nc, ns, shift, margin, balance) in ((6000, 120, (0.0, 0.0), 0.15, True), (3001, 64, (17.3, -9.1), 0.4, False)):
        xc, tc = disk_mesh(nc); xs, ts = square_mesh(ns)
        xc = xc + np.array([shift[0], shift[1], 0.0])
        parts = make_partition(xs, ts, world, xc, balance, "partitioned", shape)
        owner = np.zeros(len(xs), dtype=np.int32)
        for r, p in enumerate(parts):
            owner[p.owned] = r
        if shape == "strips" and not balance:          # the strips' own node -> rank map out of the C ABI
            L = B.load()
            x64 = np.ascontiguousarray(xs, dtype=np.float64); t32 = np.ascontiguousarray(ts, dtype=np.int32)
            h = C.c_void_p()
            assert L.wcfe_strip_partition(len(xs), x64.ctypes.data_as(B._dp), len(ts), t32.ctypes.data_as(B._ip), None, world, 0, C.byref(h)) == 0
            nat_owner = np.zeros(len(xs), dtype=np.int32)
            assert L.wcfe_strip_owner(h, nat_owner.ctypes.data_as(B._ip)) == 0 and L.wcfe_strip_free(h) == 0
            assert np.array_equal(nat_owner, owner)
        cps = cell_partition(parts, xc, tc, xs, margin)
        for r in range(world):
            peers, slot, pl, po, en, ep, imp = cell_connect_plan(cps, r)
            nat = native_cell_plan(xs, owner, xc, tc, world, r, margin)
            assert list(nat["peer_rank"]) == peers and list(nat["my_slot_at_peer"]) == slot, (world, shape, r)
            assert np.array_equal(nat["patch_list"], pl) and np.array_equal(nat["patch_owned"], po), (world, shape, r)
            assert np.array_equal(nat["export_node"], en) and np.array_equal(nat["export_peer"], ep), (world, shape, r)
            assert np.array_equal(nat["import_node"], imp), (world, shape, r)
    # argument checks
    with pytest.raises(ValueError):
        native_cell_plan(xs, owner, xc, tc, world, world, 0.15)                 # rank outside [0, world)
    with pytest.raises(ValueError):
        native_cell_plan(xs, owner + world, xc, tc, world, 0, 0.15)             # an owner outside [0, world)
    with pytest.raises(ValueError):
        native_cell_plan(xs, owner, xc, tc, 33, 0, 0.15)                        # rank sets are 32-bit masks


@pytest.mark.parametrize("shape", ["strips", "sectors"])
def test_cpp_host_partition_plan_equals_python(shape, tmp_path):
    """wholecellfe_b200/csrc/host/partition_plan: a C++ host driving the host-only partition entry points of the C ABI (weights, cut,
    strip, cell plan) end to end, no GPU, no Python in the loop -- its per-rank sizes are those of partition.py make_partition +
    cell_partition on the same meshes."""
    import re
    import subprocess
    from wholecellfe_b200.meshio import write_ecomsol
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "wholecellfe_b200", "csrc", "host")
    subprocess.check_call(["make", "-s", "-C", host])
    xc, tc = disk_mesh(4000); xs, ts = square_mesh(90)
    write_ecomsol(str(tmp_path / "cell.txt"), xc, tc); write_ecomsol(str(tmp_path / "subs.txt"), xs, ts)
    world = 4
    r = subprocess.run([os.path.join(host, "partition_plan"), str(tmp_path / "cell.txt"), str(tmp_path / "subs.txt"), str(world), shape, "0.15"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [[int(v) for v in re.findall(r"-?\d+", line)] for line in r.stdout.strip().split("\n")]
    assert len(rows) == world
    parts = make_partition(xs, ts, world, xc, True, "partitioned", shape)
    cps = cell_partition(parts, xc, tc, xs, 0.15)
    for q, (p, c) in enumerate(zip(parts, cps)):
        _, _, rows_from, en, _, _, gh = connect_plan(parts, q)
        want = [q, len(p.local_nodes), len(p.owned), len(gh), len(p.elem_gid), len(neighbours(p)), len(en),
                len(c.patch_list), int(c.patch_owned.sum()), sum(len(v) for v in c.exports.values()), len(c.imports)]
        assert rows[q] == want, (shape, q, rows[q], want)
    bad = subprocess.run([os.path.join(host, "partition_plan"), str(tmp_path / "nope.txt"), str(tmp_path / "subs.txt"), "2"], capture_output=True, text=True)
    assert bad.returncode != 0


def test_c_abi_strip_partition_rejects_bad_arguments():
    import ctypes as C
    from wholecellfe_b200 import _lib as B
    L = B.load()
    xs, ts = square_mesh(5)
    xs = np.ascontiguousarray(xs); ts = np.ascontiguousarray(ts, dtype=np.int32)
    h = C.c_void_p()
    args = (len(xs), xs.ctypes.data_as(B._dp), len(ts), ts.ctypes.data_as(B._ip), None)
    assert L.wcfe_strip_partition(*args, 2, 2, C.byref(h)) == B.ERR_ARG            # rank outside [0, world)
    assert L.wcfe_strip_partition(*args, 0, 0, C.byref(h)) == B.ERR_ARG
    bad = ts.copy(); bad[0, 0] = len(xs)
    assert L.wcfe_strip_partition(len(xs), xs.ctypes.data_as(B._dp), len(bad), bad.ctypes.data_as(B._ip), None, 2, 0, C.byref(h)) == B.ERR_ARG
    assert L.wcfe_strip_sizes(None, None, None, None, None, None, None) == B.ERR_ARG and L.wcfe_strip_free(None) == B.ERR_ARG


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xs, ts = square_mesh(30)
    part = strip_partition(xs, ts, world)[rank]
    # table[c, local node] = f(c, global id) on owned rows, garbage on ghost rows
    table = torch.full((6, len(part.local_nodes) + 3), -777.0, dtype=torch.float64)
    gid = torch.as_tensor(part.local_nodes)
    for c in range(6):
        table[c, part.owned_local] = (1000.0 * c + gid[part.owned_local]).double()
    exchange_rows(dist, table, part, [0, 1, 3], torch.device("cpu"))
    ok = True
    for c in (0, 1, 3):
        ok &= bool(torch.equal(table[c, : len(gid)], (1000.0 * c + gid).double()))
    for c in (2, 4, 5):                                                    # untouched components keep the garbage on ghosts
        ghost_local = np.setdiff1d(np.arange(len(gid)), part.owned_local)
        ok &= bool((table[c, ghost_local] == -777.0).all())
    q.put((rank, ok))
    dist.barrier(); dist.destroy_process_group()


def test_halo_exchange_over_gloo_two_ranks():
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    assert all(ok for _, ok in res)


def _initial_strain(nn):
    # a frozen substrate that starts strain-free stays exactly strain-free (e_s == averaged e for F == I), so give it a
    # non-trivial planar strain field: the exchange then carries data that matters
    rng = np.random.default_rng(11)
    ea = np.zeros((nn, 6))
    ea[:, [0, 1, 3]] = 1e-3 * rng.normal(size=(nn, 3))
    return ea


def _gpu_worker(rank, world, port, backend, halo, q, cell="replicated"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from wholecellfe_b200.partition import PartitionedSim
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    sim = PartitionedSim("k31", (xc, tc), (xs, ts), dist, dev, rng=(1, 5), halo=halo, cell=cell)
    sim.set_subs_field("ea", _initial_strain(len(xs)))
    sim.step_n(12, 2e-4)
    out = {}
    for f in ("V", "A", "ea"):
        ids, vals = sim.owned_field(f)
        out[f] = (ids, vals)
    out["cell_x"] = sim.sim.get(0, "x"); out["cell_cC"] = sim.sim.get(0, "cC")
    for f in ("x", "V", "ea", "cC", "fa_u"):
        out["cell_owned_" + f] = sim.owned_cell_field(f)
    q.put((rank, out))
    sim.close()
    dist.barrier(); dist.destroy_process_group()


def _single_gpu_reference(steps=12):
    import wholecellfe_b200 as W
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 5)
    g.set(W.SUBS, "ea", _initial_strain(len(xs)))
    g.step_n(steps, 2e-4, 0.0)
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_in_process_strips_with_fused_peer_halo_equal_single_gpu_bitwise(world):
    """the fused halo exchange (border strains stored into the neighbours' ghost rows by k_subs_ring itself, arrival counters
    awaited in-kernel) on strips that live in one process: the owned nodes of all strips tile the single-GPU result bit for bit"""
    import wholecellfe_b200 as W
    from wholecellfe_b200.partition import InProcessStrips
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    ndev = torch.cuda.device_count()
    run = InProcessStrips("k31", (xc, tc), (xs, ts), [r % ndev for r in range(world)], rng=(1, 5))
    run.set_subs_field("ea", _initial_strain(len(xs)))
    run.step_n(5, 2e-4); run.step_n(7, 2e-4)
    g = _single_gpu_reference()
    for f in ("V", "A", "ea"):
        ref = g.get(W.SUBS, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = run.owned_field(r, f)
            got[ids] = vals
        assert np.array_equal(got, ref), f
    for sim in run.sims:
        assert np.array_equal(sim.get(W.CELL, "x"), g.get(W.CELL, "x"))
    assert np.abs(g.get(W.SUBS, "ea")).max() > 1e-5 and np.abs(g.get(W.SUBS, "A")).max() > 0
    run.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,shape", [(2, "strips"), (3, "strips"), (5, "strips"), (3, "sectors"), (8, "sectors")])
def test_partitioned_cell_in_process_equals_single_gpu_bitwise(world, shape):
    """BASELINE configs[3] with the CELL partitioned as well: every strip computes only the cell patches that can reach its
    substrate strip, the rows its patches read from elsewhere are pushed by their owners through peer memory, every rank adds
    the same area / centroid partial sums in the single-GPU order.  The cell rows every rank OWNS tile the single-GPU cell bit
    for bit (positions, velocities, strains, bond state), and so do the substrate rows; 12 steps of k31, Philox."""
    import wholecellfe_b200 as W
    from wholecellfe_b200.partition import InProcessStrips
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    ndev = torch.cuda.device_count()
    run = InProcessStrips("k31", (xc, tc), (xs, ts), [r % ndev for r in range(world)], rng=(1, 5), cell="partitioned", shape=shape)
    assert sum(len(c.owned_patches) for c in run.cparts) == (len(xc) + 127) // 128
    assert min(len(c.patch_list) for c in run.cparts) < (len(xc) + 127) // 128          # somebody really computes less than the whole cell
    assert max(len(neighbours(p)) for p in run.parts) <= 8
    run.set_subs_field("ea", _initial_strain(len(xs)))
    run.step_n(5, 2e-4); run.step_n(7, 2e-4)
    g = _single_gpu_reference()
    for f in ("x", "V", "A", "ea", "cC", "fa_u"):
        ref = g.get(W.CELL, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = run.owned_cell_field(r, f)
            got[ids] = vals
        assert np.array_equal(got, ref), (f, int((got != ref).sum()))
    for f in ("V", "A", "ea"):
        ref = g.get(W.SUBS, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = run.owned_field(r, f)
            got[ids] = vals
        assert np.array_equal(got, ref), f
    d0 = run.sims[0].diagnostics()
    assert np.array_equal(d0[:4], g.diagnostics()[:4])               # total area and centroid: the same sum of the same partials
    run.close()


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["peer", "dist"])
def test_two_rank_partition_equals_single_gpu_bitwise(halo):
    import wholecellfe_b200 as W
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, backend, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
    g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 5)
    g.set(W.SUBS, "ea", _initial_strain(len(xs)))
    g.step_n(12, 2e-4, 0.0)
    for f in ("V", "A", "ea"):
        ref = g.get(W.SUBS, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = res[r][f]
            got[ids] = vals
        assert np.array_equal(got, ref), f                   # owned nodes of all ranks tile the substrate, bit for bit
    for r in range(world):                                    # the replicated cell evolves identically everywhere
        assert np.array_equal(res[r]["cell_x"], g.get(W.CELL, "x")) and np.array_equal(res[r]["cell_cC"], g.get(W.CELL, "cC"))
    assert np.abs(g.get(W.SUBS, "ea")).max() > 1e-5 and np.abs(g.get(W.SUBS, "A")).max() > 0


@pytest.mark.gpu
def test_two_rank_partitioned_cell_across_processes_equals_single_gpu_bitwise():
    """the partitioned cell between two PROCESSES (CUDA IPC mappings of the peer's cell buffers; NVLink where the box has two
    GPUs, one GPU shared otherwise): owned cell rows and owned substrate rows tile the single-GPU run bit for bit"""
    import wholecellfe_b200 as W
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, backend, "peer", q, "partitioned")) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    g = _single_gpu_reference()
    for f in ("x", "V", "ea", "cC", "fa_u"):
        ref = g.get(W.CELL, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = res[r]["cell_owned_" + f]
            got[ids] = vals
        assert np.array_equal(got, ref), f
    for f in ("V", "A", "ea"):
        ref = g.get(W.SUBS, f)
        got = np.full_like(ref, np.nan)
        for r in range(world):
            ids, vals = res[r][f]
            got[ids] = vals
        assert np.array_equal(got, ref), f


@pytest.mark.gpu
def test_halo_connect_argument_checks_and_bounded_wait():
    """wcfe_halo_connect refuses inconsistent tables, and a neighbour that never delivers makes the step fail with WCFE_ERR_PEER
    after the bounded in-kernel wait (10 s) instead of hanging the GPU."""
    import time
    import wholecellfe_b200 as W
    from wholecellfe_b200 import _lib as B
    from wholecellfe_b200.ulstep import WcfeError
    xc, tc = disk_mesh(600); xs, ts = square_mesh(40)
    parts = strip_partition(xs, ts, 2)
    part = parts[0]
    g = W.ULSim("k1"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs[part.local_nodes], part.tri_local); g.set_rng(1, 3)
    nb, slot, rows, en, ep, er, gh = connect_plan(parts, 0)
    with pytest.raises(WcfeError) as e:                       # no wcfe_halo_info_get before
        g.halo_connect([bytes(B.HaloInfo())], [0], [1], en, ep, er, gh)
    assert e.value.code == B.ERR_ARG
    info = g.halo_info()
    with pytest.raises(WcfeError) as e:                       # a ghost cannot be exported
        g.halo_connect([info], [0], rows, gh[:1], [0], [0], gh)
    assert e.value.code == B.ERR_ARG
    info = g.halo_info()
    with pytest.raises(WcfeError) as e:                       # more neighbours than WCFE_HALO_MAX_PEERS
        g.halo_connect([info] * 9, [0] * 9, [1] * 9, en, ep, er, gh)
    assert e.value.code == B.ERR_ARG
    # the strip wired to ITSELF as its neighbour, but told to expect one row more than is ever pushed
    info = g.halo_info()
    er_self = np.minimum(er, len(part.local_nodes) - 1)
    g.halo_connect([info], [0], [len(en) + 1], en, ep, er_self, gh)
    t = g.step_n(1, 1e-4, 0.0)                                # first step after the connect waits for nothing
    t0 = time.time()
    with pytest.raises(WcfeError) as e:
        g.step_n(2, 1e-4, t)
    assert e.value.code == B.ERR_PEER and "halo" in str(e.value)
    assert 5.0 < time.time() - t0 < 60.0                      # the bounded wait, once: later waits return at once
    g.close()


def test_c_abi_morton_order_equals_meshgen_and_keeps_the_mesh():
    """wcfe_morton_order (host-only C++ in libwcfe.so) numbers nodes and sorts triangles exactly as meshgen.py does for the synthetic
    configurations; the renumbered mesh is the same mesh (coordinates of every triangle's vertices, orientation) and its patches of
    128 consecutive nodes are compact."""
    import ctypes as C
    from wholecellfe_b200 import _lib as B
    from wholecellfe_b200.build import build
    from wholecellfe_b200.meshgen import _renumber
    build()
    L = B.load()
    rng = np.random.default_rng(5)
    for xyz, tri in (disk_mesh(5000, order=None), square_mesh(57, order=None, jitter=0.2)):
        shuffle = rng.permutation(len(xyz))                      # a numbering without any locality
        inv = np.empty_like(shuffle); inv[shuffle] = np.arange(len(xyz))
        xyz = np.ascontiguousarray(xyz[shuffle]); tri = np.ascontiguousarray(inv[tri][rng.permutation(len(tri))], dtype=np.int32)
        perm = np.zeros(len(xyz), dtype=np.int32); tout = np.zeros_like(tri)
        assert L.wcfe_morton_order(len(xyz), xyz.ctypes.data_as(B._dp), len(tri), tri.ctypes.data_as(B._ip), perm.ctypes.data_as(B._ip), tout.ctypes.data_as(B._ip)) == 0
        want_xyz, want_tri = _renumber(xyz, tri, "morton")
        assert np.array_equal(xyz[perm], want_xyz) and np.array_equal(tout, want_tri)
        assert sorted(perm.tolist()) == list(range(len(xyz)))
        a = np.sort(xyz[perm][tout].reshape(len(tri), -1), axis=0); b = np.sort(xyz[tri].reshape(len(tri), -1), axis=0)
        assert np.array_equal(a, b)                               # the same set of triangles, vertex order kept
        x2 = xyz[perm]
        ext = lambda pts: max(np.ptp(pts[:, 0]), np.ptp(pts[:, 1]))
        before = np.median([ext(xyz[k:k + 128]) for k in range(0, len(xyz) - 128, 128)])
        after = np.median([ext(x2[k:k + 128]) for k in range(0, len(xyz) - 128, 128)])
        assert after < 0.5 * before, (before, after)
    assert L.wcfe_morton_order(0, None, 0, None, None, None) == B.ERR_ARG
