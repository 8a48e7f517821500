"""Host logic of the N > 1 path (ensembles of independent cells, no data-path collective) on CPU with gloo, world 2."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from wholecellfe_b200.ensemble import sweep_grid, partition, max_over_ranks, gather_counts


def test_partition_is_a_disjoint_balanced_cover():
    for n in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            shards = [partition(n, world, r) for r in range(world)]
            ids = [i for s in shards for i in s]
            assert ids == list(range(n))
            sizes = [len(s) for s in shards]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition(10, 2, 2)


def test_sweep_grid_is_deterministic_and_covers_laws():
    a = sweep_grid(4096); b = sweep_grid(4096)
    assert a == b and len(a) == 4096
    assert {c["variant"] for c in a} == {"k1", "k31", "kMD"}
    rhos = sorted({c["rho_i_max"] for c in a})
    assert abs(rhos[0] - 1.0) < 1e-12 and abs(rhos[-1] - 100.0) < 1e-9
    assert len({(c["variant"], c["rho_i_max"], c["seed"]) for c in a}) == 4096


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cells, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cells = sweep_grid(n_cells)
    mine = partition(n_cells, world, rank)
    # every rank "runs" its shard: here the work is a checksum of the parameters it owns
    local = sum(int(c["seed"]) for c in cells[mine.start:mine.stop])
    total_cells = gather_counts(len(mine), dist)
    fake_seconds = 1.0 + rank                           # rank 1 is slower
    t = max_over_ranks(fake_seconds, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.start, mine.stop, local))
    q.put((rank, total_cells, t, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_sharding_and_timing_reduction():
    world, n_cells = 2, 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cells, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cells = sweep_grid(n_cells)
    for rank, total, t, gathered in out:
        assert total == n_cells
        assert t == 2.0                                  # max over ranks
        spans = sorted((a, b) for a, b, _ in gathered)
        assert spans[0][0] == 0 and spans[-1][1] == n_cells and spans[0][1] == spans[1][0]
        assert sum(s for _, _, s in gathered) == sum(c["seed"] for c in cells)


def test_c_abi_sweep_and_shard_equal_python():
    """wcfe_sweep_cell / wcfe_ensemble_shard (host-only C++ in libwcfe.so: what a C++ / MPI host calls for config 5) against
    ensemble.py sweep_grid / partition: law, seed and shard bounds exactly, the log-spaced ligand density to a few ulps (numpy's
    geomspace and pow round differently), every other parameter the law's defaults."""
    import ctypes as C
    import wholecellfe_b200 as W
    from wholecellfe_b200 import _lib as B
    from wholecellfe_b200.build import build
    build()
    L = B.load()
    laws = (C.c_int * 3)(B.VARIANTS["k1"], B.VARIANTS["k31"], B.VARIANTS["kMD"])
    for n in (1, 7, 512, 4096, 4097):
        cells = sweep_grid(n)
        for c in cells[:: max(1, n // 97)] + cells[-3:]:
            p = B.Params(); seed = C.c_uint()
            assert L.wcfe_sweep_cell(n, 3, laws, 1.0, 100.0, c["id"], C.byref(p), C.byref(seed)) == 0
            d = W.default_params(c["variant"])
            assert seed.value == c["seed"] and p.law == d.law and p.kint_force == d.kint_force and p.kint_bond == d.kint_bond, (n, c)
            assert abs(p.rho_i_max - c["rho_i_max"]) <= 1e-14 * c["rho_i_max"], (n, c, p.rho_i_max)
            for f, _ in B.Params._fields_:
                if f not in ("rho_i_max", "md_px", "md_py", "md_k"):
                    assert getattr(p, f) == getattr(d, f), f
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo = C.c_int(); hi = C.c_int()
                assert L.wcfe_ensemble_shard(n, world, r, C.byref(lo), C.byref(hi)) == 0
                want = partition(n, world, r)
                assert (lo.value, hi.value) == (want.start, want.stop)
    p = B.Params()
    assert L.wcfe_sweep_cell(10, 3, laws, 1.0, 100.0, 10, C.byref(p), None) == B.ERR_ARG
    assert L.wcfe_ensemble_shard(10, 2, 2, C.byref(C.c_int()), C.byref(C.c_int())) == B.ERR_ARG
