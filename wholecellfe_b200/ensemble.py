"""Ensembles of independent cells (BASELINE.json configs[4]): the path shards by whole cells, with no
data-path communication -- the reference's own ensemble mechanism is a loop of independent processes
(s_sim_031021_1.sub).  This module holds the host logic: the parameter sweep (integrin stiffness law x ligand density
x RNG seed), its partition over ranks, and the max-over-ranks timing reduction used by bench.py.
"""
import itertools
import numpy as np


def sweep_grid(n_cells, laws=("k1", "k31", "kMD"), rho_min=1.0, rho_max=100.0):
    """n_cells (variant, rho_i_max, seed) triples: laws x log-spaced ligand densities (rho_i_max, UL:1231) x seeds.

    Deterministic: cell id -> parameters, so every rank can rebuild the table without communication.  The default range
    stops at the reference's own rho_i_max = 100: with the example scripts' dt = 0.005 the explicit step is unstable for
    denser, i.e. stiffer, adhesion (omega^2 ~ kint rho c_C)."""
    n_rho = max(1, int(round((n_cells / len(laws)) ** 0.5)))
    rhos = np.geomspace(rho_min, rho_max, n_rho)
    base = list(itertools.product(laws, rhos))
    cells = []
    seed = 1
    while len(cells) < n_cells:
        for law, rho in base:
            if len(cells) == n_cells:
                break
            cells.append({"id": len(cells), "variant": law, "rho_i_max": float(rho), "seed": seed})
        seed += 1
    return cells


def partition(n_cells, world, rank):
    """Contiguous, balanced shard of cell ids for `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_cells, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return range(lo, hi)


def max_over_ranks(seconds, dist=None, device=None):
    """The time a multi-rank job took is the slowest rank's (bench contract); dist = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, dist=None, device=None):
    """Units processed by all ranks (whole-job throughput numerator)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(count)
    import torch
    t = torch.tensor([count], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


class EnsembleRunner:
    """This rank's shard of an ensemble.  Cells never communicate.

    mode "batch" (default): the cells share one batched device context per group -- every kernel of the step is launched once
    with gridDim.y = number of cells (wcfe_create_batch), so that thousands of config-1-size problems fill the GPU instead of
    queueing as tiny launches.  With the counter-based generator each cell keeps its own seed inside one group; with the glibc
    replay tape the cells are grouped by seed (one tape per context).
    mode "streams": one context and CUDA stream per cell, stepped asynchronously (the pre-batching path, kept for comparison)."""

    def __init__(self, cells, cell_mesh, subs_mesh, device=0, rng_mode=1, mode="batch"):
        from . import ulstep
        from ._lib import CELL, SUBS, RNG_PHILOX
        self.cells = list(cells)
        self.mode = mode
        self.sims = []          # contexts
        self.where = []         # cell -> (context index, replica index)
        xc, tc = cell_mesh
        xs, ts = subs_mesh

        def params_of(c):
            p = ulstep.default_params(c["variant"])
            p.rho_i_max = c["rho_i_max"]                      # ligand density (UL:1231)
            return p

        if mode == "streams":
            for c in self.cells:
                sim = ulstep.ULSim(c["variant"], device=device, params=params_of(c))
                sim.upload_mesh(CELL, xc, tc)
                sim.upload_mesh(SUBS, xs, ts)
                sim.set_rng(rng_mode, c["seed"])
                self.where.append((len(self.sims), 0))
                self.sims.append(sim)
        elif mode == "batch":
            groups = {}
            for i, c in enumerate(self.cells):
                groups.setdefault(0 if rng_mode == RNG_PHILOX else c["seed"], []).append(i)
            self.where = [None] * len(self.cells)
            for key in sorted(groups):
                ids = groups[key]
                sim = ulstep.ULSim(self.cells[ids[0]]["variant"], device=device, batch=[params_of(self.cells[i]) for i in ids])
                sim.upload_mesh(CELL, xc, tc)
                sim.upload_mesh(SUBS, xs, ts)
                sim.set_rng(rng_mode, self.cells[ids[0]]["seed"])
                if rng_mode == RNG_PHILOX and len(ids) > 1:
                    for r, i in enumerate(ids):
                        sim.set_replica_seed(r, self.cells[i]["seed"])
                for r, i in enumerate(ids):
                    self.where[i] = (len(self.sims), r)
                self.sims.append(sim)
        else:
            raise ValueError("mode must be 'batch' or 'streams'")
        self.t = [0.0] * len(self.sims)
        self.elements_per_cell = len(tc) + len(ts)

    def advance(self, n_steps, dt, chunk=8):
        """n_steps for every cell, interleaved in chunks (one frame of the example scripts = 8 steps)."""
        done = 0
        while done < n_steps:
            m = min(chunk, n_steps - done)
            for i, sim in enumerate(self.sims):
                self.t[i] = sim.step_n_async(m, dt, self.t[i])
            done += m
        for sim in self.sims:
            sim.sync()

    def get(self, cell_index, surface, field):
        """State of one cell of the shard."""
        k, r = self.where[cell_index]
        sim = self.sims[k]
        if sim.replicas > 1:
            sim.select(r)
        return sim.get(surface, field)

    def diagnostics(self, cell_index):
        k, r = self.where[cell_index]
        sim = self.sims[k]
        if sim.replicas > 1:
            sim.select(r)
        return sim.diagnostics()

    def close(self):
        for sim in self.sims:
            sim.close()
        self.sims = []
