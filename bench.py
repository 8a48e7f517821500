#!/usr/bin/env python
"""bench.py -- element-steps/s of the explicit UL step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], "synthetic 1M-triangle cell on 4M-node substrate" -- the
configuration the metric's roofline target is quoted on; it fits one GPU (1.1 GB of state).  A "step" is one full UL
step (predict, bond kinetics + point location, cell forces + update, substrate forces + update) over that pair of
meshes with the state resident in HBM.  N > 1: every rank advances its own independent cell+substrate replica (the
reference's only parallel mode is independent processes, s_sim_031021_1.sub) -- no data-path collective, weak scaling.
The state (1.1 GB) is ~9x the 126 MB L2, so successive steps stream from HBM: no explicit L2 flush is needed.

The JSON line carries, beside the base contract:
  roofline      the dominant kernel (k_subs_ring): algorithmic bytes per launch (216 B/substrate node + 12 B/element,
                SURVEY.md 8d) / its CUDA-event duration measured live, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle's C restatement ("port", grid point location) timed on this box's host on a bounded sample
  e2e           the same metric through the adapter-mode C-ABI call with host round trips every step
`--impl reference` times the reference CPU path instead (oracle/_ref where it was built, else the oracle port) on all
host cores as independent replicas, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

METRIC = "element_steps_per_sec"
UNIT = "element-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cell-tris", type=int, default=1000000)
    ap.add_argument("--subs-side", type=int, default=2000, help="substrate nodes per side (2000 -> 4M nodes, 8M triangles)")
    ap.add_argument("--variant", default="k31", choices=["k1", "k31", "kMD"])
    ap.add_argument("--dt", type=float, default=2e-5, help="explicit step below the substrate CFL limit h*sqrt(rho/mu) = 5e-5")
    ap.add_argument("--mode", default="ensemble", choices=["ensemble", "partition"],
                    help="N > 1: independent replicas (weak scaling, default) or ONE workload with the substrate cut into strips and a "
                         "per-step NCCL halo exchange (strong scaling)")
    ap.add_argument("--halo", default="peer", choices=["peer", "dist"],
                    help="--mode partition: border strains pushed into the neighbours' ghost rows by the substrate kernel itself over NVLink peer "
                         "memory (default), or a host-driven torch.distributed exchange after every step (the baseline)")
    ap.add_argument("--cell", default="partitioned", choices=["partitioned", "replicated"],
                    help="--mode partition: every rank computes only the cell patches that reach its strip (default), or the whole cell")
    ap.add_argument("--shape", default="strips", choices=["strips", "sectors"],
                    help="--mode partition: substrate cut into vertical strips (default; measured faster at 8 GPUs: the ranks under the cell do the cell's work while "
                         "the outer ranks carry more substrate) or into sectors around the cell's centroid (every rank gets an equal share of both)")
    ap.add_argument("--cell-cost", type=float, default=4.0, help="--mode partition, --cell partitioned: weight of a cell node in substrate-node units when the strips are balanced")
    ap.add_argument("--no-partition-record", action="store_true", help="N > 1, ensemble mode: skip the strong-scaling sub-record (`partition`)")
    ap.add_argument("--balance", type=int, default=1, help="--mode partition: strips of equal cost (nodes under the cell weigh 3x) instead of equal node count")
    ap.add_argument("--order", default="morton", choices=["morton", "morton_rows"], help="node numbering of the synthetic meshes")
    ap.add_argument("--diagonals", default="alternate", choices=["alternate", "same"], help="substrate triangulation: union jack (valence 4/8, default) or uniform valence 6 (diagnostic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the small side measurements (configs 1/2 latency, config 5 batch)")
    ap.add_argument("--cpu-sample-div", type=int, default=16, help="CPU baseline sample = workload / this (both meshes)")
    return ap.parse_args()


def workload_meshes(cell_tris, subs_side, order="morton", diagonals="alternate"):
    from wholecellfe_b200.meshgen import disk_mesh, square_mesh
    xc, tc = disk_mesh(cell_tris, order=order)
    xs, ts = square_mesh(subs_side, order=order, diagonals=diagonals)
    return xc, tc, xs, ts


def workload_name(a, sizes=None):
    """config.workload, the same string on both arms: the sizes follow from the generators (rings of 6k nodes for the disk, a
    structured square), so the reference arm does not have to build the full meshes to name them."""
    if sizes is None:
        nr = max(2, int(round(math.sqrt(a.cell_tris / 6.0))))
        sizes = (6 * nr * nr, 1 + 3 * nr * (nr + 1), a.subs_side ** 2, 2 * (a.subs_side - 1) ** 2)
    return ("synthetic %d-triangle cell (%d nodes) on %d-node substrate (%d triangles), BASELINE configs[3]; one independent replica per GPU" % sizes)


def algorithmic_bytes(nn_c, ne_c, nn_s, ne_s):
    """SURVEY.md 8d / BASELINE.md 3.4: every live state word read once + written once, int32 connectivity read once."""
    return 304 * nn_c + 216 * nn_s + 12 * (ne_c + ne_s)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the GPU is under load (B200_PROFILING.md recipe): one persistent
    `nvidia-smi -lms` process, started before the warm-up and stopped after the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        busy = [r for r in rows if r[6].isdigit() and int(r[6]) > 0] or rows
        sm = sorted(float(r[0]) for r in busy)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in busy)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows), "samples_under_load": len(busy)}


def time_oracle_port(variant, xc, tc, xs, ts, dt, steps):
    from oracle.oracle import OracleSim
    o = OracleSim(variant, xc, tc, xs, ts, use_grid=True)
    o.srand(1)
    t = 0.0
    o.step(dt, t); t += dt        # warm-up (page in, first grid build)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(dt, t); t += dt
    return (len(tc) + len(ts)) * steps / (time.perf_counter() - t0)


def _replica_worker(args):
    kind, variant, cell_tris, subs_side, dt, steps, warmup = args
    xc, tc, xs, ts = workload_meshes(cell_tris, subs_side)
    if kind == "reference":
        from oracle.refrun import RefSim
        sim = RefSim(variant, xc, tc, xs, ts); sim.srand(1)
    else:
        from oracle.oracle import OracleSim
        sim = OracleSim(variant, xc, tc, xs, ts, use_grid=True); sim.srand(1)
    t = 0.0
    for _ in range(warmup):
        sim.step(dt, t); t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step(dt, t); t += dt
    return time.perf_counter() - t0, len(tc) + len(ts)


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1048576.0
    except Exception:
        pass
    return 8.0


def run_reference_arm(a):
    """The reference's CPU implementation of the path, all host cores as independent replicas (its only parallel mode,
    s_sim_031021_1.sub).  Two measurements in one line:
      value / cpu_baseline   the UNMODIFIED reference (oracle/_ref, brute-force point location, UL:1309, :2026) on a 1/256-size
                             sample -- as large as its O(Nn_cell x Ne_subs) search allows -- exactly `steps` timed steps;
      same_algorithm         the oracle's C restatement with the GRID point location our kernels use, same workload at the
                             largest fraction of the full size the host's memory allows (1/1 where it fits), all cores:
                             the like-for-like CPU number (the reference value flatters nobody: its algorithm is quadratic)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle.refrun import ref_available
    cores = os.cpu_count() or 1
    div = a.cpu_sample_div
    use_ref = ref_available(a.variant)
    if use_ref:
        kind, cell_tris, subs_side = "reference", max(600, a.cell_tris // 256), max(24, a.subs_side // 16)
    else:
        kind, cell_tris, subs_side = "port", max(600, a.cell_tris // div), max(24, int(a.subs_side / math.sqrt(div)))
    steps = max(1, min(a.steps, 60))            # every reported step is run; above 60 the sample would outlast "a few minutes"
    warm = max(1, min(a.warmup, 3))
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_replica_worker, [(kind, a.variant, cell_tris, subs_side, a.dt, steps, warm)] * cores)
        wall = max(r[0] for r in res)
        value = sum(r[1] for r in res) * steps / wall
        sample = ("%d independent replicas (one per core) x %d steps of a %d-triangle cell on a %d-node substrate (%s point location): 1/%d of the workload's cell"
                  % (cores, steps, cell_tris, subs_side * subs_side, "brute-force" if use_ref else "grid", max(1, a.cell_tris // max(1, cell_tris))))
        # the like-for-like arm: same algorithm (grid search), same mesh generator; ~3 GB of host memory per full-size replica
        same = None
        try:
            frac = 1
            while frac < 64 and cores * 3.0 / frac > 0.6 * _mem_available_gb():
                frac *= 4
            ct, ss = max(600, a.cell_tris // frac), max(24, int(a.subs_side / math.sqrt(frac)))
            psteps = 2 if frac == 1 else 4
            pres = pool.map(_replica_worker, [("port", a.variant, ct, ss, a.dt, psteps, 1)] * cores)
            pwall = max(r[0] for r in pres)
            same = {"value": sum(r[1] for r in pres) * psteps / pwall, "unit": UNIT, "cores": cores, "kind": "port", "fraction_of_workload": "1/%d" % frac,
                    "ms_per_step_per_replica": 1e3 * pwall / psteps,
                    "sample": "%d replicas (one per core) x %d steps of a %d-triangle cell on a %d-node substrate, oracle C restatement with grid point location"
                              % (cores, psteps, ct, ss * ss)}
        except Exception as e:                  # side measurement: never fatal
            same = {"skipped": str(e)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": workload_name(a), "variant": a.variant, "dt": a.dt,
                       "note": "the CPU arm runs a bounded sample of this workload (cpu_baseline.sample); steps = timed steps actually run (requested %d)" % a.steps},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "same_algorithm": same,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_partition_mode(a, dist, rank, world, local, cell_mesh, subs_mesh, steps=None, one_gpu_ms=None):
    """Strong scaling (BASELINE configs[3]): ONE cell + substrate over `world` GPUs.  The substrate is cut into strips of equal cost;
    with --cell partitioned (default) every rank also computes only the cell patches that reach its strip, rows and partial sums
    travel through peer memory inside the kernels (include/wcfe.h); --cell replicated keeps the whole cell on every rank.
    value counts every element once.  Returns the record (rank 0) or None."""
    import torch
    from wholecellfe_b200.partition import PartitionedSim
    steps = steps or a.steps
    sim = PartitionedSim(a.variant, cell_mesh, subs_mesh, dist, local, rng=(1, 1), halo=a.halo, balance=bool(a.balance), cell=a.cell, shape=a.shape, cell_cost=a.cell_cost)
    ne = len(cell_mesh[1]) + len(subs_mesh[1])
    sim.step_n(a.warmup, a.dt)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = sim.sim.kernel_launches()
    e0.record()
    sim.step_n(steps, a.dt)
    e1.record()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    l1 = sim.sim.kernel_launches()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    halo = sum(len(v) for v in sim.part.send.values())
    cell_rows = sum(len(v) for v in sim.cpart.exports.values()) if sim.cpart is not None else 0
    # per-kernel durations of this rank's share (event pairs around every launch: the PDL chain is off in this pass)
    sim.sim.set_profile(True)
    sim.step_n(10, a.dt)
    kms, kcnt = sim.sim.kernel_ms()
    sim.sim.set_profile(False)
    names = ["k_predict_cell", "k_bond", "k_cell_patch", "k_boundary_finish", "k_subs_ring", "k_bucket_reduce"]
    per_kernel = {n: kms[i] / kcnt[i] for i, n in enumerate(names) if kcnt[i]}
    pk = torch.tensor([per_kernel.get(n, 0.0) for n in names], dtype=torch.float64, device="cuda")
    pk_all = [torch.zeros_like(pk) for _ in range(world)]
    dist.all_gather(pk_all, pk)
    # parity at full size: every rank repeats the same steps on the WHOLE mesh on its own GPU and compares its owned rows bit for bit
    import wholecellfe_b200 as W
    total_steps = a.warmup + steps + 10
    ids, ea_part = sim.owned_field("ea"); _, v_part = sim.owned_field("V")
    cid, xc_part = sim.owned_cell_field("x"); _, cc_part = sim.owned_cell_field("cC"); _, cv_part = sim.owned_cell_field("V")
    part0 = {"substrate_nodes_local": int(len(sim.part.local_nodes)), "owned": int(len(sim.part.owned)), "ghosts": int(len(sim.part.ghosts)),
             "cell_patches_computed": int(len(sim.cpart.patch_list)) if sim.cpart is not None else None,
             "cell_patches_owned": int(len(sim.cpart.owned_patches)) if sim.cpart is not None else None}
    npl = torch.tensor([len(sim.cpart.patch_list) if sim.cpart is not None else 0], dtype=torch.int64, device="cuda")
    npl_all = [torch.zeros_like(npl) for _ in range(world)]
    dist.all_gather(npl_all, npl)
    sim.close()                                  # collective: sync + barrier before any rank unmaps its halo buffers
    one = W.ULSim(a.variant, device=local)
    one.upload_mesh(W.CELL, cell_mesh[0], cell_mesh[1]); one.upload_mesh(W.SUBS, subs_mesh[0], subs_mesh[1]); one.set_rng(1, 1)
    one.step_n(total_steps, a.dt, 0.0)
    one_ms = one.last_step_ms() / total_steps
    ok = bool(np.array_equal(one.get(W.SUBS, "ea")[ids], ea_part) and np.array_equal(one.get(W.SUBS, "V")[ids], v_part) and
              np.array_equal(one.get(W.CELL, "x")[cid], xc_part) and np.array_equal(one.get(W.CELL, "cC")[cid], cc_part) and np.array_equal(one.get(W.CELL, "V")[cid], cv_part))
    one.close()
    okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank != 0:
        return None
    base_ms = one_gpu_ms if one_gpu_ms else one_ms
    cellwhat = ("cell partitioned by patches (replicated storage; %d cell rows x 72 B pushed per step by rank 0, patches computed per rank %s of %d)"
                % (cell_rows, [int(v.item()) for v in npl_all], (len(cell_mesh[0]) + 127) // 128)) if a.cell == "partitioned" else "cell replicated"
    return {"metric": METRIC, "value": ne * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": a.warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ONE synthetic %d-triangle cell on %d-node substrate, substrate in %d %s, %s"
                                   % (len(cell_mesh[1]), len(subs_mesh[0]), world, a.shape, cellwhat), "variant": a.variant, "dt": a.dt,
                       "parallelism": ("substrate partition x%d, halo exchange fused into the kernels over NVLink peer memory (%d border nodes x 24 B per step from rank 0)" if a.halo == "peer"
                                       else "substrate partition x%d + per-step torch.distributed (NCCL) halo exchange (%d border nodes x 24 B per step from rank 0)") % (world, halo)},
            "one_gpu_ms_per_step": base_ms, "speedup_vs_one_gpu": base_ms / (ms / steps),
            "gpu_launches": l1 - l0, "per_kernel_ms_by_rank": {n: [float(v[i].item()) for v in pk_all] for i, n in enumerate(names) if any(float(v[i].item()) > 0 for v in pk_all)},
            "bitwise_equal_to_single_gpu": {"value": bool(int(okt.item())), "steps": total_steps,
                                            "what": "owned substrate rows (ea, V) and owned cell rows (x, V, cC) of every rank against the same steps of the whole mesh on one GPU"},
            "pieces": ("%s of equal cost (nodes under the cell x3%s)" % (a.shape, ", + %g per cell node above" % a.cell_cost if a.cell == "partitioned" else "")) if a.balance else "%s of equal node count" % a.shape,
            "rank0": part0}


def other_configs(W):
    """Side measurements on the reference's shipped example meshes (BASELINE configs 0-1: 296/542 cell, 320/578 substrate, dt 0.005)
    and on a batch of 512 of them (configs[4] per GPU).  Wall clock around synchronous calls, a few hundred ms in total."""
    from wholecellfe_b200.ensemble import sweep_grid, EnsembleRunner
    path = os.path.join(ROOT, "tests", "golden", "example_meshes.npz")
    if not os.path.exists(path):
        return None
    d = np.load(path)
    xc, tc, xs, ts = d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]
    out = {"mesh": "shipped example: %d-triangle cell on %d-triangle substrate, dt 0.005, glibc replay RNG" % (len(tc), len(ts))}
    for var in ("k1", "k31", "kMD"):
        g = W.ULSim(var); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
        t = g.step_n(80, 0.005, 0.0)
        g.step_n(800, 0.005, t)
        out["config1_%s_us_per_step" % var] = g.last_step_ms() / 800 * 1e3
        out["config1_%s_element_steps_per_sec" % var] = (len(tc) + len(ts)) * 800 / (g.last_step_ms() * 1e-3)
        g.close()
    # BASELINE configs[2]: kMD law on a ~100k-triangle synthetic disk over a refined substrate (node spacing 0.25 um <= 4 x the cell's),
    # dt 5e-5 as in the reference's JAP scripts (scripts/cell_JAP.txt:5), counter-based RNG
    try:
        from wholecellfe_b200.meshgen import disk_mesh, square_mesh
        xc3, tc3 = disk_mesh(100000); xs3, ts3 = square_mesh(400)
        g = W.ULSim("kMD"); g.upload_mesh(W.CELL, xc3, tc3); g.upload_mesh(W.SUBS, xs3, ts3); g.set_rng(1, 1)
        t3 = g.step_n(50, 5e-5, 0.0)
        g.step_n(500, 5e-5, t3)
        out["config3_kMD_mesh"] = "%d-triangle cell on %d-triangle substrate" % (len(tc3), len(ts3))
        out["config3_kMD_ms_per_step"] = g.last_step_ms() / 500
        out["config3_kMD_element_steps_per_sec"] = (len(tc3) + len(ts3)) * 500 / (g.last_step_ms() * 1e-3)
        g.close()
    except Exception as e:                                   # side measurement: never fatal
        out["config3_kMD"] = "skipped: %s" % e
    # the whole application (reference main.cpp + El Topo + VTK frames every step), with the reference's own step and with ours
    try:
        out.update(whole_application(xc, tc, xs, ts))
    except Exception as e:                                   # side measurement: never fatal
        out["whole_application"] = "skipped: %s" % e
    n = 512
    run = EnsembleRunner(sweep_grid(n), (xc, tc), (xs, ts))
    run.advance(40, 0.005)
    w0 = time.perf_counter(); run.advance(400, 0.005); w1 = time.perf_counter()
    out["config5_batch_cells_per_gpu"] = n
    out["config5_batch_element_steps_per_sec"] = n * run.elements_per_cell * 400 / (w1 - w0)
    out["config5_us_per_cell_step"] = (w1 - w0) / (400 * n) * 1e6
    run.close()
    return out


def config5_sweep(W, dist, rank, world, local, per_gpu=512, steps=400):
    """BASELINE configs[4]: per_gpu x world independent cells of the shipped example size (4 096 at 8 GPUs) over stiffness law x
    ligand density x seed (ensemble.sweep_grid), each rank's shard as batched contexts (gridDim.y = cells), no communication.
    Wall clock between barriers, max over ranks."""
    import torch
    from wholecellfe_b200.ensemble import sweep_grid, EnsembleRunner, partition
    path = os.path.join(ROOT, "tests", "golden", "example_meshes.npz")
    d = np.load(path)
    xc, tc, xs, ts = d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]
    cells = sweep_grid(per_gpu * world)
    mine = [cells[i] for i in partition(len(cells), world, rank)]
    run = EnsembleRunner(mine, (xc, tc), (xs, ts), device=local)
    run.advance(40, 0.005)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter(); run.advance(steps, 0.005); torch.cuda.synchronize(); w = time.perf_counter() - w0
    tt = torch.tensor([w], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wmax = float(tt.item())
    finite = bool(np.isfinite(run.get(0, W.CELL, "x")).all() and np.isfinite(run.get(len(mine) - 1, W.CELL, "x")).all())
    run.close()
    return {"cells": len(cells), "cells_per_gpu": per_gpu, "n_gpus": world, "steps": steps, "element_steps_per_sec": len(cells) * run.elements_per_cell * steps / wmax,
            "us_per_cell_step": wmax / (steps * per_gpu) * 1e6, "finite": finite,
            "what": "laws k1/k31/kMD x log-spaced rho_i_max 1..100 x seeds, %d-triangle cell on %d-triangle substrate each, dt 0.005, Philox, batched (one launch per kernel per step per law group)"
                    % (len(tc), len(ts))}


def pcie_concurrent(torch, dist, world, nbytes, reps=20):
    """Every rank copies `nbytes` (the cell's Vec3d image) pinned host -> device and back, all ranks at once: GB/s per rank (slowest)."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); dv = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, src, dst in (("h2d", h, dv), ("d2h", dv, h)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out[name + "_gbs_per_rank"] = nbytes * reps / float(tt.item()) / 1e9
    out["bytes"] = nbytes; out["ranks_at_once"] = world
    return out


def whole_application(xc, tc, xs, ts, end_t=0.32):
    """oracle/_ref/cellBP_{ref,gpu}_k1 (oracle/build_app.sh): the unmodified reference application on config 1, El Topo on; wall
    time of the whole process divided by its steps.  Startup (CUDA context ~1 s) is inside the GPU number: measured twice."""
    import subprocess
    import tempfile
    from wholecellfe_b200.meshio import write_ecomsol
    exe = {k: os.path.join(ROOT, "oracle", "_ref", "cellBP_%s_k1" % k) for k in ("ref", "gpu")}
    if not all(os.path.exists(v) for v in exe.values()):
        return {"whole_application": "skipped: oracle/_ref/cellBP_{ref,gpu}_k1 not built"}
    tpl = open(os.path.join(ROOT, "tests", "golden", "reference_script_template.txt")).read()
    res = {}
    with tempfile.TemporaryDirectory() as w:
        os.makedirs(os.path.join(w, "scripts"))
        write_ecomsol(os.path.join(w, "scripts", "disk.txt"), xc, tc); write_ecomsol(os.path.join(w, "scripts", "subs.txt"), xs, ts)
        for k in ("ref", "gpu"):
            t = {}
            for e in (end_t, 7 * end_t):
                t[e] = 1e30
                for rep in range(2):          # best of two: process start-up (CUDA context, page-in) varies by hundreds of ms
                    open(os.path.join(w, "cell.txt"), "w").write(tpl.format(end_t=e, max_edge=2, improve=1, topo=1, collide=1, mesh="scripts/disk.txt"))
                    open(os.path.join(w, "subs.txt"), "w").write(tpl.format(end_t=e, max_edge=8, improve=1, topo=1, collide=1, mesh="scripts/subs.txt"))
                    out = os.path.join(w, "out_%s_%g_%d" % (k, e, rep)); os.makedirs(out)
                    w0 = time.perf_counter()
                    subprocess.run([exe[k], "cell.txt", "subs.txt", out], cwd=w, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, timeout=600)
                    t[e] = min(t[e], time.perf_counter() - w0)
            steps = round(6 * end_t / 0.005)            # 384 steps between the two runs: start-up noise of ~0.3 s is < 1 ms per step
            res["config1_whole_app_%s_ms_per_step" % k] = (t[7 * end_t] - t[end_t]) / steps * 1e3     # startup cancels
    res["whole_app_what"] = "unmodified reference application (main.cpp, El Topo integrate + improve_mesh, two .vtk files per step) on the shipped example, k1; ref = its own UL_growth_explicit.o on one host core, gpu = ul_adapter.cpp + libwcfe (state up/downloaded every step)"
    return res


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    import wholecellfe_b200 as W

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the UL step has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()

    xc, tc, xs, ts = workload_meshes(a.cell_tris, a.subs_side, a.order, a.diagonals)
    nn_c, ne_c, nn_s, ne_s = len(xc), len(tc), len(xs), len(ts)
    if a.mode == "partition" and world > 1:
        rec = run_partition_mode(a, dist, rank, world, local, (xc, tc), (xs, ts))
        if rank == 0:
            print(json.dumps(rec), flush=True)
        dist.destroy_process_group()
        return
    sim = W.ULSim(a.variant, device=local, stream=stream.cuda_stream)
    sim.upload_mesh(W.CELL, xc, tc); sim.upload_mesh(W.SUBS, xs, ts)
    sim.set_rng(1, 1 + rank)                     # counter-based Philox, one stream per replica
    t = 0.0
    launches0 = sim.kernel_launches()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    t = sim.step_n(a.warmup, a.dt, t)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = sim.kernel_launches()
    e0.record(stream)
    t = sim.step_n(a.steps, a.dt, t)
    e1.record(stream)
    barrier()
    sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = sim.kernel_launches() - l0
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = (ne_c + ne_s) * a.steps * world / (ms_max * 1e-3)

    # per-kernel durations (CUDA events on the launching stream), separate pass so the events do not perturb `value`
    sim.set_profile(True)
    t = sim.step_n(max(3, min(a.steps, 10)), a.dt, t)
    kms, kcnt = sim.kernel_ms()
    sim.set_profile(False)
    info = sim.path_info()
    names = (["k_predict_cell", "k_bond", "k_cell_patch", "k_boundary_finish", info["subs_kernel"], "k_bucket_reduce"] if info["planar"]
             else ["k_predict_cell", "k_bond", "k_cell_node", "k_boundary_finish", "k_subs_node"])
    per_kernel = {n: kms[i] / kcnt[i] for i, n in enumerate(names) if kcnt[i]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    dom = names[4]
    dom_bytes = 216 * nn_s + 12 * ne_s
    dom_ms = per_kernel[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_bytes = algorithmic_bytes(nn_c, ne_c, nn_s, ne_s)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "per_kernel_ms": per_kernel, "share_of_step": dom_ms / max(sum(per_kernel.values()), 1e-30),
                "path": "planar patch kernels (TMA-staged)" if info["planar"] else "general 3-D node-centric kernels",
                "patch_elements_per_element": {"cell": info["cell_patch_elem_ratio"], "substrate": info["subs_patch_elem_ratio"]},
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_max / a.steps * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_max / a.steps * 1e-3) / 1e9 / peak}}
    # ncu dram__bytes per launch (profiles/traffic.json, regenerated from the round's --set full capture): the planar path never
    # moves the z words the contract bytes count, so the fraction of the HBM peak actually USED is printed beside the contract one
    tr = {}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    roofline["traffic"] = tr.get(dom)
    if tr.get(dom) and dom_ms > 0:
        roofline["frac_traffic"] = tr[dom] / (dom_ms * 1e-3) / 1e9 / peak
    # bytes the planar kernels cannot avoid: in-plane x, V, A + 3 strain words (+ c_C, fa_u in-plane on the cell), read and written
    # once, 12 B of connectivity per element
    planar_subs = 128 * nn_s + 12 * ne_s
    planar_cell = 192 * nn_c + 12 * ne_c
    roofline["planar_compulsory_bytes_per_launch"] = planar_subs
    roofline["frac_planar_compulsory"] = planar_subs / (dom_ms * 1e-3) / 1e9 / peak if dom_ms > 0 else None
    if info["planar"] and per_kernel.get("k_cell_patch"):
        cms = per_kernel["k_cell_patch"]
        cell_bytes = 304 * nn_c + 12 * ne_c
        roofline["cell_kernel"] = {"kernel": "k_cell_patch", "kernel_ms": cms, "algorithmic_bytes_per_launch": cell_bytes,
                                   "achieved": cell_bytes / (cms * 1e-3) / 1e9, "frac": cell_bytes / (cms * 1e-3) / 1e9 / peak,
                                   "traffic": tr.get("k_cell_patch"),
                                   "frac_traffic": (tr["k_cell_patch"] / (cms * 1e-3) / 1e9 / peak) if tr.get("k_cell_patch") else None,
                                   "planar_compulsory_bytes_per_launch": planar_cell,
                                   "target": "north_star: >= 0.5 for the fused force+update step on the 1M-triangle cell (SURVEY 8d: 164 MB in <= 50 us)"}

    # end to end: a host integrator in the loop -- positions to pinned host memory and back EVERY step (wall clock, max over ranks)
    def time_roundtrip(both):
        nonlocal t
        steps = max(2, min(a.steps, 20 if both else 100))
        t, _, _ = sim.step_n_roundtrip(2, a.dt, t, both_surfaces=both)      # warm-up: pins the host staging buffer, first-touches it
        barrier()
        w0 = time.perf_counter()
        t, h2d, d2h = sim.step_n_roundtrip(steps, a.dt, t, both_surfaces=both)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        te = torch.tensor([w1 - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return (ne_c + ne_s) * steps * world / float(te.item()), h2d, d2h, steps
    v0, h2d, d2h, e2e_steps = time_roundtrip(False)
    v1, h2d1, d2h1, steps1 = time_roundtrip(True)
    e2e = {"value": v0, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "what": "wcfe_step_n_roundtrip, every step: predict -> D2H of the cell's predicted positions (array of Vec3d as the reference holds it, pinned) -> "
                   "(host integrator) -> H2D of the accepted positions -> update -> D2H of the diagnostics; the frozen substrate's positions (main.cpp:299) "
                   "do not travel, and the hand-over of step q+1 overlaps the substrate half of step q",
           "both_surfaces_serial": {"value": v1, "h2d_bytes_per_step": h2d1, "d2h_bytes_per_step": d2h1, "steps": steps1,
                                    "what": "the same with the substrate's 4M positions also crossing the bus both ways and no overlap (ul_adapter.cpp's traffic pattern)"}}
    x = sim.get(W.CELL, "x")
    if not np.isfinite(x).all():
        raise SystemExit("bench.py: non-finite positions after the run")

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        div = a.cpu_sample_div
        cxc, ctc, cxs, cts = workload_meshes(max(600, a.cell_tris // div), max(24, int(a.subs_side / math.sqrt(div))))
        v = time_oracle_port(a.variant, cxc, ctc, cxs, cts, a.dt, 2)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "2 steps of a 1/%d-size copy of the workload (%d-triangle cell on %d-node substrate), oracle C restatement with grid point location, 1 thread"
                         % (div, len(ctc), len(cxs))}
    other = None
    if rank == 0 and world == 1 and not a.no_other_configs:
        other = other_configs(W)
    # BASELINE configs[4] as stated: 512 independent config-1 cells per GPU (4 096 over 8 GPUs), stiffness law x ligand density x seed,
    # batched per GPU, no communication; and what the bus gives every rank when all N move their 12 MB at once (explains e2e scaling)
    config5 = pcie = None
    if not a.no_other_configs:
        try:
            config5 = config5_sweep(W, dist, rank, world, local)
        except Exception as e:
            config5 = {"skipped": str(e)}
        try:
            pcie = pcie_concurrent(torch, dist, world, 3 * 8 * nn_c)
        except Exception as e:
            pcie = {"skipped": str(e)}
    # strong scaling of the SAME workload over the N GPUs of this run (BASELINE configs[3]), as a sub-record the driver's scaling run sees
    partition = None
    if world > 1 and not a.no_partition_record:
        sim.close()
        try:
            partition = run_partition_mode(a, dist, rank, world, local, (xc, tc), (xs, ts), steps=min(max(a.steps, 100), 200), one_gpu_ms=ms_max / a.steps)
        except Exception as e:                               # never lose the main line
            partition = {"failed": str(e)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a, (ne_c, nn_c, nn_s, ne_s)), "variant": a.variant, "dt": a.dt, "rng": "philox", "parallelism": "ensemble x%d (no comms)" % world,
                           "l2": "state %.2f GB >> 126 MB L2, no flush" % (step_bytes / 2 / 1e9)},
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
        if cpu:
            line["cpu_baseline"] = cpu
        if other:
            line["other_configs"] = other
        if config5:
            line["config5_sweep"] = config5
        if pcie:
            line["pcie_concurrent"] = pcie
        if partition:
            line["partition"] = partition
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
