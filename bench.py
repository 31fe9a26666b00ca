#!/usr/bin/env python
"""bench.py -- node-ODE hot path throughput on N B200s (one process per GPU).

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line.

BASELINE.json's metric has two halves, "trees/sec & node-ODE steps/sec"; both are measured here, on the workloads
BASELINE.json's configs name:

  node arm   (`value`, `e2e`; metric node_ode_steps_per_s): a seeded batch of 10^6 node records per GPU -- the node-evolve
             calls of configs[1] (testSuite benchmark-milkyWay: 10^3 Milky-Way-mass trees) taken out of their trees, so
             that a "step" is ONE pass of the hot path over one resident batch: every record is one call of
             mergerTreeNodeEvolverStandard%evolve over its own time interval with the full quickTest operator list.
               value : accepted RKCK steps/s (successful iterations of the driver2.c:190 loop) summed over ranks, records
                       resident in HBM, timed with CUDA events on the evolver's stream, max over ranks;
               e2e   : the same through the reference-facing C-ABI call glc_evolve_batch with HOST buffers (H2D copy +
                       layout transpose + kernels + D2H inside the timed region).
  tree arm   (`trees_per_s`): configs[1] itself -- 10^3 Milky-Way-mass trees per GPU through glc_forest_evolve (host
             scheduler + batched node evolver, host buffers, end to end) -- and configs[3]'s shape (`volume`): halo-mass-
             function-sampled Monte Carlo trees, 12 500 per GPU (10^5 over 8 GPUs), at the quickTest mass resolution.
  roofline   : the dominant kernel of the node arm, and both kernels separately (`roofline_kernels`): machine_kernel and
             drain_kernel, each with its own device time and counters from THIS run (glc_last_phase_stats): algorithmic
             bytes (3*8*n_y per accepted step + one read and one write of every node record) against the
             measured HBM peak, and FP64 flop (evaluations x ncu-measured flop per evaluation, profiles/) against a
             DFMA-chain peak measured in this process.
  cpu_baseline: the CPU checker (oracle/, OpenMP over nodes resp. trees like the reference's OpenMP over trees, built
             with the reference's optimisation flags) on a bounded sample of the SAME workloads on this box's cores:
             the first records of the node batch, the first trees of the forests (at least 16 trees per thread).
``--impl reference`` times that CPU implementation alone on the same bounded samples (the Fortran reference cannot be
built in this image: no gfortran/GSL/HDF5).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "node_ode_steps_per_s"
UNIT = "accepted RKCK node-ODE steps/s"
# FP64 flop per evaluation of the rate function: DADD + DMUL + 2*DFMA thread-level SASS instructions of every
# machine_kernel and drain_kernel launch of one whole pass divided by the pass's RHS counter (ncu,
# profiles/r01f_fp64_ops_whole_pass.txt: 2.217e11 / 13 693 801).
FLOP_PER_RHS = 16.2e3
# DRAM bytes per evaluation from the ncu --set full captures of the round-2 build: machine_kernel = second bulk slice
# (profiles/r02r_machine_kernel_bulk_slice.txt: 45.68 + 24.09 GB over the slice's 5 665 006 evaluations); drain_kernel = the
# dense pass, which is where the drain spends its time (profiles/r02aj_drain_kernel_dense_pass_coop_qag.txt: 0.67 + 1.84 GB over
# the pass's ~4.46e6 evaluations; the last, one-node-per-warp passes work out of L1/L2: 71 B per evaluation, r02r)
DRAM_BYTES_PER_RHS = {"machine_kernel": (45.677725e9 + 24.085883e9) / 5665006.0, "drain_kernel": (0.666549e9 + 1.837163e9) / 4.46e6}
N_Y = 24
BLACK_HOLE_FRACTION = 0.7
MW_ROOT_MASS, MW_RESOLUTION = 1.52e12, 1.0e9  # testSuite/parameters/benchmark_milkyWay.xml:30-42
VOLUME_RESOLUTION = 5.0e9                      # mergerTreeMassResolution fixed default (quickTest.xml does not override it)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000, help="node records per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="node records of the CPU baseline sample")
    ap.add_argument("--seed", type=int, default=219)
    ap.add_argument("--trees", type=int, default=1000, help="Milky-Way-mass trees per GPU of the tree arm (0 = skip)")
    ap.add_argument("--volume-trees", type=int, default=12500,
                    help="mass-function-sampled trees per GPU of the volume arm, configs[3] (0 = skip)")
    ap.add_argument("--cpu-trees-per-thread", type=int, default=16, help="trees per host thread of the CPU tree-walk samples")
    return ap.parse_args()


def workload(n, seed):
    from galacticus_b200 import abi, synthetic
    from galacticus_b200.evolver import params_default

    p = params_default(abi.GLC_MODEL_STANDARD)  # operatorMask = GLC_OP_ALL: the full nodeOperator list of quickTest.xml
    synthetic.finalize_params(p)
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=seed, black_hole_fraction=BLACK_HOLE_FRACTION)
    return p, props, flags, t_end


def params_only():
    from galacticus_b200 import abi, synthetic

    try:
        from galacticus_b200.evolver import params_default

        p = params_default(abi.GLC_MODEL_STANDARD)
    except Exception:  # the reference arm must not need the CUDA library
        from oracle import orc

        p = orc.params_default(abi.GLC_MODEL_STANDARD)
    return synthetic.finalize_params(p)


WORKLOAD_NAME = ("node-evolve calls of testSuite benchmark-milkyWay (10^3 MW-mass trees) as a resident batch: %d node records per GPU "
                 "over the quickTest mass range (1e10-1e13 Msun), each evolved over its own 0.05-0.8 Gyr interval; full quickTest "
                 "nodeOperator list (incl. black-hole seed/accretion/winds and jet-power CGM heating; 70 percent of the nodes start "
                 "with a black hole, the others are seeded by interrupt), hotHaloRamPressureStripping=%s, synthetic CIE and ADAF tables")


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ CPU implementation
def cpu_oracle(p):
    from galacticus_b200 import synthetic
    from oracle import orc

    orc.build()
    o = orc.Oracle(fast=True)
    synthetic.install(o, p)
    return o


def cpu_nodes(o, props, flags, t_end, cores):
    """One pass of the CPU implementation over the sample; returns (accepted steps, seconds)."""
    pp, ff = props.copy(), flags.copy()
    t0 = time.perf_counter()
    _, _, c = o.evolve_batch(pp, ff, t_end, n_threads=cores)
    return c["steps_accepted"], time.perf_counter() - t0


def cpu_forest(o, forest, n_sample, cores, what):
    """The CPU checker's tree walk (oracle/orc_tree.c, OpenMP over trees with dynamic scheduling) on the FIRST n_sample trees
    of the forest the GPU arm evolves."""
    from galacticus_b200 import synthetic

    n_all = int(forest["tree"].max()) + 1
    n_sample = max(1, min(n_sample, n_all))
    sub = synthetic.forest_subset(forest, n_sample)
    t0 = time.perf_counter()
    _, _, _, fc, c = o.forest_evolve(sub, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n_sample / dt, "unit": "merger trees/s", "trees": n_sample, "trees_per_thread": n_sample / cores,
            "nodes": int(sub["parent"].shape[0]), "seconds": dt, "cores": cores,
            "node_ode_steps_per_s": c["steps_accepted"] / dt, "sample": "the first %d trees of the %s" % (n_sample, what)}


def mw_forest(p, n_trees, seed):
    from galacticus_b200 import synthetic

    return synthetic.binary_split_forest(p, n_trees, MW_ROOT_MASS, MW_RESOLUTION, seed=seed)


def volume_forest(p, n_trees, seed):
    from galacticus_b200 import synthetic

    return synthetic.mass_function_forest(p, n_trees, VOLUME_RESOLUTION, seed=seed)


MW_NAME = ("configs[1]: %d Milky-Way-mass binary-split trees per GPU (root %.3g Msun, resolution %.3g Msun: the masses of "
           "testSuite/parameters/benchmark_milkyWay.xml), quickTest physics")
VOLUME_NAME = ("configs[3]: %d Monte Carlo trees per GPU with roots drawn from dn/dlnM ~ M^-0.9 exp(-M/1e14) on [1e10, 1e14] Msun "
               "(10^5 trees over 8 GPUs), resolution %.3g Msun, quickTest physics")


def run_reference(args):
    """Reference arm: the CPU implementation of the path (the oracle port; the Fortran reference cannot be compiled in this
    image) on the host cores, each step a bounded sample of the GPU arm's own workload (its first records / trees)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from galacticus_b200 import sharding as _sh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    seed = _sh.forest_seed(args.seed, 0) if world > 1 else args.seed
    ns = min(args.cpu_sample, args.nodes)
    p, props, flags, t_end = workload(args.nodes, seed)  # the GPU arm's records; the sample is their first ns
    props, flags, t_end = props[:ns], flags[:ns], t_end[:ns]
    cores = os.cpu_count() or 1
    o = cpu_oracle(p)
    times, steps = [], 0
    for it in range(args.warmup + args.steps):
        steps, dt = cpu_nodes(o, props, flags, t_end, cores)
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = steps * len(times) / tot
    sample = "the first %d node records of the workload per step; %.0f nodes/s" % (ns, ns * len(times) / tot)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME % (args.nodes, "font2008" if getattr(p, "hotHaloRamPressureStripping", 0) else "virialRadius"),
                   "nodes_per_gpu": args.nodes, "seed": args.seed, "sample_nodes_per_step": ns},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    nt = args.cpu_trees_per_thread * cores
    if args.trees > 0:
        line["trees_per_s"] = cpu_forest(o, mw_forest(p, args.trees, seed), nt, cores, MW_NAME % (args.trees, MW_ROOT_MASS, MW_RESOLUTION))
        line["cpu_baseline"]["trees_per_s"] = line["trees_per_s"]
    if args.volume_trees > 0:
        line["volume"] = cpu_forest(o, volume_forest(p, args.volume_trees, seed), 4 * nt, cores,
                                    VOLUME_NAME % (args.volume_trees, VOLUME_RESOLUTION))
        line["cpu_baseline"]["volume"] = line["volume"]
    print(json.dumps(line))


def forest_arm(ev, forest, barrier, torch):
    barrier()
    t1 = time.perf_counter()
    _, _, fstate, ffc, fcnt = ev.forest_evolve(forest)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t1
    from galacticus_b200 import abi

    info = {"nodes_per_gpu": int(forest["parent"].shape[0]), "rounds": ffc["rounds"], "evolve_calls": ffc["evolve_calls"],
            "promotions": ffc["promotions"], "node_mergers": ffc["node_mergers"], "failed_evolves": ffc["failed_evolves"],
            "node_ode_steps": fcnt["steps_accepted"], "rhs_evaluations": fcnt["rhs_evaluations"],
            "galaxies_at_final_time": int((fstate != abi.GLC_FOREST_NODE_PROMOTED).sum())}
    return secs, info


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from galacticus_b200 import abi, synthetic
    from galacticus_b200 import sharding as _sh
    from galacticus_b200.evolver import Evolver

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.nodes
    # independent forests shard naturally: every rank owns its own forest queue (different seed), no data-path collective
    seed = _sh.forest_seed(args.seed, rank) if world > 1 else args.seed
    p, props, flags, t_end = workload(n, seed)
    ev = Evolver(local_rank)
    synthetic.install(ev, p)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ("value")
    ev.arena_upload(props, flags, t_end)
    ev.arena_snapshot(n)
    for _ in range(args.warmup):
        ev.arena_restore(n)
        ev.evolve_arena(n)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches1 = ev.kernel_launch_count()
    t0 = time.perf_counter()
    kernel_ms, counters, phases = [], None, []
    for _ in range(args.steps):
        ev.arena_restore(n)
        counters, ms = ev.evolve_arena(n)  # synchronises the evolver's stream; ms = CUDA events around the kernels
        kernel_ms.append(ms)
        phases.append(ev.last_phase_stats())
    barrier()
    wall = time.perf_counter() - t0
    launches = ev.kernel_launch_count() - launches1
    clocks = sampler.stop()
    dev_time = sum(kernel_ms) * 1e-3  # device time of the timed kernels on the launching stream
    final_props, _, st, _ = ev.arena_download(n)

    # ---------------- end-to-end arm through the C-ABI with host buffers
    pin = torch.empty((n, abi.NPROP), dtype=torch.float64).pin_memory()
    host_props = pin.numpy()
    e2e_times = []
    for it in range(1 + max(1, min(args.steps, 3))):
        host_props[:] = props
        ff = flags.copy()
        barrier()
        t1 = time.perf_counter()
        s, i, c2 = ev.evolve_batch(host_props, ff, t_end)
        torch.cuda.synchronize()
        if it > 0:
            e2e_times.append(time.perf_counter() - t1)
    e2e_time = float(np.mean(e2e_times))
    h2d = n * (abi.NPROP * 8 + 4 + 8)
    d2h = n * (abi.NPROP * 8 + 4 + 4 + 4)

    # ---------------- tree arms: configs[1] (Milky-Way trees) and configs[3] (mass-function-sampled volume), end to end
    mw, vol = None, None
    mw_s = vol_s = 0.0
    mw_info = vol_info = None
    if args.trees > 0 or args.volume_trees > 0:
        ev.forest_evolve(synthetic.binary_split_forest(p, 4, MW_ROOT_MASS, 1.0e10, seed=1))  # warm-up
    if args.trees > 0:
        mw = mw_forest(p, args.trees, seed)
        mw_s, mw_info = forest_arm(ev, mw, barrier, torch)
    if args.volume_trees > 0:
        vol = volume_forest(p, args.volume_trees, seed)
        vol_s, vol_info = forest_arm(ev, vol, barrier, torch)
    fp64_peak = ev.fp64_peak_tflops()  # after the runs: the device is warm

    # ---------------- reduce over ranks: max time, summed work, NCCL all-reduce of an output statistic
    steps_acc = counters["steps_accepted"]
    stats = torch.tensor([dev_time, wall, e2e_time, float(steps_acc), float(counters["rhs_evaluations"]),
                          float(counters["steps_rejected"]), float(n), mw_s, float(args.trees),
                          float(mw_info["node_ode_steps"]) if mw_info else 0.0, vol_s, float(args.volume_trees),
                          float(vol_info["node_ode_steps"]) if vol_info else 0.0], dtype=torch.float64, device="cuda")
    tmax, tsum = stats.clone(), stats.clone()
    per_rank = [stats.clone() for _ in range(world)]
    # stellar mass function histogram of the evolved batch (mirrors output/analyses/volume_function_1d.F90:986-987)
    mstar = final_props[:, abi.P["DISK_MASS_STELLAR"]] + final_props[:, abi.P["SPH_MASS_STELLAR"]]
    hist = np.histogram(np.log10(np.maximum(mstar, 1.0)), bins=30, range=(5.0, 12.5))[0].astype(np.float64)
    from galacticus_b200 import sharding

    hist_t = sharding.reduce_statistics(torch.from_numpy(hist).cuda())  # NCCL all-reduce over NVLink when world > 1
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_gather(per_rank, stats)
    tmax, tsum = tmax.cpu().numpy(), tsum.cpu().numpy()
    per_rank = [r.cpu().numpy() for r in per_rank]
    ok_frac = float((st == 0).mean())

    if rank == 0:
        dev_time_max, e2e_max = float(tmax[0]), float(tmax[2])
        total_steps = float(tsum[3]) * args.steps
        value = total_steps / dev_time_max
        ms_per_step = 1e3 * dev_time_max / args.steps
        e2e_value = float(tsum[3]) / e2e_max
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # ---- rooflines per kernel from this run's own phase counters (rank 0, mean over the timed passes)
        ph = {k: float(np.mean([q[k] for q in phases])) for k in phases[0]}
        kernels = {}
        for name, key in (("machine_kernel", "machine"), ("drain_kernel", "drain")):
            ms_k = ph[key + "_ms"]
            if ms_k <= 0.0:
                continue
            bytes_k = ph[key + "_steps"] * 3 * 8 * N_Y + ph[key + "_nodes"] * (2 * abi.NPROP * 8 + 8 + 3 * 4)
            gbs = bytes_k / (ms_k * 1e-3) / 1e9
            tfl = ph[key + "_rhs"] * FLOP_PER_RHS / (ms_k * 1e-3) / 1e12
            kernels[name] = {"ms_per_pass": ms_k, "share_of_pass": ms_k / (ph["machine_ms"] + ph["drain_ms"]),
                             "rhs_evaluations": ph[key + "_rhs"], "accepted_steps": ph[key + "_steps"], "nodes_finished": ph[key + "_nodes"],
                             "rhs_per_s": ph[key + "_rhs"] / (ms_k * 1e-3),
                             "traffic": DRAM_BYTES_PER_RHS.get(name, 0.0) * ph[key + "_rhs"],
                             "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak},
                             "fp64": {"achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (tfl / fp64_peak) if fp64_peak else None}}
        dominant = max(kernels, key=lambda k: kernels[k]["ms_per_pass"]) if kernels else "machine_kernel"
        dk = kernels.get(dominant, {"hbm": {"achieved": 0.0, "frac": 0.0}, "rhs_evaluations": 0.0})
        ms_kernel = float(np.mean(kernel_ms))
        rhs_total = float(counters["rhs_evaluations"])
        ach_tflops = rhs_total * FLOP_PER_RHS / (ms_kernel * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD_NAME % (n, "font2008" if getattr(p, "hotHaloRamPressureStripping", 0) else "virialRadius"),
                "nodes_per_gpu": n, "seed": args.seed,
                "l2": "inputs (%.0f MB/GPU of node records + %.0f MB of per-slot continuations) larger than L2"
                      % (n * abi.NPROP * 8 / 1e6, 148 * 2048 * 4.2e3 / 1e6),
                "nodes_per_s": float(tsum[6]) * args.steps / dev_time_max,
                "rhs_evaluations_per_s": float(tsum[4]) * args.steps / dev_time_max,
                "rejected_step_fraction": float(tsum[5]) / max(float(tsum[3]) + float(tsum[5]), 1.0),
                "status_ok_fraction": ok_frac,
                "wall_s_timed_region": float(tmax[1]),
                "per_rank": {"ms_per_step": [1e3 * float(r[0]) / args.steps for r in per_rank],
                             "e2e_s": [float(r[2]) for r in per_rank],
                             "trees_s": [float(r[7]) for r in per_rank], "volume_s": [float(r[10]) for r in per_rank]},
                "stellar_mass_function_counts": hist_t.cpu().numpy().tolist(),
            },
            "roofline": {"bound": "hbm", "achieved": dk["hbm"]["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": dk["hbm"]["frac"],
                         # DRAM traffic of that kernel: ncu --set full capture (bytes per evaluation, profiles/) x its evaluations
                         "traffic": DRAM_BYTES_PER_RHS.get(dominant, 0.0) * dk["rhs_evaluations"],
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                         "kernel": dominant,
                         "note": "the path is FP64-ALU / latency bound (SURVEY 8d: >15 flop per algorithmic byte): see roofline_fp64 "
                                 "and roofline_kernels; algorithmic bytes = 576 B per accepted step + one read and one write of "
                                 "each node record the kernel finishes"},
            "roofline_fp64": {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (ach_tflops / fp64_peak) if fp64_peak else None, "flop_per_rhs": FLOP_PER_RHS,
                              "scope": "whole pass (machine_kernel + drain_kernel)",
                              "peak_source": "DFMA-chain microbenchmark in this process (glc_measure_fp64_peak_tflops)"},
            "roofline_kernels": kernels,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            # the other half of BASELINE.json's metric, end to end through the C-ABI (host scheduler + batched node evolver,
            # host buffers): all ranks' trees / slowest rank's time
            "trees_per_s": ({"value": float(tsum[8]) / float(tmax[7]), "unit": "merger trees/s",
                             "node_ode_steps_per_s": float(tsum[9]) / float(tmax[7]), "seconds": float(tmax[7]),
                             "workload": MW_NAME % (args.trees, MW_ROOT_MASS, MW_RESOLUTION), "trees_per_gpu": args.trees,
                             **mw_info} if mw_info else None),
            "volume": ({"value": float(tsum[11]) / float(tmax[10]), "unit": "merger trees/s",
                        "node_ode_steps_per_s": float(tsum[12]) / float(tmax[10]), "seconds": float(tmax[10]),
                        "workload": VOLUME_NAME % (args.volume_trees, VOLUME_RESOLUTION), "trees_per_gpu": args.volume_trees,
                        **vol_info} if vol_info else None),
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        # CPU baseline (rank 0, N=1 only): the CPU implementation on bounded samples of the same workloads
        if world == 1:
            try:
                o = cpu_oracle(p)
                ns = min(args.cpu_sample, n)
                steps_c, secs = cpu_nodes(o, props[:ns], flags[:ns], t_end[:ns], cores)
                line["cpu_baseline"] = {"value": steps_c / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": "the first %d node records of the GPU workload, %.1f s, %.0f nodes/s"
                                                  % (ns, secs, ns / secs)}
                nt = args.cpu_trees_per_thread * cores
                if mw is not None:
                    line["cpu_baseline"]["trees_per_s"] = cpu_forest(o, mw, nt, cores, "GPU arm's forest")
                if vol is not None:
                    line["cpu_baseline"]["volume"] = cpu_forest(o, vol, 4 * nt, cores, "GPU arm's volume")
            except Exception as e:  # the checker is optional for the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
