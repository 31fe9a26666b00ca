#!/usr/bin/env python
"""bench.py -- node-ODE hot path throughput on N B200s (one process per GPU).

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line.

  workload   : a seeded batch of synthetic node records per GPU -- the stand-in for BASELINE.json configs[1]
               (testSuite benchmark-milkyWay: 10^3 Milky-Way-mass trees ~ 10^6 node-evolve calls); every record is
               one call of mergerTreeNodeEvolverStandard%evolve over its own time interval with the quickTest
               operator set (black-hole operators not restated yet) and synthetic CIE tables.
  step       : one pass of the hot path over the whole batch: every node is evolved to its end time.
  value      : accepted RKCK node-ODE steps/s (successful iterations of the driver2.c:190 loop) summed over all
               ranks, inputs already resident in HBM (device arena), timed with CUDA events on the evolver's
               stream, max over ranks.
  e2e        : the same metric through the reference-facing C-ABI call glc_evolve_batch with HOST buffers
               (H2D copy + layout transpose + kernels + D2H inside the timed region).
  roofline   : the dominant kernel (machine_kernel) against the HBM roofline with the ALGORITHMIC bytes of
               SURVEY.md 8(d) (state streamed once per accepted step: 3*8*n_y bytes, plus one read and one write
               of every node record per launch), and -- because the path is FP64-ALU/latency bound, not HBM
               bound -- `roofline_fp64`: FP64 flop/s from the kernel's own RHS counter times the ncu-measured
               flop per RHS evaluation (profiles/) against a DFMA-chain peak measured in this process.
  cpu_baseline: the CPU checker (oracle/, OpenMP over nodes like the reference's OpenMP over trees, built with the
               reference's own optimisation flags) on a bounded sample of the same workload on this box's cores.
``--impl reference`` times that CPU implementation alone (the Fortran reference cannot be built in this image:
no gfortran/GSL/HDF5).
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
# machine_kernel and drain_kernel launch of one whole pass (300 000 node records, full operator list) divided by the
# pass's RHS counter: 2.217e11 / 13 693 801 (ncu, profiles/r01f_fp64_ops_whole_pass.txt).  The first bulk slice alone
# gives 7.8e3 (cheap, disk-less nodes are queued first), the second 16.1e3.
FLOP_PER_RHS = 16.2e3
# DRAM bytes per evaluation: ncu --set full of the second bulk slice of machine_kernel (profiles/r01e_machine_kernel_bulk_slice.txt:
# 74.28 GB read + 43.91 GB written for 7 702 358 evaluations)
DRAM_BYTES_PER_RHS = (74.277716e9 + 43.909916e9) / 7702358.0
N_Y = 24
BLACK_HOLE_FRACTION = 0.7
MW_ROOT_MASS, MW_RESOLUTION = 1.52e12, 1.0e9  # testSuite/parameters/benchmark_milkyWay.xml:30-42


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000, help="node records per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="node records of the CPU baseline sample")
    ap.add_argument("--seed", type=int, default=219)
    ap.add_argument("--trees", type=int, default=1000, help="Milky-Way-mass trees per GPU of the tree-level arm (0 = skip)")
    ap.add_argument("--cpu-trees", type=int, default=64, help="trees of the CPU tree-walk sample")
    return ap.parse_args()


def workload(n, seed):
    from galacticus_b200 import abi, synthetic
    from galacticus_b200.evolver import params_default

    p = params_default(abi.GLC_MODEL_STANDARD)  # operatorMask = GLC_OP_ALL: the full nodeOperator list of quickTest.xml
    synthetic.finalize_params(p)
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=seed, black_hole_fraction=BLACK_HOLE_FRACTION)
    return p, props, flags, t_end


WORKLOAD_NAME = ("node-batch stand-in for testSuite benchmark-milkyWay (10^3 MW-mass trees ~ 10^6 node-evolve calls): "
                 "%d node records per GPU over the quickTest mass range (1e10-1e13 Msun), each evolved over its own "
                 "0.05-0.8 Gyr interval; full quickTest nodeOperator list (incl. black-hole seed/accretion/winds and "
                 "jet-power CGM heating; 70 percent of the nodes start with a black hole, the others are seeded by interrupt), "
                 "hotHaloRamPressureStripping=virialRadius, synthetic CIE and ADAF tables")


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


def cpu_run(p, props, flags, t_end, repeats):
    """The CPU implementation of the path on all host cores; returns (steps/s, nodes/s, seconds per pass, cores)."""
    from galacticus_b200 import synthetic
    from oracle import orc

    orc.build()
    cores = os.cpu_count() or 1
    o = orc.Oracle(fast=True)
    synthetic.install(o, p)
    times, steps = [], 0
    for _ in range(repeats):
        pp, ff = props.copy(), flags.copy()
        t0 = time.perf_counter()
        _, _, c = o.evolve_batch(pp, ff, t_end, n_threads=cores)
        times.append(time.perf_counter() - t0)
        steps = c["steps_accepted"]
    best = float(np.mean(times))
    return steps / best, props.shape[0] / best, best, cores


def cpu_forest(p, n_trees, seed, cores):
    """The CPU checker's tree walk (oracle/orc_tree.c, OpenMP over trees) on a sample of the same trees."""
    from galacticus_b200 import synthetic
    from oracle import orc

    sub = synthetic.binary_split_forest(p, n_trees, MW_ROOT_MASS, MW_RESOLUTION, seed=seed)
    o = orc.Oracle(fast=True)
    synthetic.install(o, p)
    t0 = time.perf_counter()
    _, _, _, fc, c = o.forest_evolve(sub, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n_trees / dt, "unit": "merger trees/s", "trees": n_trees, "seconds": dt,
            "node_ode_steps_per_s": c["steps_accepted"] / dt}


def run_reference(args):
    """Reference arm: the CPU implementation of the path (the oracle port; the Fortran reference cannot be
    compiled in this image) on the host cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_sample
    p, props, flags, t_end = workload(n, args.seed)
    cores = os.cpu_count() or 1
    from galacticus_b200 import synthetic
    from oracle import orc

    orc.build()
    o = orc.Oracle(fast=True)
    synthetic.install(o, p)
    times, steps = [], 0
    for it in range(args.warmup + args.steps):
        pp, ff = props.copy(), flags.copy()
        t0 = time.perf_counter()
        _, _, c = o.evolve_batch(pp, ff, t_end, n_threads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            steps = c["steps_accepted"]
    tot = sum(times)
    value = steps * len(times) / tot
    sample = f"{n} node records of the same generator/seed per step; {n * len(times) / tot:.0f} nodes/s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME % args.nodes, "nodes_per_gpu": args.nodes, "seed": args.seed,
                   "sample_nodes_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.trees > 0 and args.cpu_trees > 0:
        line["trees_per_s"] = cpu_forest(p, args.cpu_trees, args.seed, cores)
        line["cpu_baseline"]["trees_per_s"] = line["trees_per_s"]
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from galacticus_b200 import abi, synthetic
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
    from galacticus_b200 import sharding as _sh

    p, props, flags, t_end = workload(n, _sh.forest_seed(args.seed, rank) if world > 1 else args.seed)
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
    kernel_ms, counters = [], None
    for _ in range(args.steps):
        ev.arena_restore(n)
        counters, ms = ev.evolve_arena(n)  # synchronises the evolver's stream; ms = CUDA events around the kernels
        kernel_ms.append(ms)
    barrier()
    wall = time.perf_counter() - t0
    launches = ev.kernel_launch_count() - launches1
    clocks = sampler.stop()
    dev_time = sum(kernel_ms) * 1e-3  # device time of the timed kernels on the launching stream
    final_props, _, st, _ = ev.arena_download(n)

    # ---------------- one instrumented pass in time slices: how the pass divides into bulk and straggler tail
    profile = {}
    try:
        ev.set_option(abi.GLC_OPT_SLICE_BUDGET, 4096)
        ev.arena_restore(n)
        tb = time.perf_counter()
        c_sl, ms_sl = ev.evolve_arena(n)
        profile = {"sliced_pass_ms": ms_sl, "slices": None}
        ev.set_option(abi.GLC_OPT_SLICE_BUDGET, 0)
    except Exception as e:  # informational only
        profile = {"error": repr(e)}
        ev.set_option(abi.GLC_OPT_SLICE_BUDGET, 0)

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
    # ---------------- tree-level arm: BASELINE configs[1], 10^3 Milky-Way-mass trees through glc_forest_evolve (trees/s)
    forest_s, forest_info = 0.0, None
    if args.trees > 0:
        fseed = _sh.forest_seed(args.seed, rank) if world > 1 else args.seed
        forest = synthetic.binary_split_forest(p, args.trees, MW_ROOT_MASS, MW_RESOLUTION, seed=fseed)
        ev.forest_evolve(synthetic.binary_split_forest(p, 4, MW_ROOT_MASS, 1.0e10, seed=1))  # warm-up
        barrier()
        t1 = time.perf_counter()
        _, _, fstate, ffc, fcnt = ev.forest_evolve(forest)
        torch.cuda.synchronize()
        forest_s = time.perf_counter() - t1
        forest_info = {"trees_per_gpu": args.trees, "nodes_per_gpu": int(forest["parent"].shape[0]),
                       "root_mass": MW_ROOT_MASS, "mass_resolution": MW_RESOLUTION,
                       "rounds": ffc["rounds"], "evolve_calls": ffc["evolve_calls"], "promotions": ffc["promotions"],
                       "node_mergers": ffc["node_mergers"], "failed_evolves": ffc["failed_evolves"],
                       "node_ode_steps": fcnt["steps_accepted"],
                       "galaxies_at_final_time": int((fstate != abi.GLC_FOREST_NODE_PROMOTED).sum())}
    fp64_peak = ev.fp64_peak_tflops()  # after the runs: the device is warm

    # ---------------- reduce over ranks: max time, summed work, NCCL all-reduce of an output statistic
    steps_acc = counters["steps_accepted"]
    stats = torch.tensor([dev_time, wall, e2e_time, float(steps_acc), float(counters["rhs_evaluations"]),
                          float(counters["steps_rejected"]), float(n), forest_s, float(args.trees),
                          float(forest_info["node_ode_steps"]) if forest_info else 0.0], dtype=torch.float64, device="cuda")
    tmax, tsum = stats.clone(), stats.clone()
    # stellar mass function histogram of the evolved batch (mirrors output/analyses/volume_function_1d.F90:986-987)
    mstar = final_props[:, abi.P["DISK_MASS_STELLAR"]] + final_props[:, abi.P["SPH_MASS_STELLAR"]]
    hist = np.histogram(np.log10(np.maximum(mstar, 1.0)), bins=30, range=(5.0, 12.5))[0].astype(np.float64)
    from galacticus_b200 import sharding

    hist_t = sharding.reduce_statistics(torch.from_numpy(hist).cuda())  # NCCL all-reduce over NVLink when world > 1
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    tmax, tsum = tmax.cpu().numpy(), tsum.cpu().numpy()
    ok_frac = float((st == 0).mean())

    if rank == 0:
        dev_time_max, e2e_max = float(tmax[0]), float(tmax[2])
        total_steps = float(tsum[3]) * args.steps
        value = total_steps / dev_time_max
        ms_per_step = 1e3 * dev_time_max / args.steps
        e2e_value = float(tsum[3]) / e2e_max
        # roofline of the dominant kernel (rank 0's launches)
        ms_kernel = float(np.mean(kernel_ms))
        rhs_total = float(counters["rhs_evaluations"])
        alg_bytes = steps_acc * 3 * 8 * N_Y + n * (2 * abi.NPROP * 8 + 8 + 3 * 4)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ach_gbs = alg_bytes / (ms_kernel * 1e-3) / 1e9
        ach_tflops = rhs_total * FLOP_PER_RHS / (ms_kernel * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD_NAME % n,
                "nodes_per_gpu": n, "seed": args.seed,
                "l2": "inputs (%.0f MB/GPU of node records + %.0f MB of per-slot continuations) larger than L2"
                      % (n * abi.NPROP * 8 / 1e6, 148 * 2048 * 4.2e3 / 1e6),
                "trees_per_s_equivalent": float(tsum[6]) * args.steps / dev_time_max / 1.0e3,
                "nodes_per_s": float(tsum[6]) * args.steps / dev_time_max,
                "rhs_evaluations_per_s": float(tsum[4]) * args.steps / dev_time_max,
                "rejected_step_fraction": float(tsum[5]) / max(float(tsum[3]) + float(tsum[5]), 1.0),
                "status_ok_fraction": ok_frac,
                "wall_s_timed_region": float(tmax[1]),
                "sliced_pass": profile,
                "stellar_mass_function_counts": hist_t.cpu().numpy().tolist(),
            },
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         # DRAM traffic of the dominant kernel: ncu --set full capture of one machine_kernel launch
                         # (15.3 kB per evaluation: continuation records streaming through L2), scaled to this pass's
                         # evaluation count
                         "traffic": DRAM_BYTES_PER_RHS * rhs_total,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                         "kernel": "machine_kernel (+ its drain_kernel continuation)",
                         "note": "the path is FP64-ALU / latency bound (SURVEY 8d: >15 flop per algorithmic byte): "
                                 "see roofline_fp64; algorithmic bytes = 576 B per accepted step + one read and one "
                                 "write of each node record"},
            "roofline_fp64": {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (ach_tflops / fp64_peak) if fp64_peak else None, "flop_per_rhs": FLOP_PER_RHS,
                              "peak_source": "DFMA-chain microbenchmark in this process (glc_measure_fp64_peak_tflops)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            # the other half of BASELINE.json's metric ("trees/sec & node-ODE steps/sec"), end to end through the C-ABI
            # (host scheduler + batched node evolver, host buffers): all ranks' trees / slowest rank's time
            "trees_per_s": ({"value": float(tsum[8]) / float(tmax[7]), "unit": "merger trees/s",
                             "node_ode_steps_per_s": float(tsum[9]) / float(tmax[7]), "seconds": float(tmax[7]),
                             "workload": "Milky-Way-mass binary-split trees (root %.3g Msun, resolution %.3g Msun: the masses of "
                                         "testSuite/parameters/benchmark_milkyWay.xml), quickTest physics" % (MW_ROOT_MASS, MW_RESOLUTION),
                             **forest_info} if forest_info else None),
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        # CPU baseline (rank 0, N=1 only): the CPU implementation on a bounded sample
        if world == 1:
            try:
                ns = min(args.cpu_sample, n)
                v, nps, secs, cores = cpu_run(p, props[:ns], flags[:ns], t_end[:ns], repeats=1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": "first %d node records of the GPU workload, %.1f s, %.0f nodes/s"
                                                  % (ns, secs, nps)}
                if forest_info and args.cpu_trees > 0:
                    line["cpu_baseline"]["trees_per_s"] = cpu_forest(p, args.cpu_trees, args.seed, cores)
            except Exception as e:  # the checker is optional for the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
