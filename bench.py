#!/usr/bin/env python
"""Benchmark of the action-evaluation hot path (BASELINE.json metric: bead-pair action evals/s
+ MC sweeps/s, UEG N=256 M=128, 1024 clones per B200).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation

A step = one full energy evaluation of every clone: rebuild rho_k from the positions (K2),
IlkkaPairAction::DActionDBeta over all pairs and slices with the long-range r-space
subtraction (K1), the Ewald k-space sum and constants (K3 + finalize).  One bead-pair action
eval = one DrDrpDrrp + one CalcdUdBeta, 273 flop by the stated convention.

`value`   inputs resident in HBM, K steps timed with CUDA events on the library's stream.
`e2e`     the same step through the C ABI with HOST buffers: pimc_positions_upload from pinned
          host memory + pimc_action_dbeta returning host doubles, every step.
Both arms print one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EVAL = 273.0   # Ilkka bead-pair eval with long-range subtraction
N_PART, N_SLICE = 256, 128
BISECT_LEVEL = 3


def pair_evals_per_clone(N=N_PART, M=N_SLICE):
    return N * (N - 1) // 2 * M


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def ensure_samples(self, step, sync, n_min=2, max_s=3.0):
        """Keep the GPU under the same load (untimed extra steps) until nvidia-smi has delivered
        n_min samples: a timed region shorter than its sampling period would otherwise have none."""
        if not self.proc:
            return
        t0 = time.perf_counter()
        while len(self.lines) < n_min and time.perf_counter() - t0 < max_s:
            step()
            sync()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": "warm-up + timed steps (+ identical untimed steps until nvidia-smi has delivered 2 samples)"}


# ---------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    """One process = one walker (the reference's MPI model, framework_class.h:44-53)."""
    kind, clone, n_eval, N, M = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from simpimc_b200 import system as S
    cfg = S.ueg_config(N=N, M=M)
    R = S.synthetic_paths(cfg, 0, clone)
    n_att, t_att = 0, 0.0
    if kind == "reference":
        from oracle import refsim
        sim = refsim.RefSim(cfg, fast=refsim.available(fast=True))
        sim.set_positions(0, R)
        f = lambda: sim.dbeta(0)
        # the reference's own Bisect::DoEvent (Kinetic + IlkkaPairAction, n_level as the GPU sweep)
        cfg_mc = S.ueg_config(N=N, M=M, with_kinetic=True)
        cfg_mc.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": BISECT_LEVEL}]
        mc = refsim.RefSim(cfg_mc, seed=1000 + clone, fast=refsim.available(fast=True))
        mc.set_positions(0, R)
        mc.move_do(0, 20)
        n_att = 100 * max(1, n_eval)
        t0 = time.perf_counter()
        mc.move_do(0, n_att)
        t_att = time.perf_counter() - t0
        mc.close()
    else:
        from oracle import oracle as O
        sim = O.Oracle(cfg)
        sim.set_positions(0, R)
        f = lambda: sim.dbeta(0)
    f()
    t0 = time.perf_counter()
    val = 0.0
    for _ in range(n_eval):
        val = f()
    return time.perf_counter() - t0, val, n_att, t_att


def cpu_baseline(n_eval=1, N=N_PART, M=N_SLICE, cores=None):
    """DActionDBeta() of the CPU implementation on every host core at once: `cores` independent
    single-clone processes, each timing n_eval full evaluations after one warm-up."""
    import multiprocessing as mp
    from oracle import refsim
    kind = "reference" if refsim.available() else "port"
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, c, n_eval, N, M) for c in range(cores)])
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    evals = cores * n_eval * pair_evals_per_clone(N, M)
    out = {"value": evals / slowest, "unit": "bead-pair action evals/s", "cores": cores, "kind": kind,
           "sample": "%d clone(s) x %d DActionDBeta() of UEG N=%d M=%d, one process per core, %.1f s wall incl. setup"
                     % (cores, n_eval, N, M, wall)}
    out["_values"] = [r[1] for r in res]   # DActionDBeta() of walker c as the CPU implementation computed it (parity block)
    if res[0][2] > 0:
        slowest_mc = max(r[3] for r in res)
        out["mc_sweeps_per_s"] = cores * res[0][2] / (N * M // (1 << BISECT_LEVEL)) / slowest_mc
        out["mc_sample"] = "%d reference Bisect::DoEvent (n_level=%d, Kinetic + IlkkaPairAction) per core" % (res[0][2], BISECT_LEVEL)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_evals = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(n_eval=1)
    base = cpu_baseline(n_eval=steps_evals)
    base.pop("_values", None)
    ms = 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    line = {"impl": "reference", "metric": "bead-pair action evals/s", "value": base["value"], "unit": "bead-pair action evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UEG N=256 M=128 IlkkaPairAction use_long_range=1: DActionDBeta() per clone, one clone per host core",
                       "step": "one DActionDBeta() per core"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "bead-pair action evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from simpimc_b200 import host, moves, system as S, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C = args.clones
    cfg = S.ueg_config(N=N_PART, M=N_SLICE)
    path = host.Path(cfg, n_clones=C, device=local)
    act = path.actions[0]
    lib = path.L
    # synthetic walkers: clone index offset by rank so every GPU holds different walkers
    R_pinned = torch.empty((C, N_PART, N_SLICE, 3), dtype=torch.float64, pin_memory=True)
    R = R_pinned.numpy()
    for c in range(C):
        R[c] = S.synthetic_paths(cfg, 0, rank * C + c)
    path.SetPositions(0, R)
    out_dev = torch.zeros(C, dtype=torch.float64, device="cuda")
    out_host = np.zeros(C)
    stream = torch.cuda.ExternalStream(lib.pimc_ctx_stream(path.h), device=torch.device("cuda", local))

    def step_resident():
        capi.check(lib.pimc_rhok_rebuild(path.h, 0))
        capi.check(lib.pimc_action_dbeta_device(act.h, out_dev.data_ptr()))

    # end-to-end pipeline: the same walkers in N_PIPE contexts (own stream each), so that the
    # host->device copy of one group overlaps the kernels of the previous one; every call is the
    # public C ABI with HOST buffers (pimc_positions_upload from pinned memory, pimc_action_dbeta
    # returning host doubles)
    n_pipe = 1 if args.resident_only else max(1, min(args.pipeline, C))
    while C % n_pipe:
        n_pipe -= 1
    Cq = C // n_pipe
    pipes = [path] if n_pipe == 1 else [host.Path(cfg, n_clones=Cq, device=local) for _ in range(n_pipe)]
    out_dev_e2e = torch.zeros(C, dtype=torch.float64, device="cuda")

    def step_e2e():
        for i, pp in enumerate(pipes):   # asynchronous on each context's stream
            capi.check(lib.pimc_positions_upload(pp.h, 0, 0, Cq, R[i * Cq:(i + 1) * Cq].ctypes.data))
            capi.check(lib.pimc_action_dbeta_device(pp.actions[0].h, out_dev_e2e[i * Cq:(i + 1) * Cq].data_ptr()))
        for i, pp in enumerate(pipes):   # results back to the host
            capi.check(lib.pimc_ctx_sync(pp.h))
        out_host[:] = out_dev_e2e.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        path.Sync()

    fp64_peak = path.Fp64Peak()
    # clocks are sampled from the warm-up on (same load as the timed steps): nvidia-smi needs a few
    # hundred ms for its first line, longer than a short timed region
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        step_resident()
    path.Sync()
    # ---- device-resident timing ------------------------------------------------------------
    path.SetTiming(True)
    barrier()
    launches0 = path.LaunchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    e1.synchronize()
    barrier()
    launches = path.LaunchCount() - launches0
    ms_total = e0.elapsed_time(e1)
    k1_ms, k1_n = path.KernelTime(1)
    k2_ms, k2_n = path.KernelTime(2)
    k3_ms, k3_n = path.KernelTime(3)
    path.SetTiming(False)
    clocks.ensure_samples(step_resident, path.Sync)
    clk = clocks.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    evals_step = C * pair_evals_per_clone()
    value = world * evals_step * args.steps / (ms_max * 1e-3)
    if args.resident_only:   # profiling aid (ncu launch list of the step kernels alone): not a bench line
        if rank == 0:
            print(json.dumps({"resident_only": True, "value": value, "ms_per_step": ms_max / args.steps,
                              "kernel_ms": {"K1": k1_ms / max(1, k1_n), "K2": k2_ms / max(1, k2_n), "K3": k3_ms / max(1, k3_n)},
                              "kernel_share_of_step": {"K1": k1_ms / ms_total, "K2": k2_ms / ms_total, "K3": k3_ms / ms_total}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end to end through the C ABI with host buffers -------------------------------------
    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    # several streams are involved: the timed region is bracketed by full device synchronisations
    # and measured on the host clock (which then cannot be shorter than the device time)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    barrier()
    te = torch.tensor([1e3 * (w1 - w0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * evals_step * args.steps / (float(te.item()) * 1e-3)
    # correctness guard for the bench itself: both paths computed the same energies
    assert np.allclose(out_dev.cpu().numpy(), out_host, rtol=1e-12, atol=0), "resident and e2e energies differ"
    # ---- MC sweeps: device-resident bisection (pimc_bisect_sweep: Levy sampling, kinetic + pair +
    # long-range action deltas, Metropolis, commit -- no host round trip per attempt) ---------------
    n_att = args.attempts
    path.BisectSweep(0, BISECT_LEVEL, 8, 1234 + rank, attempt0=0)   # warm-up
    barrier()
    path.SetTiming(True)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_mc0 = path.LaunchCount()
    m0.record(stream)
    n_acc = path.BisectSweep(0, BISECT_LEVEL, n_att, 1234 + rank, attempt0=8)
    m1.record(stream)
    m1.synchronize()
    barrier()
    launches_mc = path.LaunchCount() - launches_mc0
    k4_ms, k4_n = path.KernelTime(4)
    path.SetTiming(False)
    ts = torch.tensor([m0.elapsed_time(m1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    attempts_per_sweep = N_PART * N_SLICE // (1 << BISECT_LEVEL)
    sweeps_per_s = world * C * n_att / attempts_per_sweep / float(ts.item())
    window_evals_per_s = world * C * n_att * 2 * (N_PART - 1) * (1 << BISECT_LEVEL) / float(ts.item())
    accept_ratio = float(n_acc.sum()) / (C * n_att)
    # device-resident DisplaceParticle (whole-path shift of one particle per clone and attempt)
    n_disp = max(1, min(16, n_att))
    path.DisplaceSweep(0, cfg.L / 10.0, 2, 4321 + rank, attempt0=0)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record(stream)
    n_acc_d = path.DisplaceSweep(0, cfg.L / 10.0, n_disp, 4321 + rank, attempt0=2)
    d1.record(stream)
    d1.synchronize()
    barrier()
    td = torch.tensor([d0.elapsed_time(d1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    displace = {"ms_per_attempt": 1e3 * float(td.item()) / n_disp, "attempts_timed": n_disp, "accept_ratio": float(n_acc_d.sum()) / (C * n_disp),
                "pair_evals_per_s": world * C * n_disp * 2 * (N_PART - 1) * N_SLICE / float(td.item()),
                "driver": "device-resident pimc_displace_sweep (step L/10; pair + long-range deltas over all slices, Metropolis, commit)"}
    # device-resident permuting bisection (PermBisectIterative::DoEvent: cycle selection from the permutation table,
    # the members' Levy bridges, pair + long-range deltas over the listed labels, Metropolis, relabelling)
    n_pb = max(1, min(16, n_att))
    path.PermBisectSweep(0, BISECT_LEVEL, 2, 777 + rank, attempt0=0)
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_pb0 = path.LaunchCount()
    q0.record(stream)
    n_acc_p, att_len, acc_len = path.PermBisectSweep(0, BISECT_LEVEL, n_pb, 777 + rank, attempt0=2)
    q1.record(stream)
    q1.synchronize()
    barrier()
    tp = torch.tensor([q0.elapsed_time(q1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    displace["perm_bisect"] = {"ms_per_attempt": 1e3 * float(tp.item()) / n_pb, "attempts_timed": n_pb, "launches": int(path.LaunchCount() - launches_pb0),
                               "accept_ratio": float(n_acc_p.sum()) / (C * n_pb),
                               "cycles_attempted_by_length": [int(x) for x in att_len.sum(axis=0)], "cycles_accepted_by_length": [int(x) for x in acc_len.sum(axis=0)],
                               "driver": "device-resident pimc_perm_bisect_sweep (kernel-per-phase: select, sample, pair OLD / NEW, long range, decide + commit, relabel)"}
    # back to the unpermuted walkers for what follows
    path.SetPermutation(0, np.arange(N_PART, dtype=np.int32))
    path.SetPositions(0, R)
    # the same move driven from the host through propose / GetAction(OLD, NEW) / commit (the
    # reference-shaped call sequence), a few attempts for comparison
    rng = np.random.default_rng(1234 + rank)
    bis = moves.Bisect(path, rng, 0, BISECT_LEVEL)
    bis.DoEvent()
    path.Sync()
    s0 = time.perf_counter()
    for _ in range(4):
        bis.DoEvent()
    path.Sync()
    host_driven_sweeps_per_s = C * 4 / attempts_per_sweep / (time.perf_counter() - s0)
    # ---- estimators: g(r) histogram (K5) and S(k) (K6) over every clone ---------------------------
    gr = host.PairCorrelation(path, 0, 0, 0.0, cfg.L / 2.0, 100)
    sk = host.StructureFactor(path, 0, 0, cfg.k_cut)
    gr.Accumulate()
    sk.Accumulate()
    path.SetTiming(True)
    for _ in range(2):
        gr.Accumulate()
        sk.Accumulate()
    k5_ms, k5_n = path.KernelTime(5)
    k6_ms, k6_n = path.KernelTime(6)
    path.SetTiming(False)
    estimators = {"gofr_kernel_ms": k5_ms / max(1, k5_n), "gofr_pair_slices_per_s": C * pair_evals_per_clone() / (k5_ms / max(1, k5_n) * 1e-3),
                  "sofk_kernel_ms": k6_ms / max(1, k6_n), "sofk_gbs": C * N_SLICE * 182 * 16 / (k6_ms / max(1, k6_n) * 1e-3) / 1e9,
                  "unit": "one PairCorrelation::Accumulate / StructureFactor::Accumulate over all clones (kernel time, CUDA events)"}
    if rank != 0:
        # BASELINE config C5 (below) is collective: every rank runs it, under the same watchdog as rank 0
        if not args.no_sharded:
            _run_sharded_leg(args, rank, world, local, None)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- the other two action families at the same shape (whole-path kernel times, 256 clones) ----
    families = None
    if world == 1 and args.cpu_evals > 0:   # skipped in the profiling runs (--cpu-evals 0)
        families = {"unit": "kernel ms per whole-path evaluation of %d clones (CUDA events), UEG N=%d M=%d" % (min(C, 256), N_PART, N_SLICE)}
        try:
            Cf = min(C, 256)
            for fam, kw in (("BarePairAction", dict(use_long_range=True)), ("DavidPairAction", dict(use_long_range=False))):
                fcfg = S.ueg_config(N=N_PART, M=N_SLICE, action=fam, **kw)
                fpath = host.Path(fcfg, n_clones=Cf, device=local)
                fpath.SetPositions(0, R[:Cf])
                fact = fpath.actions[0]
                res = {}
                for name, fn in (("DActionDBeta", fact.DActionDBeta), ("Potential", fact.Potential)):
                    fn()
                    fpath.Sync()
                    fpath.SetTiming(True)
                    fn()
                    fn()
                    fpath.Sync()
                    kms, kn = fpath.KernelTime(1)
                    fpath.SetTiming(False)
                    res[name + "_kernel_ms"] = kms / max(1, kn)
                    if name == "DActionDBeta":
                        res["evals_per_s"] = Cf * pair_evals_per_clone() / (kms / max(1, kn) * 1e-3)
                families[fam] = res
                fpath.close()
        except Exception as e:   # never let the side measurement take the headline line down
            families["error"] = repr(e)[:200]
    # ---- roofline of the dominant kernel (K1) ------------------------------------------------
    k1_avg_s = (k1_ms / max(1, k1_n)) * 1e-3
    achieved = evals_step * FLOP_PER_EVAL / k1_avg_s / 1e12
    traffic = None
    tf = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp64", "kernel": "pair_full_fast_kernel (IlkkaPairAction::CalcdUdBeta over all pairs x slices)", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                "peak_source": "DFMA micro-benchmark (pimc_fp64_peak) measured in this run; MEASURED_PEAKS.json holds no FP64 figure",
                "bound_note": "north_star names the FP64-or-HBM roofline: 273 flop per 0.19 algorithmic bytes puts this kernel on the FP64 side (DRAM at 0.2 % of peak in profiles/); no tensor-core work on this path",
                "algorithmic_flop_per_launch": evals_step * FLOP_PER_EVAL,
                "algorithmic_bytes_per_launch": C * N_SLICE * N_PART * 24,
                "kernel_ms": {"K1_pair_full": k1_ms / max(1, k1_n), "K2_rhok_build": k2_ms / max(1, k2_n), "K3_ksum": k3_ms / max(1, k3_n)},
                "kernel_share_of_step": {"K1": k1_ms / ms_total, "K2": k2_ms / ms_total, "K3": k3_ms / ms_total}}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm = peaks.get("hbm_gbs")
    except Exception:
        hbm = 6650.0
    roofline["hbm_gbs_peak"] = hbm
    roofline["k1_hbm_gbs_algorithmic"] = C * N_SLICE * N_PART * 24 / k1_avg_s / 1e9
    base = cpu_baseline(n_eval=args.cpu_evals) if (args.cpu_evals > 0 and world == 1) else None
    parity = None
    if base:
        # the CPU implementation evaluated the walkers 0..cores-1 the GPU holds as clones 0..cores-1: compare the
        # energies the timed e2e step returned with the reference's own numbers (north_star tolerance 1e-10)
        cpu_vals = np.array(base.pop("_values"))
        n_cmp = min(len(cpu_vals), C)
        rel = np.abs(out_host[:n_cmp] - cpu_vals[:n_cmp]) / np.abs(cpu_vals[:n_cmp])
        parity = {"clones": int(n_cmp), "max_rel_err": float(rel.max()), "tolerance": 1e-10, "quantity": "DActionDBeta() per clone: C-ABI e2e result vs cpu_baseline (kind=%s)" % base["kind"],
                  "ok": bool(rel.max() <= 1e-10)}
        assert parity["ok"], "bench parity: GPU and CPU DActionDBeta differ by %.3e" % rel.max()
    line = {"metric": "bead-pair action evals/s", "value": value, "unit": "bead-pair action evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UEG N=256 M=128, IlkkaPairAction use_long_range=1 (n_k=182), %d clones per GPU" % C,
                       "step": "rho_k rebuild + DActionDBeta (pair action over all pairs x slices + Ewald k-sum + constants)",
                       "clones_per_gpu": C, "parallelism": "independent clones per GPU, no data-path collective",
                       "l2": "inputs larger than L2 (positions %.0f MB + rho_k %.0f MB per GPU)" % (R.nbytes / 1e6, C * N_SLICE * 182 * 16 / 1e6),
                       "bisect_n_level": BISECT_LEVEL},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "bead-pair action evals/s", "h2d_bytes_per_step": int(R.nbytes),
                    "d2h_bytes_per_step": int(out_host.nbytes),
                    "pipeline": "%d contexts x %d clones, upload of one overlapping the kernels of the previous" % (n_pipe, Cq)},
            "gpu_launches": int(launches),
            "mc_sweeps_per_s": sweeps_per_s,
            "mc": {"unit": "clone-sweeps/s (one sweep = N*M/2^n_level bisection attempts)", "attempts_timed": n_att,
                   "accept_ratio": accept_ratio,
                   "accept_note": "an accepted attempt is the expensive branch (positions and rho_k of the window are committed); a rejected one skips phase E and, when it dies above level 0, the pair and k-space phases too -- a high ratio is the conservative timing",
                   "window_pair_evals_per_s": window_evals_per_s,
                   "ms_per_attempt": 1e3 * float(ts.item()) / n_att, "launches": int(launches_mc),
                   "pair_window_kernel_ms_per_attempt": k4_ms / max(1, n_att),
                   "driver": "device-resident pimc_bisect_sweep (Philox stream; kinetic + Ilkka pair + long-range deltas, Metropolis, commit)",
                   "host_driven_sweeps_per_s_per_gpu": host_driven_sweeps_per_s, "displace": {k: v for k, v in displace.items() if k != "perm_bisect"},
                   "perm_bisect": displace["perm_bisect"]},
            "estimators": estimators,
            "roofline": roofline}
    if families:
        line["other_families"] = families
    if base:
        line["cpu_baseline"] = base
    if parity:
        line["parity"] = parity
    # ---- BASELINE config C5: ONE large path sharded by imaginary-time slice over the ranks (NCCL halo +
    # all-reduce behind the C ABI), printed inside the same line.  It runs last and under a watchdog: a
    # collective that never completes must not cost the headline line ----
    if not args.no_sharded:
        line["sharded"] = _run_sharded_leg(args, rank, world, local, line)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _run_sharded_leg(args, rank, world, local, line):
    """c5_leg under a watchdog.  If it does not return within --sharded-timeout seconds (a hung collective), rank 0
    prints the line it has with sharded = {"error": ...} and every rank leaves with os._exit(0)."""
    def bail():
        if rank == 0 and line is not None:
            line["sharded"] = {"error": "slice-sharded leg did not finish within %d s (watchdog)" % args.sharded_timeout}
            print(json.dumps(line), flush=True)
        sys.stdout.flush()
        os._exit(0)

    dog = threading.Timer(args.sharded_timeout, bail)
    dog.daemon = True
    dog.start()
    try:
        block = c5_leg(args, rank, world, local)
    except Exception as e:   # setup errors are identical on every rank; never take the headline line down
        block = {"error": repr(e)[:300]}
    dog.cancel()
    return block


# --------------------------------------------- config C5: one large path, slices sharded over the GPUs
def c5_leg(args, rank, world, local, with_mc=True):
    """BASELINE config C5: dense hydrogen plasma, 1024 e + 1024 p, M = 512, three pair actions
    (e-e, e-p, p-p; Ilkka tables with long range), ONE path whose time slices are sharded over
    the ranks (strong scaling).  Step = rho_k rebuild of both species + DActionDBeta of the three
    actions on the local slices + ONE NCCL all-reduce of the partial sums -- every collective is the
    library's own (C ABI: pimc_sharded_evaluate / pimc_halo_exchange / pimc_rotate), the whole step
    replayed from one captured CUDA graph.  The end-to-end leg uploads the shard's positions from
    pinned host memory, fills the halo slice over the NCCL ring and reads the energies back.
    Collective: every rank calls it; returns the result block (meaningful on rank 0)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from simpimc_b200 import host, sharded, system as S, capi

    Ne, M, C = args.c5_n, args.c5_m, args.c5_clones
    cfg = S.plasma_config(Ne=Ne, Np=Ne, M=M, n_xy=100, n_r_long=1000, pp_action="IlkkaPairAction")
    sp_path = sharded.ShardedPath(cfg, C, local, rank, world)
    path, lib = sp_path.path, sp_path.path.L
    sh = sp_path.sh
    fulls, shards, pinned = [], [], []
    for sp in range(2):
        full = np.stack([S.synthetic_paths(cfg, sp, c, 777) for c in range(C)])   # same walkers on every rank
        fulls.append(full)
        t = torch.empty((C, Ne, path.n_store, 3), dtype=torch.float64, pin_memory=True)
        t.zero_()
        if world == 1:
            t.numpy()[...] = full
        else:
            t.numpy()[:, :, :sh.n_local, :] = full[:, :, sh.lo:sh.hi, :]      # the halo slot is filled by the ring exchange
        pinned.append(t)
        shards.append(t.numpy())
    n_act = 3
    out_dev = torch.zeros((n_act, C), dtype=torch.float64, device="cuda")
    out_host = np.zeros((n_act, C))
    stream = sp_path.stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # 2 x L2

    def upload_and_halo():
        for sp in range(2):
            path.SetPositions(sp, shards[sp])
            sp_path.ExchangeHalo(sp)

    def step_eager():
        sp_path.RebuildRhoK()
        sp_path.DActionDBetaAllDevice(out_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        path.Sync()

    upload_and_halo()
    path.Sync()
    fp64_peak = path.Fp64Peak()
    step_eager()
    path.Sync()
    energies_eager = out_dev.cpu().numpy().copy()
    step_graph = sp_path.CaptureStep(out_dev)

    def step_e2e():
        upload_and_halo()
        step_graph()
        with torch.cuda.stream(stream):
            out_host[...] = out_dev.cpu().numpy()

    n_steps = max(1, args.steps) * args.c5_mult
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(3, args.warmup)):
        step_graph()
    path.Sync()

    def timed(step, n, flush_l2):
        """Sum over n iterations of the device time of `step` (event pair per iteration on the context's
        stream; the L2 flush between iterations sits outside the pairs), max over ranks."""
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            if flush_l2:
                with torch.cuda.stream(stream):
                    flush.zero_()
            a.record(stream)
            step()
            b.record(stream)
        evs[-1][1].synchronize()
        barrier()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    launches0 = path.LaunchCount()
    ms_graph = timed(step_graph, n_steps, True)
    launches = path.LaunchCount() - launches0
    ms_eager = timed(step_eager, n_steps, True)
    # per-kernel shares (eager steps, per-kernel CUDA events: not part of the timed figure)
    path.SetTiming(True)
    for _ in range(5):
        step_eager()
    path.Sync()
    k1_ms, _ = path.KernelTime(1)
    k2_ms, _ = path.KernelTime(2)
    k3_ms, _ = path.KernelTime(3)
    path.SetTiming(False)
    # keep the same load on until nvidia-smi has had time for a few samples.  Every step holds a collective, so the
    # number of extra steps must be the SAME on every rank: it is derived from the all-reduced time, never from
    # this rank's own sample count (ClockSampler.ensure_samples would deadlock the ranks against each other)
    n_extra = min(2000, int(700.0 / max(ms_graph / n_steps, 1e-3)) + 1)
    for _ in range(n_extra):
        step_graph()
    path.Sync()
    clk = clocks.stop()
    pairs = Ne * (Ne - 1) // 2 * 2 + Ne * Ne
    evals_step = C * pairs * M            # whole path, all ranks together
    # the step replayed from the CUDA graph and the same step launched eagerly compute the same numbers (checked below);
    # the line reports the faster of the two and says which (the graph wins while the step is launch-bound, eager launches
    # run ahead of the GPU once K1 dominates and avoid the node-to-node gaps of a replay)
    ms_best = min(ms_graph, ms_eager)
    value = evals_step * n_steps / (ms_best * 1e-3)
    energies = out_dev.cpu().numpy().copy()
    # NCCL may pick another reduction order for the captured all-reduce than for the eager one: equal to rounding, not bit for bit
    graph_vs_eager = float(np.max(np.abs(energies - energies_eager) / np.abs(energies_eager)))
    # ---- end to end: host buffers in, host doubles out, every step ----
    bytes0 = sp_path.BytesSent()
    for _ in range(2):
        step_e2e()
    barrier()
    halo_bytes = (sp_path.BytesSent() - bytes0) // 2
    n_e2e = max(5, n_steps // 4)
    w0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    barrier()
    te = torch.tensor([1e3 * (w1 - w0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = evals_step * n_e2e / (float(te.item()) * 1e-3)
    assert np.allclose(energies, out_host, rtol=1e-12, atol=0), "resident and e2e energies differ"
    # ---- identity check: the all-reduced shard sums against the unsharded evaluation on rank 0; on one
    # GPU, eight in-process slice shards (the 8-GPU layout) summed against the whole path ----
    identity = None
    if rank == 0:
        if world > 1:
            whole = host.Path(cfg, n_clones=C, device=local)
            for sp in range(2):
                whole.SetPositions(sp, fulls[sp])
            ref = np.stack([a.DActionDBeta() for a in whole.actions])
            whole.close()
            identity = {"against": "unsharded evaluation of the same path on rank 0", "max_rel_err": float(np.max(np.abs(energies - ref) / np.abs(ref)))}
        else:
            parts = np.zeros_like(energies)
            for g in range(8):
                shg = sharded.SliceSharding(M, 8, g)
                pg = host.Path(cfg, n_clones=C, device=local, slice_lo=shg.lo, slice_hi=shg.hi)
                for sp in range(2):
                    pg.SetPositions(sp, shg.shard_positions(fulls[sp]))
                parts += np.stack([a.DActionDBeta() for a in pg.actions])
                pg.close()
            identity = {"against": "sum of 8 in-process slice shards (the 8-GPU layout)", "max_rel_err": float(np.max(np.abs(parts - energies) / np.abs(energies)))}
        identity["ok"] = bool(identity["max_rel_err"] <= 1e-10)
    # ---- moves on the sharded path: shard-interior bisection windows on every rank at once (no
    # communication), then one ring rotation of the slices over NCCL (positions, halos, rho_k) ----
    mc = None
    if with_mc and sh.n_local >= (1 << BISECT_LEVEL) and args.attempts > 0:
        n_att = max(4, min(args.attempts, 64))        # rounds: every disjoint window of the shard is attempted in each
        n_win = 1
        for sp in range(2):
            _, n_win = sp_path.BisectSweepWindows(sp, BISECT_LEVEL, 2, 99, attempt0=0)
        barrier()
        w0 = time.perf_counter()
        n_acc = 0
        for sp in range(2):
            n_acc += int(sp_path.BisectSweepWindows(sp, BISECT_LEVEL, n_att, 99, attempt0=2)[0].sum())
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        sp_path.Rotate(BISECT_LEVEL + 1)
        torch.cuda.synchronize()
        w2 = time.perf_counter()
        barrier()
        tm = torch.tensor([w1 - w0, w2 - w1], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        attempts_per_sweep = 2 * Ne * M // (1 << BISECT_LEVEL)
        mc = {"attempts_per_s": world * C * 2 * n_att * n_win / float(tm[0].item()), "sweeps_per_s": world * C * 2 * n_att * n_win / attempts_per_sweep / float(tm[0].item()),
              "attempts_timed_per_rank": 2 * n_att * C * n_win, "windows_per_launch": n_win, "accept_ratio_rank0": n_acc / (2.0 * n_att * C * n_win), "rotate_ms": 1e3 * float(tm[1].item()),
              "driver": "pimc_bisect_sweep_windows (kernel-per-phase path: 2 species, 3 actions): every disjoint window of %d slices of each rank's shard attempted in the same launches; pimc_rotate over NCCL" % (1 << BISECT_LEVEL)}
    k1_step_s = k1_ms / 5 * 1e-3          # the three K1 launches of a step on this rank
    achieved = (evals_step / world) * FLOP_PER_EVAL / k1_step_s / 1e12 if k1_step_s > 0 else None
    block = {"metric": "bead-pair action evals/s", "value": value, "unit": "bead-pair action evals/s", "n_gpus": world,
             "steps": n_steps, "ms_per_step": ms_best / n_steps, "step_mode": "graph" if ms_graph <= ms_eager else "eager",
             "graph_ms_per_step": ms_graph / n_steps, "eager_ms_per_step": ms_eager / n_steps, "scaling": "strong",
             "config": {"workload": "C5 dense hydrogen plasma %d e + %d p, M=%d, 3 Ilkka pair actions with long range (n_k=%d), %d path(s), slices sharded %d per GPU"
                                    % (Ne, Ne, M, path.n_k, C, sh.n_local),
                        "step": "rho_k rebuild (2 species) + DActionDBeta of 3 actions on the local slices + one NCCL all-reduce of %d doubles; timed both replayed from one CUDA graph (%d nodes) and launched eagerly, the faster reported (step_mode)" % (n_act * C, step_graph.n_nodes),
                        "parallelism": "slice sharding; ring halo of one slice per species (ncclSend/ncclRecv) + all-reduce (ncclAllReduce), both issued by the library behind the C ABI on the context's stream",
                        "l2": "flushed between timed iterations (256 MB memset outside the per-iteration event pairs)"},
             "clocks": clk,
             "e2e": {"value": e2e_value, "unit": "bead-pair action evals/s", "h2d_bytes_per_step": int(2 * shards[0].nbytes),
                     "d2h_bytes_per_step": int(out_host.nbytes), "steps": n_e2e},
             "halo_bytes": int(halo_bytes) if world > 1 else 0,
             "halo_note": "payload this rank hands to ncclSend / ncclAllReduce per e2e step (2 species x one slice of positions + the partial sums)",
             "gpu_launches": int(launches),
             "energies_match": identity,
             "graph_replay_max_rel_diff_vs_eager": graph_vs_eager,
             "energies": {"dU/dbeta per action (clone 0)": [float(x) for x in energies[:, 0]]},
             "mc": mc,
             "roofline": {"bound": "fp64", "kernel": "pair_full_fast_kernel x 3 actions", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": achieved / fp64_peak if (fp64_peak and achieved) else None, "traffic": None,
                          "kernel_ms_per_step": {"K1_pair_full": k1_ms / 5, "K2_rhok_build": k2_ms / 5, "K3_ksum": k3_ms / 5}}}
    sp_path.close()
    del flush
    return block


def run_c5(args):
    """`--workload c5`: the slice-sharded leg alone, as its own bench line."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    block = c5_leg(args, rank, world, local)
    if rank == 0:
        block.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
        print(json.dumps(block), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clones", type=int, default=int(os.environ.get("BENCH_CLONES", "1024")))
    ap.add_argument("--attempts", type=int, default=256, help="bisection attempts per clone timed for the MC-sweep figure")
    ap.add_argument("--pipeline", type=int, default=16, help="contexts the end-to-end leg splits the clones over (H2D/compute overlap)")
    ap.add_argument("--cpu-evals", type=int, default=8, help="DActionDBeta() calls per core for cpu_baseline (0 = skip)")
    ap.add_argument("--resident-only", action="store_true", help="time the resident steps only and print a reduced line (profiling aid)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"], help="c3: UEG N=256 M=128 x clones (headline); c5: plasma 1024+1024, M=512, slices sharded over the GPUs")
    ap.add_argument("--c5-n", type=int, default=1024)
    ap.add_argument("--c5-m", type=int, default=512)
    ap.add_argument("--c5-clones", type=int, default=1)
    ap.add_argument("--c5-mult", type=int, default=10, help="evaluation steps of the slice-sharded leg per --steps (its step is ~1.5-11 ms)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the slice-sharded C5 leg of the default line (profiling runs)")
    ap.add_argument("--sharded-timeout", type=int, default=240, help="watchdog of the slice-sharded leg in seconds")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
