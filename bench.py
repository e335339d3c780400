#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): edit-distance cell-updates/s (GCUPS), cells = sum over pairs of
ref_len * hyp_len on the valid (eos-terminated) lengths; N-best hyps/s is reported
beside it.  Workload (BASELINE.json configs[1], "MinimumErrorRateLoss /
prefix_error_rates: batch 64 x 8-best word hyps, T=100, vocab 10k"): the per-pair shape
is the named one; the batch is replicated to 16384 utterances x 8-best = 131072 pairs
per GPU so the chip is full and the inputs (212 MB int64) exceed the 126 MB L2
(SURVEY 8d-iii).  The literal 64 x 8 batch is timed too and reported under "literal".

A step = one `prefix_error_rates(ref (x) 8, hyp, eos=0)` call = 7 kernel launches: pack ref,
pack hyp (+ length histogram), bucketing (scan + scatter), the wavefront DP kernel in its
two builds (32-bit / packed 16x2; the device-side token range picks one, the other exits at
once), the prefix finalize (normalise + transpose, 128-bit stores) and the stand-by 64-bit
token kernel (exits at once).

  value     device-resident inputs, CUDA events, max over ranks (whole job, all GPUs)
  e2e       the same public call with HOST (pinned) int64 tensors: H2D of the inputs and
            D2H of the (H+1, N) result inside the timed region
  roofline  the DP kernel alone, timed with CUDA events on the launch stream (b200lev_profile):
            achieved = cells/s x 5 INT32 ops against the INT32 issue rate
            measured live by the library's microbenchmark kernel
  cpu_baseline  the oracle port (C, OpenMP) on the box's host cores, same workload
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

T_LEN, V, NBEST = 100, 10000, 8
OPS_PER_CELL = 5  # compare, select/add, add, add, 3-input min


def make_batch(n_utts, seed):
    """cfg2 synthetic batch: ref (T+1, n_utts), hyp (T+1, n_utts*8) int64; len ~ U{50..100}
    with one eos(0) at len-1, tail = 0 (eos-padded); hyps independent of refs."""
    rng = np.random.default_rng(seed)
    T = T_LEN + 1

    def seqs(n):
        tok = rng.integers(1, V, size=(T, n), dtype=np.int64)
        lens = rng.integers(50, T_LEN + 1, size=n)
        pos = np.arange(T)[:, None]
        tok[pos >= (lens - 1)[None, :]] = 0
        return tok, lens

    ref, rl = seqs(n_utts)
    hyp, hl = seqs(n_utts * NBEST)
    # include_eos=True: valid lengths include the eos token
    cells = int((np.repeat(rl, NBEST).astype(np.int64) * hl.astype(np.int64)).sum())
    return ref, hyp, cells


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
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
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(ref, hyp, cells, min_seconds=3.0, max_runs=5):
    """The oracle port on the host cores (all OpenMP threads), same call, same tensors."""
    from oracle import oracle as O

    O.use_all_cores()
    refx = np.repeat(ref, NBEST, axis=1)
    O.prefix_error_rates(refx[:, :64], hyp[:, :64], eos=0)  # build + warm
    best, runs, t_total = None, 0, 0.0
    while runs < max_runs and (runs < 2 or t_total < min_seconds):
        t0 = time.perf_counter()
        O.prefix_error_rates(refx, hyp, eos=0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        t_total += dt
        runs += 1
    return {"value": cells / best / 1e9, "unit": "GCUPS", "cores": O.num_threads(),
            "kind": "port",
            "sample": f"{hyp.shape[1]} pairs of the same workload, best of {runs} "
                      f"({best * 1e3:.1f} ms); oracle/lev_oracle.c, OpenMP over pairs",
            "pairs_per_s": hyp.shape[1] / best}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (the reference itself is pure Python on torch CPU ops and cannot travel to the box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_utts = 2048  # bounded sample: 16384 pairs of the same per-pair shape
    ref, hyp, cells = make_batch(n_utts, seed=3)
    from oracle import oracle as O

    O.use_all_cores()
    refx = np.repeat(ref, NBEST, axis=1)
    O.prefix_error_rates(refx[:, :64], hyp[:, :64], eos=0)
    for _ in range(max(args.warmup, 1)):
        O.prefix_error_rates(refx, hyp, eos=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.prefix_error_rates(refx, hyp, eos=0)
    dt = (time.perf_counter() - t0) / args.steps
    val = cells / dt / 1e9
    line = {
        "impl": "reference", "metric": "edit-distance cell-updates/s", "value": val,
        "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 prefix_error_rates N-best: 8-best word hyps, T=100, "
                               "vocab 10k; bounded sample of 2048 utterances x 8 = 16384 pairs",
                   "pairs": int(hyp.shape[1]), "T": T_LEN, "nbest": NBEST, "vocab": V},
        "hyps_per_s": hyp.shape[1] / dt,
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": O.num_threads(), "kind": "port",
                         "sample": "16384 pairs per step; oracle/lev_oracle.c, OpenMP over pairs"},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=16384, help="utterances per GPU (x8-best)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import b200lev.functional as F
    from b200lev import _abi, _ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU: keep this rank's host buffers on its GPU's NUMA node
        from b200lev.dist import bind_host_to_gpu
        numa_node = bind_host_to_gpu(local)
    L = _abi.lib()

    ref_np, hyp_np, cells = make_batch(args.utts, seed=100 + rank)
    refx_np = np.repeat(ref_np, NBEST, axis=1)
    ref = torch.from_numpy(refx_np).to(dev)
    hyp = torch.from_numpy(hyp_np).to(dev)
    P = hyp.shape[1]
    in_bytes = ref.numel() * 8 + hyp.numel() * 8

    def step():
        return F.prefix_error_rates(ref, hyp, eos=0, warn=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, whole public call ------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    # K steps of 0.3 ms are shorter than one nvidia-smi period: keep the same step running
    # (untimed) under the sampler until it has seen the clocks this load settles at
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.6:
        for _ in range(20):
            out = step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the timed K steps plus 0.6 s of the same step, untimed"

    # ---- roofline: per-kernel CUDA-event times of the same public call --------------------
    # (b200lev_profile brackets every phase with events on the launch stream)
    prof = np.zeros(8, dtype=np.float64)
    buf = (ctypes.c_float * 8)()
    _abi.check(L.b200lev_profile(1))
    for _ in range(3):
        step()
    nprof = max(5, min(args.steps, 20))
    for _ in range(nprof):
        out = step()
        _abi.check(L.b200lev_profile_read(buf, 8))
        prof += np.array([max(x, 0.0) for x in buf])
    _abi.check(L.b200lev_profile(0))
    prof /= nprof
    ms_pack = max(float(prof[0] + prof[1]), 1e-6)
    bitvec = prof[7] > prof[3]  # which DP ran: the bit-vector kernels or the wavefront kernels
    ms_dp = float(prof[7] if bitvec else prof[3])
    phases = {"pack_ref": prof[0], "pack_hyp": prof[1], "bucketing": prof[2], "dp": prof[3],
              "prefix_finalize": prof[4], "standby_wide": prof[5], "bitvec_uid": prof[6],
              "bitvec_dp": prof[7]}

    # the same call with the bit-vector kernels switched off: north_star's target is quoted on
    # the wavefront kernel, so it is measured live next to the path that actually ships
    wave = None
    if bitvec:
        saved = os.environ.get("B200LEV_BITVEC")
        os.environ["B200LEV_BITVEC"] = "0"
        try:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            for _ in range(nprof):
                step()
            w1.record()
            torch.cuda.synchronize()
            wave = {"ms_per_step": w0.elapsed_time(w1) / nprof}
            wprof = np.zeros(8, dtype=np.float64)
            _abi.check(L.b200lev_profile(1))
            for _ in range(nprof):
                step()
                _abi.check(L.b200lev_profile_read(buf, 8))
                wprof += np.array([max(x, 0.0) for x in buf])
            _abi.check(L.b200lev_profile(0))
            wprof /= nprof
            wave["kernel_ms"] = float(wprof[3])
            wave["pack_ms"] = float(wprof[0] + wprof[1])
        finally:
            if saved is None:
                del os.environ["B200LEV_BITVEC"]
            else:
                os.environ["B200LEV_BITVEC"] = saved

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    st = torch.cuda.current_stream(dev).cuda_stream
    outp = out

    # INT32 issue-rate peak, measured live (variant 2 = VIADDMNMX, 0 = IADD3, 3 = DP cell mix)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    peaks = {}
    for name, var in (("iadd3", 0), ("viaddmnmx", 2), ("dp_cell_mix", 3)):
        ops = ctypes.c_double(0)

        def k(var=var):
            _abi.check(L.b200lev_int32_peak_kernel(var, 148 * 8, 4096, sink.data_ptr(),
                                                   ctypes.byref(ops), st))

        t = timed(k, 5)
        peaks[name] = ops.value / (t * 1e-3) / 1e12  # T int32-op/s
    # denominator = the ALU-pipe rate (VIADDMNMX / VIMNMX / ISETP live there): SURVEY 8(d)'s
    # "64 lanes/clk/SM".  Plain IADD3 also dual-issues on the FMA pipe (the iadd3 figure).
    int32_peak = peaks["viaddmnmx"]

    # ---- e2e: host (pinned) tensors through the public API -----------------------------
    ref_h = torch.from_numpy(refx_np).pin_memory()
    hyp_h = torch.from_numpy(hyp_np).pin_memory()
    # warm-up in the timed loop's own pattern (the previous result is still referenced while
    # the next call runs), so that the page-locked result blocks of the steady state exist
    # before the clock starts: a first-time cudaHostAlloc of 53 MB costs tens of ms
    res = None
    for _ in range(max(args.warmup, 3)):
        res = F.prefix_error_rates(ref_h, hyp_h, eos=0, warn=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        res = F.prefix_error_rates(ref_h, hyp_h, eos=0, warn=False)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / e2e_steps
    assert res.device.type == "cpu"

    # ---- literal BASELINE shape (64 x 8): latency of the public call --------------------
    lref_np, lhyp_np, lcells = make_batch(64, seed=7)
    lref = torch.from_numpy(np.repeat(lref_np, NBEST, axis=1)).to(dev)
    lhyp = torch.from_numpy(lhyp_np).to(dev)
    ms_lit = timed(lambda: F.prefix_error_rates(lref, lhyp, eos=0, warn=False), 50)

    # ---- reduce over ranks ---------------------------------------------------------------
    stats = torch.tensor([ms, ms_e2e, ms_dp, ms_pack], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(cells), float(P)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e, ms_dp_max, ms_pack_max = stats.tolist()
    cells_all, pairs_all = tot.tolist()

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        achieved_ops = cells / (ms_dp * 1e-3) * OPS_PER_CELL / 1e12
        traffic_all = {}
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):  # dram bytes per launch from the last `ncu --set full` captures
            traffic_all = json.load(open(tf))
        pack_gbs = (in_bytes + in_bytes // 2) / (ms_pack * 1e-3) / 1e9
        peak_note = ("measured live (b200lev_int32_peak_kernel), ALU pipe = viaddmnmx; T int32-op/s: "
                     f"{ {k: round(v, 2) for k, v in peaks.items()} }")
        if bitvec:
            # n-best shaped batch: the bit-vector kernels took the call on the device.  The
            # dominant kernel is the uid pre-pass (it reads the raw tokens once: HBM-bound);
            # the DP kernel is reported next to it against the INT32 issue rate.
            ms_uid = float(prof[6])
            R1 = T_LEN + 1
            uid_bytes = in_bytes + 2 * P * ((R1 + 15) // 16) * 16 + 2 * 4 * P + P
            roofline = {"bound": "hbm", "kernel": "lev_bv_uid_kernel<int64>",
                        "achieved": uid_bytes / (ms_uid * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "frac": uid_bytes / (ms_uid * 1e-3) / 1e9 / hbm_peak,
                        "traffic": traffic_all.get("lev_bv_uid_kernel"), "kernel_ms": ms_uid,
                        "algorithmic_bytes": uid_bytes, "peak_source": hbm_src,
                        "note": "reads both raw int64 token tensors once, writes 1 uid byte per "
                                "token + lengths; share of the step = kernel_ms / ms_per_step"}
            roofline_dp = {"bound": "int32_issue", "kernel": "lev_bv_dp_kernel<W=4,PREFIX>",
                           "achieved": achieved_ops, "peak": int32_peak, "unit": "Tint32op/s",
                           "frac": achieved_ops / int32_peak,
                           "traffic": traffic_all.get("lev_bv_dp_kernel"),
                           "ops_per_cell": OPS_PER_CELL, "kernel_ms": ms_dp,
                           "kernel_gcups": cells / (ms_dp * 1e-3) / 1e9, "peak_source": peak_note,
                           "note": "algorithmic 5 INT32 ops/cell (SURVEY 8d); Myers' bit-vector "
                                   "recurrence advances 32 cells with ~17 instructions, so frac "
                                   "exceeds 1; ncu: ALU pipe 78 % busy"}
            launches = 10  # 2 bit-vector kernels + 8 wavefront kernels standing by (exit at once)
        else:
            roofline = {"bound": "int32_issue",
                        "kernel": "lev_group_kernel<cost,PREFIX,packed16> (+ its 32-bit twin's "
                                  "immediate exit)",
                        "achieved": achieved_ops, "peak": int32_peak, "unit": "Tint32op/s",
                        "frac": achieved_ops / int32_peak,
                        "traffic": traffic_all.get("lev_group_kernel"),
                        "ops_per_cell": OPS_PER_CELL, "kernel_ms": ms_dp,
                        "kernel_gcups": cells / (ms_dp * 1e-3) / 1e9, "peak_source": peak_note,
                        "note": "algorithmic 5 INT32 ops/cell (SURVEY 8d); the kernel issues 2.5 "
                                "instructions per cell (2 cells per 16x2 DPX instruction), so "
                                "frac can exceed 1"}
            roofline_dp = None
            # 2 pack, 1 bucketing, 2 DP builds (one exits at once), 1 prefix finalize, 1 stand-by
            # 64-bit-token kernel (+ 2 bit-vector kernels that vetoed, when they were eligible)
            launches = 7 + (3 if prof[6] > 0 else 0)
        line = {
            "metric": "edit-distance cell-updates/s", "value": cells_all / (ms * 1e-3) / 1e9,
            "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "cfg2 prefix_error_rates N-best (8-best word hyps, T=100, vocab "
                                   "10k, include_eos, norm), batch replicated to "
                                   f"{args.utts} utts x 8 = {P} pairs per GPU",
                       "pairs_per_gpu": P, "T": T_LEN, "nbest": NBEST, "vocab": V,
                       "cells_per_gpu": cells, "host_numa_node": numa_node,
                       "l2": f"inputs {in_bytes / 1e6:.0f} MB per step exceed the 126 MB L2"},
            "hyps_per_s": pairs_all / (ms * 1e-3),
            "e2e": {"value": cells_all / (ms_e2e * 1e-3) / 1e9, "unit": "GCUPS",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(outp.numel() * 4),
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps,
            "roofline": roofline,
            "phases_ms": {k: round(float(v), 5) for k, v in phases.items()},
            "literal": {"workload": "64 utts x 8-best = 512 pairs (BASELINE configs[1] as written)",
                        "ms_per_call": ms_lit, "gcups": lcells / (ms_lit * 1e-3) / 1e9,
                        "hyps_per_s": 512 / (ms_lit * 1e-3)},
            "clocks": clocks,
        }
        if roofline_dp is not None:
            line["roofline_dp"] = roofline_dp
            if wave is not None and wave["kernel_ms"] > 0:
                wops = cells / (wave["kernel_ms"] * 1e-3) * OPS_PER_CELL / 1e12
                line["roofline_wavefront"] = {
                    "bound": "int32_issue",
                    "kernel": "lev_group_kernel<cost,PREFIX,packed16> (B200LEV_BITVEC=0: the "
                              "wavefront path, taken by every batch the bit-vector path declines)",
                    "achieved": wops, "peak": int32_peak, "unit": "Tint32op/s",
                    "frac": wops / int32_peak, "traffic": traffic_all.get("lev_group_kernel"),
                    "ops_per_cell": OPS_PER_CELL, "kernel_ms": wave["kernel_ms"],
                    "kernel_gcups": cells / (wave["kernel_ms"] * 1e-3) / 1e9,
                    "whole_call_ms": wave["ms_per_step"],
                    "whole_call_gcups": cells / (wave["ms_per_step"] * 1e-3) / 1e9,
                    "pack_ms": wave["pack_ms"],
                    "pack_gbs": (in_bytes + in_bytes // 2) / (wave["pack_ms"] * 1e-3) / 1e9,
                    "note": "algorithmic 5 INT32 ops/cell (SURVEY 8d); the kernel issues 2.5 "
                            "instructions per cell (2 cells per 16x2 DPX instruction)"}
        else:
            line["roofline_pack"] = {"bound": "hbm",
                                     "kernel": "lev_pack_seqfirst_kernel<int64> x2 (ref, hyp)",
                                     "achieved": pack_gbs, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": pack_gbs / hbm_peak, "kernel_ms": ms_pack,
                                     "peak_source": hbm_src}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(ref_np, hyp_np, cells)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
