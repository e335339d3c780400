#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): edit-distance cell-updates/s (GCUPS), cells = sum over pairs of
ref_len * hyp_len on the valid (eos-terminated) lengths; hyps/s is reported beside it.

Workloads = BASELINE.json's configs, each at its per-pair shape with the batch
replicated until the chip is full ("saturated"); the literal batch is timed too ("literal"):

  --config 1  error_rate, char pairs T~50, V=30                    65 536 pairs / GPU   weak
  --config 2  prefix_error_rates, 8-best word hyps, T=100, V=10k  131 072 pairs / GPU   weak   (default;
              the configuration the metric is quoted on)
  --config 3  optimal_completion, T=200, V=32, include_eos          8 192 pairs / GPU   weak
  --config 4  bulk WER: error_rate + device error sums + ONE all-reduce of the 24-byte fp64
              totals inside the timed step, 1 000 000 pairs sharded over the ranks      strong
  --config 5  prefix_edit_distances, T=2000 ragged, costs 3/3/4     1 184 pairs / GPU   weak

A step = one public call on one batch.

  value         device-resident inputs, CUDA events, max over ranks (whole job, all GPUs)
  e2e           the same public call with HOST (pinned) int64 tensors: H2D of the inputs and
                D2H of the result inside the timed region
  roofline      the dominant kernel alone, timed with CUDA events on the launch stream
                (b200lev_profile): INT32-issue bound kernels as cells/s x 5 INT32 ops
                against the issue rate measured live by the library's microbenchmark kernel,
                HBM-bound kernels as algorithmic bytes/s against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (C, OpenMP) on the box's host cores, bounded sample

--impl reference times the UNMODIFIED reference (sdrobert/pydrobert-pytorch installed into
baseline/_ref by oracle/make_ref.sh): its own torch implementation of the same public call on
the host cores (torch CPU ops, all threads), same per-pair shape, bounded pair count; and, when
a GPU is present, the same reference code on CUDA tensors next to it ("reference_on_gpu").
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

OPS_PER_CELL = 5  # compare, select/add, add, add, 3-input min
NBEST = 8
PROF_SLOTS = ("pack_ref", "pack_hyp", "bucketing", "dp", "prefix_finalize", "standby_wide",
              "bitvec_probe_or_uid", "bitvec_dp", "completion_uid", "completion_fill", "err_sum", "mask_packed")


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def _seqs(rng, T, n, V, lo, hi, eos, pad):
    """(T, n) int64 tokens in [1, V), lengths ~ U{lo..hi} INCLUDING one eos at len-1, tail pad."""
    tok = rng.integers(1, V, size=(T, n), dtype=np.int64)
    lens = rng.integers(lo, hi + 1, size=n)
    pos = np.arange(T)[:, None]
    tok[pos == (lens - 1)[None, :]] = eos
    tok[pos > (lens - 1)[None, :]] = pad
    return tok, lens


class Workload:
    """One BASELINE config: data, the public call (ours and the reference's), the oracle call."""

    def __init__(self, cfg):
        self.cfg = cfg
        (self.name, self.T, self.V, self.lo, self.hi, self.eos, self.pad, self.sat_pairs,
         self.lit_pairs, self.scaling, self.ref_pairs, self.port_pairs) = {
            1: ("cfg1 error_rate: char-level pairs, T~50, vocab 30, eos-padded, unit costs",
                51, 30, 25, 50, 0, 0, 65536, 32, "weak", 2048, 65536),
            2: ("cfg2 prefix_error_rates N-best: 8-best word hyps, T=100, vocab 10k, include_eos, norm",
                101, 10000, 50, 100, 0, 0, 131072, 512, "weak", 512, 131072),
            3: ("cfg3 optimal_completion: char seqs T=200, V=32, include_eos",
                201, 32, 100, 200, 0, 0, 8192, 128, "weak", 16, 8192),
            4: ("cfg4 bulk WER scoring: error_rate(norm=False) + error sums, word pairs T~30, vocab 10k, "
                "sharded over the ranks, one all-reduce of the totals",
                31, 10000, 10, 30, -1, -2, 1000000, 1000000, "strong", 10000, 1000000),
            5: ("cfg5 long-form prefix_edit_distances: T=2000 chars, ragged, costs ins=3 del=3 sub=4",
                2001, 64, 200, 2000, 0, 0, 1184, 256, "weak", 2, 256),
        }[cfg]
        self.nbest = NBEST if cfg == 2 else 1

    # include_eos of the call decides which lengths count as cells
    @property
    def include_eos(self):
        return self.cfg in (2, 3, 5)

    def make(self, pairs, seed):
        rng = np.random.default_rng(seed)
        n_ref = pairs // self.nbest
        ref, rl = _seqs(rng, self.T, n_ref, self.V, self.lo, self.hi, self.eos, self.pad)
        hyp, hl = _seqs(rng, self.T, pairs, self.V, self.lo, self.hi, self.eos, self.pad)
        if self.nbest > 1:  # prefix_error_rates(ref (x) 8, hyp): the reference physically repeated
            ref, rl = np.repeat(ref, self.nbest, axis=1), np.repeat(rl, self.nbest)
        d = 0 if self.include_eos else 1
        cells = int(((rl - d).astype(np.int64) * (hl - d).astype(np.int64)).sum())
        return ref, hyp, cells

    def kwargs(self):
        return {1: dict(eos=0), 2: dict(eos=0), 3: dict(eos=0),
                4: dict(eos=-1, include_eos=False, norm=False),
                5: dict(eos=0, ins_cost=3.0, del_cost=3.0, sub_cost=4.0)}[self.cfg]

    def fn_name(self):
        return {1: "error_rate", 2: "prefix_error_rates", 3: "optimal_completion", 4: "error_rate",
                5: "prefix_edit_distances"}[self.cfg]

    def call(self, F, ref, hyp):
        return getattr(F, self.fn_name())(ref, hyp, warn=False, **self.kwargs())

    def oracle_call(self, O, ref, hyp):
        return getattr(O, self.fn_name())(ref, hyp, **self.kwargs())

    def config(self, pairs_per_gpu, extra=None):
        c = {"workload": self.name, "cfg": self.cfg, "call": self.fn_name(), "T": self.T - 1,
             "vocab": self.V, "nbest": self.nbest, "pairs_per_gpu": pairs_per_gpu}
        c.update(extra or {})
        return c


def make_batch(n_utts, seed):
    """cfg2 synthetic batch with the reference NOT repeated: ref (T+1, n_utts), hyp (T+1, n_utts*8)
    int64, and the cells of the 8-best pairs (the tests build their own n-best / unrelated
    pairings from it)."""
    wl = Workload(2)
    rng = np.random.default_rng(seed)
    ref, rl = _seqs(rng, wl.T, n_utts, wl.V, wl.lo, wl.hi, wl.eos, wl.pad)
    hyp, hl = _seqs(rng, wl.T, n_utts * NBEST, wl.V, wl.lo, wl.hi, wl.eos, wl.pad)
    cells = int((np.repeat(rl, NBEST).astype(np.int64) * hl.astype(np.int64)).sum())
    return ref, hyp, cells


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).

    Read through NVML from a thread of this process (nvidia_ml_py): one `nvidia-smi -lms` child
    takes ~150 ms to come up and every one of its polls stalls kernel launches for milliseconds
    -- more than a whole timed region of K short steps -- so it can only ever bracket the region,
    never sample inside it.  An NVML clock query from inside the process costs tens of
    microseconds.  Falls back to the nvidia-smi child when the module is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.01):
        self.idx = gpu_index
        self.period = period_s
        self.proc = None
        self.lines = []
        self.samples = []  # (sm_mhz, max_mhz, reasons bitmask)
        self.nvml = None
        self._stop = threading.Event()
        self._paused = False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the torch device
            import torch

            props = torch.cuda.get_device_properties(self.idx)
            bus = "{:08x}:{:02x}:{:02x}.0".format(props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            if self._paused:
                self._stop.wait(0.001)
                continue
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    why = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except AttributeError:
                    why = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(mx), int(why)))
            except Exception:
                pass
            self._stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Number of samples so far (to tell which ones fell inside the timed region)."""
        return len(self.samples)

    def pause(self, on):
        """No polls while `on`: ONE poll inside a region of K short steps stalls kernel launches
        for milliseconds (measured: 0.35 ms/step with a poll inside 20 steps of 0.17 ms)."""
        self._paused = on

    def stop(self, first=0, last=None):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            n = self.nvml
            names = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
            inside = self.samples[first:last] or self.samples
            reasons = sorted({nm for _, _, why in self.samples for nm, bit in names if why & bit})
            return {"sm_mhz": float(np.median([x[0] for x in inside])) if inside else None,
                    "sm_max_mhz": float(max(x[1] for x in inside)) if inside else None,
                    "samples": len(self.samples), "samples_in_timed_region": len(self.samples[first:last]),
                    "reasons": reasons, "how": f"NVML polled every {self.period * 1e3:.0f} ms from this process"}
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
                "samples": len(sm), "reasons": sorted(reasons), "how": "nvidia-smi -lms 100 child"}


# ---------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------
def cpu_baseline(wl, seed, min_seconds=3.0, max_runs=5):
    """The oracle port on the host cores (all OpenMP threads), same call, bounded sample."""
    from oracle import oracle as O

    O.use_all_cores()
    ref, hyp, cells = wl.make(wl.port_pairs, seed)
    wl.oracle_call(O, ref[:, :64], hyp[:, :64])  # build + warm
    best, runs, t_total = None, 0, 0.0
    while runs < max_runs and (runs < 2 or t_total < min_seconds):
        t0 = time.perf_counter()
        wl.oracle_call(O, ref, hyp)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        t_total += dt
        runs += 1
    return {"value": cells / best / 1e9, "unit": "GCUPS", "cores": O.num_threads(), "kind": "port",
            "sample": f"{hyp.shape[1]} pairs of the same workload, best of {runs} "
                      f"({best * 1e3:.1f} ms); oracle/lev_oracle.c, OpenMP over pairs",
            "pairs_per_s": hyp.shape[1] / best}


def _import_reference():
    """The unmodified reference from baseline/_ref (installed by oracle/make_ref.sh)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "pydrobert", "torch")):
        return None, "baseline/_ref is absent (oracle/make_ref.sh needs /root/reference)"
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import pydrobert.torch.functional as RF  # noqa: N812
    except Exception as e:  # pragma: no cover
        return None, f"import of the reference failed: {e!r}"
    return RF, None


def run_reference(args):
    """--impl reference: the reference's own implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    wl = Workload(args.config)
    RF, why = _import_reference()
    if RF is None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    pairs = wl.ref_pairs
    ref, hyp, cells = wl.make(pairs, seed=3)
    tr, th = torch.from_numpy(ref), torch.from_numpy(hyp)
    kw = dict(wl.kwargs())
    steps_of = None
    if wl.cfg == 5:
        # a full T=2000 call of the reference takes ~1 h: time the first 16 hypothesis
        # steps (the cost per step does not depend on the step, SM:286-318) and scale
        steps_of = 16
        th = th[:steps_of].contiguous()
    fn = getattr(RF, wl.fn_name())

    def step():
        if wl.cfg == 4:  # the reference command scores in batches of 100 (command_line.py:1124)
            tot = 0.0
            for a in range(0, pairs, 100):
                tot += float(fn(tr[:, a:a + 100], th[:, a:a + 100], warn=False, **kw).sum())
            return tot
        return fn(tr, th, warn=False, **kw)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    scale = 1.0
    sample = f"{pairs} pairs per step, reference torch-CPU ops, {cores} threads"
    if steps_of is not None:
        scale = (wl.T) / steps_of
        sample += f"; first {steps_of} of {wl.T} hypothesis steps timed, scaled x{scale:.1f} (extrapolated)"
    dt_full = dt * scale
    val = cells / dt_full / 1e9
    line = {
        "impl": "reference", "metric": "edit-distance cell-updates/s", "value": val,
        "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt_full * 1e3, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config(pairs, {"sample": "bounded pair count, same per-pair shape"}),
        "hyps_per_s": pairs / dt_full,
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the same reference code on the B200 (CUDA tensors), when there is one
    if torch.cuda.is_available() and not args.no_reference_on_gpu:
        dev = torch.device("cuda", 0)
        cr, ch = tr.to(dev), th.to(dev)

        def gstep():
            if wl.cfg == 4:
                tot = torch.zeros((), device=dev)
                for a in range(0, pairs, 100):
                    tot += fn(cr[:, a:a + 100], ch[:, a:a + 100], warn=False, **kw).sum()
                return tot
            return fn(cr, ch, warn=False, **kw)

        for _ in range(2):
            gstep()
        torch.cuda.synchronize()
        reps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            gstep()
        torch.cuda.synchronize()
        gdt = (time.perf_counter() - t0) / reps * scale
        line["reference_on_gpu"] = {"value": cells / gdt / 1e9, "unit": "GCUPS", "ms_per_step": gdt * 1e3,
                                    "device": torch.cuda.get_device_name(0),
                                    "note": "the reference's torch code on CUDA tensors, same sample"}
    # the C/OpenMP port beside it, for continuity with round 1's reference arm
    if not args.no_cpu_baseline:
        line["cpu_baseline_port"] = cpu_baseline(wl, seed=3, min_seconds=2.0, max_runs=3)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measure_sharded_scoring(torch, dist, D, dev, rank, world, steps):
    """BASELINE config 4 as a secondary line of a multi-GPU run of another config: the million
    pairs cut over the ranks, every step = shard -> error_rate + device sums -> ONE all-reduce of
    the 24-byte fp64 totals.  Same protocol as the main measurement (warm-up, barrier, K steps
    between events, max over ranks); the totals are checked in the run."""
    wl = Workload(4)
    lo, hi = D.shard_bounds(wl.sat_pairs, rank, world)
    ref_np, hyp_np, cells = wl.make(hi - lo, seed=900 + rank)
    ref, hyp = torch.from_numpy(ref_np).to(dev), torch.from_numpy(hyp_np).to(dev)
    acc = None
    for _ in range(5):
        er, acc = D.bulk_error_rate(ref, hyp, eos=-1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        er, acc = D.bulk_error_rate(ref, hyp, eos=-1)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    c = torch.tensor([float(cells)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    tot = acc.cpu().tolist()
    ok = int(tot[2]) == wl.sat_pairs
    return {"workload": wl.name, "scaling": "strong", "pairs_total": wl.sat_pairs, "ms_per_step": ms,
            "pairs_per_s": wl.sat_pairs / (ms * 1e-3), "gcups": float(c.item()) / (ms * 1e-3) / 1e9,
            "collective": "one all_reduce(sum) of 3 fp64 values per step on the NCCL stream",
            "all_reduced_pair_count_is_global": ok, "wer": tot[0] / max(tot[1], 1.0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (cfg4: in total); 0 = the config's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-on-gpu", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnosis only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import b200lev.functional as F
    from b200lev import _abi
    from b200lev import dist as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        import datetime

        # (a short collective timeout: a rank that falls out of step must fail, not hang the box)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
        # one process per GPU: keep this rank's host buffers on its GPU's NUMA node
        numa_node = D.bind_host_to_gpu(local)
    L = _abi.lib()
    wl = Workload(args.config)
    total_pairs = args.pairs or wl.sat_pairs
    if wl.scaling == "strong":  # cfg4: the million pairs are cut into contiguous shards
        lo, hi = D.shard_bounds(total_pairs, rank, world)
        my_pairs = hi - lo
    else:
        my_pairs = total_pairs
    ref_np, hyp_np, cells = wl.make(my_pairs, seed=100 + rank)
    ref = torch.from_numpy(ref_np).to(dev)
    hyp = torch.from_numpy(hyp_np).to(dev)
    P = hyp.shape[1]
    in_bytes = ref.numel() * 8 + hyp.numel() * 8
    # L2 (126 MB): when one batch does not exceed it, the timed steps rotate through several
    # independent batches of the same shape whose inputs together do, so no step finds its
    # inputs in L2 (the roofline / e2e legs use the first batch)
    batches = [(ref, hyp, cells)]
    if in_bytes < (160 << 20):
        n_sets = min(8, -(-(200 << 20) // max(in_bytes, 1)))
        for k in range(1, n_sets):
            r_k, h_k, c_k = wl.make(my_pairs, seed=100 + rank + 1000 * k)
            batches.append((torch.from_numpy(r_k).to(dev), torch.from_numpy(h_k).to(dev), c_k))
    totals = {}
    turn = [0]

    def step(r=None, h=None):
        if r is None:
            r, h, _ = batches[turn[0] % len(batches)]
            turn[0] += 1
        if wl.cfg == 4:
            # shard -> error_rate + device sums (K7) -> ONE all-reduce of [sum err, sum ref
            # tokens, #pairs] (fp64, 24 bytes) on the NCCL stream, inside the step
            er, acc = D.bulk_error_rate(r, h, eos=-1)
            totals["acc"] = acc
            return er
        return wl.call(F, r, h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    debug = os.environ.get("B200LEV_BENCH_DEBUG") == "1"

    def timed_loop(fn, reps):
        """K back-to-back calls between two events on the launch stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if debug:
            st0 = torch.cuda.memory_stats(dev)
            walls = []
        e0.record()
        for _ in range(reps):
            if debug:
                tw = time.perf_counter()
            out = fn()
            if debug:
                walls.append(round((time.perf_counter() - tw) * 1e6))
        e1.record()
        torch.cuda.synchronize()
        if debug:
            st1 = torch.cuda.memory_stats(dev)
            keys = ("num_device_alloc", "num_device_free", "num_alloc_retries", "reserved_bytes.all.current")
            sys.stderr.write(f"timed_loop: {e0.elapsed_time(e1) / reps:.4f} ms/step; enqueue us {walls}; "
                             f"allocator {[(k, st0.get(k), st1.get(k)) for k in keys]}\n")
        return e0.elapsed_time(e1) / reps, out

    # ---- value: device-resident, whole public call ------------------------------------
    # W untimed warm-up steps, then a short soak of the same step with the clock sampler already
    # running (K steps of 0.2 ms are shorter than one nvidia-smi period, and the first calls
    # after an idle GPU run below the clocks the load settles at), then EXACTLY K timed steps.
    # (warm-up, soak and timed steps all run through timed_loop: its local `out` is the only
    # reference to a result, so the caching allocator ping-pongs between the same two blocks in
    # all three phases -- a result kept alive outside would cost the timed region a cudaMalloc
    # of a third block, which takes 1 to 70 ms on these boxes)
    timed_loop(step, max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    soak_s = 0.5

    def soak(seconds):
        """The same step, untimed, for about `seconds` -- by a step COUNT every rank agrees on (a
        wall-clock loop would run a different number of steps, hence of all-reduces, per rank)."""
        t10, _ = timed_loop(step, 10)
        t = torch.tensor([t10], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = int(min(max(seconds * 1e3 / max(float(t.item()), 1e-3), 10), 20000))
        done = 0
        while done < n:
            timed_loop(step, min(100, n - done))
            done += min(100, n - done)

    soak(soak_s)
    barrier()
    # the sampler polls up to the first timed step and again from the last one on, while the same
    # step keeps running (untimed): the clocks are those of this load, read within milliseconds of
    # the timed region on both sides; inside it a poll would be the largest thing measured
    sampler.pause(True)
    m0 = sampler.mark()
    first_turn = turn[0]
    ms, out = timed_loop(step, args.steps)
    cells_step = sum(batches[(first_turn + k) % len(batches)][2] for k in range(args.steps)) / args.steps
    sampler.pause(False)
    soak(0.25)
    barrier()
    clocks = sampler.stop(0, None) if rank == 0 else None
    if clocks is not None:
        clocks["samples_before_timed_region"] = m0
        clocks.pop("samples_in_timed_region", None)
        clocks["window"] = (f"{soak_s} s of the same step before and 0.25 s after the timed K steps, polling "
                            "paused during them (one NVML poll stalls launches for milliseconds)")
    out_bytes = int(out.numel() * out.element_size())

    # ---- per-kernel CUDA-event times of the same public call (b200lev_profile) -------------
    nslots = len(PROF_SLOTS)
    buf = (ctypes.c_float * nslots)()

    def profile(nprof):
        acc = np.zeros(nslots, dtype=np.float64)
        _abi.check(L.b200lev_profile(1))
        for _ in range(3):
            step(ref, hyp)
        for _ in range(nprof):
            step(ref, hyp)
            _abi.check(L.b200lev_profile_read(buf, nslots))
            acc += np.array([max(x, 0.0) for x in buf])
        _abi.check(L.b200lev_profile(0))
        return acc / nprof

    nprof = max(5, min(args.steps, 20))
    prof = profile(nprof)
    phases = {k: round(float(v), 5) for k, v in zip(PROF_SLOTS, prof)}
    bitvec = prof[7] > prof[3]  # which DP ran: the bit-vector kernel(s) or the wavefront kernels

    # cfg2: the same call with the bit-vector kernels switched off -- north_star's target is
    # quoted on the wavefront kernel, so it is measured live next to the path that ships
    wave = None
    if bitvec and wl.cfg == 2:
        saved = os.environ.get("B200LEV_BITVEC")
        os.environ["B200LEV_BITVEC"] = "0"
        try:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            wms, _ = timed_loop(step, nprof)
            wprof = profile(nprof)
            wave = {"ms_per_step": wms, "kernel_ms": float(wprof[3]), "pack_ms": float(wprof[0] + wprof[1])}
        finally:
            if saved is None:
                del os.environ["B200LEV_BITVEC"]
            else:
                os.environ["B200LEV_BITVEC"] = saved

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        return timed_loop(fn, reps)[0]

    st = torch.cuda.current_stream(dev).cuda_stream
    # INT32 issue-rate peak, measured live (variant 2 = VIADDMNMX, 0 = IADD3, 3 = DP cell mix)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    peaks = {}
    for name, var in (("iadd3", 0), ("viaddmnmx", 2), ("dp_cell_mix", 3)):
        ops = ctypes.c_double(0)

        def k(var=var, ops=ops):
            _abi.check(L.b200lev_int32_peak_kernel(var, 148 * 8, 4096, sink.data_ptr(),
                                                   ctypes.byref(ops), st))

        for _ in range(3):
            k()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            k()
        b.record()
        torch.cuda.synchronize()
        peaks[name] = ops.value / (a.elapsed_time(b) / 5 * 1e-3) / 1e12  # T int32-op/s
    # denominator = the ALU-pipe rate (VIADDMNMX / VIMNMX / ISETP / LOP3 live there): SURVEY 8(d)'s
    # "64 lanes/clk/SM".  Plain IADD3 also dual-issues on the FMA pipe (the iadd3 figure).
    int32_peak = peaks["viaddmnmx"]

    # ---- e2e: host (pinned) tensors through the public API -----------------------------
    ref_h = torch.from_numpy(ref_np).pin_memory()
    hyp_h = torch.from_numpy(hyp_np).pin_memory()
    # warm-up in the timed loop's own pattern (the previous result is still referenced while
    # the next call runs), so that the page-locked result blocks of the steady state exist
    # before the clock starts: a first-time cudaHostAlloc of 53 MB costs tens of ms
    res = None
    for _ in range(max(args.warmup, 3)):
        res = step(ref_h, hyp_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        res = step(ref_h, hyp_h)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / e2e_steps
    assert res.device.type == "cpu"
    # the same call with the tokens held as int16 on the host (every config's vocabulary, eos and
    # padding values fit): the kernels read 2-, 4- and 8-byte tokens, and the reference's API
    # takes any integer dtype too, so a caller who stores tokens narrow moves a quarter of the bytes
    ms_e2e_16 = None
    if max(abs(int(ref_np.min())), abs(int(ref_np.max())), abs(int(hyp_np.min())), abs(int(hyp_np.max()))) < 32768:
        ref_h16 = torch.from_numpy(ref_np.astype(np.int16)).pin_memory()
        hyp_h16 = torch.from_numpy(hyp_np.astype(np.int16)).pin_memory()
        for _ in range(max(args.warmup, 3)):
            res16 = step(ref_h16, hyp_h16)
        barrier()
        e0.record()
        for _ in range(e2e_steps):
            res16 = step(ref_h16, hyp_h16)
        e1.record()
        barrier()
        ms_e2e_16 = e0.elapsed_time(e1) / e2e_steps
        assert torch.equal(res16, res), "int16 host tokens changed the result"

    # ---- literal BASELINE shape: latency of the public call --------------------------------
    lit = None
    if wl.lit_pairs != total_pairs and world == 1:
        lref_np, lhyp_np, lcells = wl.make(wl.lit_pairs, seed=7)
        lref, lhyp = torch.from_numpy(lref_np).to(dev), torch.from_numpy(lhyp_np).to(dev)
        ms_lit = timed(lambda: wl.call(F, lref, lhyp), 30)
        lit = {"workload": f"{wl.lit_pairs} pairs (the BASELINE config as written)", "ms_per_call": ms_lit,
               "gcups": lcells / (ms_lit * 1e-3) / 1e9, "hyps_per_s": wl.lit_pairs / (ms_lit * 1e-3)}

    # ---- cfg4: the reduced totals against the oracle on a slice ------------------------------
    check = None
    if wl.cfg == 4:
        from oracle import oracle as O

        n_chk = min(4096, P)
        want = np.asarray(O.error_rate(ref_np[:, :n_chk], hyp_np[:, :n_chk], eos=-1, norm=False))
        # one more step on the FIRST batch, on every rank (it carries the all-reduce): the timed
        # steps rotate through several batches when a shard is smaller than L2
        chk = step(ref, hyp)
        barrier()
        got = chk[:n_chk].cpu().numpy()
        acc = totals["acc"].cpu().numpy()
        check = {"slice_pairs": n_chk, "slice_matches_oracle": bool(np.array_equal(got, want)),
                 "sum_errors": float(acc[0]), "sum_ref_tokens": float(acc[1]), "pairs": float(acc[2]),
                 "wer": float(acc[0] / max(acc[1], 1.0))}
        assert check["slice_matches_oracle"], "cfg4: per-pair errors differ from the oracle"
        assert int(acc[2]) == total_pairs, "cfg4: the all-reduced pair count is not the global one"

    # ---- multi-GPU runs of the other configs also carry config 4 with its all-reduce ---------
    sharded = None
    if world > 1 and wl.cfg != 4:
        sharded = measure_sharded_scoring(torch, dist, D, dev, rank, world, max(args.steps, 20))

    # ---- reduce over ranks ---------------------------------------------------------------
    dom_ms = float(prof[7] if bitvec else prof[3])
    stats = torch.tensor([ms, ms_e2e, dom_ms, ms_e2e_16 or 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(cells_step), float(P), float(cells)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e, _, ms_e2e_16 = stats.tolist()
    cells_all, pairs_all, cells_e2e_all = tot.tolist()

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        traffic_all = {}
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):  # dram bytes per launch from the committed `ncu --set full` captures
            traffic_all = json.load(open(tf))
        peak_note = ("measured live (b200lev_int32_peak_kernel), ALU pipe = viaddmnmx; T int32-op/s: "
                     f"{ {k: round(v, 2) for k, v in peaks.items()} }")

        def int32_roofline(kernel, kernel_ms, note, traffic_key):
            ops = cells / (kernel_ms * 1e-3) * OPS_PER_CELL / 1e12
            return {"bound": "int32_issue", "kernel": kernel, "achieved": ops, "peak": int32_peak,
                    "unit": "Tint32op/s", "frac": ops / int32_peak,
                    "traffic": traffic_all.get(traffic_key), "traffic_source": "profiles/traffic.json (ncu --set full)",
                    "ops_per_cell": OPS_PER_CELL, "kernel_ms": kernel_ms,
                    "kernel_gcups": cells / (kernel_ms * 1e-3) / 1e9, "peak_source": peak_note, "note": note}

        def hbm_roofline(kernel, kernel_ms, nbytes, note, traffic_key):
            gbs = nbytes / (kernel_ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbs / hbm_peak, "traffic": traffic_all.get(traffic_key),
                    "traffic_source": "profiles/traffic.json (ncu --set full)", "kernel_ms": kernel_ms,
                    "algorithmic_bytes": int(nbytes), "peak_source": hbm_src, "note": note}

        extra = {}
        pack_ms = float(prof[0] + prof[1])
        short = bitvec and wl.T <= 64 and os.environ.get("B200LEV_BV_SHORT", "1") != "0"
        if short:
            # R <= 64: one kernel, lane = pair, reference in registers, straight from the raw tokens
            roofline = int32_roofline(
                f"lev_bv_short_kernel<int64,W={1 if wl.T <= 32 else 2},{'PREFIX' if wl.cfg == 2 else 'FINAL'}>",
                float(prof[7]),
                "algorithmic 5 INT32 ops/cell (SURVEY 8d); the match masks come from 16-bit packed compares "
                "against the reference registers (one DPX add-and-clamp + one FMA-pipe shift per two "
                "positions), the recurrence is Myers' bit-vector step; the same kernel reads both raw int64 "
                "token tensors (roofline_hbm)" + ("; config 4: the totals are accumulated in this kernel" if wl.cfg == 4 else ""),
                "lev_bv_short_kernel" if wl.T <= 32 else "lev_bv_short_kernel_cfg1")
            extra["roofline_hbm"] = hbm_roofline(
                "lev_bv_short_kernel (its memory side)", float(prof[7]), in_bytes + out_bytes + 8 * P,
                "reads both raw int64 token tensors once, writes the results + lengths",
                "lev_bv_short_kernel" if wl.T <= 32 else "lev_bv_short_kernel_cfg1")
            launches = 1  # (config 4 included: b200lev_final_sums accumulates inside the kernel)
        elif bitvec:
            fused = os.environ.get("B200LEV_BV_FUSED", "1") != "0"
            if fused:
                # one kernel: lengths, run detection, hash tables, Myers' recurrence, output rows
                roofline = int32_roofline(
                    "lev_bv_fused_kernel<int64,W=4,PREFIX>", float(prof[7]),
                    "algorithmic 5 INT32 ops/cell (SURVEY 8d); Myers' bit-vector recurrence advances 128 "
                    "cells with ~50 ALU instructions, so frac exceeds 1; the same kernel reads both raw "
                    "int64 token tensors and writes the (H+1, N) fp32 rows (roofline_hbm)",
                    "lev_bv_fused_kernel")
                extra["roofline_hbm"] = hbm_roofline(
                    "lev_bv_fused_kernel (its memory side)", float(prof[7]), in_bytes + out_bytes + 8 * P,
                    "reads both raw int64 token tensors once, writes the output rows + lengths",
                    "lev_bv_fused_kernel")
                launches = 10  # probe + fused kernel + 8 wavefront kernels standing by (exit at once)
            else:
                R1 = wl.T
                uid_bytes = in_bytes + 2 * P * ((R1 + 15) // 16) * 16 + 2 * 4 * P + P
                roofline = hbm_roofline("lev_bv_uid_kernel<int64>", float(prof[6]), uid_bytes,
                                        "two-kernel form (B200LEV_BV_FUSED=0): reads both raw int64 token "
                                        "tensors once, writes 1 uid byte per token + lengths",
                                        "lev_bv_uid_kernel")
                extra["roofline_dp"] = int32_roofline("lev_bv_dp_kernel<W=4,PREFIX>", float(prof[7]),
                                                      "Myers' recurrence on the uid bytes", "lev_bv_dp_kernel")
                launches = 10
            if wave is not None and wave["kernel_ms"] > 0:
                w = int32_roofline(
                    "lev_group_kernel<cost,PREFIX,packed16> (B200LEV_BITVEC=0: the wavefront path, taken "
                    "by every batch the bit-vector path declines)", wave["kernel_ms"],
                    "algorithmic 5 INT32 ops/cell (SURVEY 8d); the kernel issues 2 instructions per cell "
                    "(3 ALU-pipe DPX + 1 FMA-pipe instruction per two cells)", "lev_group_kernel")
                w.update({"whole_call_ms": wave["ms_per_step"],
                          "whole_call_gcups": cells / (wave["ms_per_step"] * 1e-3) / 1e9,
                          "pack_ms": wave["pack_ms"],
                          "pack_gbs": (in_bytes + in_bytes // 2) / (wave["pack_ms"] * 1e-3) / 1e9})
                extra["roofline_wavefront"] = w
        else:
            kname = {1: "lev_group_kernel<cost,FINAL,packed16>", 3: "lev_warp_kernel<int,MASK> (two passes)",
                     4: "lev_group_kernel<cost,FINAL,packed16>", 5: "lev_cta_kernel<int,PREFIX> (TMA-staged)",
                     2: "lev_group_kernel<cost,PREFIX,packed16>"}[wl.cfg]
            dp_ms = float(prof[3])
            if wl.cfg == 3 and len(prof) > 11 and prof[11] > dp_ms:
                # mask mode on the packed kernel (lev_warp_kernel only takes what it leaves)
                kname, dp_ms = "lev_mask16_kernel<8> (two passes, two pairs per warp, 16x2 DPX)", float(prof[11])
            roofline = int32_roofline(kname, dp_ms,
                                      "algorithmic 5 INT32 ops/cell (SURVEY 8d) on the wavefront DP kernel",
                                      kname.split("<")[0])
            if pack_ms > 0:
                extra["roofline_pack"] = hbm_roofline("lev_pack_*_kernel x2 (ref, hyp)", pack_ms,
                                                      in_bytes + in_bytes // 2,
                                                      "reads the raw int64 tokens, writes the packed int32 "
                                                      "(+ 16-bit hypothesis) tables", "lev_pack_kernel")
            if wl.cfg == 3 and prof[9] > 0:
                # north_star item 4: achieved HBM GB/s of the kernel that writes the targets
                extra["roofline_target_writer"] = hbm_roofline(
                    "lev_completion_fill_kernel", float(prof[9]), out_bytes + 4 * (out_bytes // 8 // max(out.shape[-1], 1)),
                    "writes the (H', N, U) int64 targets (128-bit staged stores), reads one bitmap word "
                    "per row", "lev_completion_fill_kernel")
            launches = int(sum(1 for v in prof if v > 0)) + 3
        line = {
            "metric": "edit-distance cell-updates/s", "value": cells_all / (ms * 1e-3) / 1e9,
            "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "soak_s": soak_s,
            "ms_per_step": ms, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": wl.config(P, {
                "pairs_total": int(pairs_all), "cells_per_gpu": cells, "host_numa_node": numa_node,
                "l2": (f"inputs {in_bytes / 1e6:.0f} MB per step exceed the 126 MB L2" if len(batches) == 1 else
                       f"inputs {in_bytes / 1e6:.0f} MB per step: the timed steps rotate through "
                       f"{len(batches)} independent batches ({len(batches) * in_bytes / 1e6:.0f} MB > 126 MB L2)")}),
            "hyps_per_s": pairs_all / (ms * 1e-3),
            "e2e": {"value": cells_e2e_all / (ms_e2e * 1e-3) / 1e9, "unit": "GCUPS",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps,
            "roofline": roofline,
            "phases_ms": phases,
            "clocks": clocks,
        }
        line.update(extra)
        if ms_e2e_16:
            line["e2e_int16_tokens"] = {
                "value": cells_e2e_all / (ms_e2e_16 * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms_e2e_16,
                "h2d_bytes_per_step": in_bytes // 4, "d2h_bytes_per_step": out_bytes,
                "note": "same public call, same pairs, host tensors of dtype int16 instead of int64 "
                        "(results asserted equal)"}
        if sharded is not None:
            line["sharded_scoring_cfg4"] = sharded
        if lit is not None:
            line["literal"] = lit
        if check is not None:
            line["check"] = check
            line["collective"] = ("one all_reduce(sum) of 3 fp64 values per step on the NCCL stream" if world > 1
                                  else "none (1 rank)")
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, seed=100)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
