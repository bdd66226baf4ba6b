#!/usr/bin/env python
"""bench.py -- GCUPS of the ssw_align hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W  the reference's own SSE2 path on the host cores

Workload (BASELINE.json configs[1], SURVEY 8(d) "config 2"): 1,000 synthetic 150 bp DNA reads x one
5 Mbp reference, DNA matrix 2/-2 (N = 0), gap open 3 / extend 1, score_size 2, flag 0, maskLen 75.
With N GPUs every rank aligns its own 1,000 reads against the same reference (weak scaling; the pairs
are independent, nothing is exchanged during the fill; results are gathered to rank 0 over NCCL).
GCUPS = sum(readLen * refLen) / time / 1e9, cells counted once per pair.

One JSON line is printed by rank 0.  `value` is timed with the sequences already resident in HBM;
`e2e` goes through the host-buffer C ABI call (ssw_align_batch: H2D of all sequences + D2H of results).
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALG_BYTES_NOTE = "refLen + readLen + n*n + 40 bytes per pair (SURVEY 8(d))"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=1000, help="reads per GPU")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--ref-len", type=int, default=5_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0: 4 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inst", type=int, default=-1, help="experiment: force a fill-kernel instance")
    ap.add_argument("--chunk", type=int, default=0, help="experiment: reference chunk length")
    ap.add_argument("--opt", action="append", default=[], help="experiment: engine option name=value (repeatable)")
    ap.add_argument("--clock-sampler", choices=["smi", "nvml"], default="smi",
                    help="how clocks / throttle reasons are sampled during the timed region")
    ap.add_argument("--lib", default="libssw.so", help="experiment: alternative build of the library")
    return ap.parse_args()


def workload(args, rank):
    import common as C
    ref, reads = C.make_dna_workload(args.ref_len, args.reads, args.read_len, seed_ref=1001, seed_reads=2002 + 7919 * rank)
    return ref, reads, C.dna_matrix(2, 2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler:
    """The same quantities read in-process through NVML (nvidia_ml_py) every 100 ms: no nvidia-smi process polling the
    driver while the timed region runs (measured: the -lms 200 poller costs the timed steps several ms each)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        seen = 0
        for r in self.rows:
            seen |= int(r[2])
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "reasons": sorted(k for k, b in bits.items() if seen & b), "samples": len(self.rows), "source": "nvml, 100 ms period"}


def cpu_reference_rate(ref, reads, mat, n_threads, sample):
    """Time the reference's SSE2 path (oracle/_ref/libssw_ref.so; the scalar oracle port if that is absent)
    on `sample` reads with `n_threads` host threads.  Returns (GCUPS, kind, seconds)."""
    import common as C
    if C.have_ref():
        lib, kind = C.load_ref(), "reference"
    else:
        lib, kind = C.load_oracle(), "port"
    refp = C.i8ptr(ref)
    matp = C.i8ptr(mat)

    def one(q):
        p = lib.ssw_init(C.i8ptr(q), len(q), matp, 5, 2)
        r = lib.ssw_align(p, refp, len(ref), 3, 1, 0, 0, 0, len(q) // 2)   # ctypes releases the GIL during the call
        s = r.contents.score1
        lib.align_destroy(r)
        lib.init_destroy(p)
        return s

    qs = [np.ascontiguousarray(reads[i % len(reads)]) for i in range(sample)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        list(ex.map(one, qs))
    dt = time.perf_counter() - t0
    cells = float(sum(len(q) for q in qs)) * len(ref)
    return cells / dt / 1e9, kind, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref, reads, mat = workload(args, 0)
    cores = os.cpu_count() or 1
    if args.ref_len > 1_000_000 and not __import__("common").have_ref():
        sample = max(2, cores // 8)          # scalar port is ~20x slower than SSE2
    else:
        sample = args.cpu_sample or 4 * cores
    vals = []
    for s in range(args.warmup + args.steps):
        v, kind, dt = cpu_reference_rate(ref, reads, mat, cores, sample)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {"impl": "reference", "metric": "GCUPS (DP cell updates/s), ssw_align forward path", "value": value, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 (SSE2 16-lane saturating)", "data": "synthetic",
            "config": {"workload": "config2: %d x %d bp reads vs %d bp reference, byte-score path, flag 0 (bounded sample of %d reads per step)"
                                   % (args.reads, args.read_len, args.ref_len, sample)},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": "%d reads x %d bp reference per step, %d host threads" % (sample, args.ref_len, cores)},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import common as C
    from __graft_entry__ import load_package
    L = load_package()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ssw_b200_dist", os.path.join(ROOT, "complete-striped-smith-waterman-library_b200", "ssw_dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ref, reads, mat = workload(args, rank)
    cells_rank = float(sum(len(q) for q in reads)) * len(ref)
    eng = L.BatchAligner(device=local, lib_name=args.lib)
    if args.inst >= 0:
        eng.set_option("inst", args.inst)
    if args.chunk:
        eng.set_option("chunk", args.chunk)
    for kv in args.opt:
        name, val = kv.split("=")
        eng.set_option(name, int(val))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        res, _ = eng.align(mat, 5, 3, 1, flag=0, mask_len=args.read_len // 2, score_size=2)
        return res

    def gather(res):
        # fixed-size result records of every rank -> rank 0 (NCCL gather over NVLink), the only collective of a step
        return D.gather_records(res, rank, world, device="cuda:%d" % local)

    # ---- device-resident timing (`value`) ----
    eng.set_sequences(reads, [ref])
    for _ in range(args.warmup):
        gather(step_resident())
    sampler = {"smi": ClockSampler, "nvml": NvmlSampler}[args.clock_sampler](local)
    if rank == 0:
        sampler.start()
    tot_ms = fill_ms = 0.0
    launches = 0
    fill_launches = 0
    wall = 0.0
    for _ in range(args.steps):
        flush.zero_()                       # flush L2 between timed iterations (untimed)
        barrier()
        t0 = time.perf_counter()
        res = step_resident()
        gather(res)
        barrier()
        wall += time.perf_counter() - t0
        tm = eng.timing()
        tot_ms += tm["total_ms"]
        fill_ms += tm["fill_forward_ms"]
        launches += tm["fill_forward_launches"] + tm["other_launches"]
        fill_launches += tm["fill_forward_launches"]
    clocks = sampler.stop() if rank == 0 else None
    # device time (CUDA events on the engine's stream) and the driver-visible wall clock; max over ranks
    tt = torch.tensor([tot_ms, wall * 1e3, fill_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, fill_ms_max = [float(x) for x in tt.tolist()]
    step_ms = max(dev_ms, wall_ms) / args.steps         # the slower of the two clocks: never flatter than the driver's own
    value = cells_rank * world / (step_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer C ABI (`e2e`) ----
    qc, qo = L.concat(reads)
    rc, ro = L.concat([ref])
    out = (ct.c_void_p * len(reads))()
    lib = eng.lib
    lib.ssw_align_batch.argtypes = [ct.c_void_p, ct.POINTER(L.BatchParams), ct.c_int32, ct.POINTER(ct.c_int8), ct.POINTER(ct.c_int64),
                                    ct.c_int32, ct.POINTER(ct.c_int8), ct.POINTER(ct.c_int64), ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p]
    lib.ssw_align_batch.restype = ct.c_int
    lib.align_destroy.argtypes = [ct.c_void_p]
    mat_c = np.ascontiguousarray(mat, dtype=np.int8)
    P = L.BatchParams(mat_c.ctypes.data_as(ct.POINTER(ct.c_int8)), 5, 3, 1, 0, 0, 0, args.read_len // 2, 2)

    def step_e2e():
        rv = lib.ssw_align_batch(eng.h, ct.byref(P), len(reads), qc.ctypes.data_as(ct.POINTER(ct.c_int8)), qo.ctypes.data_as(ct.POINTER(ct.c_int64)),
                                 1, rc.ctypes.data_as(ct.POINTER(ct.c_int8)), ro.ctypes.data_as(ct.POINTER(ct.c_int64)), len(reads), None, None, out)
        assert rv == 0
        sc = [ct.cast(out[i], ct.POINTER(C.SAlign)).contents.score1 for i in range(len(reads))]
        for i in range(len(reads)):
            lib.align_destroy(out[i])
        return sc

    step_e2e()
    e2e_wall = 0.0
    for _ in range(max(2, args.steps // 2)):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        barrier()
        e2e_wall += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_wall / max(2, args.steps // 2)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = cells_rank * world / float(e2e_t.item()) / 1e9

    if rank == 0:
        hbm_peak, peak_src = peaks()
        n = 5
        alg_bytes_launch = sum(len(ref) + len(q) + n * n + 40 for q in reads)          # per rank, one fill launch
        fill_s = fill_ms_max / 1e3 / args.steps          # all forward-fill launches of one step (byte pass + word re-run of overflows)
        achieved = alg_bytes_launch / fill_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_fill.json")
        if os.path.exists(tp) and args.reads == 1000 and args.ref_len == 5_000_000:
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")      # ncu dram__bytes_read.sum + dram__bytes_write.sum, full-size launch
        line = {"metric": "GCUPS (DP cell updates/s), ssw_align forward path", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s16x2 (byte-score semantics in 16-bit DPX lanes)", "data": "synthetic",
                "config": {"workload": "config2: %d x %d bp reads per GPU vs one %d bp reference, byte-score path, flag 0, maskLen %d"
                                       % (args.reads, args.read_len, args.ref_len, args.read_len // 2),
                           "l2": "256 MB buffer written between timed steps (L2 flush); 10 GB column-maximum scratch rewritten every step",
                           "timing": "per step: max(CUDA-event time on the engine stream, wall clock between synchronised barriers), max over ranks"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "GCUPS",
                        "h2d_bytes_per_step": int(len(qc) + len(rc) + len(mat_c)), "d2h_bytes_per_step": int(36 * len(reads)),
                        "note": "ssw_align_batch() on pageable host buffers: H2D of all sequences, kernels, D2H of results, malloc'd s_align records"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": ALG_BYTES_NOTE,
                             "kernel": "ssw_fill_kernel<8,20,+1> (forward fill: byte pass + word re-fill of the byte overflows)", "kernel_ms_per_step": fill_s * 1e3,
                             "kernel_launches_per_step": fill_launches / args.steps,
                             "note": "integer-issue bound recurrence (150 cells per reference byte): the HBM fraction is reported as required, "
                                     "the meaningful efficiency is alu_roofline"},
                "alu_roofline": alu_roofline(cells_rank, fill_s)}
        if not args.no_cpu_baseline and world == 1:            # reported baseline, rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or 4 * cores
            if not C.have_ref():
                sample = max(2, cores // 8)
            v, kind, dt = cpu_reference_rate(ref, reads, mat, cores, sample)
            line["cpu_baseline"] = {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind,
                                    "sample": "%d reads x %d bp reference, %d host threads, %.1f s" % (sample, args.ref_len, cores, dt)}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def alu_roofline(cells, fill_s):
    """Cell updates/s of the fill kernel against the measured DPX issue peak (profiles/dpx_peak.json, written by
    tools/microbench on the B200): 5.5 packed-s16x2 ops per 2 cells."""
    p = os.path.join(ROOT, "profiles", "dpx_peak.json")
    peak = None
    if os.path.exists(p):
        with open(p) as f:
            peak = json.load(f).get("gcups_peak_5p5_ops_per_cellpair")
    ach = cells / fill_s / 1e9
    return {"achieved_gcups": ach, "peak_gcups": peak, "frac": (ach / peak) if peak else None,
            "unit": "GCUPS", "basis": "measured VIADDMNMX.S16x2 issue rate x 148 SMs / 2.75 ops per cell"}


if __name__ == "__main__":
    main()
