#!/usr/bin/env python
"""bench.py -- headline benchmark of the swga hot path on B200.

Metric (BASELINE.json): Gbases/s of all-k (k = 5..12) both-strand counting.  One "step" = one pass of
the counting path over the workload: pack + mark + accumulate (+ NCCL allreduce at N > 1) + finalize +
fold for the foreground AND the background genome, producing the canonical tables of both.

Workload at every N: BASELINE.json configs[1] (= C2): synthetic 4.4 Mbp GC-rich foreground, 3.1 Gbp
human-sized background (SURVEY.md 8d generator).  At N > 1 the background is sharded by contiguous
chunk with an 11-base halo (strong scaling: total work fixed), the foreground is replicated.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KMIN, KMAX = 5, 12
METRIC = "allk_count_throughput"
UNIT = "Gbases/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--bg-bases", type=float, default=None, help="override background size (debug only)")
    ap.add_argument("--cpu-sample-bases", type=int, default=600_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own DSK binary (oracle/_ref/dsk) + the restated Python record parser
# ------------------------------------------------------------------------------------------------

def _dsk_one(args):
    fa, k, tmp = args
    from oracle import swga_oracle as O
    out = O.run_dsk(fa, k, tmp)
    n = sum(1 for _ in O.parse_kmer_binary(out))          # swga/kmers.py:52-54 builds the dict from it
    os.remove(out)
    return n


def cpu_reference_pass(workload, sample_bases, parallel):
    """One pass of the reference CPU path (dsk k=5..12 + record parsing) over a bounded sample of the
    workload's background genome.  Returns (seconds, bases, kind, cores)."""
    from oracle import swga_oracle as O
    from swga_b200 import synth
    seed, n, gc, n_rec = synth.CONFIGS[workload + "_bg"]
    sample = min(sample_bases, n)
    seq = synth.bases(seed, 0, sample, gc)
    have_dsk = os.access(os.path.join(ROOT, "oracle", "_ref", "dsk"), os.X_OK)
    with tempfile.TemporaryDirectory() as tmp:
        if have_dsk:
            fa = os.path.join(tmp, "sample.fa")
            synth.write_fasta(fa, seq, synth.equal_records(sample, 1))
            jobs = [(fa, k, tmp) for k in range(KMIN, KMAX + 1)]
            t0 = time.perf_counter()
            if parallel > 1:
                import multiprocessing as mp
                with mp.get_context("fork").Pool(parallel) as pool:
                    pool.map(_dsk_one, jobs)
            else:
                for j in jobs:
                    _dsk_one(j)
            dt = time.perf_counter() - t0
            return dt, sample, "reference", max(1, parallel)
        rs = synth.equal_records(sample, 1)
        t0 = time.perf_counter()
        O.count_dense_allk(seq, rs, KMIN, KMAX)
        return time.perf_counter() - t0, sample, "port", 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(KMAX - KMIN + 1, os.cpu_count() or 1)
    vals = []
    for i in range(args.warmup + args.steps):
        dt, bases, kind, used = cpu_reference_pass(args.workload, args.cpu_sample_bases, cores)
        if i >= args.warmup:
            vals.append(dt)
    ms = 1e3 * sum(vals) / len(vals)
    value = bases / (ms * 1e-3) / 1e9
    sample = ("%d-base prefix of the %s background, dsk k=5..12 as %d concurrent processes (the reference runs "
              "them one after another on one core) + Python record parser" % (bases, args.workload, used))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "sample_bases": bases},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(w):
    from swga_b200 import synth
    fg, bg = synth.CONFIGS[w + "_fg"], synth.CONFIGS[w + "_bg"]
    return "%s: synthetic %.1f Mbp fg (gc %.3f) + %.1f Mbp bg (gc %.2f, %d records), all-k %d..%d both strands" % (
        w, fg[1] / 1e6, fg[2], bg[1] / 1e6, bg[2], bg[3], KMIN, KMAX)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML every ~2 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for pre in ("GPU-", ""):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID((pre + uuid).encode())
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((clk, mx, int(r)))
                time.sleep(0.002)
        except Exception as e:                      # NVML missing: fall back to nvidia-smi (slow, few samples)
            self.error = repr(e)
            q = "clocks.sm,clocks.max.sm"
            while not self.stop_flag:
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, timeout=5).stdout.decode()
                    f = [float(x) for x in out.strip().split(",")]
                    self.samples.append((f[0], f[1], 0))
                except Exception:
                    pass
                time.sleep(0.05)

    def summary(self):
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        sm = [s[0] for s in self.samples]
        reasons = sorted({name for s in self.samples for name, bit in bits.items() if s[2] & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(s[1] for s in self.samples) if sm else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated (first touch
    then places the buffers on the GPU's NUMA node): with 8 ranks copying at once, buffers on the far
    socket cap the aggregate H2D rate.  Best effort; returns the node or None."""
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        h = None
        for pre in ("GPU-", ""):
            try:
                h = nv.nvmlDeviceGetHandleByUUID((pre + uuid).encode())
                break
            except Exception:
                h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(index)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                      # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        with open(base + "/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return None


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as td
    from swga_b200 import _lib, dist, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            td.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib, ctx = _lib.load(), _lib.context(local)
    P = _lib.ptr

    fg_seed, n_fg, fg_gc, fg_nrec = synth.CONFIGS[args.workload + "_fg"]
    bg_seed, n_bg, bg_gc, bg_nrec = synth.CONFIGS[args.workload + "_bg"]
    if args.bg_bases:
        n_bg = int(args.bg_bases)
    words = int(lib.swga_tables_words(KMIN, KMAX))

    class Dev(object):
        """one genome (or one rank's shard of it) resident in HBM"""

        def __init__(self, seed, n_total, gc, n_rec, shard):
            rs = synth.equal_records(n_total, n_rec)
            a, b = dist.shard_bounds(n_total, rank, world) if shard else (0, n_total)
            e = min(n_total, b + KMAX - 1)
            self.n, self.n_starts, self.a = e - a, b - a, a
            self.mode_b = bool(np.diff(rs.astype(np.int64))[:10001].max() > 1000000)
            inner = rs[(rs > a) & (rs < e)] - np.uint64(a)
            self.rs = np.concatenate([[0], inner, [e - a]]).astype(np.uint64)
            ta, tc, tg = synth.thresholds(gc)
            self.ascii = torch.empty(max(self.n, 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.swga_synth_bases(ctx.h, seed, a, self.n, ta, tc, tg, P(self.ascii), None))
            self.d_rs = torch.from_numpy(self.rs.view(np.int64).copy()).to(dev)
            self.packed = torch.empty(int(lib.swga_packed_words(self.n)), dtype=torch.int32, device=dev)
            self.brk = torch.empty(int(lib.swga_brk_words(self.n)), dtype=torch.int32, device=dev)
            self.fwd = torch.empty(words, dtype=torch.int32, device=dev)
            self.canon = torch.empty(words, dtype=torch.int32, device=dev)
            self.launches = 0

        def accumulate(self, ev=None):
            """zero + pack + mark + swga_count_accumulate: this rank's additive table block"""
            self.fwd.zero_()
            _lib.check(lib.swga_pack(ctx.h, P(self.ascii), self.n, 0 if self.mode_b else 1, P(self.packed), P(self.brk), None))
            k = 1
            if len(self.rs) > 2:
                _lib.check(lib.swga_mark_records(ctx.h, P(self.d_rs[1:]), len(self.rs) - 2, self.n, P(self.brk), None))
                k += 1
            if ev is not None:
                ev[0].record()
            _lib.check(lib.swga_count_accumulate(ctx.h, P(self.packed), P(self.brk), self.n_starts, KMIN, KMAX, P(self.fwd), None))
            if ev is not None:
                ev[1].record()
            # accumulate = k_bucket_hist (sample) + k_bucket_layout + k_partition + k_item_prefix + k_bucket_count
            # when the partitioned path applies (>= 14 Mi starts, kmax 11 or 12), else the single k_count
            acc = 5 if (self.n_starts >= (14 << 20) and 8 <= 2 * KMAX - 14 <= 10) else 1
            # + k_marginalize per level + k_fold (k < 10) + k_fold_tiled per k >= 10
            self.launches = k + acc + (KMAX - KMIN) + (1 if KMIN < 10 else 0) + max(0, KMAX - max(KMIN, 10) + 1)

        def finish(self):
            """finalize + fold of the (allreduced) block"""
            _lib.check(lib.swga_count_finalize(ctx.h, KMIN, KMAX, P(self.fwd), None))
            _lib.check(lib.swga_fold_canonical(ctx.h, KMIN, KMAX, P(self.fwd), P(self.canon), None))

    fg = Dev(fg_seed, n_fg, fg_gc, fg_nrec, shard=False)
    bg = Dev(bg_seed, n_bg, bg_gc, bg_nrec, shard=True)
    torch.cuda.synchronize()

    def step(ev=None):
        # the background's allreduce (NCCL, its own stream) runs while this rank counts the replicated foreground
        bg.accumulate(ev)
        work = td.all_reduce(bg.fwd, op=td.ReduceOp.SUM, async_op=True) if world > 1 else None
        fg.accumulate()
        fg.finish()
        if work is not None:
            work.wait()
        bg.finish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    e0.record()
    for i in range(args.steps):
        step(kev[i])
    e1.record()
    barrier()
    sampler.stop_flag = True
    total_ms = e0.elapsed_time(e1)
    count_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([total_ms, count_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    total_ms, count_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = (n_fg + n_bg) / (ms_per_step * 1e-3) / 1e9
    sampler.join(timeout=2)
    clocks = sampler.summary()

    # ---- per-kernel times of the count stage: CUDA events recorded inside swga_count_accumulate, on its stream ----
    part_ms = hist_ms = bcount_ms = None
    _lib.check(lib.swga_ctx_set_stage_timing(ctx.h, 1))
    acc = [0.0, 0.0, 0.0]
    n_st = max(3, min(args.steps, 10))
    for _ in range(n_st):
        bg.accumulate()
        st = (ctypes.c_float * 3)()
        _lib.check(lib.swga_count_stage_ms(ctx.h, st))      # waits for this background count
        for i in range(3):
            acc[i] += st[i] / n_st
    _lib.check(lib.swga_ctx_set_stage_timing(ctx.h, 0))
    torch.cuda.synchronize()
    t = torch.tensor(acc, dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    hist_ms, part_ms, bcount_ms = (float(x) for x in t)
    partitioned = part_ms > 0

    # ---- roofline of the dominant kernel on this rank's background shard ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    tr = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_count_stage_traffic.json")) as f:
            tr = json.load(f)
    except Exception:
        pass
    if partitioned:
        # k_partition reads the packed genome once: 1/4 B per base (+ 1/32 B of break mask); the 2-byte bucket
        # elements it writes are this implementation's own intermediate, counted under `traffic`, not here
        k_bytes = (0.25 + 1.0 / 32.0) * bg.n_starts
        k_ms, k_name = part_ms, "k_partition"
        k_traffic = tr.get("k_partition_dram_bytes_per_base")
    else:
        k_bytes = (0.25 + 1.0 / 32.0) * bg.n_starts + 4.0 * 4 ** KMAX
        k_ms, k_name = count_ms, "k_count"
        k_traffic = None
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    stage_bytes = 0.25 * bg.n_starts + 4.0 * 4 ** KMAX         # packed genome read once + top table written once
    roofline = {"kernel": k_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": k_traffic * bg.n_starts if k_traffic else None,
                "peak_source": peak_src, "ms_per_launch": k_ms, "share_of_step": k_ms / ms_per_step,
                "dram_gbs_actual": (k_traffic * bg.n_starts / (k_ms * 1e-3) / 1e9) if k_traffic else None,
                "count_stage": {"kernels": "k_bucket_hist(sample) + k_bucket_layout + k_partition + k_item_prefix + k_bucket_count"
                                if partitioned else "k_count",
                                "ms": count_ms, "share_of_step": count_ms / ms_per_step,
                                "ms_hist_layout": hist_ms, "ms_partition": part_ms, "ms_bucket_count": bcount_ms,
                                "bytes_alg": stage_bytes, "achieved": stage_bytes / (count_ms * 1e-3) / 1e9,
                                "frac": stage_bytes / (count_ms * 1e-3) / 1e9 / hbm_peak,
                                "traffic": tr.get("dram_bytes_per_base", 0) * bg.n_starts or None},
                "note": "k_partition is bound by the L1/shared-memory data pipe (ncu: l1tex data-pipe wavefronts 77-85 % of peak, "
                        "DRAM 21 %), not by HBM: every element costs one shared atomic and one 2-byte shared store with "
                        "~3.8-way bank serialisation; see 'atomic' and profiles/r01_summary.md"}
    path_bytes = 1.5 * (fg.n + bg.n) + 2 * 2 * 4.0 * words       # 1.5 B/base + 2T per genome
    path = {"bytes_alg": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s"}
    path["frac"] = path["achieved"] / hbm_peak

    # ---- A_peak: random red.global.add.u32 on a 64 MB table; shared atomics on a 64 KB table ----
    atomic = None
    if rank == 0 and not args.no_extras:
        ms = ctypes.c_float()
        tab = torch.zeros(1 << 24, dtype=torch.int32, device=dev)
        _lib.check(lib.swga_bench_red_global(ctx.h, P(tab), 1 << 24, 1 << 31, ctypes.byref(ms)))
        red_peak = (1 << 31) / (ms.value * 1e-3) / 1e9
        _lib.check(lib.swga_bench_atom_shared(ctx.h, 1 << 14, 1 << 33, ctypes.byref(ms)))
        sh_peak = (1 << 33) / (ms.value * 1e-3) / 1e9
        ach = bg.n_starts / (count_ms * 1e-3) / 1e9
        atomic = {"achieved_gops": ach, "red_global_64MB_peak_gops": red_peak, "frac": ach / red_peak,
                  "atom_shared_64KB_peak_gops": sh_peak, "ops_alg_per_base": 1,
                  "note": "1 table update per base (top table k=12 + tails; k<12 by marginalisation). The partitioned "
                          "path spends 2 shared-memory atomics/base (staging slot in k_partition, bin in k_bucket_count) "
                          "instead of 1 L2 red/base"}
        del tab

    # ---- e2e: host buffers through the C-ABI host calls, H2D + D2H inside the timed region ------
    e2e = None
    if not args.no_e2e:
        h_fg = torch.empty(fg.n, dtype=torch.uint8, pin_memory=True)
        h_bg = torch.empty(bg.n, dtype=torch.uint8, pin_memory=True)
        h_fg.copy_(fg.ascii[:fg.n])
        h_bg.copy_(bg.ascii[:bg.n])
        h_canon_fg = torch.empty(words, dtype=torch.int32, pin_memory=True)
        h_canon_bg = torch.empty(words, dtype=torch.int32, pin_memory=True)
        cs = torch.cuda.ExternalStream(lib.swga_ctx_compute_stream(ctx.h), device=dev)
        fg_rs, bg_rs = np.ascontiguousarray(fg.rs), np.ascontiguousarray(bg.rs)

        def e2e_step():
            _lib.check(lib.swga_count_genome_host(ctx.h, P(h_fg), fg.n, P(fg_rs), len(fg_rs) - 1, int(fg.mode_b),
                                                  KMIN, KMAX, P(h_canon_fg)))
            with torch.cuda.stream(cs):
                bg.fwd.zero_()
                _lib.check(lib.swga_count_stream_host(ctx.h, P(h_bg), bg.n, P(bg_rs), len(bg_rs) - 1, int(bg.mode_b),
                                                      KMIN, KMAX, bg.n_starts, P(bg.fwd)))
                if world > 1:
                    td.all_reduce(bg.fwd, op=td.ReduceOp.SUM)
                _lib.check(lib.swga_count_finalize(ctx.h, KMIN, KMAX, P(bg.fwd), cs.cuda_stream))
                _lib.check(lib.swga_fold_canonical(ctx.h, KMIN, KMAX, P(bg.fwd), P(bg.canon), cs.cuda_stream))
                h_canon_bg.copy_(bg.canon, non_blocking=True)
            cs.synchronize()

        n_e2e = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(dt, op=td.ReduceOp.MAX)
        e2e = {"value": (n_fg + n_bg) / float(dt[0]) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(fg.n + bg.n + 8 * (len(fg_rs) + len(bg_rs))),
               "d2h_bytes_per_step": int(2 * 4 * words), "ms_per_step": float(dt[0]) * 1e3, "steps": n_e2e,
               "api": "swga_count_genome_host / swga_count_stream_host (pinned host buffers)"}
        # parity of the two paths on the full-size workload: a checksum of the tables
        same = bool(torch.equal(h_canon_bg.to(dev), bg.canon))
        e2e["tables_equal_device_path"] = same

    # ---- size-independent parity properties of the full-size tables (no CPU oracle at 3.1 Gbp) ----
    props = None
    if world == 1 and not args.no_extras:
        step()
        torch.cuda.synchronize()
        ok = True
        lens = np.diff(synth.equal_records(n_bg, bg_nrec).astype(np.int64))
        for k in range(KMIN, KMAX + 1):
            o0, o1 = int(lib.swga_table_offset(KMIN, k)), int(lib.swga_table_offset(KMIN, k + 1))
            expect = int(np.maximum(0, lens - k + 1).sum())
            ok &= int(bg.fwd[o0:o1].sum(dtype=torch.int64)) == expect          # every window counted once
            ok &= int(bg.canon[o0:o1].sum(dtype=torch.int64)) == expect        # fold loses nothing
            if k < KMAX:                                                       # table k = marginal of k+1 + one tail per record
                up = bg.fwd[o1:o1 + 4 ** (k + 1)].view(-1, 4).sum(1, dtype=torch.int64)
                diff = bg.fwd[o0:o1].to(torch.int64) - up
                ok &= bool((diff >= 0).all()) and int(diff.sum()) == bg_nrec
        props = {"bg_tables_sum_and_marginals_ok": bool(ok)}

    # ---- the count -> primer-row merge of `swga count` on the tables just built (BASELINE config 2: count + filter) ----
    merge = None
    if rank == 0 and not args.no_extras:
        cap = 1 << 22
        rows = torch.empty(cap * 4, dtype=torch.int32, device=dev)
        n_rows = torch.zeros(1, dtype=torch.int32, device=dev)
        min_fg, max_bg = int(1e-5 * n_fg), int(6.7e-6 * n_bg)         # the defaults `swga init` derives (init.py:34-35)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            n_rows.zero_()
            f0.record()
            _lib.check(lib.swga_filter_primers(ctx.h, P(fg.canon), P(bg.canon), None, KMIN, KMAX, min_fg, max_bg, 3, 1,
                                               P(rows), cap, P(n_rows), None))
            f1.record()
            torch.cuda.synchronize()
            ms = f0.elapsed_time(f1)
            best = ms if best is None else min(best, ms)
        merge = {"kernel": "k_filter_primers", "ms": best, "rows": int(n_rows.item()), "min_fg_bind": min_fg,
                 "max_bg_bind": max_bg, "max_dimer_bp": 3, "codes_tested": int(words)}
        del rows

    scoring = None
    if rank == 0 and not args.no_extras:
        try:
            scoring = scoring_extra(dev)
        except Exception as e:                      # the counting line must survive a failure of the extra
            scoring = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        dt, bases, kind, cores = cpu_reference_pass(args.workload, args.cpu_sample_bases, 1)
        cpu = {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d-base prefix of the background: dsk k=5..12 run one after another (as "
                         "swga/commands/count.py:84-91 does) + record parser; %.1f s" % (bases, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "fg_bases": n_fg, "bg_bases": n_bg,
                       "kmin": KMIN, "kmax": KMAX, "sharding": "bg contiguous chunks + 11-base halo, fg replicated"
                       if world > 1 else "single GPU", "l2": "inputs (%.2f GB/GPU) larger than L2" % (bg.n / 1e9)},
            "roofline": roofline, "path_roofline": path, "atomic": atomic, "cpu_baseline": cpu, "e2e": e2e,
            "clocks": clocks, "gpu_launches": (fg.launches + bg.launches) * args.steps, "numa_node": numa,
            "scoring": scoring, "count_merge": merge, "full_size_properties": props,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def scoring_extra(dev):
    """Second headline of BASELINE.json: candidate sets scored per second on one B200 (kernel (d)), device
    resident site lists and sets, CUDA events.  Workload C4 of SURVEY.md 8(d) on a 2x10^5-set sample of the
    10^7 (the 200 most frequent 8..12-mers of the 23 Mbp AT-rich foreground: ~2x10^5 unique sites per
    set), plus a 'typical' variant with 200 12-mers of ~400 sites each (the scale of swga's default
    min_fg_bind = 1e-5 * fg_length), 10^6 sets."""
    import numpy as np
    import torch
    from swga_b200 import fasta, kmers, locate, score, synth, workloads
    seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
    g = fasta.Genome(synth.bases(seed, 0, n, gc), synth.equal_records(n, n_rec), ["chr%d" % (i + 1) for i in range(n_rec)])
    tabs = kmers.count_all(g, 5, 12)
    gi = locate.GenomeIndex(g)
    out = {"unit": "sets/s", "fg_bases": n, "timing": "CUDA events, best of 3, device-resident inputs; e2e = host API (numpy in, numpy out), wall clock, best of 3"}
    t12 = tabs.table(12)
    order = np.argsort(-t12.astype(np.int64), kind="stable")
    lo = int(np.searchsorted(-t12[order].astype(np.int64), -400))
    variants = {
        "C4_dense": ([kmers.codes_to_kmers([c], k)[0] for k, c, _ in workloads.top_primers(tabs, 5, 12, 200)], 200_000),
        "typical_400_sites": (kmers.codes_to_kmers(order[lo:lo + 200], 12), 1_000_000),
    }
    for name, (seqs, S) in variants.items():
        sl = score.SiteLists.from_index(gi, seqs)
        sets_, _ = workloads.c4_sets(S)
        mapped = np.where(sets_ >= 0, sl.list_of_primer[np.clip(sets_, 0, 199)], -1).astype(np.int32)
        mapped_host = sets_
        d_sets = torch.from_numpy(mapped).to(dev)
        bg = torch.full((S,), 1000.0, dtype=torch.float64, device=dev)
        row = {"sets": S}
        for label, max_fg in (("max_fg_36000", 36000), ("no_reject", 2 ** 31 - 1)):
            best = None
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = score.score_sets(sl, None, bg, max_fg, d_sets=d_sets, return_device=True)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            L = res.n_sites.cpu().numpy().view(np.uint32)
            st = res.status.cpu().numpy()
            # end to end through the host API: set array and bg_dist_mean from host memory, results back to numpy
            h_bg = np.full(S, 1000.0)
            e2e_s = None
            for _ in range(3):                                   # best of 3, like the device timing above
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                score.score_sets(sl, sets_, h_bg, max_fg)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                e2e_s = dt if e2e_s is None else min(e2e_s, dt)
            n_pass = int((st & 1).sum())
            # score.score_sets: bulk calls copy back status + max_dist of every set and the rest for the passing ones
            d2h = S * 5 + n_pass * 44 if (S > (1 << 17) and 2 * n_pass <= S) else S * 41
            row[label] = {"sets_per_s": S / best * 1e3, "ms": best, "mean_unique_sites_per_set": float(L.mean()),
                          "passed_frac": float((st & 1).mean()),
                          "merged_sites_per_s": float(sl.counts()[:-1].mean()) * 8 * S / best * 1e3,
                          "e2e_sets_per_s": S / e2e_s, "e2e_h2d_bytes": int(mapped_host.nbytes + h_bg.nbytes),
                          "e2e_d2h_bytes": int(d2h)}
        out[name] = row
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
