#!/usr/bin/env python
"""bench.py -- headline benchmark of the swga hot path on B200.

Metric (BASELINE.json): Gbases/s of all-k (k = 5..12) both-strand counting.  One "step" = one pass of
the counting path over the workload: pack + mark + accumulate (+ the cross-GPU merge at N > 1) + finalize +
fold for the foreground AND the background genome, producing the canonical tables of both on every GPU.

Workload at every N: BASELINE.json configs[1] (= C2): synthetic 4.4 Mbp GC-rich foreground, 3.1 Gbp
human-sized background (SURVEY.md 8d generator).  At N > 1 the background is sharded by contiguous
chunk with an 11-base halo (strong scaling: total work fixed), the foreground is replicated.  The same
step on BASELINE configs[2] (C3: 23 Mbp AT-rich foreground) is timed as the per-N extra `c3`.

The second half of BASELINE.json's metric -- candidate sets scored per second (configs[3], C4) -- is the
`scoring` object of the same line: sets sharded by set index over the ranks, no collective.

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
CPU_SAMPLE_DEFAULT = 10_000_000          # = the C1 background: DSK's ~1 s per invocation is < 10 % of a run


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--bg-bases", type=float, default=None, help="override background size (debug only)")
    ap.add_argument("--cpu-sample-bases", type=int, default=None,
                    help="bases of the background prefix the CPU arm counts per step (default: 10 Mbp in the "
                         "reference arm, 5 Mbp in the cpu_baseline leg of the b200 arm)")
    ap.add_argument("--cpu-budget-s", type=float, default=900.0,
                    help="reference arm: shrink the sample so that all timed steps fit this many seconds")
    ap.add_argument("--score-sets", type=float, default=1e7, help="candidate sets of the 'typical' scoring workload")
    ap.add_argument("--score-sets-dense", type=float, default=1e6, help="candidate sets of the C4-dense scoring workload")
    ap.add_argument("--score-sets-thin", type=float, default=1e6, help="candidate sets of the thin-set scoring workload (~7,200 sites per set)")
    ap.add_argument("--merge", default="peer", choices=["peer", "multicast", "nccl"],
                    help="cross-GPU merge: the fused kernel over peer pointers (default; measured fastest), the same kernel with "
                         "the reduction inside the NVSwitch (multimem.ld_reduce / multimem.st), or NCCL all_reduce + finalize "
                         "on every rank")
    ap.add_argument("--via-pack", action="store_true", help="A/B: k_pack + partitioned kernels on its output (round-1 data flow)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--verify-full", action="store_true",
                    help="also compare every table of the full-size background with the CPU oracle (the checker) "
                         "counted on the host cores in parallel; result in full_size_properties.oracle_equal")
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
            jobs = [(fa, k, tmp) for k in range(KMAX, KMIN - 1, -1)]         # longest job first
            t0 = time.perf_counter()
            if parallel > 1:
                import multiprocessing as mp
                with mp.get_context("fork").Pool(parallel) as pool:
                    pool.map(_dsk_one, jobs, chunksize=1)
            else:
                for j in jobs:
                    _dsk_one(j)
            dt = time.perf_counter() - t0
            return dt, sample, "reference", max(1, parallel)
        rs = synth.equal_records(sample, 1)
        t0 = time.perf_counter()
        O.count_dense_allk(seq, rs, KMIN, KMAX)
        return time.perf_counter() - t0, sample, "port", 1


def cpu_full_runs():
    """Full C1 / C5 runs of the reference path (tools/cpu_full_runs.py, tens of minutes, measured once)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_cpu_full_runs.json")) as f:
            d = json.load(f)
        return {"source": "profiles/r02_cpu_full_runs.json (tools/cpu_full_runs.py; host %s, %s cpus; dsk k=5..12 one after "
                          "another on 1 core + record parser, full genomes)" % (d.get("host"), d.get("cpus")),
                "runs": {c: {k: v for k, v in r.items() if k != "per_run"} for c, r in d.get("runs", {}).items()}}
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(KMAX - KMIN + 1, os.cpu_count() or 1)
    want = args.cpu_sample_bases or CPU_SAMPLE_DEFAULT
    # Warm-up steps run a 1 Mbp sample (they only warm the page cache and the binary); the first one also sizes the
    # timed sample: DSK costs ~1 s per invocation plus a part linear in the bases, and all timed steps have to fit
    # --cpu-budget-s (the default 10 Mbp sample is kept when it does).
    t_small = None
    for _ in range(args.warmup):
        dt, b, _, _ = cpu_reference_pass(args.workload, min(want, 1_000_000), cores)
        t_small = dt if t_small is None else min(t_small, dt)
    sample = want
    if t_small is not None and want > 1_000_000:
        per_mbp = max(t_small - 1.0, 0.05 * t_small)
        predicted = 1.0 + per_mbp * want / 1e6
        if predicted * args.steps > args.cpu_budget_s:
            sample = int(max(2_000_000, min(want, (args.cpu_budget_s / args.steps - 1.0) / per_mbp * 1e6)))
    vals = []
    for _ in range(args.steps):
        dt, bases, kind, used = cpu_reference_pass(args.workload, sample, cores)
        vals.append(dt)
    ms = 1e3 * sum(vals) / len(vals)
    value = bases / (ms * 1e-3) / 1e9
    desc = ("%d-base prefix of the %s background per step, dsk k=5..12 as %d concurrent processes (the reference runs "
            "them one after another on one core) + Python record parser; warm-up steps use 1 Mbp"
            % (bases, args.workload, used))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "sample_bases": bases},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": desc,
                         "full_runs": cpu_full_runs()},
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
        self.nv = self.h = self.mx = None
        # NVML is opened HERE, before the timed region: at N = 8 ten steps last 11 ms, and a sampler that first had to
        # initialise the library and look its device up came back with no sample at all
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
            self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.nv, self.h = nv, h
        except Exception as e:
            self.error = repr(e)

    def run(self):
        try:
            nv, h, mx = self.nv, self.h, self.mx
            if nv is None:
                raise RuntimeError(getattr(self, "error", "NVML unavailable"))
            while not self.stop_flag:
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((clk, mx, int(r)))
                time.sleep(0.001)
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


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


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
    if args.via_pack:
        _lib.check(lib.swga_ctx_set_count_via_pack(ctx.h, 1))
    words = int(lib.swga_tables_words(KMIN, KMAX))
    merge = dist.table_merge(words, KMIN, KMAX, dev, prefer=args.merge) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def rank_max(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return [float(x) for x in t]

    class Dev(object):
        """one genome (or one rank's shard of it) resident in HBM"""

        def __init__(self, seed, n_total, gc, n_rec, shard, ascii_t=None, rs=None):
            rs = synth.equal_records(n_total, n_rec) if rs is None else rs
            a, b = dist.shard_bounds(n_total, rank, world) if shard else (0, n_total)
            e = min(n_total, b + KMAX - 1)
            self.n, self.n_starts, self.a = e - a, b - a, a
            self.mode_b = bool(np.diff(rs.astype(np.int64))[:10001].max() > 1000000)
            inner = rs[(rs > a) & (rs < e)] - np.uint64(a)
            self.rs = np.concatenate([[0], inner, [e - a]]).astype(np.uint64)
            if ascii_t is None:
                ta, tc, tg = synth.thresholds(gc)
                self.ascii = torch.empty(max(self.n, 16), dtype=torch.uint8, device=dev)
                _lib.check(lib.swga_synth_bases(ctx.h, seed, a, self.n, ta, tc, tg, P(self.ascii), None))
            else:
                self.ascii = ascii_t[a:e]
            self.d_rs = torch.from_numpy(self.rs.view(np.int64).copy()).to(dev)
            self.packed = torch.empty(int(lib.swga_packed_words(self.n)), dtype=torch.int32, device=dev)
            self.brk = torch.empty(int(lib.swga_brk_words(self.n)), dtype=torch.int32, device=dev)
            self.fwd = merge.block() if (shard and merge is not None) else torch.empty(words, dtype=torch.int32, device=dev)
            self.canon = torch.empty(words, dtype=torch.int32, device=dev)

        def accumulate(self, ev=None):
            """zero + pack + mark + swga_count_accumulate: this rank's additive table block"""
            _lib.check(lib.swga_memset(ctx.h, P(self.fwd), 0, words * 4, None))
            if ev is not None:
                ev[0].record()
            # ASCII bases in: the partitioned kernels pack them in registers (no pack pass); small genomes: k_pack + k_count
            n_marks = len(self.rs) - 2
            _lib.check(lib.swga_count_accumulate_ascii(ctx.h, P(self.ascii), self.n, P(self.d_rs[1:]) if n_marks else None, n_marks,
                                                       0 if self.mode_b else 1, self.n_starts, KMIN, KMAX, P(self.packed),
                                                       P(self.brk), P(self.fwd), None))
            if ev is not None:
                ev[1].record()

        def finish(self):
            """finalize + fold of the (merged) block"""
            _lib.check(lib.swga_count_finalize(ctx.h, KMIN, KMAX, P(self.fwd), None))
            _lib.check(lib.swga_fold_canonical(ctx.h, KMIN, KMAX, P(self.fwd), P(self.canon), None))

    def make_step(fg, bg):
        def step(ev=None, mev=None):
            # the background's cross-GPU merge runs while this rank counts the replicated foreground
            bg.accumulate(ev)
            if merge is not None:
                merge.start(bg.fwd, mev)
            fg.accumulate()
            fg.finish()
            if merge is not None:
                merge.finish(bg.fwd, bg.canon, mev)           # leaves the finalized forward block and the canonical block
            else:
                bg.finish()
        return step

    def time_steps(step, n_steps, n_warm):
        for _ in range(n_warm):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        mev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n_steps)] if merge is not None else None
        barrier()
        l0 = int(lib.swga_launch_count())
        e0.record()
        for i in range(n_steps):
            step(kev[i], mev[i] if mev else None)
        e1.record()
        barrier()
        launches = int(lib.swga_launch_count()) - l0
        total_ms = e0.elapsed_time(e1)
        count_ms = sum(a.elapsed_time(b) for a, b in kev) / n_steps
        merge_ms = (sum(a.elapsed_time(b) for a, b in mev) / n_steps) if mev else 0.0
        total_ms, count_ms, merge_ms = rank_max([total_ms, count_ms, merge_ms])
        return total_ms / n_steps, count_ms, merge_ms, launches

    fg_seed, n_fg, fg_gc, fg_nrec = synth.CONFIGS[args.workload + "_fg"]
    bg_seed, n_bg, bg_gc, bg_nrec = synth.CONFIGS[args.workload + "_bg"]
    if args.bg_bases:
        n_bg = int(args.bg_bases)
    fg = Dev(fg_seed, n_fg, fg_gc, fg_nrec, shard=False)
    bg = Dev(bg_seed, n_bg, bg_gc, bg_nrec, shard=True)
    torch.cuda.synchronize()
    step = make_step(fg, bg)

    n_warm = max(3, args.warmup)
    for _ in range(n_warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_per_step, count_ms, merge_ms, launches = time_steps(step, args.steps, 0)
    sampler.stop_flag = True
    value = (n_fg + n_bg) / (ms_per_step * 1e-3) / 1e9
    sampler.join(timeout=2)
    clocks = sampler.summary()

    # ---- the merge kernel itself (events on its own stream, between the two barriers) ----
    merge_kernel_ms = merge_alone_ms = None
    if merge is not None and hasattr(merge, "time_kernel"):
        merge.time_kernel = True
        for _ in range(5):
            step()
        barrier()
        merge_kernel_ms = rank_max([sum(a.elapsed_time(b) for a, b in merge.kernel_events) / max(1, len(merge.kernel_events))])[0]
        # ... and with nothing else on the GPU (in the step the replicated foreground count shares the SMs with it)
        merge.kernel_events = []
        for _ in range(5):
            bg.accumulate()
            torch.cuda.synchronize()
            merge.start(bg.fwd, None)
            merge.finish(bg.fwd, bg.canon, None)
        barrier()
        merge.time_kernel = False
        merge_alone_ms = rank_max([sum(a.elapsed_time(b) for a, b in merge.kernel_events) / max(1, len(merge.kernel_events))])[0]

    # ---- per-kernel times of the count stage: CUDA events recorded inside swga_count_accumulate, on its stream ----
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
    hist_ms, part_ms, bcount_ms = rank_max(acc)
    partitioned = part_ms > 0

    # ---- shared-memory operation peak (the ceiling the count stage runs against) and the L2 red peak ----
    ms = ctypes.c_float()
    _lib.check(lib.swga_bench_atom_shared(ctx.h, 1 << 14, 1 << 33, ctypes.byref(ms)))
    sh_peak = (1 << 33) / (ms.value * 1e-3) / 1e9            # G random 4-byte shared atomics / s, 64 KB table
    red_peak = None
    if rank == 0 and not args.no_extras:
        tab = torch.zeros(1 << 24, dtype=torch.int32, device=dev)
        _lib.check(lib.swga_bench_red_global(ctx.h, P(tab), 1 << 24, 1 << 31, ctypes.byref(ms)))
        red_peak = (1 << 31) / (ms.value * 1e-3) / 1e9
        del tab

    # ---- roofline of the dominant kernel on this rank's background shard ----
    peaks = load_json("MEASURED_PEAKS.json")
    hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (
        6650.0, "fallback of B200_PROFILING.md")
    tr = load_json("profiles", "r02_count_stage_traffic.json") or load_json("profiles", "r01_count_stage_traffic.json")
    if partitioned:
        # k_partition reads the genome once: 1 B per base of ASCII (+ 1/32 B of break mask) -- SURVEY 8(d)'s "ASCII in";
        # with --via-pack 1/4 B per base of packed genome.  The bucket elements it writes are this implementation's
        # own intermediate, counted under `traffic`, not here
        k_bytes = ((0.25 if args.via_pack else 1.0) + 1.0 / 32.0) * bg.n_starts
        k_ms, k_name = part_ms, "k_partition"
        k_traffic = tr.get("k_partition_dram_bytes_per_base")
        smem_ops = 0.5                                        # slot atomics per base: one per element = per two bases
    else:
        k_bytes = (0.25 + 1.0 / 32.0) * bg.n_starts + 4.0 * 4 ** KMAX
        k_ms, k_name = count_ms, "k_count"
        k_traffic, smem_ops = None, 0.0
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    stage_bytes = (0.25 if args.via_pack else 1.0) * bg.n_starts + 4.0 * 4 ** KMAX     # genome read once + top table written once
    stage_ops = 1.0                                            # shared atomics per base, whole stage: slot + histogram bin per element
    roofline = {
        "kernel": k_name, "bound": "smem" if partitioned else "l2_atomic",
        "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "traffic": k_traffic * bg.n_starts if k_traffic else None, "peak_source": peak_src,
        "ms_per_launch": k_ms, "share_of_step": k_ms / ms_per_step,
        "dram_gbs_actual": (k_traffic * bg.n_starts / (k_ms * 1e-3) / 1e9) if k_traffic else None,
        # the ceiling this kernel runs against: random shared-memory atomics (one slot atomic per element in k_partition,
        # one bin atomic per element in k_bucket_count; an element = two bases) against the live microbenchmark
        "smem_atomics_per_base": smem_ops,
        "smem_gops": smem_ops * bg.n_starts / (k_ms * 1e-3) / 1e9 if partitioned else None,
        "smem_peak_gops": sh_peak,
        "smem_frac": (smem_ops * bg.n_starts / (k_ms * 1e-3) / 1e9 / sh_peak) if partitioned else None,
        "count_stage": {
            "kernels": "k_bucket_hist(sample) + k_bucket_layout + k_partition + k_item_prefix + k_bucket_count"
                       if partitioned else "k_count",
            "ms": count_ms, "share_of_step": count_ms / ms_per_step, "ms_hist_layout": hist_ms,
            "ms_partition": part_ms, "ms_bucket_count": bcount_ms, "bytes_alg": stage_bytes,
            "achieved": stage_bytes / (count_ms * 1e-3) / 1e9,
            "frac": stage_bytes / (count_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": tr.get("dram_bytes_per_base", 0) * bg.n_starts or None,
            "updates_per_s_g": bg.n_starts / (count_ms * 1e-3) / 1e9,
            "smem_frac": stage_ops * bg.n_starts / (count_ms * 1e-3) / 1e9 / sh_peak if partitioned else None,
            "red_global_64MB_peak_gops": red_peak},
        "note": "k_partition is bound by the L1/shared-memory data pipe (ncu: l1tex data-pipe wavefronts, DRAM ~20 %), not by "
                "HBM: `frac` is the algorithmic HBM fraction the contract asks for, `smem_frac` the fraction of the ceiling "
                "the kernel actually runs against; see profiles/r02_summary.md"}
    path_bytes = 1.5 * (fg.n + bg.n) + 2 * 2 * 4.0 * words       # 1.5 B/base + 2T per genome
    path = {"bytes_alg": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s"}
    path["frac"] = path["achieved"] / hbm_peak

    # ---- e2e: host buffers through the C-ABI host calls, H2D + D2H inside the timed region ------
    # The result is wanted on ONE host: rank 0 counts the (replicated) foreground and copies both canonical blocks
    # back; the other ranks stream their background shard and take part in the merge.
    e2e = None
    if not args.no_e2e:
        h_bg = torch.empty(bg.n, dtype=torch.uint8, pin_memory=True)
        h_bg.copy_(bg.ascii[:bg.n])
        cs = torch.cuda.ExternalStream(lib.swga_ctx_compute_stream(ctx.h), device=dev)
        bg_rs = np.ascontiguousarray(bg.rs)
        if rank == 0:
            h_fg = torch.empty(fg.n, dtype=torch.uint8, pin_memory=True)
            h_fg.copy_(fg.ascii[:fg.n])
            h_canon_fg = torch.empty(words, dtype=torch.int32, pin_memory=True)
            h_canon_bg = torch.empty(words, dtype=torch.int32, pin_memory=True)
            fg_rs = np.ascontiguousarray(fg.rs)
            # the two genomes are independent calls: the foreground goes through a context of its own (own streams and
            # workspaces) on a second host thread, so that its 89.5 MB of results travel device -> host while the
            # background's bases travel host -> device (PCIe is full duplex; ctypes releases the GIL during the call)
            import threading
            ctx_fg = _lib.Context(local)

        def fg_job(err):
            try:
                _lib.check(lib.swga_count_genome_host(ctx_fg.h, P(h_fg), fg.n, P(fg_rs), len(fg_rs) - 1, int(fg.mode_b),
                                                      KMIN, KMAX, P(h_canon_fg)))
            except Exception as e:                              # surfaced by the step
                err.append(e)

        def e2e_step():
            th, err = None, []
            if rank == 0:
                th = threading.Thread(target=fg_job, args=(err,))
                th.start()
            with torch.cuda.stream(cs):
                _lib.check(lib.swga_memset(ctx.h, P(bg.fwd), 0, words * 4, cs.cuda_stream))
                _lib.check(lib.swga_count_stream_host(ctx.h, P(h_bg), bg.n, P(bg_rs), len(bg_rs) - 1, int(bg.mode_b),
                                                      KMIN, KMAX, bg.n_starts, P(bg.fwd)))
                if merge is not None:
                    merge.start(bg.fwd, None)
                    merge.finish(bg.fwd, bg.canon, None)
                else:
                    _lib.check(lib.swga_count_finalize(ctx.h, KMIN, KMAX, P(bg.fwd), cs.cuda_stream))
                    _lib.check(lib.swga_fold_canonical(ctx.h, KMIN, KMAX, P(bg.fwd), P(bg.canon), cs.cuda_stream))
                if rank == 0:
                    h_canon_bg.copy_(bg.canon, non_blocking=True)
            cs.synchronize()
            if th is not None:
                th.join()
                if err:
                    raise err[0]

        n_e2e = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = rank_max([(time.perf_counter() - t0) / n_e2e])[0]
        t = torch.tensor([float(bg.n + 8 * len(bg_rs))], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.SUM)
        h2d_total = int(t[0]) + int(fg.n + 8 * len(fg.rs))
        if rank == 0:
            e2e = {"value": (n_fg + n_bg) / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d_total,
                   "d2h_bytes_per_step": int(2 * 4 * words), "ms_per_step": dt * 1e3, "steps": n_e2e,
                   "h2d_gbs_aggregate": h2d_total / dt / 1e9,
                   "api": "swga_count_genome_host (foreground, own context and host thread) concurrently with "
                          "swga_count_stream_host (background); pinned host buffers; results to rank 0's host",
                   "tables_equal_device_path": None}
        # parity of the two paths on the full-size workload
        step()
        torch.cuda.synchronize()
        if rank == 0:
            e2e["tables_equal_device_path"] = bool(torch.equal(h_canon_bg.to(dev), bg.canon)) and \
                bool(torch.equal(h_canon_fg.to(dev), fg.canon))
            del h_fg, h_canon_fg, h_canon_bg
            ctx_fg.close()
        del h_bg

    # ---- size-independent parity properties of the full-size tables, and the oracle itself on request ----
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
        props = {"bg_tables_sum_and_marginals_ok": bool(ok),
                 "oracle_equal": "see GPU test test_full_size_c2_background_vs_oracle (every table of the 3.1 Gbp background "
                                 "against the CPU oracle); --verify-full repeats it here"}
        if args.verify_full:
            from oracle import swga_oracle as O                 # the checker
            seq = np.empty(bg.n, dtype=np.uint8)
            for o in range(0, bg.n, 1 << 28):
                seq[o:o + (1 << 28)] = bg.ascii[o:o + (1 << 28)].cpu().numpy()
            t0 = time.perf_counter()
            want = O.count_dense_allk_parallel(seq, synth.equal_records(n_bg, bg_nrec), KMIN, KMAX, mode_b=int(bg.mode_b))
            props["oracle_equal"] = bool(np.array_equal(bg.canon.cpu().numpy().view(np.uint32), want))
            props["oracle_s"] = time.perf_counter() - t0
            del seq, want

    # ---- the count -> primer-row merge of `swga count` on the tables just built (BASELINE config 2: count + filter) ----
    count_merge = None
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
            ms_ = f0.elapsed_time(f1)
            best = ms_ if best is None else min(best, ms_)
        count_merge = {"kernel": "k_filter_primers", "ms": best, "rows": int(n_rows.item()), "min_fg_bind": min_fg,
                       "max_bg_bind": max_bg, "max_dimer_bp": 3, "codes_tested": int(words)}
        del rows

    # ---- BASELINE configs[2]: the same step with the 23 Mbp AT-rich foreground (partitioned path on every rank) ----
    c3 = None
    if not args.no_extras and args.workload == "C2" and not args.bg_bases:
        s3, n3, g3, r3 = synth.CONFIGS["C3_fg"]
        fg3 = Dev(s3, n3, g3, r3, shard=False)
        ms3, cnt3, mrg3, _ = time_steps(make_step(fg3, bg), max(3, min(args.steps, 10)), 2)
        c3 = {"workload": workload_name("C3"), "ms_per_step": ms3, "value": (n3 + n_bg) / (ms3 * 1e-3) / 1e9, "unit": UNIT,
              "bg_count_stage_ms": cnt3, "merge_ms": mrg3}
        del fg3

    # ---- low-entropy input: a genome built to make the sampled bucket layout wrong (N = 1 only) ----
    hostile = None
    if world == 1 and not args.no_extras and not args.bg_bases:
        try:
            hostile = hostile_extra(dev, lib, ctx, Dev, time_steps, fg)
        except Exception as e:
            hostile = {"error": repr(e)}

    del bg.ascii, bg.packed, bg.brk
    torch.cuda.empty_cache()

    scoring = None
    if not args.no_extras and not args.no_scoring:
        try:
            scoring = scoring_extra(args, dev, rank, world, rank_max, barrier, sh_peak, hbm_peak)
        except Exception as e:                      # the counting line must survive a failure of the extra
            import traceback
            traceback.print_exc()
            scoring = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = min(KMAX - KMIN + 1, os.cpu_count() or 1)
        dt, bases, kind, cores = cpu_reference_pass(args.workload, args.cpu_sample_bases or 5_000_000, cores)
        cpu = {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d-base prefix of the background: the reference's dsk binary for k=5..12 as %d concurrent processes "
                         "(swga/commands/count.py:84-91 runs them one after another on one core) + the record parser of "
                         "swga/kmers.py:94-115; %.1f s" % (bases, cores, dt),
               "full_runs": cpu_full_runs()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "fg_bases": n_fg, "bg_bases": n_bg,
                       "kmin": KMIN, "kmax": KMAX, "sharding": "bg contiguous chunks + 11-base halo, fg replicated"
                       if world > 1 else "single GPU", "l2": "inputs (%.2f GB/GPU) larger than L2" % (bg.n / 1e9)},
            "roofline": roofline, "path_roofline": path, "cpu_baseline": cpu, "e2e": e2e,
            "clocks": clocks, "gpu_launches": launches, "numa_node": numa,
            "merge": {"impl": merge.name, "exposed_ms": merge_ms, "kernel_ms": merge_kernel_ms,
                      "kernel_alone_ms": merge_alone_ms} if merge is not None else None,
            "scoring": scoring, "c3": c3, "low_entropy": hostile, "count_merge": count_merge,
            "full_size_properties": props,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def hostile_extra(dev, lib, ctx, Dev, time_steps, fg):
    """No throughput cliff when the sample is wrong: the genome of tests/test_count_gpu.py::
    test_partitioned_kernel_on_hostile_composition scaled to 1.2 Gbp -- composition drifting along the genome
    (gc 0.15 / 0.80 / 0.45 thirds), a 15 Mbp dinucleotide repeat, a 10 Mbp poly-A run, 7.5 Mbp of N (mode B: N -> G),
    and a 14-base-period repeat over 30 % of the genome placed only where the sampler does not look."""
    import numpy as np
    import torch
    from swga_b200 import _lib, synth
    P = _lib.ptr
    n = 1_200_000_000
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    third = n // 3
    for i, (seed, gc) in enumerate(((51, 0.15), (52, 0.80), (53, 0.45))):
        lo = i * third
        m = third if i < 2 else n - 2 * third
        ta, tc, tg = synth.thresholds(gc)
        _lib.check(lib.swga_synth_bases(ctx.h, seed, 0, m, ta, tc, tg, P(a[lo:lo + m]), None))
    ac = torch.tensor([65, 67], dtype=torch.uint8, device=dev).repeat(7_500_000)
    a[100_000_000:115_000_000] = ac
    a[250_000_000:260_000_000] = 65
    a[450_000_000:457_500_000] = 78
    unit = torch.from_numpy(np.frombuffer(b"ACGTTGCAAGGCTT" * 1200, dtype=np.uint8).copy()).to(dev)
    stride = min(64, max(1, ((n + 31) // 32) >> 15))
    for lo in range(800_000_000, 1_160_000_000, 1 << 26):
        hi = min(1_160_000_000, lo + (1 << 26))
        pos = torch.arange(lo, hi, device=dev)
        hidden = (pos % (1024 * stride)) >= 1100
        seg = a[lo:hi]
        seg[hidden] = unit[pos[hidden] % unit.numel()]
    rs = synth.equal_records(n, 10)
    g = Dev(0, n, 0.0, 10, shard=False, ascii_t=a, rs=rs)

    def step(ev=None, mev=None):
        g.accumulate(ev)
        g.finish()
    ms, cnt, _, _ = time_steps(step, 5, 2)
    stats = np.zeros(4, dtype=np.uint64)
    _lib.check(lib.swga_count_partition_stats(ctx.h, stats.ctypes.data_as(_lib.c_u64p)))
    lens = np.diff(rs.astype(np.int64))
    o12 = int(lib.swga_table_offset(KMIN, KMAX))
    ok = int(g.canon[o12:].sum(dtype=torch.int64)) == int(np.maximum(0, lens - KMAX + 1).sum())
    return {"bases": n, "ms_per_step": ms, "value": n / (ms * 1e-3) / 1e9, "unit": UNIT, "count_stage_ms": cnt,
            "partition_exits": {"staging_full": int(stats[0]), "bucket_full": int(stats[1]), "uniform_groups": int(stats[2]),
                                "break_groups": int(stats[3])},
            "top_table_sum_ok": bool(ok),
            "composition": "gc 0.15 / 0.80 / 0.45 thirds, 15 Mbp (AC)n, 10 Mbp poly-A, 7.5 Mbp N, 14-periodic repeat hidden from "
                           "the sampler over 30 % of the genome; 10 records (mode B)"}


def scoring_extra(args, dev, rank, world, rank_max, barrier, sh_peak, hbm_peak):
    """Second half of BASELINE.json's metric: candidate sets scored per second (kernel (d)); device-resident site
    lists and sets, CUDA events, best of 3; sets sharded by set index over the ranks (score.shard_sets), no
    collective.  Workloads on the 23 Mbp AT-rich foreground (configs[3]):
      C4_dense  SURVEY 8(d)'s C4 primers: the 200 most frequent 8..12-mers (~2 x 10^5 unique sites per set)
      typical   200 12-mers of ~400 sites each (the scale of swga's default min_fg_bind = 1e-5 * fg_length)
      thin      200 12-mers of ~1,000 sites each (~7,200 sites per set: the dense kernel's coarse tiles)
    each with the default max_fg_bind_dist = 36000 and with no early reject (every set needs Gini)."""
    import numpy as np
    import torch
    from swga_b200 import _lib, fasta, kmers, locate, score, synth, workloads
    lib, ctx = _lib.load(), _lib.context(dev.index)
    P = _lib.ptr
    seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
    names = ["chr%d" % (i + 1) for i in range(n_rec)]
    g = fasta.Genome(synth.bases(seed, 0, n, gc), synth.equal_records(n, n_rec), names)
    tabs = kmers.count_all(g, 5, 12)
    out = {"unit": "sets/s", "fg_bases": n, "n_gpus": world,
           "timing": "CUDA events, best of 3, device-resident site lists and set array; sets sharded by set index, slowest "
                     "rank counts; e2e = host API (numpy set array + bg_dist_mean in, numpy results out), wall clock"}
    t12 = tabs.table(12)
    order = np.argsort(-t12.astype(np.int64), kind="stable")
    lo = int(np.searchsorted(-t12[order].astype(np.int64), -400))
    lo_thin = int(np.searchsorted(-t12[order].astype(np.int64), -1000))
    c4 = workloads.top_primers(tabs, 5, 12, 200)
    c4_seqs = [kmers.codes_to_kmers([c], k)[0] for k, c, _ in c4]

    # ---- locate: CSR index build + both-strand locate + strand union of the 200 C4 primers ----
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gi = locate.GenomeIndex(g)
    sl_dense = score.SiteLists.from_index(gi, c4_seqs)
    torch.cuda.synchronize()
    t_loc = time.perf_counter() - t0
    loc = {"primers": len(c4_seqs), "k_values": sorted({len(s) for s in c4_seqs}), "seconds": t_loc,
           "primers_per_s": len(c4_seqs) / t_loc, "sites_located": int(sl_dense.counts()[:-1].sum()),
           "includes": "H2D of the 23 Mbp genome, strict pack, one CSR index per k, locate of both strands, strand union, "
                       "tile offsets (wall clock, one pass, cold)"}
    variants = {
        "C4_dense": (c4_seqs, int(args.score_sets_dense), sl_dense),
        "typical_400_sites": (kmers.codes_to_kmers(order[lo:lo + 200], 12), int(args.score_sets), None),
        # between the two: ~7,200 sites per set, above the sparse kernel's 4,096 (thin sets in the dense kernel: coarse tiles)
        "thin_1000_sites": (kmers.codes_to_kmers(order[lo_thin:lo_thin + 200], 12), int(args.score_sets_thin), None),
    }
    recs = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import swga_oracle as O                   # cpu_baseline leg: the restated reference, timed
        recs = [(names[r], bytes(g.seq[int(g.rec_starts[r]):int(g.rec_starts[r + 1])]).decode("ascii")) for r in range(n_rec)]
        t0 = time.perf_counter()
        n_cpu = 24
        for s_ in c4_seqs[:n_cpu]:
            O.binding_sites(s_, recs)
        dt = time.perf_counter() - t0
        loc["cpu_baseline"] = {"value": n_cpu / dt, "unit": "primers/s", "cores": 1, "kind": "port",
                               "sample": "restated locate.binding_sites (swga/locate.py:39-51,76-90) for the first %d of the 200 "
                                         "C4 primers on the 23 Mbp foreground, %.1f s; 200 primers extrapolated: %.0f s"
                                         % (n_cpu, dt, dt / n_cpu * 200)}
    out["locate"] = loc

    for name, (seqs, S, sl) in variants.items():
        if sl is None:
            sl = score.SiteLists.from_index(gi, seqs)
        lo_s, hi_s = score.shard_sets(S, rank, world)
        S_r = hi_s - lo_s
        d_raw = torch.empty((max(S_r, 1), 10), dtype=torch.int32, device=dev)
        _lib.check(lib.swga_synth_sets(ctx.h, 8, lo_s, S_r, 200, P(d_raw), None))
        d_raw = d_raw[:S_r]
        lop = torch.from_numpy(np.ascontiguousarray(sl.list_of_primer, dtype=np.int32)).to(dev)
        d_sets = torch.where(d_raw >= 0, lop[d_raw.clamp(0, 199).long()], d_raw.new_full((), -1))
        bg = torch.full((S_r,), 1000.0, dtype=torch.float64, device=dev)
        # algorithmic bytes per set: 4 (sum of the set's site lists + 2R) + 4 m + 8 + 41
        cnt = torch.from_numpy(sl.counts().astype(np.int64)).to(dev)
        n_bounds = int(cnt[-1])
        m = (d_raw >= 0).sum(1)
        l_raw = torch.where(d_sets >= 0, cnt[d_sets.clamp(0).long()], torch.zeros((), dtype=torch.int64, device=dev)).sum(1) + n_bounds
        bytes_alg = float((4 * l_raw + 4 * m + 49).sum())
        merged = float(l_raw.sum())
        row = {"sets": S, "sets_this_rank": S_r, "mean_raw_sites_per_set": merged / max(S_r, 1)}
        for label, max_fg in (("max_fg_36000", 36000), ("no_reject", 2 ** 31 - 1)):
            best = None
            for _ in range(3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = score.score_sets(sl, None, bg, max_fg, d_sets=d_sets, return_device=True)
                e1.record()
                torch.cuda.synchronize()
                ms = rank_max([e0.elapsed_time(e1)])[0]
                best = ms if best is None else min(best, ms)
            st = res.status
            L = res.n_sites
            r = {"sets_per_s": S / best * 1e3, "ms": best,
                 "mean_unique_sites_per_set": float(L.to(torch.float64).mean()) if S_r else 0.0,
                 "passed_frac": float((st & 1).to(torch.float64).mean()) if S_r else 0.0,
                 "merged_sites_per_s": merged / best * 1e3,
                 "roofline": {"kernel": "k_score_sort" if name == "typical_400_sites" else "k_score_dense", "bound": "smem",
                              "bytes_alg_per_set": bytes_alg / max(S_r, 1),
                              "achieved": bytes_alg / (best * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": bytes_alg / (best * 1e-3) / 1e9 / hbm_peak,
                              "smem_frac": merged / (best * 1e-3) / 1e9 / sh_peak,
                              "note": "site lists are L2-resident; smem_frac = merged sites/s against the shared-memory "
                                      "operation peak (one operation per merged site is the floor)" + (
                                          "; ncu (profiles/r02_k_score_dense_phases.txt): 351 warp-instructions per 2^13-position "
                                          "tile (round 1: 524), issue slots 75 % and the shared-memory data pipe 79 % busy"
                                          if name == "C4_dense" else
                                          "; ncu (profiles/r02_k_score_sort_ncu_raw.csv): routed front end 10.5 k warp-instructions "
                                          "per set (merge of round 1: 16.8 k), issue slots 54 % busy at 25 % occupancy"
                                          if name == "typical_400_sites" else
                                          "; thin sets: coarse tiles of 8 fine ones, 142 k warp-instructions per set (per-tile cost "
                                          "is fixed; round 1 walked 2,808 fine tiles: 1.2 M sets/s)")}}
            if world == 1:
                # end to end through the host API: set array and bg_dist_mean from host memory, results back to numpy
                h_sets = d_raw.cpu().numpy()
                h_bg = np.full(S_r, 1000.0)
                e2e_s = None
                for _ in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    hres = score.score_sets(sl, h_sets, h_bg, max_fg)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    e2e_s = dt if e2e_s is None else min(e2e_s, dt)
                n_pass = int(hres.passed.sum())
                d2h = S_r * 5 + n_pass * 44 if (S_r > (1 << 17) and 2 * n_pass <= S_r) else S_r * 41
                r["e2e"] = {"value": S_r / e2e_s, "unit": "sets/s", "h2d_bytes_per_step": int(h_sets.nbytes + h_bg.nbytes),
                            "d2h_bytes_per_step": int(d2h)}
                del h_sets, hres
            row[label] = r
        # CPU baseline: the restated score_set on the first sets of the workload, bounded to ~10 s
        if recs is not None:
            from oracle import swga_oracle as O
            ends = O.chromosome_ends(recs)
            locs = [gi.binding_sites(s_) for s_ in seqs]        # Primer.locations as `swga filter` stored them
            sets_h, _ = workloads.c4_sets(4000)
            for label, max_fg in (("max_fg_36000", 36000), ("no_reject", 2 ** 31 - 1)):
                t0 = time.perf_counter()
                done = 0
                for row_ in sets_h:
                    O.score_set([locs[i] for i in row_ if i >= 0], max_fg, 1000.0, ends)
                    done += 1
                    if time.perf_counter() - t0 > 6.0:
                        break
                dt = time.perf_counter() - t0
                row[label]["cpu_baseline"] = {
                    "value": done / dt, "unit": "sets/s", "cores": 1, "kind": "port",
                    "sample": "restated score.score_set + stats (swga/score.py:73-98, stats.py:10-54, Python-2 arithmetic) on the "
                              "first %d sets, %.1f s; %d sets extrapolated: %.0f s" % (done, dt, S, S / (done / dt))}
        out[name] = row
        del d_sets, d_raw, bg
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
