#!/usr/bin/env python3
"""bench.py -- Gbp/s scanned on BASELINE.json's headline config (cfg3: synthetic 3.1 Gbp, 24 chromosomes,
100 000 planted STS, -W 11 -N 1 -X 1 -M 50).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm (CUDA, one process per GPU)
    python bench.py --impl reference [...]                         # the reference's CPU algorithm (C port of it)

One "step" = one pass of the hot path (scan + verify + emit + sort + read-back of the hit count) over the rank's
resident genome shard.  `value` is whole-job throughput with the packed genome and the table resident in HBM, steps
queued four deep (the host reads step k's count while the next ones run; `config.ms_per_step_host_synced` is the same
step with a host synchronisation after every one); `e2e` repeats the step from pinned HOST bytes (host-side nibble
packing / H2D + plane building + scan + sort + D2H of the hits inside the timed region); `e2e_file` starts from a
FASTA file in the page cache and ends with the output text on disk.
N > 1: strong scaling by default -- ONE 3.1 Gbp genome cut into bp-balanced ranges (+ halos), one per rank, through
the engine's (rank, world) sharding (BASELINE.json config 3); there is no collective on the scan path, NCCL only
carries the timing barrier / max-over-ranks.  `--scaling weak` gives every rank its own genome copy instead.

The CPU baseline is oracle/merpcr_oracle.c (a C restatement of the reference's algorithm; the Python reference
itself cannot travel to the GPU box), timed on a bounded sample with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import synth  # noqa: E402

SEED = 1003
PARAMS = dict(wordsize=11, margin=50, mismatches=1, three_prime_match=1, iupac_mode=0)
METRIC = "Gbp/s scanned (3.1 Gbp genome, 100k STS, N=1)"
ALGO_BYTES_PER_BP = 0.75   # plane2 0.25 + plane4 0.5, each read once
ALGO_BYTES_PER_HIT = 16.0
IN_FLIGHT = 4              # steps queued ahead of the host's read-back of their hit counts


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong",
                    help="strong (default): ONE 3.1 Gbp genome sharded over the GPUs -- bp-balanced ranges with halos, "
                         "BASELINE.json config 3 and the production multi-GPU mode; weak: one genome copy per GPU")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; recorded in config)")
    ap.add_argument("--as-shard", default="", metavar="R/W",
                    help="tuning only (recorded in config): on ONE GPU, run rank R's share of a W-way sharded genome -- "
                         "the per-GPU step of an N = W run without W GPUs")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true", help="skip the FASTA file -> output text leg")
    return ap.parse_args()


def workload(scale: float):
    lengths = [max(2000, int(L * scale)) for L in synth.GRCH38_LENGTHS]
    n_sts = max(50, int(round(100000 * scale)))
    sts = synth.make_sts_set(SEED + 1, n_sts, 18, 25, 100, 1000)
    return lengths, n_sts, sts


def contig_seed(copy: int, ci: int) -> int:
    return SEED * 1000003 + copy * 1009 + ci


def plan_writes(lengths, sts, copy):
    expected, writes = synth.plant_amplicons(SEED + 2 + 7919 * copy, list(lengths), sts, PARAMS["margin"],
                                             sub_mode="cfg3")
    return expected, writes


def apply_writes_numpy(arr, ci, writes, lo=0):
    for c, off, b in writes:
        if c != ci:
            continue
        a, e = off - lo, off - lo + len(b)
        if e <= 0 or a >= len(arr):
            continue
        s = max(a, 0)
        arr[s: min(e, len(arr))] = b[s - a: min(e, len(arr)) - a]


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None
        self.offset = 0
        self.end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark(self):
        """Samples after this point are the ones taken under load (the timed region starts now)."""
        try:
            self.offset = os.path.getsize(self.path)
        except OSError:
            self.offset = 0

    def mark_end(self):
        try:
            self.end = os.path.getsize(self.path)
        except OSError:
            self.end = None

    def wait_first_sample(self, timeout: float = 3.0):
        """nvidia-smi needs a moment to start; the timed region (tens of ms) must not begin before it samples."""
        t0 = time.time()
        while self.proc and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.01)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            text = open(self.path).read()
            lines = text[self.offset: self.end].splitlines()
            if not any(len(x.split(",")) >= 6 for x in lines):      # region shorter than a sampling period
                lines = text.splitlines()
            for line in lines:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[2:6]):
                    if v == "Active":
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference algorithm)
# ---------------------------------------------------------------------------------------------------------
def cpu_sample_search(sts_text: bytes, sample: np.ndarray, threads: int, repeats: int = 1):
    """Time the oracle's search of one in-memory sequence with `threads` host threads. Returns (s/step, hits)."""
    from oracle.oracle import Oracle
    o = Oracle(**PARAMS)
    assert o.load_sts_text(sts_text)
    best, hits = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        hits = o.search_count_buffer(sample.ctypes.data, int(sample.size), threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, hits, o


def cpu_genome_pass(orc, samples, threads: int):
    """One pass of the reference's search loop (engine.py:373-423) over a list of sequences: record after record,
    each cut into `threads` overlapping chunks searched concurrently (the C restatement of -T).  Returns (s, hits)."""
    t0 = time.perf_counter()
    hits = 0
    for a in samples:
        hits += orc.search_count_buffer(a.ctypes.data, int(a.size), threads)
    return time.perf_counter() - t0, hits


def sample_fraction(rate_bp_s: float, seconds: float, total_bp: int) -> float:
    """Share of every contig (its first f * L bases) one CPU step covers within `seconds`."""
    return float(min(1.0, max(1e-3, rate_bp_s * seconds / total_bp)))


def host_genome(lengths, sts, frac: float):
    """The first frac * L bases of every contig of copy 0 -- the same bytes the GPU arm scans -- built on the host
    (numpy, contigs generated concurrently)."""
    from concurrent.futures import ThreadPoolExecutor
    _, writes = plan_writes(lengths, sts, 0)
    by = {}
    for ci, off, b in writes:
        by.setdefault(ci, []).append((ci, off, b))

    def one(ci):
        n = max(2000, int(lengths[ci] * frac)) if frac < 1.0 else lengths[ci]
        n = min(n, lengths[ci])
        arr = synth.dna_chunked(contig_seed(0, ci), n)
        apply_writes_numpy(arr, ci, by.get(ci, ()))
        return arr

    with ThreadPoolExecutor(max_workers=min(len(lengths), os.cpu_count() or 1)) as ex:
        return list(ex.map(one, range(len(lengths))))


def sample_text(frac: float, samples, cores: int) -> str:
    bp = int(sum(a.size for a in samples))
    what = "the whole genome" if frac >= 1.0 else f"the first {frac:.4f} of every contig"
    return (f"{what} ({len(samples)} contigs, {bp} bp) x all STS per step, record after record, each cut into "
            f"{cores} overlapping chunks on {cores} threads (oracle/merpcr_oracle.c restating engine.py:381-423)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lengths, n_sts, sts = workload(args.scale)
    sts_text = synth.sts_lines(sts)
    cores = os.cpu_count() or 1
    total = int(sum(lengths))
    calib = synth.dna_chunked(contig_seed(0, 0), min(lengths[0], 4_000_000))   # rate calibration only
    dt, _, orc = cpu_sample_search(sts_text, calib, cores)
    rate = calib.size / dt
    budget = 150.0 / max(1, args.steps + args.warmup)
    frac = sample_fraction(rate, min(budget, args.cpu_seconds), total)
    samples = host_genome(lengths, sts, frac)
    for _ in range(args.warmup):
        cpu_genome_pass(orc, samples, cores)
    total_s, hits = 0.0, 0
    for _ in range(args.steps):
        dt, hits = cpu_genome_pass(orc, samples, cores)
        total_s += dt
    ms = 1e3 * total_s / args.steps
    bp = int(sum(a.size for a in samples))
    value = bp / (ms * 1e-3) / 1e9
    line = dict(metric=METRIC, value=value, unit="Gbp/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="u8",
                data="synthetic", impl="reference",
                config=dict(workload="cfg3: 24-chromosome synthetic genome (GRCh38 lengths) x 100k planted STS, "
                                     "-W 11 -N 1 -X 1 -M 50",
                            scale=args.scale, n_sts=n_sts, sample_bp=bp, sample_fraction=frac, hits_in_sample=int(hits),
                            note="the Python reference cannot travel to the GPU box (and runs ~0.0013 Gbp/s, "
                                 "profiles/r2_python_reference_rate.json); this is its algorithm restated in C"),
                cpu_baseline=dict(value=value, unit="Gbp/s", cores=cores, kind="port", sample=sample_text(frac, samples, cores)),
                e2e=dict(value=value, unit="Gbp/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local: int):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off: with 8 ranks pulling
    3 GB each per end-to-end step, cross-socket traffic is what the host side can least afford."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return "none exposed (sysfs numa_node = -1: a single-node VM)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None



def run_b200(args):
    import torch
    import torch.distributed as dist
    from merpcr_b200 import MerPCR, _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local)   # pinned host buffers then live next to this GPU's PCIe root
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lengths, n_sts, sts = workload(args.scale)
    sts_text = synth.sts_lines(sts)
    ncont = len(lengths)
    # weak: world-sized layout, copy k of the 24 chromosomes is owned by rank k; strong: one copy, sharded by range
    strong = args.scaling == "strong"
    copy = 0 if strong else rank
    base_ci = 0 if strong else rank * ncont
    all_lengths = lengths if strong else lengths * world
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(sts_text)
        sts_path = f.name
    shard_rw = (rank, world)
    if args.as_shard:
        if world != 1:
            raise SystemExit("--as-shard is a single-GPU tuning mode")
        shard_rw = tuple(int(x) for x in args.as_shard.split("/"))
    eng = MerPCR(**PARAMS, device=local, shard=shard_rw)
    try:
        assert eng.load_sts_file(sts_path)
    finally:
        os.unlink(sts_path)
    layout = eng.make_layout(all_lengths)

    # ---- synthetic genome of this rank's copy, generated and planted directly in HBM
    expected, writes = plan_writes(lengths, sts, copy)
    dev_contigs = []
    by_contig = {}
    for ci, off, b in writes:
        by_contig.setdefault(ci, []).append((off, b))
    for ci, L in enumerate(lengths):
        t = synth.dna_torch(contig_seed(copy, ci), 0, L, dev)
        w = by_contig.get(ci)
        if w:
            idx = np.concatenate([np.arange(off, off + len(b), dtype=np.int64) for off, b in w])
            val = np.concatenate([b for _, b in w])
            t[torch.from_numpy(idx).to(dev)] = torch.from_numpy(val).to(dev)
        dev_contigs.append(t)
    seqs_dev = [None] * len(all_lengths)
    for ci in range(ncont):
        seqs_dev[base_ci + ci] = dev_contigs[ci]
    if strong:   # bases of this rank's range of the padded coordinate
        cg = layout["contigs"]
        my_bp = int(sum(max(0, min(int(c["gstart"]) + int(c["length"]), layout["end"]) - max(int(c["gstart"]), layout["begin"]))
                        for c in cg))
    else:
        my_bp = int(sum(lengths))
    shard = eng.upload(layout, seqs_dev)
    torch.cuda.synchronize()

    # ---- resident (device-timed) region
    lib, ctx = eng._be.lib, eng._ctx
    n_hits = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples every 20 ms from the warm-up to the end of the timed region (GPU under load)
    # Steps are queued IN_FLIGHT deep: the host reads step k's hit count (and, per slot, its kernel times) while the next
    # ones run, so the GPU never waits for the host round trip between steps (a sharded step takes 0.45 ms on eight GPUs;
    # one scheduling hiccup of a host thread is several steps long).  Every step does all of its work -- scan, verify,
    # ordering, count read-back -- and every count is checked; nothing is skipped or cached.
    scan_ms, verify_ms = [], []

    def run_steps(k: int, record: bool) -> int:
        from collections import deque
        pending, n = deque(), 0

        def finish_oldest():
            h = pending.popleft()
            _, n_ = eng.scan_finish(layout, shard, h)
            if record:
                scan_ms.append(float(lib.mpcr_slot_scan_ms(ctx, h[0])))
                verify_ms.append(float(lib.mpcr_slot_verify_ms(ctx, h[0])))
            return n_

        for i in range(k):
            if len(pending) == IN_FLIGHT:
                n = finish_oldest()
            pending.append(eng.scan_device_async(layout, shard, slot=i % IN_FLIGHT))
        while pending:
            n = finish_oldest()
        return n

    run_steps(IN_FLIGHT, False)   # untimed: every pipeline slot allocates its hit buffer and pinned result words here
    n_hits = run_steps(args.warmup, False)
    if rank == 0:
        sampler.wait_first_sample()
        run_steps(2, False)      # keep the GPU busy while the sampler gets going (untimed)
    barrier()
    if rank == 0:
        sampler.mark()
    launches0 = eng.gpu_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()            # a collector pause on the host is longer than a sharded step
    ev0.record()
    n_hits = run_steps(args.steps, True)
    ev1.record()
    torch.cuda.synchronize()
    gc.enable()
    if rank == 0:
        sampler.mark_end()
    t_ms = ev0.elapsed_time(ev1)
    launches = eng.gpu_launches - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_rank = None
    if world > 1:    # every rank's own numbers (the headline is the slowest rank's)
        mine = dict(rank=rank, ms_per_step=round(t_ms / args.steps, 4), scan_kernel_ms=round(float(np.mean(scan_ms)), 4),
                    verify_kernel_ms=round(float(np.mean(verify_ms)), 4), hits=int(n_hits))
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    t_ms = max_over_ranks(t_ms)
    ms_per_step = t_ms / args.steps
    total_bp = sum_over_ranks(float(my_bp))
    total_hits = sum_over_ranks(float(n_hits))
    value = total_bp / (ms_per_step * 1e-3) / 1e9
    kern_ms = float(np.mean(scan_ms))

    # the same step with a host synchronisation after EVERY step (one call, one round trip): what a caller sees who
    # needs each result before issuing the next step
    sync_steps = max(3, min(args.steps, 10))
    eng.scan_device(layout, shard)      # untimed: this path's own hit buffer grows to size here
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(sync_steps):
        eng.scan_device(layout, shard)
    torch.cuda.synchronize()
    ms_synced = max_over_ranks((time.perf_counter() - t0) / sync_steps * 1e3)

    # planted truth + ordering sanity on the resident result (not timed)
    hits_t, n = eng.scan_device(layout, shard)
    hits = hits_t[: n * _capi.HIT_DTYPE.itemsize].cpu().numpy().view(_capi.HIT_DTYPE)
    all_hits = hits
    if strong and world > 1:   # the planted truth is checked on the merged hit lists of all shards
        gathered = [None] * world
        dist.all_gather_object(gathered, np.array(hits))
        all_hits = np.concatenate(gathered)
    found = set(zip((all_hits["contig"] - base_ci).tolist(), all_hits["pos1"].tolist(), all_hits["pos2"].tolist()))
    planted_ok = all((c, a, b) in found for c, a, b, _, _ in expected) and (not strong or len(all_hits) == len(found))
    key = np.stack([hits["contig"], hits["pos1"]], axis=1).astype(np.int64)
    sorted_ok = bool(np.all((key[1:, 0] > key[:-1, 0]) | ((key[1:, 0] == key[:-1, 0]) & (key[1:, 1] >= key[:-1, 1]))))

    # ---- end-to-end from pinned host bytes
    e2e = None
    host_contigs = None
    if not args.no_e2e:
        host_contigs = [torch.empty(L, dtype=torch.uint8).pin_memory() for L in lengths]
        for h, d in zip(host_contigs, dev_contigs):
            h.copy_(d)
        torch.cuda.synchronize()
        seqs_host = [None] * len(all_lengths)
        for ci in range(ncont):
            seqs_host[base_ci + ci] = host_contigs[ci]
        del dev_contigs, seqs_dev
        e2e_steps = max(2, min(args.steps, 5))
        sh2 = None
        d2h = 0
        for it in range(1 + e2e_steps):            # 1 warm-up
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            sh2, hits_t, n_e2e = eng.upload_and_scan(layout, seqs_host, shard=sh2)   # copy | pack | scan pipelined
            out = eng._hits_to_host(hits_t, n_e2e, copy=False)                        # sorted hits back on the host (pinned)
            d2h = out.nbytes + 8
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        e2e = dict(value=total_bp / dt / 1e9, unit="Gbp/s", h2d_bytes_per_step=int(sum_over_ranks(float(eng.last_h2d_bytes))),
                   d2h_bytes_per_step=int(sum_over_ranks(float(d2h))), ms_per_step=dt * 1e3, steps=e2e_steps,
                   host_pack_ms=round(1e3 * float(eng.last_timing.get("host_pack_s", 0.0)), 3),
                   host_pack_threads=int(eng.last_timing.get("host_pack_threads", 0)),
                   wire="4-bit packed on the host cores (0.5 B/bp)" if eng.host_pack else "ASCII (1 B/bp)",
                   hits=int(sum_over_ranks(float(len(out)))))
        assert len(out) == n_hits, "e2e and resident hit counts differ"

    # ---- file -> text: the reference's whole user-visible path (io/fasta.py:19-71 + engine.py:365-451) on a FASTA file
    # in the page cache: load_fasta_file (parallel chunked read | H2D | text ingest on the device) -> search -> output
    # file.  N > 1 (sharded mode): every rank ingests only its own byte range of the file (rank-local ingest), scans the
    # positions that range starts, and rank 0 gathers the hits and writes the text.
    e2e_file = None
    if not args.no_e2e and not args.no_e2e_file and not args.as_shard and (strong or world == 1):
        import shutil
        paths = [None, None]
        if rank == 0:
            tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2.5 * sum(lengths) \
                else tempfile.gettempdir()
            paths = [os.path.join(tmpdir, f"merpcr_b200_bench_{os.getpid()}.fa"),
                     os.path.join(tmpdir, f"merpcr_b200_bench_{os.getpid()}.out")]
        if world > 1:
            dist.broadcast_object_list(paths, src=0)
        fa_path, out_path = paths
        try:
            if rank == 0:
                with open(fa_path, "wb") as f:
                    for ci, name in enumerate(synth.GRCH38_NAMES):
                        arr = host_contigs[ci].numpy()
                        f.write(b">%s synthetic cfg3 contig\n" % name.encode())
                        body = arr[: len(arr) // 60 * 60].reshape(-1, 60)
                        for s0 in range(0, body.shape[0], 1 << 18):
                            blk = body[s0: s0 + (1 << 18)]
                            buf = np.empty((blk.shape[0], 61), dtype=np.uint8)
                            buf[:, :60] = blk
                            buf[:, 60] = 10
                            f.write(buf.tobytes())
                        tail = arr[len(arr) // 60 * 60:]
                        if tail.size:
                            f.write(tail.tobytes() + b"\n")
            barrier()
            file_bytes = os.path.getsize(fa_path)
            reps, runs, rank_local = 3, [], False
            for it in range(reps):
                barrier()
                t0 = time.perf_counter()
                recs = eng.load_fasta_file(fa_path)
                t1 = time.perf_counter()
                n_file = eng.search(recs, out_path)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                rank_local = hasattr(recs, "owned_range")
                if it:      # the first pass warms up (allocations, pinned staging)
                    runs.append((max_over_ranks(t2 - t0), max_over_ranks(t1 - t0), max_over_ranks(t2 - t1)))
                del recs
            if rank == 0:
                text = open(out_path, "rb").read()
                # text-exact against the oracle for the smallest chromosome (its block of lines in the output file)
                from oracle.oracle import Oracle
                orc2 = Oracle(**PARAMS)
                assert orc2.load_sts_text(sts_text)
                small = int(np.argmin(lengths))
                label = synth.GRCH38_NAMES[small]
                _, want_text = orc2.search_text(label, host_contigs[small].numpy().tobytes(), threads=1)
                got_text = b"".join(ln for ln in text.splitlines(keepends=True) if ln.startswith(label.encode() + b"\t"))
                dt, t_load, t_search = min(runs)
                e2e_file = dict(value=float(sum(lengths)) / dt / 1e9, unit="Gbp/s", ms=dt * 1e3, load_fasta_ms=t_load * 1e3,
                                search_and_write_ms=t_search * 1e3, file_bytes=int(file_bytes), text_bytes=len(text),
                                hits=int(n_file), lines=text.count(b"\n"), file_in=os.path.dirname(fa_path),
                                rank_local_ingest=bool(rank_local),
                                text_exact_vs_oracle=dict(contig=label, ok=got_text == want_text.encode("latin-1"),
                                                          lines=got_text.count(b"\n")),
                                what="FASTA file (page cache) -> load_fasta_file -> search -> output file, STS table "
                                     "resident; max over ranks")
                assert text.count(b"\n") == n_file, "file -> text: line count differs from the hit count"
            barrier()
        finally:
            if rank == 0:
                for pth in (fa_path, out_path):
                    try:
                        os.unlink(pth)
                    except OSError:
                        pass

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) + bit-exact parity against the GPU result
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import fullsize
        cores = os.cpu_count() or 1
        total = int(sum(lengths))
        get = (lambda ci: host_contigs[ci].numpy()) if host_contigs is not None else (lambda ci: dev_contigs[ci].cpu().numpy())
        calib = get(0)[: min(lengths[0], 4_000_000)]
        dt, _, orc = cpu_sample_search(sts_text, calib, cores)
        frac = sample_fraction(calib.size / dt, args.cpu_seconds, total)
        samples = [get(ci)[: (lengths[ci] if frac >= 1.0 else max(2000, int(lengths[ci] * frac)))] for ci in range(ncont)]
        dt, cpu_hits = cpu_genome_pass(orc, samples, cores)
        nb = int(sum(a.size for a in samples))
        # parity: the COMPLETE ordered hit lists of whole chromosomes (every contig of at most 65 Mbp at full scale: chr19
        # .. chr22 and chrY, 278 Mbp) from the single-threaded oracle (== reference -T 1), one chromosome per host thread
        whole = [ci for ci in range(ncont) if lengths[ci] <= 65_000_000 * max(args.scale, 1e-9)] or [ncont - 1]
        jobs = [(ci, 0, lengths[ci], lengths[ci]) for ci in whole]
        t0 = time.perf_counter()
        want = np.concatenate(fullsize.oracle_rows(PARAMS, sts_text, jobs, lambda ci, a, b: get(ci)[a:b]))
        rows = fullsize.gpu_rows(eng, hits[np.isin(hits["contig"], whole)])
        parity_ok = want.shape == rows.shape and bool(np.array_equal(want, rows))
        cpu = dict(value=nb / dt / 1e9, unit="Gbp/s", cores=cores, kind="port", sample=sample_text(frac, samples, cores),
                   seconds=dt, hits=int(cpu_hits), sample_fraction=frac,
                   parity=dict(bit_exact=parity_ok, contigs=[int(c) for c in whole], bp=int(sum(lengths[c] for c in whole)),
                               hits=int(len(want)), seconds=round(time.perf_counter() - t0, 2),
                               what="complete ordered hit list of whole chromosomes, GPU vs single-threaded oracle"))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        algo_bytes = ALGO_BYTES_PER_BP * my_bp + ALGO_BYTES_PER_HIT * n_hits
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        # DRAM bytes per launch of the dominant kernel: from the committed `ncu --set full` capture of the same workload
        # (a number taken under a profiler cannot be measured inside this run).  It is only quoted while that capture
        # still describes the kernel being timed: same scale, whole genome on one GPU, and a kernel duration within 5 %
        # of the one measured here; otherwise null, with the reason.
        traffic, traffic_source = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_ncu.json")))
            dur = float(prof.get("duration_ms", 0.0))
            if abs(args.scale - float(prof.get("scale", 1.0))) > 1e-9 or world != 1 or args.as_shard:
                traffic_source = "none: the committed capture is of the whole genome on one GPU at scale 1"
            elif kern_ms <= 0 or abs(dur - kern_ms) > 0.05 * kern_ms:
                traffic_source = (f"none: committed capture ran {dur:.3f} ms per launch, this run {kern_ms:.3f} ms "
                                  "(the kernel changed; re-capture with scripts/gpu/profile.sh)")
            else:
                traffic = float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])
                traffic_source = ("profiles/scan_kernel_ncu.json (ncu --set full capture of this kernel, "
                                  f"{dur:.3f} ms per launch there vs {kern_ms:.3f} ms here)")
        except Exception as e:  # noqa: BLE001
            traffic_source = f"none: {e}"
        line = dict(
            metric=METRIC, value=value, unit="Gbp/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_per_step, higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None,
            dtype="u8",
            data="synthetic", impl="b200",
            config=dict(workload="cfg3: 24-chromosome synthetic genome (GRCh38 lengths) x 100k planted STS, "
                                 "-W 11 -N 1 -X 1 -M 50; " +
                                 ("ONE genome sharded over the GPUs (bp-balanced ranges + halos)" if strong
                                  else "one genome copy per GPU"),
                        scale=args.scale, bp_per_gpu=my_bp, n_sts=n_sts, hits_per_gpu=int(n_hits),
                        **({"emulated_shard": args.as_shard, "note": "tuning run: one rank's share only, not a bench line"}
                           if args.as_shard else {}),
                        l2="inputs (2.7 GB of planes per GPU) exceed the 126 MB L2; no flush needed",
                        steps_in_flight=IN_FLIGHT, ms_per_step_host_synced=ms_synced,
                        numa_node=numa, host_cpus=len(os.sched_getaffinity(0)),
                        planted_found=planted_ok, sorted=sorted_ok, **({"per_rank": per_rank} if per_rank else {})),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak if peak else None,
                          traffic=traffic, traffic_source=traffic_source, kernel="scan_kernel", kernel_ms=kern_ms,
                          verify_kernel_ms=float(np.mean(verify_ms)),
                          algorithmic_bytes_per_launch=algo_bytes,
                          peak_source="MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650"),
            cpu_baseline=cpu, e2e=e2e, e2e_file=e2e_file, gpu_launches=int(launches), clocks=clocks,
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
