#!/usr/bin/env python
"""bench.py -- headline benchmark of the vocabulary-tree hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port of the reference)

Primary metric: descriptors/s quantised through a k=10, L=6 (1 M-leaf) tree; one step = one pass
of the fused descent kernel over this rank's shard of synthetic SIFT-shaped byte descriptors
(BASELINE configs[2]: 100 M x 128 per GPU, tree replicated, no data-path collective -> weak
scaling).  `value` is device-resident throughput, `e2e` the same metric through the host-buffer
C-ABI call (vt_quantize_host: pinned host rows in, word ids out, copies inside the timed region).
Secondary metric (same JSON line, key "secondary"): queries/s top-10 over a 1 M-image inverted
file (BASELINE configs[3], leaf-words mode), database sharded by image id across ranks.

Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_BRANCH, DEPTH, DIM = 10, 6, 128
METRIC = "descriptors/s quantized (k=10,L=6 tree)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-desc", type=int, default=100_000_000, help="descriptors per GPU")
    ap.add_argument("--e2e-desc", type=int, default=32_000_000, help="descriptors per end-to-end step per GPU")
    ap.add_argument("--db-images", type=int, default=1_000_000, help="database images (whole job)")
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--words-per-image", type=int, default=1000)
    ap.add_argument("--engine", default="mma", choices=["mma", "simt"],
                    help="distance step of the sorted descent: tcgen05 INT8 tensor cores or FP32 SIMT")
    ap.add_argument("--all-nodes-images", type=int, default=1_000_000)
    ap.add_argument("--all-nodes-queries", type=int, default=2048)
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-all-nodes", action="store_true")
    ap.add_argument("--no-tfidf", action="store_true")
    ap.add_argument("--tfidf-queries", type=int, default=4096)
    ap.add_argument("--no-build", action="store_true")
    ap.add_argument("--build-desc", type=int, default=10_000_000, help="descriptors for the tree-build leg (whole job)")
    ap.add_argument("--build-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--profile-range", default="", choices=["", "quantize", "retrieval", "retrieval_all_nodes"],
                    help="cudaProfilerStart/Stop around this timed region (for `ncu --profile-from-start off` launch lists)")
    return ap.parse_args()


class ProfileRange:
    """cudaProfilerStart/Stop around a timed region when --profile-range names it: `ncu --profile-from-start off` then lists
    exactly the launches bench.py times (a no-op without a profiler attached)."""

    def __init__(self, torch, on: bool):
        self.torch, self.on = torch, on

    def __enter__(self):
        if self.on:
            self.torch.cuda.synchronize()
            self.torch.cuda.profiler.start()

    def __exit__(self, *a):
        if self.on:
            self.torch.cuda.synchronize()
            self.torch.cuda.profiler.stop()


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------- synthetic data
def full_tree_tables(vt, k, depth):
    nh = vt.tree.heap_size(k, depth)
    exists = np.ones(nh, np.uint8)
    internal = np.zeros(nh, np.uint8)
    internal[:vt.tree.level_offset(k, depth)] = 1
    ref_id, n_ref = vt.tree.reference_ids(exists, internal, k, depth)
    return exists, internal, ref_id, n_ref


def hashed_words(torch, img_ids, words_per_image, n_leaf, leaf_ref_base, device):
    """Deterministic pseudo-random leaf words for any image id (so every rank can regenerate any
    image): a 64-bit mix of (image, slot) reduced modulo the number of leaves."""
    slot = torch.arange(words_per_image, device=device, dtype=torch.int64)[None, :]
    x = img_ids.to(torch.int64)[:, None] * 1_000_003 + slot + 0x9E3779B9
    x = (x ^ (x >> 15)) * 0x2C1B3C6D
    x = (x ^ (x >> 12)) * 0x297A2D39
    x = x ^ (x >> 15)
    leaf = (x & 0x7FFFFFFFFFFF) % n_leaf
    return leaf_ref_base[leaf].to(torch.int32).reshape(-1)


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (multi-socket hosts: the pinned
    staging buffers of the end-to-end path then live on the GPU's own NUMA node).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, bits in enumerate(words) for b in range(64) if (bits >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path (oracle port, all host threads)
    timed on a bounded sample of the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)     # torchrun pins it to 1; the arm uses every core
    import torch
    import vtree_b200 as vt
    from oracle import vt_oracle as vo
    torch.set_num_threads(os.cpu_count() or 1)
    cent = vt.synth.hierarchical_tree(K_BRANCH, DEPTH, DIM, seed=7, device="cpu")
    exists, internal, ref_id, n_ref = full_tree_tables(vt, K_BRANCH, DEPTH)
    tree = dict(k=K_BRANCH, depth=DEPTH, D=DIM, DP=DIM, centroids=cent.numpy(), exists=exists, internal=internal,
                ref_id=ref_id, n_ref=n_ref)
    cores = os.cpu_count() or 1
    probe = vt.synth.descriptors_from_tree(cent, K_BRANCH, DEPTH, 20_000, seed=11).numpy()
    t0 = time.perf_counter()
    vo.descend(tree, probe)
    rate = len(probe) / (time.perf_counter() - t0)
    sample = int(max(20_000, min(5_000_000, rate * 2.0)))           # ~2 s of CPU work per step
    x = vt.synth.descriptors_from_tree(cent, K_BRANCH, DEPTH, sample, seed=12).numpy()
    for _ in range(max(1, min(args.warmup, 2))):
        vo.descend(tree, x[:max(1, sample // 4)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vo.descend(tree, x)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "descriptors/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "quantize SIFT-128 byte descriptors, k=10 L=6 (1M-leaf) tree, CPU sample of %d per step" % sample},
            "cpu_baseline": {"value": value, "unit": "descriptors/s", "cores": cores, "kind": "port",
                             "sample": "%d descriptors x %d steps, OpenMP over all host threads" % (sample, args.steps)},
            "e2e": {"value": value, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    del torch


# ----------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import vtree_b200 as vt
    from vtree_b200 import engine, _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local)      # pinned host buffers are first-touched next to the GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created; stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    nat.require_device()
    lib = nat.load()
    comm = vt.dist.Comm() if world > 1 else None

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

    # ---- tree (replicated) and this rank's descriptor shard -------------------------------------
    cent = vt.synth.hierarchical_tree(K_BRANCH, DEPTH, DIM, seed=7, device=dev)
    exists, internal, ref_id, n_ref = full_tree_tables(vt, K_BRANCH, DEPTH)
    tree_dev = dict(centroids=cent, internal=torch.from_numpy(internal).to(dev), ref_id=torch.from_numpy(ref_id).to(dev))
    n = args.n_desc
    desc = vt.synth.descriptors_from_tree(cent, K_BRANCH, DEPTH, n, seed=100 + rank)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)

    # ---- FP32 pipe peak on this box ---------------------------------------------------------------
    scratch = torch.empty(148 * 8 * 256 * 2, dtype=torch.float32, device=dev)
    peak_tf, peak_ms = C.c_double(0), C.c_double(0)
    nat.check(lib.vt_measure_fp32_peak(nat.ptr(scratch), 4096, 5, C.byref(peak_tf), C.byref(peak_ms), nat.stream_ptr()),
              "vt_measure_fp32_peak")

    # ---- device-resident throughput ---------------------------------------------------------------
    work = torch.empty(int(max(lib.vt_quantize_sorted_work_bytes(n), lib.vt_quantize_mma_work_bytes(n, K_BRANCH, DEPTH))),
                       dtype=torch.uint8, device=dev)
    method = "mma" if args.engine == "mma" else "sorted"

    def one_step():
        # level-synchronous sorted descent: 6 level kernels + the radix passes between them
        return engine.quantize(tree_dev, K_BRANCH, DEPTH, desc, stats=stats, method=method, work=work)

    for _ in range(args.warmup):
        words = one_step()
    stats.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with ClockSampler(local) as clocks, ProfileRange(torch, args.profile_range == "quantize"):
        t0 = time.perf_counter()
        for a, b in ev:
            a.record()
            words = one_step()                                  # 12.8 GB input >> 126 MB L2
            b.record()
        barrier()
        wall = time.perf_counter() - t0
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    step_ms = max_over_ranks(max(wall / args.steps * 1e3, kern_ms))
    value = n * world / (step_ms * 1e-3)
    rechecks = int(stats[0].item())

    # per-stage CUDA events (same call, measurement variant): which share of a step is the level kernel
    def stage_profile(fn, reps=3):
        lvl = (C.c_float * DEPTH)()
        srt = (C.c_float * DEPTH)()
        lvl_acc, srt_acc = np.zeros(DEPTH), np.zeros(DEPTH)
        for _ in range(reps):
            nat.check(fn(nat.ptr(desc), 0, n, DIM, nat.ptr(cent), nat.ptr(tree_dev["internal"]), nat.ptr(tree_dev["ref_id"]),
                         K_BRANCH, DEPTH, None, nat.ptr(words), None, nat.ptr(work), work.numel(), lvl, srt,
                         nat.stream_ptr()), "profile variant")
            lvl_acc += np.array(list(lvl))
            srt_acc += np.array(list(srt))
        return lvl_acc / reps, srt_acc / reps

    level_ms, sort_ms = stage_profile(lib.vt_quantize_mma_profile if method == "mma" else lib.vt_quantize_sorted_profile)
    # the two-levels-per-launch variant of the tcgen05 descent (vt_quantize_pair.cu; off by default): same ids, 3 DRAM
    # reads of every row instead of 6
    pair = None
    if method == "mma":
        prev = lib.vt_quantize_mma_pair_min(1 << 20)
        w2 = engine.quantize(tree_dev, K_BRANCH, DEPTH, desc, method="mma", work=work)
        pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pa.record()
        for _ in range(3):
            w2 = engine.quantize(tree_dev, K_BRANCH, DEPTH, desc, method="mma", work=work)
        pb.record()
        torch.cuda.synchronize()
        p_lvl, p_srt = stage_profile(lib.vt_quantize_mma_profile, reps=2)
        lib.vt_quantize_mma_pair_min(prev if prev < (1 << 62) else -1)
        pair = {"kernel": "level_pair_kernel<10> (vt_quantize_pair.cu): two tree levels per launch, TMA tile::gather4, the second "
                          "level re-gathers from L2", "ms_per_step": pa.elapsed_time(pb) / 3, "value": n / (pa.elapsed_time(pb) / 3 * 1e-3),
                "unit": "descriptors/s", "level_kernel_ms": [round(float(v), 3) for v in p_lvl if v > 0],
                "resort_ms": [round(float(v), 3) for v in p_srt if v > 0], "same_words": bool(torch.equal(w2, words)),
                "dram_row_reads_per_descent": (DEPTH + 1) // 2, "default": False}
        del w2
    other_ms, _ = stage_profile(lib.vt_quantize_sorted_profile if method == "mma" else lib.vt_quantize_mma_profile, reps=1)
    # kernels per step: one level launch and one regrouping pass (hist, scan, scatter) per level transition; the tcgen05
    # engine adds, from 2^22 rows on, the deferred output of the last level (one more pass + the id write); the byte
    # planes are prepared once per tree, outside the steps.
    launches_per_step = DEPTH + 3 * (DEPTH - 1) + (4 if method == "mma" and n >= (1 << 22) else 0)

    # float32 input (the primary feed is fp32 [N, 128]; bytes are its lossless compressed form): the staging
    # pass vt_stage_f32_to_u8 (512 B read + 128 B written per descriptor) is part of the step
    n_32 = min(n, 16_000_000)
    x32 = desc[:n_32].to(torch.float32)
    for _ in range(2):
        d8, _ = engine.stage(x32, dev)
        w32 = engine.quantize(tree_dev, K_BRANCH, DEPTH, d8, method=method, work=work)
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fa.record()
    for _ in range(3):
        d8, _ = engine.stage(x32, dev)
        w32 = engine.quantize(tree_dev, K_BRANCH, DEPTH, d8, method=method, work=work)
    fb.record()
    torch.cuda.synchronize()
    fp32_in = {"value": 3 * n_32 / (fa.elapsed_time(fb) * 1e-3), "unit": "descriptors/s", "rows": n_32,
               "bytes_per_descriptor": 4 * DIM + 4, "same_words_as_byte_input": bool(torch.equal(w32, words[:n_32])),
               "what": "float32 rows resident in HBM -> vt_stage_f32_to_u8 -> descent, staging inside the timed region"}
    del x32, d8, w32

    # the single-launch fused descent on a slice, for comparison
    n_f = min(n, 8_000_000)
    for _ in range(2):
        engine.quantize(tree_dev, K_BRANCH, DEPTH, desc[:n_f], method="fused")
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fa.record()
    engine.quantize(tree_dev, K_BRANCH, DEPTH, desc[:n_f], method="fused")
    fb.record()
    torch.cuda.synchronize()
    fused_rate = n_f / (fa.elapsed_time(fb) * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    n_e2e = min(n, args.e2e_desc)
    h_desc = torch.empty((n_e2e, DIM), dtype=torch.uint8).pin_memory()
    h_desc.copy_(desc[:n_e2e])
    h_word = torch.empty(n_e2e, dtype=torch.int32).pin_memory()
    e2e_steps = max(3, min(args.steps, 5))
    for _ in range(2):
        engine.quantize_host(tree_dev, K_BRANCH, DEPTH, h_desc, h_word)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.quantize_host(tree_dev, K_BRANCH, DEPTH, h_desc, h_word)               # blocks until h_word is filled
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
    e2e_value = n_e2e * world / (e2e_ms * 1e-3)
    e2e_ok = bool(torch.equal(h_word, words[:n_e2e].cpu()))
    # the ceiling of that path: a plain pinned host->device copy of the same bytes, issued on every rank at the same time
    # (the ranks of one box share the host's PCIe root ports and memory controllers)
    d_tmp = torch.empty((n_e2e, DIM), dtype=torch.uint8, device=dev)
    d_tmp.copy_(h_desc, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        d_tmp.copy_(h_desc, non_blocking=True)
    barrier()
    copy_ms = max_over_ranks((time.perf_counter() - t0) / 3 * 1e3)
    h2d_ceiling = {"aggregate_GBps": n_e2e * DIM * world / (copy_ms * 1e-3) / 1e9, "per_gpu_GBps": n_e2e * DIM / (copy_ms * 1e-3) / 1e9,
                   "what": "concurrent plain pinned cudaMemcpyAsync of the e2e input on every rank, max over ranks"}
    del d_tmp

    # ---- secondary metric: queries/s top-10 over the sharded inverted file ---------------------
    secondary = None
    if not args.no_retrieval:
        secondary = bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, cent, exists, internal, barrier,
                                    max_over_ranks)
        # the reference's own embedding semantics (every node on the path counts, vocabulary.py:92-100)
        # BASELINE configs[3] names TF-IDF scoring: idf = ln(N / N_i) weights, L2 normalisation (binary64 kernels, vt_weight.cu)
        if not args.no_tfidf:
            secondary["tfidf_mode"] = bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, cent, exists,
                                                      internal, barrier, max_over_ranks, weighting="tfidf",
                                                      queries=args.tfidf_queries)
        if not args.no_all_nodes:
            secondary["all_nodes_mode"] = bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, cent, exists,
                                                          internal, barrier, max_over_ranks, all_nodes=True,
                                                          db_images=args.all_nodes_images, queries=args.all_nodes_queries)

    # ---- tertiary: BASELINE configs[1], tree build k=10 L=6 on 10 M descriptors ----------------
    build = None
    if not args.no_build:
        build = bench_build(args, torch, vt, engine, dev, rank, world, comm, desc, barrier, max_over_ranks)

    # ---- CPU baseline (rank 0, single-GPU runs only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, vt, cent, exists, internal, ref_id, n_ref, desc, words)

    if rank == 0:
        flop_per_desc = 3 * K_BRANCH * DIM * DEPTH                 # sub, mul, add per element per candidate
        alg_bytes = DIM + 4                                        # byte row in, int32 word id out = 132 B per descriptor
        per_launch = n
        n_stages = int((level_ms > 0).sum())
        lvl_total_ms = float(level_ms.sum())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        other_total_ms = float(other_ms.sum())
        # DRAM bytes per descriptor and step of the descent kernels, from the committed ncu launch list of this command
        # (profiles/r02_descent_traffic.json, written by tools/summarise_launches.py from the capture); null when no
        # capture of the current kernels is committed
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_descent_traffic.json")))
            traffic, traffic_src = float(tj["level_kernel_dram_bytes_per_row"]) * per_launch, tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        simt = {"bound": "fp32", "kernel": "level_kernel<128,R> (vt_quantize_sorted.cu)", "unit": "TFLOP/s",
                "peak": peak_tf.value, "peak_source": "FFMA microbenchmark on this device (vt_measure_fp32_peak)",
                "flop_per_descriptor": flop_per_desc}
        mma = {"bound": "hbm", "kernel": "level_mma_kernel<8,10> (vt_quantize_mma.cu, tcgen05.mma kind::i8, one tree level per launch)",
               "unit": "GB/s", "peak": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
               "bytes_per_descriptor": alg_bytes}
        def fill(d, total_ms):
            if d["bound"] == "fp32":
                d["achieved"] = flop_per_desc * per_launch / (total_ms * 1e-3) / 1e12
            else:
                # ALGORITHMIC bytes of the whole descent (132 B per descriptor: the row once, the id once)
                # over the summed duration of the descent kernels -- re-reads of the row by later levels are traffic,
                # not algorithmic bytes
                d["achieved"] = alg_bytes * per_launch / (total_ms * 1e-3) / 1e9
            d["frac"] = d["achieved"] / d["peak"] if d["peak"] else None
            d["kernel_ms_per_step"] = total_ms
            return d
        main_r, other_r = (mma, simt) if method == "mma" else (simt, mma)
        roofline = fill(main_r, lvl_total_ms)
        roofline.update({
            "launches_per_step": n_stages,
            "level_kernel_ms": [round(float(v), 3) for v in level_ms[:n_stages]],
            "resort_ms": [round(float(v), 3) for v in sort_ms[:n_stages]],
            "kernel_share_of_step": lvl_total_ms / kern_ms,
            "whole_step": {"step_ms_events": kern_ms,
                           "algorithmic_GBps": alg_bytes * per_launch / (kern_ms * 1e-3) / 1e9,
                           "frac_of_hbm_peak": alg_bytes * per_launch / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                           "algorithmic_TFLOPs": flop_per_desc * per_launch / (kern_ms * 1e-3) / 1e12,
                           "hbm_bound_descriptors_per_s": hbm_peak * 1e9 / alg_bytes},
            "traffic": traffic,
            "traffic_source": traffic_src,
            # the same launches against what ONE level has to move (128 B row gather + 8 B key/permutation in + 8 B out,
            # 136 / 144 / 140 B on the first / middle / last level): how close a single launch is to HBM speed
            "per_level_launch": {"bytes_per_descriptor_level": 140.0,
                                 "achieved": 140.0 * per_launch * n_stages / (lvl_total_ms * 1e-3) / 1e9,
                                 "frac": 140.0 * per_launch * n_stages / (lvl_total_ms * 1e-3) / 1e9 / hbm_peak},
            "level_kernel_ms_note": "CUDA events around each level_mma_kernel launch alone; resort_ms = the regrouping pass after each level, its last entry the deferred output of the last level (bucket pass + id write); the per-call plane preparation is in step_ms only",
            "other_engine": dict(fill(other_r, other_total_ms), level_kernel_ms=[round(float(v), 3) for v in other_ms]),
            "fused_single_launch_descent": {"value": fused_rate, "unit": "descriptors/s", "rows": n_f},
            "fp32_input": fp32_in,
            "two_levels_per_launch": pair,
        })
        line = {
            "metric": METRIC, "value": value, "unit": "descriptors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if method == "mma" else "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: quantize %d SIFT-128 byte descriptors per GPU through a "
                                   "k=10 L=6 (1,111,111-node) tree, tree replicated, descriptors sharded" % n,
                       "l2": "input (%.1f GB) is larger than L2; nothing is flushed between steps" % (n * DIM / 1e9),
                       "tree": "synthetic hierarchical mixture, seed 7", "binary64_rechecks_per_step": rechecks // max(args.steps, 1)},
            "e2e": {"value": e2e_value, "unit": "descriptors/s", "h2d_bytes_per_step": n_e2e * DIM,
                    "d2h_bytes_per_step": n_e2e * 4, "descriptors_per_step": n_e2e, "matches_device_path": e2e_ok,
                    "api": "vt_quantize_host (engine.quantize_host)", "cpus_bound_to_gpu_node": affinity,
                    "h2d_achieved_GBps": n_e2e * DIM * world / (e2e_ms * 1e-3) / 1e9, "h2d_ceiling": h2d_ceiling},
            "gpu_launches": args.steps * launches_per_step,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "secondary": secondary,
            "build": build,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, cent, exists, internal, barrier, max_over_ranks,
                    all_nodes=False, db_images=None, queries=None, weighting="tf"):
    """queries/s top-10 through the PUBLIC API: ``Database(dataset, VocabularyTree, comm=...)``.

    The database images' words are synthetic (hashed leaves, ~1k per image: BASELINE configs[3]) and adopted as this
    rank's forward CSR; the QUERIES are descriptor arrays on the host: ``Database.retrieve_batch(paths, k)`` packs them
    into pinned memory, uploads, quantises them through the tree, encodes, scores the inverted file, merges the shard
    / rank lists and returns host ids -- the end-to-end number.  Half of the queries are database members (their
    database rows are the quantised words of the same descriptors), so the self-match must rank first with score 0.
    ``leaf-words`` (headline) counts the stop nodes only; ``all-nodes`` is the reference's embedding semantics
    (vocabulary.py:92-100: every node on the path counts)."""
    n_leaf = K_BRANCH ** DEPTH
    flat = vt.FlatTree.from_device(K_BRANCH, DEPTH, DIM, cent, torch.from_numpy(exists).to(dev),
                                   torch.from_numpy(internal).to(dev), byte_built=True)
    tdev = flat.device(dev)
    leaf_ref = tdev["ref_id"][vt.tree.level_offset(K_BRANCH, DEPTH):].to(torch.int64)
    db_images = args.db_images if db_images is None else db_images
    lo, hi = vt.dist.shard_range(db_images, rank, world)
    wpi = args.words_per_image
    q = args.queries if queries is None else queries
    voc = vt.encoders.VocabularyTree(K_BRANCH, DEPTH, descriptor=lambda im: im, all_nodes=all_nodes, device=dev).set_tree(flat)

    # ---- queries: host descriptor arrays (the same on every rank) --------------------------------------------------
    n_member = q // 2
    member = (torch.arange(n_member, device=dev, dtype=torch.int64) * 197) % max(db_images, 1)
    q_desc = vt.synth.descriptors_from_tree(cent, K_BRANCH, DEPTH, q * wpi, seed=555)
    member_words = engine.quantize(tdev, K_BRANCH, DEPTH, q_desc[:n_member * wpi])
    h_q = q_desc.cpu().numpy().reshape(q, wpi, DIM)
    del q_desc
    names = ["q%06d" % i for i in range(q)]
    qds = vt.ArrayDataset([h_q[i] for i in range(q)], names)

    # ---- this rank's database shard --------------------------------------------------------------------------------
    def shard_words(a, b):
        ids = torch.arange(a, b, device=dev)
        w = torch.empty(len(ids) * wpi, dtype=torch.int32, device=dev)
        for s0 in range(0, len(ids), 65536):
            e0 = min(len(ids), s0 + 65536)
            w[s0 * wpi:e0 * wpi] = hashed_words(torch, ids[s0:e0], wpi, n_leaf, leaf_ref, dev)
        mine = torch.nonzero((member >= a) & (member < b)).flatten()           # members: the words of their descriptors
        if mine.numel():
            dst = ((member[mine] - a)[:, None] * wpi + torch.arange(wpi, device=dev)[None, :]).reshape(-1)
            src = (mine[:, None] * wpi + torch.arange(wpi, device=dev)[None, :]).reshape(-1)
            w[dst] = member_words[src]
        return w, torch.arange(len(ids) + 1, dtype=torch.int64, device=dev) * wpi

    db = vt.Database(qds, voc, comm=comm, weighting=weighting, norm="l2")
    w0, o0 = shard_words(lo, min(hi, lo + 64))
    engine.encode(w0, o0, tdev, DEPTH, flat.n_ref, all_nodes=all_nodes)                                # warm-up
    torch.cuda.synchronize()
    if all_nodes:
        # the all-nodes forward CSR of 1 M images is 3.7 G entries: encode in batches of 64 K images, scatter the top
        # levels into the dense block, keep only the deep levels' entries for the posting lists
        builder = engine.DenseTopBuilder(tdev, DEPTH, flat.n_ref, hi - lo)
        t_encode, nnz_total, sumsq = 0.0, 0, []
        t0 = time.perf_counter()
        for b0 in range(lo, hi, 65536):
            b1 = min(hi, b0 + 65536)
            wb, ob = shard_words(b0, b1)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            cb = engine.encode(wb, ob, tdev, DEPTH, flat.n_ref, all_nodes=True)
            torch.cuda.synchronize()
            t_encode += time.perf_counter() - t1
            nnz_total += cb["nnz"]
            sumsq.append(cb["sumsq"])
            builder.add(cb, b0 - lo)
            del wb, cb
        dense = builder.finish()
        assert dense is not None, "dense counts overflowed a byte"
        db._adopt_dense(dense, torch.cat(sumsq), db_images, lo)
        torch.cuda.synchronize()
        t_invert = time.perf_counter() - t0 - t_encode
        csr = dict(nnz=nnz_total)
        del builder
    else:
        db_words, db_off = shard_words(lo, hi)
        # (one untimed pass first: the caching allocator then holds the multi-GB blocks, whose first cudaMalloc costs up to
        # 0.2 s on a fresh process and is not the encoder)
        csr = engine.encode(db_words, db_off, tdev, DEPTH, flat.n_ref, all_nodes=False, want_df=weighting == "tfidf")
        del csr
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        csr = engine.encode(db_words, db_off, tdev, DEPTH, flat.n_ref, all_nodes=False, want_df=weighting == "tfidf")
        torch.cuda.synchronize()
        t_encode = time.perf_counter() - t0
        del db_words
        t0 = time.perf_counter()
        db._adopt(csr, {}, db_images, lo)                  # inverted file of this rank's shard (Database.index minus the image loop)
        torch.cuda.synchronize()
        t_invert = time.perf_counter() - t0
    index = db._index
    topk = 10

    # ---- device-resident: the query CSR stays in HBM, per-rank score + merged top-k --------------------------------
    qcsr, back = db._query_csr(names)

    def resident_pass():
        out = db._score_local(qcsr, topk, False)
        if world == 1:
            return out["id"], out["score"]
        ids, sc, _ = vt.dist.retrieve_sharded(lambda: out, lo, topk, comm)
        return ids, sc

    for _ in range(2):
        ids, sc = resident_pass()
    steps = 3
    barrier()
    with ProfileRange(torch, args.profile_range == ("retrieval_all_nodes" if all_nodes else "retrieval")):
        t0 = time.perf_counter()
        for _ in range(steps):
            ids, sc = resident_pass()
        barrier()
        ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    back_t = torch.from_numpy(back).to(dev)
    ids_q, sc_q = ids[back_t], sc[back_t]
    ok = bool((ids_q[:n_member, 0].to(torch.int64) == member).all().item() and (sc_q[:n_member, 0] == 0).all().item())

    # ---- end to end through Database.retrieve_batch: host descriptors in, host ids out -----------------------------
    def e2e_pass():
        db._extra_csr.clear()                              # nothing cached between steps: every query is staged and encoded again
        return db.retrieve_batch(names, k=topk)

    for _ in range(2):
        h_ids, h_sc = e2e_pass()
    if os.environ.get("VT_BENCH_PROFILE_E2E"):             # developer switch: where the host time of one pass goes
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        e2e_pass()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        h_ids, h_sc = e2e_pass()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    r_ids, r_sc = ids_q.cpu().numpy(), sc_q.cpu().numpy()
    if weighting == "tf":
        e2e_ok = bool(np.array_equal(h_ids, r_ids) and np.array_equal(h_sc, r_sc))          # integer dots: bit for bit
        e2e_diff = None
    else:
        # binary64 weights accumulated with shared-memory atomics: the order of the additions is not fixed, so two runs
        # agree to rounding (1e-5 is what north_star asks of tf-idf scores); ids may only differ where scores tie to that
        same_id = h_ids == r_ids
        close = np.abs(h_sc - r_sc) <= 1e-9 * np.maximum(1.0, np.abs(r_sc))
        e2e_ok = bool(close.all() and (same_id | (np.abs(h_sc - r_sc) <= 1e-9)).all())
        e2e_diff = {"ids_differ": int((~same_id).sum()), "max_abs_score_diff": float(np.abs(h_sc - r_sc).max())}

    # ---- N > 1: rank 0 recomputes on ONE GPU and the merged result must be identical -------------------------------
    same_as_single = None
    if world > 1 and rank == 0 and not all_nodes:
        n_chk = min(q, 256)
        words_all, off_all = shard_words(0, db_images)
        csr1 = engine.encode(words_all, off_all, tdev, DEPTH, flat.n_ref, all_nodes=all_nodes)
        del words_all
        db1 = vt.Database(qds, voc, weighting=weighting, norm="l2")
        db1._adopt(csr1, {}, db_images, 0)
        ids1, sc1 = db1.retrieve_batch(names[:n_chk], k=topk)
        if weighting == "tf":
            same_as_single = bool(np.array_equal(ids1, h_ids[:n_chk]) and np.array_equal(sc1, h_sc[:n_chk]))
        else:
            # binary64 sums accumulated with shared-memory atomics: equal up to the order of the additions (see e2e above)
            d_sc = np.abs(sc1 - h_sc[:n_chk])
            same_as_single = bool((d_sc <= 1e-9 * np.maximum(1.0, np.abs(sc1))).all()
                                  and ((ids1 == h_ids[:n_chk]) | (d_sc <= 1e-9)).all())
        del db1, csr1

    out = {"metric": "queries/s top-10 @%s images" % ("1M" if db_images == 1_000_000 else db_images),
           "value": q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "queries": q, "db_images": db_images,
           "words_per_image": wpi, "mode": "all-nodes tf/L2 (reference semantics)" if all_nodes else "leaf-words %s/L2" % ("tf-idf" if weighting == "tfidf" else "tf"),
           "api": "Database(dataset, VocabularyTree%s)" % (", comm=Comm()" if world > 1 else ""),
           "entries_per_image": float(csr["nnz"]) / max(hi - lo, 1),
           "self_match_rank0": ok, "index_build_s": t_encode + t_invert, "n_shards_per_gpu": index["n_shards"],
           "e2e": {"value": q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_batch": e2e_ms,
                   "h2d_bytes_per_step": int(q) * wpi * DIM, "d2h_bytes_per_step": int(q) * topk * 12,
                   "api": "Database.retrieve_batch(paths, k=10): host descriptor arrays -> pinned staging -> H2D -> quantise -> encode -> score -> merge -> host ids",
                   "equals_device_resident_result": e2e_ok, "difference": e2e_diff},
           "equals_single_gpu_recomputation": same_as_single,
           # HBM-bound index-side stages: bytes that must move (two encode passes read the 4 B words, the fill pass
           # writes 8 B per entry; inversion = 12 B keys/entry/image out + radix passes at 20 B + 16 B gather/fill)
           "encode": {"seconds": t_encode, "images_per_s": (hi - lo) / t_encode,
                      "algorithmic_GBps": (8.0 * (hi - lo) * wpi + 8.0 * csr["nnz"]) / t_encode / 1e9},
           "invert": {"seconds": t_invert, "postings_per_s": csr["nnz"] / t_invert,
                      "algorithmic_GBps": csr["nnz"] * (12.0 + 20.0 * -(-int(index["n_shards"] * flat.n_ref).bit_length() // 8) + 16.0)
                                          / t_invert / 1e9}}
    if not all_nodes and weighting == "tf":
        alg = 4.0 * wpi * (db_images * wpi / n_leaf)          # 4-byte postings of the query's words
        hbm = {"algorithmic_bytes_per_query": alg, "achieved_GBps": alg * q / (ms * 1e-3) / 1e9}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_scorer_traffic.json")))
            hbm["traffic_bytes_per_query"], hbm["traffic_source"] = float(tj["dram_bytes_per_query"]), tj["source"]
        except (OSError, KeyError, ValueError):
            hbm["traffic_bytes_per_query"], hbm["traffic_source"] = None, None
        out["hbm"] = hbm
    del db, csr, index, qcsr
    torch.cuda.empty_cache()
    return out


def bench_build(args, torch, vt, engine, dev, rank, world, comm, desc, barrier, max_over_ranks):
    """VocabularyTree.fit at BASELINE configs[1]: k=10, L=6 on 10 M descriptors, through the public class
    (``VocabularyTree(..., comm=)``: descriptors sharded over the ranks; integer sums all-reduced on the top
    levels, subtree ownership below)."""
    # the SAME global dataset at every GPU count (rank r owns rows [lo, hi) of it), so the tree -- and the
    # checksum reported below -- must not depend on the number of GPUs
    lo, hi = vt.dist.shard_range(args.build_desc, rank, world)
    cent0 = vt.synth.hierarchical_tree(K_BRANCH, DEPTH, DIM, seed=7, device=dev)
    x_all = vt.synth.descriptors_from_tree(cent0, K_BRANCH, DEPTH, args.build_desc, seed=4242)
    x = x_all[lo:hi].contiguous()
    del cent0

    def checksum(ft):
        return int(ft.device(dev)["centroids"].view(torch.int32).to(torch.int64).sum().item())

    warm = vt.encoders.VocabularyTree(K_BRANCH, DEPTH, descriptor=None, n_iter=1, device=dev, comm=comm)
    warm.fit(x)                                            # warm-up: module load, allocator pools for every level
    voc = vt.encoders.VocabularyTree(K_BRANCH, DEPTH, descriptor=None, n_iter=args.build_iters, device=dev, comm=comm)
    barrier()
    t0 = time.perf_counter()
    voc.fit(x)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    tree, log = voc.flat, voc.build_log
    same = None
    if world > 1 and rank == 0:                             # the sharded tree must be the single-GPU tree, bit for bit
        one = vt.encoders.VocabularyTree(K_BRANCH, DEPTH, descriptor=None, n_iter=args.build_iters, device=dev).fit(x_all)
        same = bool(checksum(one.flat) == checksum(tree) and
                    torch.equal(one.flat.device(dev)["centroids"], tree.device(dev)["centroids"]))
        del one
    del x_all
    passes = sum(l["iters"] + 1 for l in log)
    member_passes = sum(l["members"] * (l["iters"] + 1) for l in log) * (world if world == 1 else 1)
    return {"metric": "tree build k=10 L=6", "seconds": dt, "api": "VocabularyTree(...%s).fit(shard)" % (", comm=Comm()" if world > 1 else ""),
            # 128 B row + 4 B permutation + 4 B node + 1 B label per member and assignment pass (this rank's passes)
            "algorithmic_GBps_rank0": 137.0 * member_passes / dt / 1e9, "descriptors": args.build_desc, "n_iter": args.build_iters,
            "tree_checksum": checksum(tree), "equals_single_gpu_recomputation": same,
            "nodes": int(tree.n_ref), "leaves_reached_depth": int((tree.level_ref == DEPTH).sum()),
            "assign_passes_rank0": passes, "descriptors_per_s": args.build_desc / dt,
            "levels": [{"level": l["level"], "iters": l["iters"], "converged": l["converged"]} for l in log]}


def cpu_baseline(args, vt, cent, exists, internal, ref_id, n_ref, desc, words):
    """The oracle (C port of the reference's descent, OpenMP over all host threads) on a bounded
    sample of the same descriptors; its output doubles as a parity check of the GPU words."""
    from oracle import vt_oracle as vo
    tree = dict(k=K_BRANCH, depth=DEPTH, D=DIM, DP=DIM, centroids=cent.cpu().numpy(), exists=exists,
                internal=internal, ref_id=ref_id, n_ref=n_ref)
    probe = desc[:20_000].cpu().numpy()
    t0 = time.perf_counter()
    vo.descend(tree, probe)
    rate = len(probe) / (time.perf_counter() - t0)
    sample = int(max(20_000, min(len(desc), rate * args.cpu_seconds)))
    x = desc[:sample].cpu().numpy()
    t0 = time.perf_counter()
    leaf = vo.descend(tree, x)
    dt = time.perf_counter() - t0
    mism = int((ref_id[leaf] != words[:sample].cpu().numpy()).sum())
    # what the reference's own Python costs per descriptor: its propagate_feature loop restated line by line
    # (float32 np.linalg.norm per child, one interpreter thread), on a small slice
    n_lit = 3000
    t0 = time.perf_counter()
    lit = vo.descend_literal(tree, x[:n_lit])
    dt_lit = time.perf_counter() - t0
    lit_mism = int((lit != leaf[:n_lit]).sum())
    return {"python_literal": {"value": n_lit / dt_lit, "unit": "descriptors/s", "cores": 1, "sample": n_lit,
                               "what": "line-by-line restatement of VocabularyTree.propagate_feature (vocabulary.py:104-124)",
                               "differs_from_binary64_oracle_on": lit_mism},
            "value": sample / dt, "unit": "descriptors/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "first %d descriptors of the GPU workload, oracle/vt_oracle.c vt_or_descend (binary64), %.1f s" % (sample, dt),
            "gpu_word_mismatches_on_sample": mism}


if __name__ == "__main__":
    main()
