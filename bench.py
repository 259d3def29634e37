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
    ap.add_argument("--all-nodes-images", type=int, default=100_000)
    ap.add_argument("--all-nodes-queries", type=int, default=500)
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-build", action="store_true")
    ap.add_argument("--build-desc", type=int, default=10_000_000, help="descriptors for the tree-build leg (whole job)")
    ap.add_argument("--build-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


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
    with ClockSampler(local) as clocks:
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
    other_ms, _ = stage_profile(lib.vt_quantize_sorted_profile if method == "mma" else lib.vt_quantize_mma_profile, reps=1)
    sort_passes = DEPTH - 1                     # packed paths: one radix pass (hist, scan, scatter) per level transition
    # (the tensor path's byte planes are prepared once per tree, outside the steps); batches of >= 2^22 rows add the
    # deferred output of the last level: one more radix pass (3 kernels) + the bucketed id write
    launches_per_step = DEPTH + 3 * sort_passes + (4 if method == "mma" and n >= (1 << 22) else 0)

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

    # ---- secondary metric: queries/s top-10 over the sharded inverted file ---------------------
    secondary = None
    if not args.no_retrieval:
        secondary = bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, ref_id, n_ref, barrier,
                                    max_over_ranks)
        # the reference's own embedding semantics (every node on the path counts): long posting lists for the
        # top 111 nodes, still walked posting by posting in round 1 -- reported on a smaller database
        secondary["all_nodes_mode"] = bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, ref_id, n_ref,
                                                      barrier, max_over_ranks, all_nodes=True,
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
        alg_bytes = DIM + 4                                        # byte row in, int32 word id out
        per_launch = n
        lvl_total_ms = float(level_ms.sum())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        other_total_ms = float(other_ms.sum())
        # bytes the level kernels must move per descriptor-level: 128 B row gather + 4 B permutation +
        # 4 B node key in, 4 B key + 4 B permutation out (the last level writes the 4 B word id instead)
        level_bytes = [DIM + (8 if l > 0 else 0) + (8 if l + 1 < DEPTH else 4) for l in range(DEPTH)]
        simt = {"bound": "fp32", "kernel": "level_kernel<128,R> (vt_quantize_sorted.cu)", "unit": "TFLOP/s",
                "peak": peak_tf.value, "peak_source": "FFMA microbenchmark on this device (vt_measure_fp32_peak)",
                "flop_per_descriptor": flop_per_desc}
        mma = {"bound": "hbm", "kernel": "level_mma_kernel (vt_quantize_mma.cu, tcgen05.mma kind::i8)", "unit": "GB/s",
               "peak": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
               "bytes_per_descriptor_level": level_bytes}
        def fill(d, total_ms):
            if d["bound"] == "fp32":
                d["achieved"] = flop_per_desc * per_launch / (total_ms * 1e-3) / 1e12
            else:
                d["achieved"] = sum(level_bytes) * per_launch / (total_ms * 1e-3) / 1e9
            d["frac"] = d["achieved"] / d["peak"] if d["peak"] else None
            d["kernel_ms_per_step"] = total_ms
            return d
        main_r, other_r = (mma, simt) if method == "mma" else (simt, mma)
        roofline = fill(main_r, lvl_total_ms)
        roofline.update({
            "launches_per_step": DEPTH,
            "level_kernel_ms": [round(float(v), 3) for v in level_ms],
            "resort_ms": [round(float(v), 3) for v in sort_ms],
            "kernel_share_of_step": lvl_total_ms / kern_ms,
            "whole_step": {"step_ms_events": kern_ms,
                           "algorithmic_GBps": alg_bytes * per_launch / (kern_ms * 1e-3) / 1e9,
                           "algorithmic_TFLOPs": flop_per_desc * per_launch / (kern_ms * 1e-3) / 1e12},
            # DRAM bytes per launch of the dominant kernel: ncu capture of this command at 100 M rows
            # (profiles/r01_launches_final.csv: dram__bytes_read.sum + dram__bytes_write.sum of the six level launches
            # = 13.6 / 14.4 / 14.4 / 14.4 / 14.5 / 15.5 GB, mean 144.6 B per row), scaled to this launch
            "traffic": (144.6 * per_launch) if method == "mma" else None,
            "traffic_source": "ncu launch list of this command at 100 M rows, mean of the six level launches, scaled by rows (not re-measured in this run)" if method == "mma" else None,
            "level_kernel_ms_note": "entry 0 includes the per-call plane preparation of the profile variant; the last entry includes the deferred output (bucket pass + id write)",
            "other_engine": dict(fill(other_r, other_total_ms), level_kernel_ms=[round(float(v), 3) for v in other_ms]),
            "fused_single_launch_descent": {"value": fused_rate, "unit": "descriptors/s", "rows": n_f},
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
                    "api": "vt_quantize_host (engine.quantize_host)", "cpus_bound_to_gpu_node": affinity},
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


def bench_retrieval(args, torch, vt, engine, nat, dev, rank, world, comm, ref_id, n_ref, barrier, max_over_ranks,
                    all_nodes=False, db_images=None, queries=None):
    """leaf-words mode (headline, ~1k words per image) or all-nodes mode (the reference's semantics,
    vocabulary.py:92-100: every node on the path counts, ~3.7k entries per image)."""
    n_leaf = K_BRANCH ** DEPTH
    leaf_ref = torch.from_numpy(ref_id[vt.tree.level_offset(K_BRANCH, DEPTH):].astype(np.int64)).to(dev)
    db_images = args.db_images if db_images is None else db_images
    lo, hi = vt.dist.shard_range(db_images, rank, world)
    wpi = args.words_per_image
    if all_nodes:        # parent / level tables of the full tree, in reference-id space
        nh = vt.tree.heap_size(K_BRANCH, DEPTH)
        heap_of_ref = np.empty(n_ref, np.int64)
        heap_of_ref[ref_id] = np.arange(nh)
        parent_ref = np.where(heap_of_ref > 0, ref_id[np.maximum(heap_of_ref - 1, 0) // K_BRANCH], -1).astype(np.int32)
        lvl = np.zeros(nh, np.uint8)
        for level in range(1, DEPTH + 1):
            o = vt.tree.level_offset(K_BRANCH, level)
            lvl[o:o + K_BRANCH ** level] = level
        dummy = dict(parent_ref=torch.from_numpy(parent_ref).to(dev), level_ref=torch.from_numpy(lvl[heap_of_ref]).to(dev))
    else:
        dummy = dict(parent_ref=torch.zeros(1, dtype=torch.int32, device=dev), level_ref=torch.zeros(1, dtype=torch.uint8, device=dev))

    def encode_ids(ids):
        words = torch.empty(len(ids) * wpi, dtype=torch.int32, device=dev)
        step = 65536
        for s in range(0, len(ids), step):
            e = min(len(ids), s + step)
            words[s * wpi:e * wpi] = hashed_words(torch, ids[s:e], wpi, n_leaf, leaf_ref, dev)
        off = torch.arange(len(ids) + 1, dtype=torch.int64, device=dev) * wpi
        return engine.encode(words, off, dummy, DEPTH, n_ref, all_nodes=all_nodes)

    # words of this rank's images (synthetic), then the two index-side stages timed separately
    db_ids = torch.arange(lo, hi, device=dev)
    db_words = torch.empty(len(db_ids) * wpi, dtype=torch.int32, device=dev)
    for s0 in range(0, len(db_ids), 65536):
        e0 = min(len(db_ids), s0 + 65536)
        db_words[s0 * wpi:e0 * wpi] = hashed_words(torch, db_ids[s0:e0], wpi, n_leaf, leaf_ref, dev)
    db_off = torch.arange(len(db_ids) + 1, dtype=torch.int64, device=dev) * wpi
    engine.encode(db_words[:wpi * 64], db_off[:65], dummy, DEPTH, n_ref, all_nodes=all_nodes)      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csr = engine.encode(db_words, db_off, dummy, DEPTH, n_ref, all_nodes=all_nodes)
    torch.cuda.synchronize()
    t_encode = time.perf_counter() - t0
    del db_words
    t0 = time.perf_counter()
    index = engine.build_index(csr, n_ref)
    torch.cuda.synchronize()
    t_invert = time.perf_counter() - t0
    t_index = t_encode + t_invert
    q = args.queries if queries is None else queries
    member = (torch.arange(q // 2, device=dev, dtype=torch.int64) * 197) % max(db_images, 1)
    fresh = db_images + torch.arange(q - q // 2, device=dev, dtype=torch.int64)
    qcsr = encode_ids(torch.cat([member, fresh]))
    topk = 10

    def one_pass():
        local_out = engine.score(index, qcsr, topk)
        if world == 1:
            return local_out["id"], local_out["score"]
        ids, sc, _ = vt.dist.retrieve_sharded(lambda: local_out, lo, topk, comm)
        return ids, sc

    for _ in range(2):
        ids, sc = one_pass()
    steps = 3
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, sc = one_pass()
    barrier()
    ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    ok = bool((ids[:q // 2, 0].to(torch.int64) == member).all().item() and (sc[:q // 2, 0] == 0).all().item())
    out = {"metric": "queries/s top-10 @%s images" % ("1M" if db_images == 1_000_000 else db_images),
           "value": q / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "queries": q, "db_images": db_images,
           "words_per_image": wpi, "mode": "all-nodes tf/L2 (reference semantics)" if all_nodes else "leaf-words tf/L2",
           "entries_per_image": float(csr["nnz"]) / max(hi - lo, 1),
           "self_match_rank0": ok, "index_build_s": t_index, "n_shards_per_gpu": index["n_shards"],
           # HBM-bound index-side stages: bytes that must move (two encode passes read the 4 B words, the fill pass
           # writes 8 B per entry; inversion = 12 B keys/entry/image out + radix passes at 20 B + 16 B gather/fill)
           "encode": {"seconds": t_encode, "images_per_s": (hi - lo) / t_encode,
                      "algorithmic_GBps": (8.0 * (hi - lo) * wpi + 8.0 * csr["nnz"]) / t_encode / 1e9},
           "invert": {"seconds": t_invert, "postings_per_s": csr["nnz"] / t_invert,
                      "algorithmic_GBps": csr["nnz"] * (12.0 + 20.0 * -(-int(index["n_shards"] * n_ref).bit_length() // 8) + 16.0)
                                          / t_invert / 1e9}}
    if not all_nodes:
        alg = 8.0 * wpi * (db_images * wpi / n_leaf)
        out["hbm"] = {"algorithmic_bytes_per_query": alg, "achieved_GBps": alg * q / (ms * 1e-3) / 1e9,
                      # ncu --set full of score_shard_tf_kernel on this workload (profiles/r01_ncu_final_scorer.txt):
                      # 219.7 GB of DRAM reads per 10 k queries = 2.7x the algorithmic bytes (32 B sectors around ~64 B
                      # posting runs and 8 B row-pointer pairs), 3.2 TB/s; issue slots 27 % busy
                      "traffic_bytes_per_query": 21.97e6, "traffic_source": "ncu capture of this workload (not re-measured in this run)"}
    return out


def bench_build(args, torch, vt, engine, dev, rank, world, comm, desc, barrier, max_over_ranks):
    """VocabularyTree.fit at BASELINE configs[1]: k=10, L=6 on 10 M descriptors (sharded over the
    ranks, integer sums all-reduced per level-iteration)."""
    # the SAME global dataset at every GPU count (rank r owns rows [lo, hi) of it), so the tree -- and the
    # checksum reported below -- must not depend on the number of GPUs
    lo, hi = vt.dist.shard_range(args.build_desc, rank, world)
    cent0 = vt.synth.hierarchical_tree(K_BRANCH, DEPTH, DIM, seed=7, device=dev)
    x = vt.synth.descriptors_from_tree(cent0, K_BRANCH, DEPTH, args.build_desc, seed=4242)[lo:hi].contiguous()
    del cent0
    log = []
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    engine.build_tree(x, DIM, K_BRANCH, DEPTH, n_iter=1, comm=comm)   # warm-up: module load, allocator pools for every level
    barrier()
    t0 = time.perf_counter()
    tree = engine.build_tree(x, DIM, K_BRANCH, DEPTH, n_iter=args.build_iters, comm=comm, stats=stats,
                             log=lambda **kw: log.append(kw))
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    passes = sum(l["iters"] + (0 if l["converged"] else 1) + (1 if l["converged"] else 0) for l in log)
    member_passes = sum(l["members"] * (l["iters"] + 1) for l in log) * world
    return {"metric": "tree build k=10 L=6", "seconds": dt,
            # 128 B row + 4 B permutation + 4 B node + 1 B label per member and assignment pass
            "algorithmic_GBps": 137.0 * member_passes / dt / 1e9, "descriptors": args.build_desc, "n_iter": args.build_iters,
            "tree_checksum": int(tree.device(dev)["centroids"].view(torch.int32).to(torch.int64).sum().item()),
            "nodes": int(tree.n_ref), "leaves_reached_depth": int((tree.level_ref == DEPTH).sum()),
            "assign_passes": passes, "member_assignments_per_s": member_passes / dt,
            "descriptors_per_s": args.build_desc / dt, "binary64_rechecks": int(stats[0].item()),
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
