#!/usr/bin/env python
"""Benchmark of the segmentation hot path (BASELINE.json metric: DP GCUPS and reads/s vs the reference CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R] [--config c2|c1]

One "step" = one pass of the hot path (Aligner.align with calc_probabilities=True: backward, forward,
posterior, posterior-Viterbi, traceback, medians) over R synthetic reads per GPU (default: the 100 000 reads of
BASELINE.json configs[1]), handed to the C ABI in batches of --batch reads; all reads are distinct and resident in
HBM before the timed region (`value`) or in pinned host memory (`e2e`).
N > 1 is launched with torchrun, one rank per GPU; reads shard across ranks with no communication (weak scaling).

Workload c2 (default, the configuration the metric is quoted on for one GPU): rna004 pore, synthetic 9-mer
model, basic mode, reads uniform 0.5-5 kb at ~30 samples/base, band 400.
GCUPS = in-band lattice cells (each counted once, whatever number of passes touches it) / second / 1e9.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (pore, model, min_len, max_len, samples/base, dwell, default reads per step per GPU)
    "c1": ("rna002", "rna002_5mer", 1000, 1000, 30.0, "geometric", 1000),
    "c2": ("rna004", "synthetic_rna004_9mer", 500, 5000, 30.0, "geometric", 100000),
    # c2 with a per-kmer sigma (what a trained model looks like): same kernels, the emission constants are per column anyway
    "c2v": ("rna004", "synthetic_rna004_9mer_varsd", 500, 5000, 30.0, "geometric", 100000),
    # config 4 (long-read stress): 50 kb reads of ~2 M samples, Gamma-4 dwell; 10 000 reads = 8 steps
    "c4": ("rna004", "synthetic_rna004_9mer", 50000, 50000, 40.0, "gamma", 2560),
    # config 5 (dynamont-train): one step = one pooled Baum-Welch iteration (expected counts of all reads, ONE all-reduce of
    # the 3*4^9+4 statistics over NCCL, M-step on the device); 1 M reads = 10 steps of 100 000 per GPU at 1 GPU
    "c5": ("rna004", "synthetic_rna004_9mer", 500, 5000, 30.0, "geometric", 100000),
}
TRAIN_CONFIGS = {"c5"}
# config 3 (resquiggle / NTK mode with kmer polishing): its own code path and its own unit (run_ntk below).  The 9-mer
# pore needs the dense T x 4^9 pre-pass: reads of ~60 bases at 12.5 samples/base (T ~ 750) are what fits a line that ends
# in minutes; the 5-mer line beside it uses realistic 1 kb reads.  The reference throws for every input in this mode
# (SURVEY.md F2); its repaired build needs ~4 min and 10 GB for ONE such 9-mer read, so no CPU arm is run here.
NTK_CONFIGS = {"c3"}
MODELS_DIR = os.path.join(ROOT, "tests", "golden", "_models")


# k_align template instances of the build variants (csrc/engine.cu): (Cfg, resident single-warp CTAs per SM)
VARIANT_KERNELS = {0: ("Cfg<13,16,4,8>", 7), 1: ("Cfg<13,8,4,8>", 10), 2: ("Cfg<13,8,4,8>", 12), 3: ("Cfg<13,8,4,8>", 8),
                   4: ("Cfg<13,8,4,8,UNI>", 8), 5: ("Cfg<13,8,4,8,UNI>", 9), 6: ("Cfg<13,8,4,8,UNI>", 10),
                   7: ("Cfg<13,8,4,8,UNI>", 11), 8: ("Cfg<13,8,4,8,UNI>", 12), 9: ("Cfg<13,8,4,8>", 9),
                   10: ("Cfg<13,8,8,8,UNI>", 8), 11: ("Cfg<13,8,8,8>", 8), 12: ("Cfg<13,8,8,8,UNI>", 1), 13: ("Cfg<13,8,8,8>", 1)}


def kernel_label(variant, lin):
    cfg, minb = VARIANT_KERNELS.get(variant, ("Cfg<?>", 0))
    if not lin:
        return f"k_align<{cfg.replace(',UNI', '')},1,{minb},LIN=0> (log2-domain FP32, 2 MUFU per cell-update)"
    uni = " uniform-sigma emission constants," if "UNI" in cfg else ""
    if variant in (12, 13):
        return f"k_align<{cfg},1,1,LIN=1,WPC=8> (linear-domain FP32,{uni} CTAs of 8 warps in step pass by pass, 1 MUFU per cell-update)"
    return f"k_align<{cfg},1,{minb},LIN=1> (linear-domain FP32,{uni} 1 MUFU per cell-update)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + sorted(NTK_CONFIGS))
    ap.add_argument("--em-iterations", type=int, default=0, help="c5: (unused; every step is one EM iteration)")
    ap.add_argument("--reads", type=int, default=0, help="reads per step per GPU (0 = config default)")
    ap.add_argument("--batch", type=int, default=20000, help="reads per C-ABI call (a step runs ceil(reads/batch) calls)")
    ap.add_argument("--seed", type=int, default=20262000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--warps-per-sm", type=int, default=0)
    ap.add_argument("--z-only", action="store_true", help="experiment: backward pass only (calc_probabilities=False)")
    ap.add_argument("--opt", action="append", default=[], help="experiment: KEY=VALUE passed to dyn_set_option")
    ap.add_argument("--variant", type=int, default=-1, help="kernel build variant (see csrc/engine.cu); -1 = library default")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------------------------------------
def gen_reads_numpy(cfg, n, seed):
    """CPU generation (numpy) of n reads of the config's distribution — used for the CPU baseline sample."""
    from dynamont_b200.synth import PORE_INFO, materialize_model, native_model, synth_read
    pore, model, lo, hi, spb, dwell, _ = cfg
    path = materialize_model(model, MODELS_DIR)
    nm, ns = native_model(path, pore)
    k = PORE_INFO[pore][1]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s, q, _ = synth_read(rng, nm, ns, k, L, spb, dwell=dwell)
        out.append((s, q))
    return path, out


def gen_reads_torch(cfg, n, seed, device):
    """GPU generation of the same distribution: returns device signal (float32), device bases (uint8 ASCII),
    host offsets."""
    import torch
    from dynamont_b200.synth import PORE_INFO, materialize_model, native_model
    pore, model, lo, hi, spb, dwell, _ = cfg
    path = materialize_model(model, MODELS_DIR)
    nm, ns = native_model(path, pore)
    k = PORE_INFO[pore][1]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    mu = torch.tensor(nm, dtype=torch.float32, device=device)
    sd = torch.tensor(ns, dtype=torch.float32, device=device)
    lens = torch.randint(lo, hi + 1, (n,), generator=g, device=device, dtype=torch.int64)
    seq_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    seq_off[1:] = torch.cumsum(lens, 0)
    B = int(seq_off[-1].item())
    digits = torch.randint(0, 4, (B,), generator=g, device=device, dtype=torch.int64)
    pos = torch.arange(B, device=device) - torch.repeat_interleave(seq_off[:-1], lens)
    digits[pos < k] = 0  # k x 'A' prefix (front end does this for RNA reads, segment.py:155-158)
    ids = torch.zeros(B, dtype=torch.int64, device=device)
    for i in range(k):
        ids = ids * 4 + torch.roll(digits, -i)
    valid = pos <= (torch.repeat_interleave(lens, lens) - k)  # kmer start positions
    if dwell == "geometric":
        u = torch.rand(B, generator=g, device=device, dtype=torch.float64).clamp_min(1e-300)
        d = torch.floor(torch.log(u) / np.log1p(-1.0 / spb)).to(torch.int64) + 1
    else:
        gam = torch.distributions.Gamma(torch.tensor(4.0, device=device), torch.tensor(4.0 / spb, device=device))
        d = torch.round(gam.sample((B,))).to(torch.int64)
    d = torch.clamp(d, min=2) * valid
    # per-read sample counts
    csum = torch.zeros(B + 1, dtype=torch.int64, device=device)
    csum[1:] = torch.cumsum(d, 0)
    sig_off = csum[seq_off]
    total = int(csum[-1].item())
    signal = torch.empty(total, dtype=torch.float32, device=device)
    chunk = 1 << 22  # kmer positions per chunk keeps repeat_interleave temporaries bounded
    for a in range(0, B, chunk):
        b = min(B, a + chunk)
        per = torch.repeat_interleave(ids[a:b], d[a:b])
        seg = signal[int(csum[a].item()):int(csum[b].item())]
        seg.copy_(mu[per] + sd[per] * torch.randn(per.numel(), generator=g, device=device))
    bases = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)[digits]
    return path, signal, bases, sig_off.cpu().numpy().astype(np.uint64), seq_off.cpu().numpy().astype(np.uint64)


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation, one single-threaded process per host core
# ----------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(model_path, pore, kind, train=False):
    import oracle
    _W["al"] = oracle.Reference(model_path, pore) if kind == "reference" else oracle.Oracle(model_path, pore)
    _W["train"] = train


def _cpu_align(item):
    sig, seq = item
    if _W.get("train"):
        # config 5: the reference's per-read Baum-Welch (NTAligner::train, NT:567-639), what dynamont-train's workers run
        return len(_W["al"].train(sig, seq)["emission_model"]["mean"])
    r = _W["al"].align(sig, seq, True)
    return len(r["signal_positions"])


def cpu_pool_plan(cfg, sample):
    import oracle
    import psutil
    pore, model, lo, hi, spb, dwell, _ = cfg
    kind = "reference" if oracle.have_reference() else "port"
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # the reference holds 8 T x B double matrices per read: ~64*T*403 bytes
    worst = 64.0 * (hi * spb + 1) * 403
    avail = psutil.virtual_memory().available
    procs = int(max(1, min(cores, (0.5 * avail) // worst)))
    if sample <= 0:
        sample = max(4 * procs, 8)  # ~4 reads per worker: bounded (tens of seconds per step) yet not dominated by the longest read
    return kind, procs, sample


def run_cpu_sample(cfg, kind, procs, reads, model_path, train=False):
    """Wall time of aligning (training on) `reads` with `procs` single-threaded worker processes (segment.py:304-324 shape)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(model_path, cfg[0], kind, train)) as pool:
        pool.map(_cpu_align, reads[:procs], chunksize=1)  # construct aligners / warm caches (untimed)
        t0 = time.perf_counter()
        list(pool.imap_unordered(_cpu_align, reads, chunksize=1))
        dt = time.perf_counter() - t0
    return dt


def cells_of(reads, model_path, pore):
    import oracle
    o = oracle.Oracle(model_path, pore)
    return sum(o.cells(len(s), len(q)) for s, q in reads)


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_ntk(args):
    """--config c3: resquiggle mode, dyn_ntk_align_batch (pre-passes + sparse 5-state stages, reads on a pool of CUDA streams).
    Unit of work: dense pre-pass cells T*N + T*K per read (never mixed with basic-mode GCUPS)."""
    import torch
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import PORE_INFO, materialize_model, native_model, synth_read
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lines = {}
    for tag, pore, model, length, spb, n in (("k9", "rna004", "synthetic_rna004_9mer", 60, 12.5, args.reads or 32),
                                             ("k5", "dna_r9", "rna004_5mer", 1000, 12.5, args.reads or 192)):
        path = materialize_model(model, MODELS_DIR)
        nm, ns = native_model(path, pore)
        k = PORE_INFO[pore][1]
        rng = np.random.default_rng(args.seed + 31 * rank + (9 if tag == "k9" else 5))
        reads = [synth_read(rng, nm, ns, k, length, spb) for _ in range(n)]
        sigs, seqs = [r[0].astype(np.float32) for r in reads], [r[1] for r in reads]
        al = Aligner(path, pore, mode="resquiggle", device=local_rank)
        cells = sum((s.size + 1) * (len(q) - k + 2) + (s.size + 1) * 4 ** k for s, q in zip(sigs, seqs))
        for _ in range(max(1, args.warmup)):
            al.align_batch(sigs, seqs, True)  # the same batch shape: the worker pool's lattices are allocated (and kept) here
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ok = 0
        for _ in range(args.steps):
            res = al.align_batch(sigs, seqs, True)
            ok = sum(isinstance(r, dict) for r in res)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        lines[tag] = {"reads": n, "reads_ok": ok, "read_length": length, "samples_per_read": int(np.mean([s.size for s in sigs])),
                      "dense_cells_per_step": int(cells), "ms_per_step": dt * 1e3, "reads_per_s": n / dt,
                      "g_dense_cells_per_s": cells / dt / 1e9}
    if rank == 0:
        k9 = lines["k9"]
        print(json.dumps({
            "metric": "ntk_dense_gcups", "value": k9["g_dense_cells_per_s"] * world, "unit": "G dense pre-pass cells/s (T*N + T*K)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": k9["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "reads_per_s": k9["reads_per_s"] * world,
            "config": {"workload": "c3: resquiggle (NTK) mode with kmer polishing, 9-mer pore (rna004 synthetic model), polyA-prefixed reads "
                                   "of 60 b at 12.5 samples/base; beside it 5-mer reads of 1 kb", "k9": k9, "k5": lines["k5"],
                       "timing": "host wall clock around dyn_ntk_align_batch (H2D, kernels on a pool of streams, D2H), per rank"},
            "roofline": None, "cpu_baseline": None,
            "note": "no reference arm: the unmodified reference throws in this mode (SURVEY.md F2); the repaired reference needs "
                    "~4 min and 10 GB for one 9-mer read of 865 samples (tools/make_golden_ntk.py big) = 0.004 reads/s per host core"}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.config in NTK_CONFIGS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference throws for every input in resquiggle mode (SURVEY.md F2)"}))
            return
        return run_ntk(args)
    cfg = CONFIGS[args.config]
    pore, model, lo, hi, spb, dwell, default_reads = cfg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    reads_per_gpu = args.reads or default_reads
    train = args.config in TRAIN_CONFIGS
    what = "dynamont-train (pooled Baum-Welch iteration: expected counts, all-reduce, M-step)" if train else "basic mode align(calc_probabilities=True)"
    workload = f"{args.config}: {pore} {model} {what}, reads {lo}-{hi} b at ~{spb:g} samples/base ({dwell} dwell), band 400"

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        kind, procs, sample = cpu_pool_plan(cfg, args.cpu_sample)
        model_path, reads = gen_reads_numpy(cfg, sample, args.seed + 17)
        cells = cells_of(reads, model_path, pore)
        if kind == "reference":
            # the workers are forked from this process: load the compiled reference here as well, so that the library the
            # arm runs (oracle/_ref/libdynamont_ref.so) is visible in this process' memory map
            import oracle
            _parent_handle = oracle.Reference(model_path, pore)  # noqa: F841
        times = []
        for i in range(args.warmup + args.steps):
            dt = run_cpu_sample(cfg, kind, procs, reads, model_path, train)
            if i >= args.warmup:
                times.append(dt)
        dt = float(np.mean(times))
        gcups = cells / dt / 1e9
        line = {
            "impl": "reference", "metric": "dp_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "reads_per_s": len(reads) / dt,
            "config": {"workload": workload, "reads_per_step": len(reads), "cells_per_step": int(cells)},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": procs, "kind": kind,
                             "sample": f"{len(reads)} reads of the workload per step, one single-threaded process per core "
                                       f"({procs} processes), " + ("train()" if train else "align(calc_probabilities=True)"),
                             "reads_per_s": len(reads) / dt},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu_baseline = None
    if world == 1 and args.gpus == 1 and not args.no_cpu_baseline:
        try:
            kind, procs, sample = cpu_pool_plan(cfg, args.cpu_sample)
            model_path, reads = gen_reads_numpy(cfg, sample, args.seed + 17)
            cells = cells_of(reads, model_path, pore)
            dt = run_cpu_sample(cfg, kind, procs, reads, model_path, train)
            cpu_baseline = {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": procs, "kind": kind,
                            "sample": f"{len(reads)} reads of the workload, one single-threaded process per core "
                                      f"({procs} processes), align(calc_probabilities=True), {dt:.1f} s wall",
                            "reads_per_s": len(reads) / dt}
        except Exception as e:  # the baseline is context, never fatal
            cpu_baseline = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dynamont_b200 import Aligner
    # one step = reads_per_gpu reads, handed to the C ABI in batches of <= args.batch reads (all distinct, all resident)
    n_batches = max(1, -(-reads_per_gpu // args.batch))
    per_batch = -(-reads_per_gpu // n_batches)
    batches = []
    for bi in range(n_batches):
        nb = min(per_batch, reads_per_gpu - bi * per_batch)
        model_path, d_sig, d_bas, so, qo = gen_reads_torch(cfg, nb, args.seed + 1000 * rank + 7 * bi, dev)
        batches.append({"sig": d_sig, "bases": d_bas, "sig_off": so, "seq_off": qo})
    al = Aligner(model_path, pore, device=local_rank)
    al.set_stream(torch.cuda.current_stream().cuda_stream)
    if args.warps_per_sm:
        al.set_option("warps_per_sm", args.warps_per_sm)
    if args.variant >= 0:
        al.set_option("variant", args.variant)
    for kv in args.opt:
        k_, v_ = kv.split("=")
        al.set_option(k_, float(v_))
    cells = sum(al.batch_cells(b["sig_off"], b["seq_off"]) for b in batches)
    n_reads = sum(b["sig_off"].size - 1 for b in batches)
    n_samples = sum(int(b["sig_off"][-1]) for b in batches)
    n_bases = sum(int(b["seq_off"][-1]) for b in batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats = torch.zeros(3 * al.num_kmers + 4, dtype=torch.float64, device=dev) if train else None
    ar_ms, mstep_ms = [], []

    def step_train():
        """One pooled EM iteration: dyn_train_accumulate over every batch (device-resident inputs, statistics accumulated
        in `stats` on the device), ONE all-reduce of `stats` in place, M-step on the device (dyn_train_mstep_device)."""
        ok, kms, nl, fb, rl = 0, 0.0, 0, 0, 0
        rib = step_device.rib
        stats.zero_()
        for b in batches:
            st = al.train_accumulate_packed(b["sig"].data_ptr(), b["sig_off"], b["bases"].data_ptr(), b["seq_off"],
                                            stats.data_ptr(), device=True)
            tm = al.last_timing()
            kms += tm["dp_ms"]
            nl += tm["launches"]
            fb += tm["log2_fallback_reads"]
            rl += tm["lin_retry_reads"]
            rib[0] += tm["ribbon_reads"]
            rib[1] += tm["ribbon_faults"]
            ok += int((st == 0).sum())
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        a1.record()
        t0m = time.perf_counter()
        al.mstep_device(stats.data_ptr())  # synchronises the handle's stream
        torch.cuda.synchronize()
        mstep_ms.append((time.perf_counter() - t0m) * 1e3)
        ar_ms.append(a0.elapsed_time(a1))
        return ok, kms, nl + 2, fb, rl

    from collections import deque
    pending = deque()
    # batches in flight: two hide the host side of a batch; config 4 keeps three, so that the seconds-long full-band hand-over of
    # the few reads whose alignment leaves the reference band overlaps the ribbon kernels of the next two batches
    depth = 3 if args.config == "c4" else 2
    acc = {"ok": 0, "kms": 0.0, "nl": 0, "fb": 0, "rl": 0}

    outsets = deque()

    def collect(item):
        job, b = item
        res, _, _, _ = outv = al.wait(job)
        outsets.append(outv)
        tm = al.last_timing()
        acc["kms"] += tm["dp_ms"]
        acc["nl"] += tm["launches"]
        acc["fb"] += tm["log2_fallback_reads"]
        acc["rl"] += tm["lin_retry_reads"]
        step_device.rib[0] += tm["ribbon_reads"]
        step_device.rib[1] += tm["ribbon_faults"]
        acc["ok"] += sum(1 for i in range(b["sig_off"].size - 1) if res[i].status == 0)

    def step_device(drain=True):
        """One step: every batch through the C ABI with device-resident inputs (dyn_align_submit_device / dyn_align_wait,
        two batches in flight: the result copy + fan-out of a batch overlaps the kernels of the next).  The counters of the
        completed batches accumulate in `acc`."""
        if train:
            ok, kms, nl, fb, rl = step_train()
            acc["ok"] += ok; acc["kms"] += kms; acc["nl"] += nl; acc["fb"] += fb; acc["rl"] += rl
            return
        for b in batches:
            out = outsets.popleft() if outsets else None  # result buffers of a completed job, rotated
            pending.append((al.submit_packed(b["sig"].data_ptr(), b["sig_off"], b["bases"].data_ptr(), b["seq_off"],
                                             not args.z_only, device=True, out=out), b))
            if len(pending) >= depth:
                collect(pending.popleft())
        while drain and pending:
            collect(pending.popleft())

    step_device.rib = [0, 0]
    for _ in range(args.warmup):
        step_device(True)
    step_device.rib = [0, 0]
    for k_ in acc:
        acc[k_] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_device(i == args.steps - 1)
    ev1.record()
    barrier()
    n_ok = acc["ok"] // args.steps
    dp_ms = [acc["kms"] / args.steps]
    launches, fallbacks, lin_retries = acc["nl"], acc["fb"], acc["rl"]
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import ctypes
        import psutil
        need = n_samples * 4 + n_bases
        distinct = len(batches) if psutil.virtual_memory().available > 3 * need * max(1, world) else 1
        host = []
        for b in batches[:distinct]:
            hs = torch.empty(b["sig"].numel(), dtype=torch.float32, pin_memory=True)
            hs.copy_(b["sig"])
            hb = torch.empty(b["bases"].numel(), dtype=torch.uint8, pin_memory=True)
            hb.copy_(b["bases"])
            host.append((hs, hb, b["sig_off"], b["seq_off"]))
        torch.cuda.synchronize()
        # the end-to-end phase owns the device: the resident copies of the inputs go (the C ABI stages its own per batch)
        for b in batches:
            b.pop("sig", None)
            b.pop("bases", None)
        del d_sig, d_bas
        torch.cuda.empty_cache()

        # the call a streaming user makes: dyn_align_submit / dyn_align_wait (Aligner.submit_packed / wait), two batches in
        # flight, so that the H2D copy of the next batch and the D2H copy + fan-out of the previous one overlap the kernels
        from collections import deque
        inflight = deque()

        def step_host_train():
            stats.zero_()
            for bi in range(len(batches)):
                hs, hb, so, qo = host[bi % len(host)]
                al.train_accumulate_packed(hs.data_ptr(), so, hb.data_ptr(), qo, stats.data_ptr(), device=False)
            if world > 1:
                dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            al.mstep_device(stats.data_ptr())
            torch.cuda.synchronize()

        def step_host(drain):
            if train:
                return step_host_train()
            for bi in range(len(batches)):
                hs, hb, so, qo = host[bi % len(host)]
                out = outsets.popleft() if outsets else None
                inflight.append(al.submit_packed(hs.data_ptr(), so, hb.data_ptr(), qo, True, out=out))
                if len(inflight) >= 2:
                    outsets.append(al.wait(inflight.popleft()))
            while drain and inflight:
                outsets.append(al.wait(inflight.popleft()))
        for _ in range(max(1, args.warmup)):
            step_host(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            step_host(i == args.steps - 1)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) / args.steps
        e2e_ms = max(e0.elapsed_time(e1) / args.steps, 0.0)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        nseg = sum(int(al._lib.dyn_count_segments(al._h, b["seq_off"].ctypes.data_as(u64p), b["seq_off"].size - 1))
                   for b in batches)
        e2e = {"ms": e2e_ms, "wall_ms": wall * 1e3, "h2d": n_samples * 4 + n_bases + (n_reads + 1) * 8 + n_reads * 48,
               "d2h": (n_reads * 52 + 2 * 4 ** 9 * 8) if train else (nseg * 12 + n_reads * 52), "distinct_host_batches": len(host)}

    # ---- reduce over ranks: time = max, work = sum
    tstats = torch.tensor([ms, e2e["ms"] if e2e else 0.0, float(np.mean(dp_ms))], dtype=torch.float64, device=dev)
    work = torch.tensor([float(cells), float(n_reads), float(n_ok)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tstats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_all, e2e_ms_all, dp_ms_all = [float(x) for x in tstats.tolist()]
    cells_all, reads_all, ok_all = [float(x) for x in work.tolist()]
    gcups = cells_all / (ms_all * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    # north_star: "achieved FP32/SFU throughput against B200 peak".  `achieved` keeps the ALGORITHMIC definition of
    # the reference's log-space forward + backward = 2 passes x 2 MUFU (ex2 + lg2) per in-band lattice cell
    # (training: + 2 for the two exp of gamma, NT:505-508), divided by the kernel time measured with CUDA events.  The
    # ribbon kernels evaluate only the window of the band in which FP32 values are non-zero (63 of ~401 columns) with
    # 1 MUFU per evaluated cell-update, so `frac` can exceed 1: `executed` / `executed_frac` are the MUFU ops actually
    # issued, and `issue` (from the committed ncu capture of the same kernel) is what really bounds it: issue slots.
    props = torch.cuda.get_device_properties(dev)
    sms = props.multi_processor_count
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    counters = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2f_ribbon_counters.json")) as fh:
            counters = json.load(fh)
    except Exception:
        pass
    clk_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    mufu_peak = 16.0 * sms * clk_mhz * 1e6 / 1e9           # G MUFU op/s at the clock seen under load
    mufu_peak_max = 16.0 * sms * (peaks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9
    dp_s = dp_ms_all * 1e-3
    lin = not any(kv.startswith("arith=") and float(kv.split("=")[1]) != 0 for kv in args.opt)
    rib_frac = step_device.rib[0] / max(1.0, float(n_reads) * args.steps)  # share of the reads the ribbon kernels kept
    ribbon_on = rib_frac > 0.5
    alg_per_cell = 6.0 if train else 4.0
    alg_mufu = alg_per_cell * cells / dp_s / 1e9
    rows = float(n_samples)
    rib_info = al.ribbon_fault_reasons()
    if ribbon_on:
        # executed: sweeps x 64 ring slots per row x 1 MUFU.  3 sweeps (backward, recomputation, forward); two-level checkpoints
        # replay the backward sweep once more; the records-free layout of long reads adds a second recomputation + forward
        # sweep (and its replay) for the path posteriors
        sweeps = 7.0 if rib_info.get("records_free_layout") else (4.0 if rib_info.get("two_level_checkpoints") else 3.0)
        exe_mufu = sweeps * 64.0 * rows / dp_s / 1e9
        kern = "k_ribbon<RCfg<2,false,16>,%d,5> (ribbon: 63-column window that follows the probability mass, groups of 16 rows, linear-domain " \
               "FP32 block floating point in packed FP32x2 (FFMA2 / FMUL2), 1 MUFU per evaluated cell-update, %d sweeps%s)" % (
                   3 if rib_info.get("records_free_layout") else (2 if train else 1), int(sweeps),
                   "; long reads: two-level checkpoints, records-free scratch, path posteriors from a second forward sweep"
                   if rib_info.get("records_free_layout") else "")
    else:
        exe_mufu = (3.0 if lin else 6.0) * cells / dp_s / 1e9
        kern = kernel_label(al.last_timing()["variant"], lin)
    ck_key = "train" if train else ("align_records_free" if rib_info.get("records_free_layout") else "align")
    ck = counters.get(ck_key, {}) if ribbon_on else {}
    traffic = ck.get("dram_bytes_per_row") * rows / max(1, len(batches)) if ck.get("dram_bytes_per_row") else None
    roofline = {
        "bound": "sfu",
        "kernel": kern,
        "achieved": alg_mufu, "peak": mufu_peak, "unit": "G MUFU op/s", "frac": alg_mufu / mufu_peak,
        "achieved_definition": "algorithmic: %g MUFU per in-band lattice cell (log-space forward + backward of the reference, SURVEY 8d) x cells / kernel time (CUDA events)" % alg_per_cell,
        "executed": exe_mufu, "executed_frac": exe_mufu / mufu_peak,
        "evaluated_cells_frac": (64.0 * rows / cells) if ribbon_on else 1.0,
        "peak_at_max_clock": mufu_peak_max, "peak_source": "16 MUFU/clk/SM x SMs x SM clock sampled by nvidia-smi during the timed region",
        "kernel_ms": dp_ms_all, "lattice_rows_per_s": 3.0 * rows / dp_s,
        "log2_fallback_reads": int(fallbacks), "lin_retry_reads": int(lin_retries),
        "ribbon_reads": int(step_device.rib[0]), "ribbon_fault_reads": int(step_device.rib[1]),
        "ribbon_fault_reasons": rib_info,
        "issue": {"source": ck.get("source"), "warp_instructions_per_lattice_row": ck.get("instr_per_row"),
                  "issue_slots_busy_pct": ck.get("issue_busy_pct"), "xu_pipe_pct": ck.get("xu_pct"),
                  "achieved_warp_instr_per_s": (ck.get("instr_per_row") * rows / dp_s) if ck.get("instr_per_row") else None,
                  "peak_warp_instr_per_s": 4.0 * sms * clk_mhz * 1e6,
                  "note": "the ribbon kernels are bound by issue slots (FP32 FMUL/FFMA + shuffles), not by the MUFU pipe"},
        "traffic": traffic,
        "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (%s), scaled per lattice row to "
                         "this launch (one launch = one batch)" % ck.get("source")) if traffic else "no ncu capture for this kernel",
        "hbm": {"achieved": (ck.get("dram_bytes_per_row") * rows / dp_s / 1e9) if ck.get("dram_bytes_per_row") else None,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "note": "measured DRAM traffic of the capture (checkpoints, row headers, posterior records, signal) over the kernel time"},
    }
    line = {
        "metric": "dp_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_all, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "reads_per_s": reads_all / (ms_all * 1e-3),
        "config": {"workload": workload, "reads_per_step_per_gpu": int(n_reads), "batches_per_step": len(batches),
                   "cells_per_step": int(cells_all), "samples_per_step_per_gpu": n_samples, "reads_ok": int(ok_all),
                   "l2": "inputs larger than L2 (signal %.1f GB per step per GPU, all batches distinct and resident)" % (n_samples * 4 / 1e9)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if train:
        line["train"] = {"stats_bytes_allreduced": (3 * al.num_kmers + 4) * 8, "allreduce_ms": float(np.mean(ar_ms[-args.steps:])),
                         "mstep_ms": float(np.mean(mstep_ms[-args.steps:])), "em_iterations_timed": args.steps,
                         "collective": "torch.distributed all_reduce (NCCL) on the device tensor, in place" if world > 1 else "none (1 rank)"}
    if e2e:
        line["e2e"] = {"value": cells_all / (e2e_ms_all * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms_all,
                       "reads_per_s": reads_all / (e2e_ms_all * 1e-3),
                       "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                       "api": "dyn_align_submit / dyn_align_wait (Aligner.submit_packed / wait), two batches in flight; results fanned out into three rotating sets of host arrays",
                       "host_buffers": "pinned; %d of %d batches distinct on the host (cycled if fewer)" % (e2e["distinct_host_batches"], len(batches))}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
