"""dynamont-train data-parallel step (BASELINE config 5 shape): every rank runs the training kernels on its shard of
synthetic rna004 9-mer reads, ONE NCCL all-reduce sums the pooled sufficient statistics (3*4^9 + 4 doubles), the M-step
is replicated.  Launch:  python -m torch.distributed.run --nproc-per-node N tools/train_timing.py [reads_per_rank [max_len]]
(or plain python for one GPU).  Prints one line from rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.parallel import allreduce_stats  # noqa: E402
from dynamont_b200.synth import materialize_model, native_model, synth_read  # noqa: E402
from dynamont_b200.train import m_step  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 2000  # BASELINE config 5 (as config 2): 5000
path = materialize_model("synthetic_rna004_9mer", os.path.join(ROOT, "tests", "golden", "_models"))
nm, ns = native_model(path, "rna004")
rng = np.random.default_rng(20265000 + rank)
sigs, seqs = [], []
for _ in range(n):
    s, q, _ = synth_read(rng, nm, ns, 9, int(rng.integers(500, max_len + 1)), 30.0)
    sigs.append(s.astype(np.float32))
    seqs.append(q)
al = Aligner(path, "rna004", device=local)
cells = sum(al.read_cells(len(s), len(q)) for s, q in zip(sigs, seqs))
al.train_batch(sigs[:64], seqs[:64])  # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res, pooled = al.train_batch(sigs, seqs)
t1 = time.perf_counter()
ok = [r for r in res if not isinstance(r, Exception)]
pooled["Z"] = float(sum(r["Z"] for r in ok))
pooled["n"] = float(len(ok))
pooled = allreduce_stats(pooled, torch.device("cuda", local))
torch.cuda.synchronize()
t2 = time.perf_counter()
mean0, sd0 = al.model()
mean, sd, trans = m_step(pooled, mean0, sd0)
t3 = time.perf_counter()
times = torch.tensor([t1 - t0, t2 - t1, t3 - t2, float(np.abs(mean).sum())], dtype=torch.float64, device="cuda")
if world > 1:
    tmax = times.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = times.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    agree = abs(float(tmax[3]) - float(tmin[3])) == 0.0
else:
    tmax, agree = times, True
if rank == 0:
    tk, ta, tm = [float(v) for v in tmax[:3]]
    touched = int((pooled["w"] > 0).sum())
    print("train step: %d GPUs x %d reads, %.3g cells/GPU: kernels+marshalling %.2f s (%.1f GCUPS, %.0f reads/s aggregate), "
          "all-reduce of %.1f MB %.1f ms, M-step %.0f ms; reads ok %d; kmers touched %d; m1 %.6f e2 %.6f; ranks agree: %s; "
          "log2-domain fallback reads (rank 0): %d" % (
              world, n, cells, tk, world * cells / tk / 1e9, world * n / tk, (3 * 4 ** 9 + 4) * 8 / 1e6, ta * 1e3, tm * 1e3,
              int(pooled["n"]), touched, trans["m1"], trans["e2"], agree, al.last_timing()["log2_fallback_reads"]))
if world > 1:
    dist.destroy_process_group()
