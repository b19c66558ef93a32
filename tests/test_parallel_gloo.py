"""N > 1 path on CPU: two gloo ranks shard a read set, run the (emulated) training kernels on their shards,
all-reduce the pooled sufficient statistics and must agree with a single-process run."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import MODELS_DIR, ROOT

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))


def _reads():
    from dynamont_b200.synth import materialize_model, native_model, synth_read
    path = materialize_model("rna002_5mer", MODELS_DIR)
    nm, ns = native_model(path, "rna002")
    rng = np.random.default_rng(5)
    sigs, seqs = [], []
    for L in (30, 60, 45, 80, 25, 70):
        s, q, _ = synth_read(rng, nm, ns, 5, L, 6)
        sigs.append(s.astype(np.float32))
        seqs.append(q)
    return path, sigs, seqs


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import torch.distributed as dist
    from dynamont_b200 import Aligner
    from dynamont_b200.train import PooledTrainer
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    path, sigs, seqs = _reads()
    al = Aligner(path, "rna002", _lib_path=build_emu.build())
    mean, sd, trans, stats = PooledTrainer(al, rank, world).iteration(sigs, seqs)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), mean=mean, sd=sd, w=stats["w"], n=stats["n"], m1=trans["m1"])
    # the same iteration with the statistics kept in "device" memory (the emulator's device memory is host memory) and
    # all-reduced in place: dyn_train_accumulate -> all_reduce(tensor) -> dyn_train_mstep_device
    al2 = Aligner(path, "rna002", _lib_path=build_emu.build())
    trans2, st = PooledTrainer(al2, rank, world).iteration_device(sigs, seqs)
    mean2, sd2 = al2.model()
    K = al2.num_kmers
    np.savez(os.path.join(out_dir, f"d{rank}.npz"), mean=mean2, sd=sd2, w=st[:K].numpy(), n=float(st[3 * K + 3]), m1=trans2["m1"])
    dist.destroy_process_group()


def test_two_ranks_agree_with_one(tmp_path):
    import build_emu
    from dynamont_b200 import Aligner
    from dynamont_b200.train import PooledTrainer
    build_emu.build()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    path, sigs, seqs = _reads()
    al = Aligner(path, "rna002", _lib_path=build_emu.build())
    mean, sd, trans, stats = PooledTrainer(al).iteration(sigs, seqs)
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert z["n"] == len(sigs) == stats["n"]
        np.testing.assert_allclose(z["w"], stats["w"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(z["mean"], mean, rtol=1e-12)
        np.testing.assert_allclose(z["sd"], sd, rtol=1e-9)
        assert abs(z["m1"] - trans["m1"]) < 1e-12
        d = np.load(tmp_path / f"d{r}.npz")
        assert d["n"] == len(sigs)
        np.testing.assert_allclose(d["w"], stats["w"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(d["mean"], mean, rtol=1e-12)
        np.testing.assert_allclose(d["sd"], sd, rtol=1e-9)
        assert abs(d["m1"] - trans["m1"]) < 1e-12
