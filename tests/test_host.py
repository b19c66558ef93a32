"""Host-side logic: model TSV round trip, synthetic generator, cell counting, read sharding, streaming binary."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import build_emu  # noqa: E402


def test_model_tsv_round_trip(tmp_path, models_dir):
    from dynamont_b200.synth import kmer_strings, materialize_model, native_model, read_kmer_model, write_kmer_model
    path = materialize_model("rna002_5mer", models_dir)
    kmers, mean, sd = read_kmer_model(path)
    assert kmers == kmer_strings(5) and mean.size == 1024
    out = tmp_path / "m.model"
    write_kmer_model(str(out), kmers, mean, sd)
    assert open(path).read() == open(out).read()
    nm, _ = native_model(path, "rna002")   # RNA: index of the reversed kmer (aligner.cpp:136-141)
    assert nm[1] == mean[kmers.index("CAAAA")]
    nd, _ = native_model(path, "dna_r9")
    assert nd[1] == mean[kmers.index("AAAAC")]


def test_cells_match_oracle(models_dir):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    from oracle import Oracle
    path = materialize_model("rna002_5mer", models_dir)
    al = Aligner(path, "rna002", _lib_path=build_emu.build())
    orc = Oracle(path, "rna002")
    rng = np.random.default_rng(1)
    for _ in range(200):
        L = int(rng.integers(5, 3000))
        S = int(rng.integers(2 * (L - 4), 40 * (L - 4) + 5))
        assert al.read_cells(S, L) == orc.cells(S, L)


def test_sharding_is_a_balanced_partition():
    from dynamont_b200.parallel import shard_indices
    costs = np.random.default_rng(0).integers(1, 1000, size=1001)
    for world in (1, 2, 4, 8):
        parts = [shard_indices(costs, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(costs.size))
        loads = [costs[p].sum() for p in parts]
        assert max(loads) - min(loads) <= costs.max()
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_streaming_binary_protocol(tmp_path):
    """dynamont-NT-b200 linked against the emulator build: two lines in, one line out, order preserved."""
    lib = build_emu.build()
    exe = os.path.join(os.path.dirname(lib), "dynamont-NT-emu")
    src = os.path.join(ROOT, "dynamont_b200", "csrc", "stream_main.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, src, lib, "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    sig = ",".join(repr(float(v)) for v in case.signal)
    text = f"{sig}\n{case.sequence}\n1.0,2.0\nACGTACGTAC\n{sig}\n{case.sequence}\n"
    out = subprocess.run([exe, "-m", case.model_path, "-r", case.pore, "-p", "--batch", "2"], input=text,
                         capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(out) == 3 and out[0] == out[2]
    assert out[1] == "error:Signal too short compared to sequence"
    segs, z = out[0].split("\t")
    fields = [s.split(",") for s in segs.strip(";").split(";")]
    assert [int(f[0][1:]) for f in fields] == case.sequence_positions.tolist()
    assert [int(f[1]) for f in fields] == case.signal_positions.tolist()
    assert np.allclose([float(f[2]) for f in fields], case.probabilities, atol=1e-4)
    assert abs(float(z[2:]) - case.Z) < 1e-5 * abs(case.Z)


@pytest.mark.gpu
def test_streaming_binary_protocol_gpu():
    """the product binary (dynamont_b200/csrc/dynamont-NT-b200, linked against the CUDA library): same protocol, same answers"""
    exe = os.path.join(ROOT, "dynamont_b200", "csrc", "dynamont-NT-b200")
    assert os.path.exists(exe), "dynamont-NT-b200 is built by dynamont_b200.build (graft entry build())"
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    sig = ",".join(repr(float(v)) for v in case.signal)
    text = f"{sig}\n{case.sequence}\n1.0,2.0\nACGTACGTAC\n{sig}\n{case.sequence}\n"
    out = subprocess.run([exe, "-m", case.model_path, "-r", case.pore, "-p", "--batch", "2"], input=text,
                         capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(out) == 3 and out[0] == out[2]
    assert out[1] == "error:Signal too short compared to sequence"
    segs, z = out[0].split("\t")
    fields = [s.split(",") for s in segs.strip(";").split(";")]
    assert [int(f[0][1:]) for f in fields] == case.sequence_positions.tolist()
    assert [int(f[1]) for f in fields] == case.signal_positions.tolist()
    assert np.allclose([float(f[2]) for f in fields], case.probabilities, atol=1e-4)
    assert abs(float(z[2:]) - case.Z) < 1e-5 * abs(case.Z)
