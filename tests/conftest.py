import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS_DIR = os.path.join(ROOT, "tests", "golden", "_models")
GOLDEN = os.path.join(ROOT, "tests", "golden", "nt_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing at dyn_create"""
    if have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU path; run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenCase:
    def __init__(self, z, name):
        g = lambda k: z[name + "/" + k]  # noqa: E731
        self.name = name
        self.pore = str(g("pore"))
        self.model_name = str(g("model"))
        self.signal = g("signal").astype(np.float64)
        self.sequence = str(g("sequence"))
        self.Z = float(g("Z"))
        self.Zf = float(g("Zf"))
        self.sequence_positions = g("sequence_positions")
        self.signal_positions = g("signal_positions")
        self.probabilities = g("probabilities")
        self.train_Z = float(g("train_Z"))
        self.train_trans = g("train_trans")
        self.train_kmers = g("train_kmers")
        self.train_mean = g("train_mean")
        self.train_stdev = g("train_stdev")
        self.stat_w, self.stat_x, self.stat_xx = g("stat_w"), g("stat_x"), g("stat_xx")

    @property
    def model_path(self):
        from dynamont_b200.synth import materialize_model
        return materialize_model(self.model_name, MODELS_DIR)


def load_golden():
    with np.load(GOLDEN) as z:
        return [GoldenCase(z, str(n)) for n in z["names"]]


@pytest.fixture(scope="session")
def golden_cases():
    return load_golden()


@pytest.fixture(scope="session")
def models_dir():
    os.makedirs(MODELS_DIR, exist_ok=True)
    return MODELS_DIR


def have_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# parity gates from BASELINE.json north_star
BORDER_MIN_IDENTICAL = 0.999
PROB_ATOL = 1e-4
TRAIN_RTOL = 1e-4
Z_RTOL = 1e-6


def check_alignment(ours: dict, ref_sigpos, ref_seqpos, ref_prob, ref_Z, what=""):
    assert np.array_equal(ours["sequence_positions"], ref_seqpos), what + ": sequence positions differ"
    same = ours["signal_positions"] == ref_sigpos
    frac = same.mean() if same.size else 1.0
    assert frac >= BORDER_MIN_IDENTICAL, f"{what}: only {same.sum()}/{same.size} borders identical"
    dp = np.abs(ours["probabilities"] - ref_prob)
    # a moved border legitimately changes the two adjacent segment medians; compare untouched segments
    ok = same.copy()
    ok[:-1] &= same[1:]
    assert dp[ok].max(initial=0.0) <= PROB_ATOL, f"{what}: max |dp| = {dp[ok].max():.3e}"
    assert abs(ours["Z"] - ref_Z) <= Z_RTOL * max(1.0, abs(ref_Z)), f"{what}: Z {ours['Z']} vs {ref_Z}"
    return frac, float(dp[ok].max(initial=0.0))
