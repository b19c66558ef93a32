"""Generate tests/golden/frontend_golden.npz: known answers for the front-end stages either side of the DP
(SURVEY.md 8f N1 / N2), produced by EXECUTING the reference's own Python source where it lies under /root/reference
(`hampel`, `_decode_native_state`, `segmentation_to_string` of src/dynamont/segmentation/utils.py; the module itself
cannot be imported here because it pulls in matplotlib / seaborn / the native extension).  Nothing is copied into this
repository: the function texts are sliced out of the file at run time.  Run:  python tools/make_golden_frontend.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
SRC = os.path.join(os.environ.get("DYNAMONT_REFERENCE", "/root/reference"), "src", "dynamont", "segmentation", "utils.py")


def reference_functions():
    text = open(SRC).read()
    ns = {"np": np}
    for name in ("hampel", "_decode_native_state", "segmentation_to_string"):
        m = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, text, re.S | re.M)
        exec(m.group(0), ns)
    return ns


if __name__ == "__main__":
    ns = reference_functions()
    rng = np.random.default_rng(5)
    out = {}
    # ---- N1: (raw - shift) / scale + hampel, the two parameter sets the front end uses + the reference's own tests
    cases = []
    for n, window, nsig in ((7, 3, 3.0), (8, 3, 3.0), (3, 3, 3.0), (2, 3, 3.0), (500, 3, 3.0), (500, 7, 5.0), (64, 7, 5.0), (9, 4, 3.0)):
        raw = rng.normal(90.0, 12.0, n).astype(np.float32)
        raw[rng.random(n) < 0.05] += 80.0
        cases.append((raw, 88.5, 11.25, window, nsig))
    cases.append((np.array([1, 1, 1, 10, 1, 1, 1], dtype=np.float32), 0.0, 1.0, 3, 3.0))                 # tests/test_utils.py:7-14
    cases.append((np.array([1.0, 1.0, 50.0, 1.0, 1.0, 1.0, 75.0, 1.0], dtype=np.float32), 0.0, 1.0, 3, 3.0))  # tests/test_segment.py:194
    out["n_pre"] = np.array(len(cases))
    for i, (raw, shift, scale, window, nsig) in enumerate(cases):
        sig = np.array(raw, dtype=np.float64, copy=True)   # segment.py:146-153
        sig -= shift
        sig /= scale
        ns["hampel"](sig, window, nsig)
        out["pre%d/raw" % i] = raw
        out["pre%d/params" % i] = np.array([shift, scale, window, nsig])
        out["pre%d/expected" % i] = sig
    # ---- N2: CSV lines
    fmt = []
    for rna, k, with_polish in ((False, 5, False), (True, 5, False), (True, 9, True)):
        read = "".join(rng.choice(list("ACGT"), 40))
        nseg = 40 - k + 1
        res = {"sequence_positions": np.arange(nseg, dtype=np.uint64) + k // 2,
               "signal_positions": np.cumsum(rng.integers(2, 30, nseg)).astype(np.uint64) - 2,
               "probabilities": rng.random(nseg), "states": ["M" if rng.random() < 0.8 else "P" for _ in range(nseg)],
               "polishes": (["".join(rng.choice(list("ACGT"), k)) if rng.random() < 0.7 else "" for _ in range(nseg)]
                            if with_polish else [""] * nseg)}
        txt = ns["segmentation_to_string"](res, "read-%d" % len(fmt), "sig-%d" % len(fmt), 1234, 99999, read, k, rna)
        fmt.append((res, read, k, rna, txt))
    out["n_fmt"] = np.array(len(fmt))
    for i, (res, read, k, rna, txt) in enumerate(fmt):
        for key in ("sequence_positions", "signal_positions", "probabilities"):
            out["fmt%d/%s" % (i, key)] = res[key]
        out["fmt%d/states" % i] = np.array(res["states"])
        out["fmt%d/polishes" % i] = np.array(res["polishes"])
        out["fmt%d/read" % i] = np.array(read)
        out["fmt%d/params" % i] = np.array([k, int(rna)])
        out["fmt%d/expected" % i] = np.frombuffer(txt, dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "frontend_golden.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes;", len(cases), "preprocessing cases,", len(fmt), "formatting cases")
