"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once on golden-size reads.
usage: compute-sanitizer --tool memcheck python tools/sanitize_driver.py [quick]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from dynamont_b200 import Aligner
cases = {c.name: c for c in load_golden()}
quick = len(sys.argv) > 1
def run(tag, case, opts, train=True):
    al = Aligner(case.model_path, case.pore)
    for k, v in opts.items():
        al.set_option(k, v)
    sig = case.signal if not quick else case.signal[:1500]
    seq = case.sequence if not quick else case.sequence[:120]
    res = al.align_batch([sig.astype(np.float32)] * 3, [seq] * 3, True)
    ok = sum(not isinstance(r, Exception) for r in res)
    al.align(sig, seq, False)
    if train:
        al.train_batch([sig.astype(np.float32)], [seq])
    print(tag, "ok reads", ok, al.last_timing()["ribbon_reads"], flush=True)
band = cases["rna002_band"]
run("ribbon C=2", band, {})
run("ribbon C=2 two-level", band, {"rib_two_level": 1})
run("ribbon C=4", band, {"ribbon": 4})
run("full band v12 (uniform sigma, 8-warp CTAs)", band, {"ribbon": 0})
run("full band v13 (general)", cases["rna002_trained"], {"ribbon": 0})
run("log2 domain", cases["rna002_dinuc"], {"ribbon": 0, "arith": 1})
run("9-mer", cases["rna004_9mer"], {})
# async lanes
al = Aligner(band.model_path, band.pore)
b = ([band.signal.astype(np.float32)] * 2, [band.sequence] * 2)
print("lanes", len(list(al.align_stream([b, b, b], True))), flush=True)
if not quick:
    # NTK (resquiggle) first path
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_ntk_stages import load_ntk
    c = [x for x in load_ntk() if x.has_alignment][0]
    r = Aligner(c.model_path, c.pore, mode="resquiggle").align(c.signal, c.sequence, True)
    print("ntk segments", len(r["states"]), flush=True)
print("done")
