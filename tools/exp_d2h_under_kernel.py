"""experiment: how long does a pinned D2H copy on a side stream take while the ribbon kernel saturates the GPU?"""
import os, sys, threading, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dynamont_b200 import Aligner
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["c2"]
model_path, d_sig, d_bas, so, qo = bench.gen_reads_torch(cfg, 20000, 123, dev)
al = Aligner(model_path, cfg[0], device=0)
for opt in sys.argv[1:]:
    k, v = opt.split("="); al.set_option(k, float(v))
al.align_packed(d_sig.data_ptr(), so, d_bas.data_ptr(), qo, True, device=True)
src = torch.empty(660 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
dst = torch.empty(src.numel(), dtype=torch.float32, pin_memory=True)
side = torch.cuda.Stream()
def copy_ms():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record(); dst.copy_(src, non_blocking=True); e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)
print("idle GPU: D2H 660 MB pinned %.1f ms" % copy_ms())
th = threading.Thread(target=lambda: al.align_packed(d_sig.data_ptr(), so, d_bas.data_ptr(), qo, True, device=True))
th.start()
time.sleep(0.08)
t = [copy_ms() for _ in range(3)]
th.join()
print("during the ribbon kernel:", ["%.1f" % x for x in t], "kernel ms", al.last_timing()["dp_ms"])
# H2D during kernel
hsrc = torch.empty(src.numel(), dtype=torch.float32, pin_memory=True)
def h2d_ms():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record(); src.copy_(hsrc, non_blocking=True); e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)
th = threading.Thread(target=lambda: al.align_packed(d_sig.data_ptr(), so, d_bas.data_ptr(), qo, True, device=True))
th.start()
time.sleep(0.08)
t = [h2d_ms() for _ in range(3)]
th.join()
print("H2D during the ribbon kernel:", ["%.1f" % x for x in t])
