"""Distil the counters bench.py's roofline block cites from an Nsight Compute capture of a ribbon kernel.

usage: python tools/ncu_counters.py REP.ncu-rep PLAIN.json KEY [OUT.json]
  REP    ncu --set full capture of one k_ribbon launch (tools/gpu_prof_rib.sh)
  PLAIN  the bench.py line of the same command run without ncu (gives the lattice rows of the launch)
  KEY    "align" or "train"
Merges {KEY: {...}} into OUT (default profiles/r2_ribbon_counters.json)."""
import csv, io, json, os, subprocess, sys
rep, plain, key = sys.argv[1:4]
out_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ribbon_counters.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
line = json.loads(open(plain).read().strip().splitlines()[-1])
lattice_rows = float(line["config"]["samples_per_step_per_gpu"]) / max(1, line["config"]["batches_per_step"])
f = lambda k: float(d[k].replace(",", ""))
unit = dict(zip(rows[0], rows[1]))
def to_bytes(k):
    v, u = f(k), unit[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
entry = {
    "source": "profiles/" + {"prof_r2m_align": "r2m_k_ribbon_align_full.md", "prof_r2m_train": "r2m_k_ribbon_train_full.md",
                             "prof_r2f_align": "r2f_k_ribbon_align_full.md", "prof_r2f_train": "r2f_k_ribbon_train_full.md"}.get(os.path.basename(rep).replace(".ncu-rep", ""), os.path.basename(rep)),
    "kernel": d["Kernel Name"],
    "lattice_rows_in_capture": lattice_rows,
    "kernel_ms_in_capture": f("gpu__time_duration.sum") * ({"ms": 1.0, "msecond": 1.0, "us": 1e-3, "usecond": 1e-3, "s": 1e3, "second": 1e3, "ns": 1e-6, "nsecond": 1e-6}[unit["gpu__time_duration.sum"].lower()]),
    "instr_per_row": f("smsp__inst_executed.sum") / lattice_rows,
    "issue_busy_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "xu_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed"),
    "fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "warps_per_sm": f("sm__warps_active.avg.per_cycle_active"),
    "registers": f("launch__registers_per_thread"),
    "dram_bytes_per_row": (to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")) / lattice_rows,
    "dram_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
}
cur = {}
if os.path.exists(out_path):
    cur = json.load(open(out_path))
cur[key] = entry
json.dump(cur, open(out_path, "w"), indent=1)
print(json.dumps(entry, indent=1))
