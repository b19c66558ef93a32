"""Pack the reference's shipped 5-mer pore-model tables (data, not code) into models/pore_tables.npz.

Run once in the dev container (needs /root/reference).  The GPU box has no /root/reference, so tests and
bench.py re-materialise ``.model`` TSV files from this archive (dynamont_b200.synth.materialize_model) —
text identical to the shipped files because values are written with Python ``str(float)`` which is how the
reference itself writes models (utils.py:147-152).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from dynamont_b200.synth import read_kmer_model  # noqa: E402

REF = "/root/reference/models"
FILES = {
    "rna002_5mer": "rna/rna002/rna002_5mer.model",
    "trained_rna002_5mer": "rna/rna002/trained_rna002_5mer.model",
    "rna004_5mer": "rna/rna004/rna004_5mer.model",
}
out = {}
for name, rel in FILES.items():
    kmers, mean, sd = read_kmer_model(os.path.join(REF, rel))
    assert kmers == sorted(kmers) and len(kmers) == 1024, name
    out[name + "_mean"] = mean
    out[name + "_stdv"] = sd
np.savez_compressed(os.path.join(os.path.dirname(__file__), "..", "models", "pore_tables.npz"), **out)
print({k: v.shape for k, v in out.items()})
