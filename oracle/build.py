"""TEST INFRASTRUCTURE — build recipe for the CPU oracles (not the product).

* ``oracle/_build/liboracle_nt.so``  — the plain-C restatement (``nt_oracle.c``); always buildable.
* ``oracle/_ref/libdynamont_ref.so`` — the UNMODIFIED reference C++ compiled from the sources where
  they lie under ``/root/reference`` (aligner.cpp, NT_aligner_api.cpp, NTK_aligner_api.cpp) together
  with our harness ``ref_shim.cpp``.  Flags follow the reference's own Release build
  (CMakeLists.txt:38-45: ``-O3``, C++17, no ``-march``, no fast-math).  The reference's CMake build is
  not run; no reference source is copied into this repo.  Built only where ``/root/reference`` exists
  (the dev container); the resulting ``.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("DYNAMONT_REFERENCE", "/root/reference")
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(BUILD_DIR, "liboracle_nt.so")
REF_SO = os.path.join(REF_DIR, "libdynamont_ref.so")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "nt_oracle.c")
    if not force and _newer(ORACLE_SO, [src]):
        return ORACLE_SO
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-D_GNU_SOURCE",
           "-o", ORACLE_SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return ORACLE_SO


def reference_sources() -> list[str]:
    return [os.path.join(REFERENCE, "src", "cpp", f)
            for f in ("aligner.cpp", "NT_aligner_api.cpp", "NTK_aligner_api.cpp")]


def build_reference(force: bool = False) -> str | None:
    """Returns the path of the reference library, or None if it is neither buildable nor prebuilt."""
    srcs = reference_sources()
    shim = os.path.join(HERE, "ref_shim.cpp")
    if not all(os.path.exists(s) for s in srcs):
        return REF_SO if os.path.exists(REF_SO) else None  # GPU box: prebuilt file only
    if not force and _newer(REF_SO, srcs + [shim]):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-DNDEBUG",
           "-I", os.path.join(REFERENCE, "include"), "-o", REF_SO, shim] + srcs
    subprocess.run(cmd, check=True)
    return REF_SO


REF_NTKFIX_SO = os.path.join(REF_DIR, "libdynamont_ref_ntkfix.so")


def build_reference_ntkfix(force: bool = False) -> str | None:
    """The reference with the two-line repair of resquiggle (NTK) mode described in SURVEY.md F2 / 8c: as shipped,
    ``NTKAligner::logF`` / ``logB`` (NTK_aligner_api.cpp:443-607) write their results into the sparse map only after
    the loop, so every predecessor lookup inside the loop reads -inf and every input throws.  The patch makes each
    result visible immediately (``forAPSEI[tnk] = computed[idx];`` after :508, ``backAPSEI[tnk] = computed[idx];``
    after :600) — the evident intent, and what ``decodeMAP`` (:862) already does.  The patched translation unit is
    generated in a temporary directory and deleted; only the ``.so`` lands in ``oracle/_ref/``."""
    import shutil
    import tempfile
    srcs = reference_sources()
    shim = os.path.join(HERE, "ref_shim.cpp")
    if not all(os.path.exists(s) for s in srcs):
        return REF_NTKFIX_SO if os.path.exists(REF_NTKFIX_SO) else None
    if not force and _newer(REF_NTKFIX_SO, srcs + [shim]):
        return REF_NTKFIX_SO
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="dyn_ntkfix_")
    try:
        text = open(srcs[2]).read()
        marker = "computed[idx] = {a, p, s, e, i};"
        parts = text.split(marker)
        if len(parts) != 3:
            raise RuntimeError("NTK_aligner_api.cpp does not look like the surveyed snapshot (marker count %d)" % (len(parts) - 1))
        text = (parts[0] + marker + "\n\t\tforAPSEI[tnk] = computed[idx];" + parts[1] + marker +
                "\n\t\tbackAPSEI[tnk] = computed[idx];" + parts[2])
        patched = os.path.join(tmp, "NTK_aligner_api_ntkfix.cpp")
        with open(patched, "w") as fh:
            fh.write(text)
        cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-DNDEBUG",
               "-I", os.path.join(REFERENCE, "include"), "-o", REF_NTKFIX_SO, shim, srcs[0], srcs[1], patched]
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return REF_NTKFIX_SO


def build_all(force: bool = False) -> dict:
    return {"oracle": build_oracle(force), "reference": build_reference(force),
            "reference_ntkfix": build_reference_ntkfix(force)}


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
