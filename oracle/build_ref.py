"""Stage the UNMODIFIED reference modules for the GPU box.

TEST / BENCH INFRASTRUCTURE.  The upstream checkout (/root/reference) exists only in the authoring container; the
reference is pure Python (no build system, nothing to compile), so "building" it means copying the two importable
modules of the hot path -- SRFR_model.py and model.py -- verbatim into oracle/_ref/.  That directory is git-ignored (the
sources never enter this repository's history) but not gpurun-ignored, so it travels to the GPU box with the snapshot
like a built .so would.  bench.py --impl reference then times the real reference module on the box's host cores
(cpu_baseline.kind = "reference"); when oracle/_ref is absent it falls back to the oracle port (kind = "port").

    python oracle/build_ref.py      # also run by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRFRD_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("SRFR_model.py", "model.py", "LICENSE")


def build_ref() -> bool:
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        src = os.path.join(REF, f)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(DST, f))
    return True


def load_ref():
    """Import the staged reference modules (None if they were never staged)."""
    if not os.path.exists(os.path.join(DST, "SRFR_model.py")):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("srfrd_reference_SRFR_model", os.path.join(DST, "SRFR_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("staged" if build_ref() else f"{REF} not present: nothing staged", file=sys.stderr)
