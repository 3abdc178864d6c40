"""Recipe: stage the UNMODIFIED reference (pvjosue/CWFA, /root/reference) under oracle/_ref/ so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is ~8 kLoC of pure Python (10 top-level files + its vendored FrEIA): nothing to
compile -- the "build" is a verbatim copy of the .py files from where they lie (plus LICENSE / README for attribution).
oracle/_ref/ is listed in .gitignore (reference sources never enter this repository's history) but NOT in .gpurunignore, so the
copy is shipped with the snapshot like the built .so files.  Used by
  * bench.py --impl reference  and the cpu_baseline leg  (kind "reference": the actual networks.py / FrEIA timed on host cores),
  * tests/test_gpu_reference_swap.py (the reference's own conditional_wavelet_flow driven with the import swap of INTEGRATION.md).

    python oracle/make_ref.py            # no-op when /root/reference is absent (the GPU box uses the staged copy)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CWFA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_TOP = {"LICENSE", "README.md", "requirements.txt"}


def stage(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "networks.py")):
        if verbose:
            print(f"make_ref: {SRC} not present; keeping {DST if os.path.isdir(DST) else 'nothing'}")
        return os.path.isfile(os.path.join(DST, "networks.py"))
    shutil.rmtree(DST, ignore_errors=True)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "images")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if f.endswith(".py") or (rel == "." and f in KEEP_TOP):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                n += 1
    if verbose:
        print(f"make_ref: staged {n} files from {SRC} into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
