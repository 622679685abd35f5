"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference Python sources the CPU arm of bench.py
runs (`--impl reference`, `cpu_baseline.kind = "reference"`) into oracle/_ref/.

    python oracle/stage_reference.py          (also called by __graft_entry__.build() when /root/reference exists)

/root/reference does not exist on the GPU box; oracle/_ref/ is git-ignored (the reference's sources never enter this
repository's history) but NOT gpurun-ignored, so the staged copy travels to the box exactly like the built .so does.
Nothing under missm-benchmark_b200/ imports it.  What is staged: the two packages the hot path lives in
(`languagebind/`, `src/`, *.py only), byte for byte; oracle/ref_shim.py supplies the third-party names they import
that this image lacks."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PACKAGES = ("languagebind", "src")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"stage_reference: {SRC} not present; keeping {DST} as it is")
        return os.path.isdir(os.path.join(DST, "languagebind"))
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(root, f)
                d = os.path.join(DST, os.path.relpath(s, SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(f"{SRC} ({n} .py files of {', '.join(PACKAGES)}), unmodified; staged by oracle/stage_reference.py\n")
    if verbose:
        print(f"stage_reference: {n} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
