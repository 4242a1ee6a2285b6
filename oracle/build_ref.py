"""Recipe for `oracle/_ref/`: the reference's OWN Python modules for the hot path, copied unmodified from
/root/reference (build container only) so that they travel to the GPU box with the repo snapshot - `oracle/_ref/` is
git-ignored (no reference source enters the history) but not gpurun-ignored.

    python -m oracle.build_ref        (also run by __graft_entry__.build() when /root/reference exists)

With it present, `oracle/ref_shim.py` imports `cldm.cldm` / `ldm.*` from here when /root/reference does not exist,
and `bench.py --impl reference` / `cpu_baseline` time the reference's own ControlNet / ControlledUnetModel modules
(`kind: "reference"`) instead of the oracle port.  Test infrastructure only - nothing under editanything_b200/
imports it."""
import os
import shutil
import sys

SRC = os.environ.get("EA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGES = ("ldm", "cldm")          # pure-Python packages the step networks live in


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "cldm")):
        return None
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, pkg), dst, ignore=shutil.ignore_patterns("*.pyc", "__pycache__", "*.ipynb"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copies of the reference's ldm/ and cldm/ packages, made by oracle/build_ref.py.\n"
                "Git-ignored: not part of this repository's history.\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"oracle/_ref: {n} files from {SRC}")
    return DST


if __name__ == "__main__":
    if build(verbose=True) is None:
        sys.exit("reference tree not present: nothing to do")
