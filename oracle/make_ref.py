"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference available on the GPU box.

`/root/reference` exists only in the build container.  The reference's hot path is pure Python (no build step), so the
"build" of the reference arm is a file copy: this script copies the reference's own `modules/*.py` and `util.py`, untouched,
into `oracle/_ref/` -- git-ignored (never part of the repository's history), NOT gpurun-ignored (it travels to the GPU box
with the snapshot, like the compiled `.so`).  `oracle/ref_harness.py` then imports the reference from there with the six
shims of SURVEY 8(c), and `bench.py --impl reference` / the `cpu_baseline` leg time the reference's own nn.Module
(`kind: "reference"`) instead of the oracle port.  Run by `__graft_entry__.build()` when `/root/reference` is present.

    python oracle/make_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make(src="/root/reference"):
    if not os.path.isfile(os.path.join(src, "modules", "modeling.py")):
        return None
    os.makedirs(os.path.join(DST, "modules"), exist_ok=True)
    n = 0
    for f in sorted(os.listdir(os.path.join(src, "modules"))):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(src, "modules", f), os.path.join(DST, "modules", f))
            n += 1
    shutil.copyfile(os.path.join(src, "util.py"), os.path.join(DST, "util.py"))
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write("unmodified copy of %s/modules/*.py and util.py made by oracle/make_ref.py (%d files); not tracked by git\n" % (src, n + 1))
    return DST


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
