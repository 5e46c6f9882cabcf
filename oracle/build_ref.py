"""Recipe for oracle/_ref/: the UNMODIFIED reference package, as one archive.

TEST / BENCH INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, but
git-ignored build outputs travel there with the repo snapshot (like libb2rl.so).  This
recipe zips the reference's own `pfrl/` package, byte for byte, into
`oracle/_ref/pfrl_ref.zip` (git-ignored, never committed); `oracle/refimport.py` imports
it with zipimport.  `bench.py --impl reference` and the `cpu_baseline` leg then time the
reference itself on the box's host cores (cpu_baseline.kind = "reference") instead of
the pure-Python port.

    python oracle/build_ref.py          (run by __graft_entry__.build())
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PFRL_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "pfrl_ref.zip")


def build(force=False):
    src = os.path.join(REFERENCE_ROOT, "pfrl")
    if not os.path.isdir(src):
        return None  # GPU box: use the prebuilt archive if it travelled
    if os.path.exists(OUT) and not force:
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = OUT + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for d, _, files in sorted(os.walk(src)):
            for f in sorted(files):
                if f.endswith(".py"):
                    p = os.path.join(d, f)
                    z.write(p, os.path.relpath(p, REFERENCE_ROOT))
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
