"""Recipe for oracle/_ref/: the UNMODIFIED reference env, materialised so that it can be timed on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY - never imported by alphatank_b200/.

/root/reference exists in the build container only.  The reference arm of bench.py (`--impl reference`,
BASELINE.md section 3) has to time the reference's own Python env on the GPU box's host cores, so this recipe copies
the two packages that ARE the environment step -

    /root/reference/env/      (*.py, bots/*.py; the render-only assets/ directory is NOT copied)
    /root/reference/configs/  (*.py)

- byte for byte into oracle/_ref/reference/.  oracle/_ref/ is listed in .gitignore (nothing of the reference enters
the repository's history) and not in .gpurunignore (it travels to the GPU box like the built .so files).  The
pygame / gym stand-ins the reference needs to import headless are the repo's own tests/refstubs.

    python -m oracle.make_ref            # or: __graft_entry__.build()
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ALPHATANK_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
PACKAGES = ("env", "configs")


def _py_files(root):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x not in ("assets", "__pycache__"))
        for f in sorted(files):
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, f), root))
    return out


def available():
    return os.path.isfile(os.path.join(DST, "env", "gym_env_multi.py"))


def build(force=False):
    """Copy the reference's env/ and configs/ python sources into oracle/_ref/reference/ when /root/reference is
    present (the build container); on the GPU box the prebuilt copy is used as it is.  Returns the directory or None."""
    if not os.path.isdir(os.path.join(REF_SRC, "env")):
        return DST if available() else None
    manifest = {}
    for pkg in PACKAGES:
        src = os.path.join(REF_SRC, pkg)
        for rel in _py_files(src):
            s, d = os.path.join(src, rel), os.path.join(DST, pkg, rel)
            data = open(s, "rb").read()
            manifest["%s/%s" % (pkg, rel)] = hashlib.sha256(data).hexdigest()
            if force or not os.path.exists(d) or open(d, "rb").read() != data:
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    d = build(force="--force" in sys.argv)
    print("oracle/_ref: %s" % (d or "reference not present, nothing built"))
