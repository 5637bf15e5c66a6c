#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the GPU box (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

    python oracle/stage_ref.py            # build container: /root/reference -> oracle/_ref/*.zip

/root/reference does not exist on the GPU box and the reference is pure Python (nothing to
compile), so the "build" of the real-reference oracle is an archive of the sources it needs for
this path -- run.py and common/*.py, byte for byte -- written to `oracle/_ref/videopose3d_ref.zip`.
`oracle/_ref/` is git-ignored (no reference source ever enters the history) but not
gpurun-ignored, so the archive travels with the snapshot like a built .so does.  `reference_dir()`
hands tests/, bench.py's reference arms and tools/run_reference.py a directory to import from:
/root/reference when it exists, otherwise the archive unpacked into a per-user temp directory.
Nothing under videopose3d_b200/ imports this module.
"""
import hashlib
import os
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = "/root/reference"
ARCHIVE = os.path.join(HERE, "_ref", "videopose3d_ref.zip")
FILES = ["run.py", "LICENSE"]
DIRS = ["common"]


def _members():
    out = []
    for f in FILES:
        if os.path.exists(os.path.join(SOURCE, f)):
            out.append(f)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SOURCE, d))):
            if name.endswith(".py"):
                out.append(os.path.join(d, name))
    return out


def stage():
    """Write the archive (deterministic: fixed timestamps, sorted members).  Returns its path, or
    None when the reference checkout is not present (e.g. on the GPU box)."""
    if not os.path.exists(os.path.join(SOURCE, "run.py")):
        return None
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in _members():
            info = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))
            info.compress_type = zipfile.ZIP_DEFLATED
            info.external_attr = 0o644 << 16
            with open(os.path.join(SOURCE, rel), "rb") as f:
                z.writestr(info, f.read())
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


def reference_dir():
    """Directory holding the unmodified reference (run.py, common/), or None."""
    if os.path.exists(os.path.join(SOURCE, "run.py")):
        return SOURCE
    if not os.path.exists(ARCHIVE):
        return None
    with open(ARCHIVE, "rb") as f:
        tag = hashlib.sha1(f.read()).hexdigest()[:12]
    dst = os.path.join(tempfile.gettempdir(), f"vp3d_ref_{os.getuid()}_{tag}")
    if not os.path.exists(os.path.join(dst, "run.py")):
        part = dst + f".part{os.getpid()}"
        with zipfile.ZipFile(ARCHIVE) as z:
            z.extractall(part)
        try:
            os.rename(part, dst)
        except OSError:  # another process won the race
            import shutil
            shutil.rmtree(part, ignore_errors=True)
    return dst


def import_reference():
    """Put the reference on sys.path and return its `common.model` module (or None)."""
    d = reference_dir()
    if d is None:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import common.model as ref_model
    return ref_model


if __name__ == "__main__":
    p = stage()
    print(p if p else "reference checkout not found; nothing staged")
