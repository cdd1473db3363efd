#!/usr/bin/env python
"""Recipe that lets the UNMODIFIED reference travel to the GPU box (which has no /root/reference).

TEST INFRASTRUCTURE ONLY.  Packs the reference's Python sources (pure Python, no native code) and its configs, byte for
byte and with their directory layout, from `/root/reference` into ONE git-ignored archive `oracle/_ref/reference_src.tar.gz`,
and copies the two shipped checkpoints into `oracle/_ref/ckpt/`.  `oracle/_ref/` is listed in `.gitignore` (nothing of the
reference enters this repository's history or its source tree) but not in `.gpurunignore`, so the archive rides along with
the snapshot like built `.so` files do.

`__graft_entry__.build()` calls `stash()` whenever `/root/reference` is present; on the GPU box that is a no-op and
`unpack()` extracts the archive into a scratch directory outside the repository, which `oracle/reference_shim.py` then uses
as the reference root.  Consumers: the `-m gpu` parity-at-scale test (attribution of outlier pixels by re-running the
reference on them), `bench.py --impl reference` and the bench line's `cpu_baseline` (kind "reference").
"""
from __future__ import annotations

import io
import os
import shutil
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
ARCHIVE = os.path.join(HERE, "_ref", "reference_src.tar.gz")
CKPT = os.path.join(HERE, "_ref", "ckpt")
KEEP_EXT = (".py", ".yml", ".yaml")
SKIP_DIRS = {"outputs", "images", "__pycache__", ".git"}
MARKER = os.path.join("modeling", "layered_rfrender.py")


def _members():
    out = []
    for root, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith(KEEP_EXT):
                p = os.path.join(root, f)
                out.append((os.path.relpath(p, SRC), p))
    return out


def stash(force: bool = False) -> int:
    """Returns the number of files packed / copied (0 when the reference is absent or everything is already in place)."""
    if not os.path.isfile(os.path.join(SRC, MARKER)):
        return 0
    n = 0
    members = _members()
    newest = max(os.path.getmtime(p) for _, p in members)
    if force or not os.path.isfile(ARCHIVE) or os.path.getmtime(ARCHIVE) < newest:
        os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
        tmp = ARCHIVE + ".tmp"
        with tarfile.open(tmp, "w:gz") as tar:
            for rel, p in members:
                data = open(p, "rb").read()
                info = tarfile.TarInfo(rel)
                info.size, info.mtime, info.mode = len(data), 0, 0o644
                tar.addfile(info, io.BytesIO(data))
                n += 1
        os.replace(tmp, ARCHIVE)
    os.makedirs(CKPT, exist_ok=True)
    for scene in ("taekwondo", "walking"):
        s = os.path.join(SRC, "outputs", scene, "layered_rfnr_checkpoint_1.pt")
        d = os.path.join(CKPT, scene + ".pt")
        if os.path.isfile(s) and (force or not os.path.isfile(d)):
            shutil.copyfile(s, d)
            n += 1
    return n


def unpack():
    """Extract the archive (once per archive version) into a scratch directory; returns that directory or None."""
    if not os.path.isfile(ARCHIVE):
        return None
    st = os.stat(ARCHIVE)
    dst = os.path.join(tempfile.gettempdir(), "stnerf_reference_%d_%d" % (st.st_size, int(st.st_mtime)))
    if not os.path.isfile(os.path.join(dst, MARKER)):
        tmp = tempfile.mkdtemp(prefix="stnerf_reference_")
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            for m in tar.getmembers():
                if m.isfile() and not os.path.isabs(m.name) and ".." not in m.name.split("/"):
                    tar.extract(m, tmp)
        try:
            os.rename(tmp, dst)
        except OSError:                       # another process won the race
            shutil.rmtree(tmp, ignore_errors=True)
    return dst if os.path.isfile(os.path.join(dst, MARKER)) else None


def reference_root():
    """`$STNERF_REFERENCE_ROOT`, else /root/reference, else the unpacked archive, else None."""
    env = os.environ.get("STNERF_REFERENCE_ROOT")
    for cand in (env, SRC):
        if cand and os.path.isfile(os.path.join(cand, MARKER)):
            return cand
    return unpack()


if __name__ == "__main__":
    print("packed %d file(s) into %s" % (stash(), os.path.dirname(ARCHIVE)))
    print("reference root:", reference_root())
