"""TEST INFRASTRUCTURE ONLY.  Build recipe for ``oracle/_ref/``: the UNMODIFIED Python reference, byte-compiled.

The reference is pure Python, so "compiling it from the sources where they lie" means ``py_compile``: every module
of ``/root/reference/{environment,lib}`` and ``game.py`` is compiled IN PLACE (the sources are read from
``/root/reference``; nothing is written there) into a sourceless compiled module ``<name>.refc`` (the bytes of a ``.pyc``;
the neutral extension because build snapshots commonly drop ``*.pyc``) under ``oracle/_ref/`` with the same package
layout; ``oracle/refshim.py`` installs an import finder for that extension.  No reference source text enters this repository: ``oracle/_ref/`` holds compiled code objects only,
is git-ignored (it never enters history) and is NOT gpurun-ignored, so -- like ``libah_b200.so`` -- it travels to the
GPU box, where ``/root/reference`` does not exist.  There it is what

  * ``bench.py`` times as the reference's own CPU implementation (``cpu_baseline.kind == "reference"``,
    ``--impl reference``): BASELINE config 1, one process and a pool over all host cores;
  * ``tests/test_oracle_vs_reference.py`` diffs the C oracle against, live, on the GPU box too.

The compiled modules are tied to this image's CPython (3.12): both boxes run the same image.  ``python -m
oracle.fetch_ref`` rebuilds; ``__graft_entry__.build()`` calls ``build()`` whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE_ROOT = os.environ.get("AH_REFERENCE_SOURCE", "/root/reference")
PARTS = ("environment", "lib")          # packages on the hot path's import closure
FILES = ("game.py",)
EXT = ".refc"                           # a .pyc by content


def source_available() -> bool:
    return os.path.isfile(os.path.join(SOURCE_ROOT, "environment", "airhockey.py"))


def built() -> bool:
    return os.path.isfile(os.path.join(DEST, "environment", "airhockey" + EXT))


def _jobs():
    for part in PARTS:
        for dirpath, _, names in os.walk(os.path.join(SOURCE_ROOT, part)):
            for name in sorted(names):
                if name.endswith(".py"):
                    src = os.path.join(dirpath, name)
                    yield src, os.path.relpath(src, SOURCE_ROOT)
    for name in FILES:
        yield os.path.join(SOURCE_ROOT, name), name


def build(force: bool = False) -> str:
    """Compile the reference into ``oracle/_ref``; no-op when the sources are absent (GPU box) or up to date."""
    if not source_available():
        return DEST
    stamp = os.path.join(DEST, "BUILD.json")
    if not force and built() and os.path.exists(stamp):
        newest = max(os.path.getmtime(src) for src, _ in _jobs())
        if os.path.getmtime(stamp) >= newest:
            return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    count = 0
    for src, rel in _jobs():
        out = os.path.join(DEST, rel[:-3] + EXT)                 # sourceless layout: pkg/module.refc
        os.makedirs(os.path.dirname(out), exist_ok=True)
        py_compile.compile(src, cfile=out, dfile="<reference>/" + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        count += 1
    with open(stamp, "w") as f:
        json.dump({"what": "byte-compiled reference (code objects only, no source text)", "modules": count,
                   "python": sys.version.split()[0], "source_root": SOURCE_ROOT}, f)
    return DEST


if __name__ == "__main__":
    print(build(force=True), "built" if built() else "NOT built (reference sources absent)")
