"""CPU: repository rules -- the product package never touches the oracle or the reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, dirs, files in os.walk(os.path.join(ROOT, d)):
        dirs[:] = [x for x in dirs if x != "_ref"]          # oracle/_ref: the installed, unmodified reference (git-ignored)
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(base, f)


def test_product_does_not_import_oracle_or_reference():
    pat = re.compile(r"^\s*(from|import)\s+(oracle|apgs\b(?!_b200))", re.M)
    for path in _py_files("apgs_b200"):
        txt = open(path).read()
        assert not pat.search(txt), path
        assert "/root/reference" not in txt, path


def test_oracle_headers_say_test_only():
    for path in _py_files("oracle"):
        assert "TEST INFRASTRUCTURE ONLY" in open(path).read(), path


def test_gpu_side_files_do_not_read_reference_at_runtime():
    for name in ("bench.py", "__graft_entry__.py"):
        p = os.path.join(ROOT, name)
        if os.path.exists(p):
            txt = open(p).read()
            # build() may look for /root/reference (container only); nothing else may
            for m in re.finditer(r"/root/reference", txt):
                line = txt[:m.start()].count("\n")
                ctx = "\n".join(txt.split("\n")[max(0, line - 30):line + 1])
                assert "def build" in ctx or "REFERENCE_IN_CONTAINER" in ctx, (name, line)
