"""Import shim for running the UNMODIFIED reference on CPU (thin alias of oracle/ref_arm.py).

Used by tests/golden/make_golden.py (build container: /root/reference) and by tests that compare against the reference
directly; those run wherever the reference is present -- /root/reference here, or the git-ignored install
`oracle/_ref/` that __graft_entry__.build() makes and that travels to the GPU box.  The two arithmetic-free stubs
(torch._six, probtorch.Trace) live in oracle/ref_arm.py.
"""
from oracle import ref_arm

REF_ROOT = ref_arm.REF_SRC


def available() -> bool:
    return ref_arm.available()


def install():
    if not available():
        raise RuntimeError("reference tree not present (neither %s nor oracle/_ref)" % REF_ROOT)
    ref_arm.load()
