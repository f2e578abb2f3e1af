"""Import the real reference (pgmpy at /root/reference) as the parity oracle.

TEST INFRASTRUCTURE. Works only where /root/reference exists (the build container); the GPU box
never runs this. Recipe: SURVEY.md Appendix C.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_PYPARSING_VENDORED = "/usr/lib/python3/dist-packages/pip/_vendor/pyparsing"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pgmpy"))


def load_reference():
    """Returns the imported `pgmpy` module of the reference (numpy backend, fp64)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at /root/reference")
    if "pgmpy" in sys.modules and getattr(sys.modules["pgmpy"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["pgmpy"]
    sys.dont_write_bytecode = True  # the reference tree is read-only
    paths = [_SHIMS, REFERENCE_ROOT]
    try:
        import pyparsing  # noqa: F401
    except ImportError:
        # pip's vendored copy; link it alone (the whole _vendor dir would shadow pygments)
        link_dir = tempfile.mkdtemp(prefix="pgx_pyparsing_")
        os.symlink(_PYPARSING_VENDORED, os.path.join(link_dir, "pyparsing"))
        paths.append(link_dir)
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import torch

        torch.backends.opt_einsum.enabled = False  # keep torch away from the opt_einsum stand-in
    except Exception:  # pragma: no cover
        pass
    import pgmpy

    pgmpy.config.set_show_progress(False)
    return pgmpy
