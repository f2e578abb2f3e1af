"""Import the real reference (pgmpy 1.0.0) as the parity oracle and as the CPU arm of bench.py.

TEST INFRASTRUCTURE. Two places are searched: /root/reference (the read-only source tree of the build container) and
oracle/_ref/ (a `pip install --target` of that tree made by __graft_entry__.build(); git-ignored, but it travels to the
GPU box with the snapshot, where /root/reference does not exist). Recipe for the missing dependencies: SURVEY.md App. C.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_PYPARSING_VENDORED = "/usr/lib/python3/dist-packages/pip/_vendor/pyparsing"


def reference_root():
    """Directory to put on sys.path, or None when the reference is nowhere to be found."""
    for root in (REFERENCE_ROOT, INSTALLED_ROOT):
        if os.path.isdir(os.path.join(root, "pgmpy")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def install_reference() -> bool:
    """pip-installs the unmodified reference into oracle/_ref (no dependencies, no index). Build container only;
    returns True when oracle/_ref holds the package afterwards."""
    import shutil
    import subprocess

    if os.path.isdir(os.path.join(INSTALLED_ROOT, "pgmpy")):
        return True
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "pgmpy")):
        return False
    tmp = tempfile.mkdtemp(prefix="pgmpy_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE_ROOT, src, ignore=shutil.ignore_patterns(".git"))  # the source tree is read-only
        res = subprocess.run(
            [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
             "--find-links", "/opt/wheelhouse", "--target", INSTALLED_ROOT, src],
            capture_output=True, text=True)
        return res.returncode == 0 and os.path.isdir(os.path.join(INSTALLED_ROOT, "pgmpy"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def load_reference():
    """Returns the imported `pgmpy` module of the reference (numpy backend, fp64)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present (neither /root/reference nor oracle/_ref)")
    if "pgmpy" in sys.modules and getattr(sys.modules["pgmpy"], "__file__", "").startswith(root):
        return sys.modules["pgmpy"]
    sys.dont_write_bytecode = True  # the reference tree is read-only
    paths = [_SHIMS, root]
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
