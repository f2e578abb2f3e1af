"""Minimal `_check_soft_dependencies` so `pgmpy.global_vars` / `compat_fns` import."""
import importlib.util


def _check_soft_dependencies(*packages, severity="error", msg=None, **kwargs):
    flat = []
    for p in packages:
        if isinstance(p, (list, tuple)):
            flat.extend(p)
        else:
            flat.append(p)
    missing = []
    for p in flat:
        name = str(p).split(">")[0].split("=")[0].split("<")[0].strip().replace("-", "_")
        try:
            found = importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            found = False
        if not found:
            missing.append(p)
    if missing:
        if severity == "error":
            raise ModuleNotFoundError(msg or f"missing soft dependencies: {missing}")
        return False
    return True
