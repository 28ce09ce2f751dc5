"""Auto-imported by Python when ``shim/`` is on PYTHONPATH: installs the compatibility layer (shim/reftr_compat.py) before
the reference's main_vg.py runs, then chains to any other ``sitecustomize`` further down sys.path (so an environment's own
hook keeps working).  Usage:

    PYTHONPATH=/path/to/reftr-b200/shim:/path/to/reftr-b200:/path/to/RefTR  python main_vg.py <flags of configs/*.sh>
"""
import importlib.machinery
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
import reftr_compat  # noqa: E402

reftr_compat.install()

# chain: another sitecustomize on the path (e.g. a launcher's hook) must still run
for _p in sys.path:
    if os.path.abspath(_p or ".") == _here:
        continue
    _f = os.path.join(_p or ".", "sitecustomize.py")
    if os.path.isfile(_f):
        _spec = importlib.util.spec_from_file_location("_chained_sitecustomize", _f)
        _mod = importlib.util.module_from_spec(_spec)
        try:
            _spec.loader.exec_module(_mod)
        except Exception as _e:  # a broken foreign hook must not stop the run
            sys.stderr.write(f"reftr_b200 shim: chained sitecustomize {_f} failed: {_e}\n")
        break
