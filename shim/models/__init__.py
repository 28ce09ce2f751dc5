"""Drop-in replacement of the reference's ``models`` package entry point (models/__init__.py:4-11).

With ``shim/`` ahead of the reference on PYTHONPATH, main_vg.py:19 ``from models import build_reftr`` resolves HERE and gets
``reftr_b200.build_reftr`` -- same signature, same ``(model, criterion, postprocessors)`` triple (main_vg.py:179) -- while every
other ``models.*`` import of the reference (models.criterion, models.post_process, models.reftr_segmentation, ...) still
resolves to the reference's own files: this package's ``__path__`` is extended with the reference's ``models/`` directory.
No reference file is edited.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.append(_root)  # so that ``import reftr_b200`` works when only shim/ was put on PYTHONPATH

try:
    import reftr_compat
    reftr_compat.install()
except ImportError:  # shim/ itself not on sys.path (package imported by file location)
    sys.path.insert(0, os.path.dirname(_here))
    import reftr_compat
    reftr_compat.install()


def _reference_models_dir():
    cands = [os.environ.get("REFTR_REF")] + [p for p in sys.path]
    for p in cands:
        if not p:
            continue
        d = os.path.join(os.path.abspath(p), "models")
        if os.path.abspath(d) != _here and os.path.isfile(os.path.join(d, "reftr_transformer.py")):
            return d
    return None


_ref = _reference_models_dir()
if _ref is not None:
    __path__.append(_ref)  # models.criterion, models.post_process, models.modeling.* -> the reference's own files

from reftr_b200.api import build_reftr  # noqa: E402,F401

reftr_compat.install_torch()
