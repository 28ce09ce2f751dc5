"""Import-time compatibility layer that lets the UNMODIFIED reference (main_vg.py, engine_vg.py, util/, datasets/) import on a
current PyTorch stack (torch >= 2.0, numpy >= 1.24, no matplotlib) -- SURVEY.md section 7.2 "Drop-in without editing the
reference".  Nothing here touches the hot path; it only supplies names the reference expects:

  * ``torch._six`` (util/collate_fn.py:6 -- removed in torch 2.0): ``string_classes``, ``container_abcs``, ``int_classes``
  * ``collections.Iterable`` / ``Mapping`` / ``Sequence`` (util/transforms.py:10 -- removed in Python 3.10)
  * ``matplotlib.pyplot`` (engine_vg.py:13 -- only used by ``evaluate(visualize=True)``): a stub module when matplotlib is absent
  * ``np.bool`` / ``np.int`` / ``np.float`` (datasets/grounding_datasets/refer_dataset.py:120, :202 -- removed in numpy 1.24)

  * ``torch.load`` of the reference's own checkpoints (main_vg.py:311, :343): they hold the ``argparse.Namespace`` of the run
    (main_vg.py:383), which torch >= 2.6 refuses under its ``weights_only=True`` default -> ``install_torch()`` allow-lists it

``install()`` is idempotent and cheap: torch is NOT imported here; ``torch._six`` is served by a meta-path finder when asked for.
"""
import collections
import collections.abc
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

_INSTALLED = False


def _six_module():
    m = types.ModuleType("torch._six")
    m.string_classes = (str, bytes)
    m.int_classes = int
    m.container_abcs = collections.abc
    m.__doc__ = "reftr_b200 shim of the removed torch._six (util/collate_fn.py:6)"
    return m


class _StubLoader(importlib.abc.Loader):
    def __init__(self, factory):
        self.factory = factory

    def create_module(self, spec):
        return self.factory(spec.name)

    def exec_module(self, module):
        pass


class _CompatFinder(importlib.abc.MetaPathFinder):
    """Serves torch._six always, and matplotlib / matplotlib.pyplot only when the real package is not installed."""

    def __init__(self):
        self._probing = False

    def find_spec(self, name, path=None, target=None):
        if name == "torch._six":
            return importlib.machinery.ModuleSpec(name, _StubLoader(lambda n: _six_module()))
        if name in ("matplotlib", "matplotlib.pyplot") and not self._probing:
            self._probing = True
            try:
                real = importlib.machinery.PathFinder.find_spec(name.split(".")[0])
            finally:
                self._probing = False
            if real is not None:
                return None

            def make(n):
                m = types.ModuleType(n)
                m.__path__ = []
                m.__doc__ = "reftr_b200 stub: matplotlib is not installed (only evaluate(visualize=True) needs it)"

                def _missing(*a, **k):
                    raise ImportError("matplotlib is not installed; evaluate(visualize=True) is unavailable")
                m.imsave = m.figure = m.plot = m.savefig = _missing
                return m
            return importlib.machinery.ModuleSpec(name, _StubLoader(make), is_package=(name == "matplotlib"))
        return None


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    for n in ("Iterable", "Mapping", "MutableMapping", "Sequence", "Callable"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    sys.meta_path.insert(0, _CompatFinder())
    try:
        import numpy as np
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for n, t in (("bool", bool), ("int", int), ("float", float), ("object", object)):
                if not hasattr(np, n):
                    setattr(np, n, t)
    except ImportError:
        pass


def install_torch():
    """Called once torch is imported anyway (shim/models/__init__.py): allow-list what the reference pickles into checkpoint.pth."""
    import argparse
    import pathlib
    import torch
    if hasattr(torch.serialization, "add_safe_globals"):
        torch.serialization.add_safe_globals([argparse.Namespace, pathlib.PosixPath, pathlib.PurePosixPath])
