"""``python -m main_vg <flags>`` / ``python shim/main_vg.py <flags>`` with shim/ first on PYTHONPATH: installs the compatibility
layer, then runs the reference's UNMODIFIED main_vg.py (located through $REFTR_REF or sys.path) as ``__main__``."""
import os
import runpy
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
import reftr_compat  # noqa: E402

reftr_compat.install()


def _find_reference_main():
    for p in [os.environ.get("REFTR_REF")] + list(sys.path):
        if not p:
            continue
        f = os.path.join(os.path.abspath(p), "main_vg.py")
        if os.path.isfile(f) and os.path.dirname(f) != _here:
            return f
    raise SystemExit("reftr_b200 shim: the reference's main_vg.py was not found (set REFTR_REF or put the RefTR checkout on PYTHONPATH)")


if __name__ == "__main__":
    ref_main = _find_reference_main()
    ref_dir = os.path.dirname(ref_main)
    if ref_dir not in sys.path:
        sys.path.append(ref_dir)
    sys.argv[0] = ref_main
    runpy.run_path(ref_main, run_name="__main__")
