"""The C ABI (include/reftr_b200.h <-> reftr_b200/libreftr_b200.so <-> the ctypes binding in reftr_b200/_lib.py), checked without a
GPU: the library loads, exports every function the header declares, the ctypes mirrors of the header's structs have the C
compiler's size and field offsets, and nothing falls back to the CPU."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from reftr_b200 import _lib
    lib = _lib.lib()
    protos = _lib.header_prototypes()
    assert len(protos) >= 45
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.rb_last_error is not None
    assert lib.rb_act_dtype() in (0, 1) and lib.rb_version() >= 2


def test_ctypes_structs_match_the_header():
    from reftr_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "reftr_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(rb_geom), sizeof(rb_dropout), sizeof(rb_gemm_args), sizeof(rb_adamw_segments));
  printf("%zu %zu %zu %zu %zu\n", offsetof(rb_gemm_args, bias), offsetof(rb_gemm_args, geom), offsetof(rb_gemm_args, drop),
         offsetof(rb_gemm_args, mask_scale), offsetof(rb_adamw_segments, lr));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out[:4]]
    offs = [int(x) for x in out[4:]]
    assert sizes == [C.sizeof(_lib.Geom), C.sizeof(_lib.Dropout), C.sizeof(_lib.GemmArgs), C.sizeof(_lib.AdamwSegments)]
    G = _lib.GemmArgs
    assert offs == [G.bias.offset, G.geom.offset, G.drop.offset, G.mask_scale.offset, _lib.AdamwSegments.lr.offset]


def test_no_cpu_fallback():
    """CPU tensors are refused by the product path (reftr_b200 has no CPU / PyTorch fallback); the CPU tests of the engine run on
    tests/emu_ops.py, which the product never imports."""
    from reftr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.require_device(torch.zeros(1))
    import reftr_b200
    pkg = os.path.dirname(reftr_b200.__file__)
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "emu_ops" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_sm_limit_scope_nests_and_ignores_zero():
    """ops.sm_limit_scope sets the default rb_gemm_args.sm_limit of a region (BERT's chains beside the conv backbone); 0 leaves the
    enclosing value in place."""
    from reftr_b200 import ops
    assert ops.SM_LIMIT == 0
    with ops.sm_limit_scope(36):
        assert ops.SM_LIMIT == 36
        with ops.sm_limit_scope(0):
            assert ops.SM_LIMIT == 36
        with ops.sm_limit_scope(12):
            assert ops.SM_LIMIT == 12
        assert ops.SM_LIMIT == 36
    assert ops.SM_LIMIT == 0


def test_kernel_source_hash_is_stable_and_tracks_the_sources():
    from reftr_b200 import _lib
    a, b = _lib.kernel_source_hash(), _lib.kernel_source_hash()
    assert a == b and len(a) == 16
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
    if os.path.exists(path):  # the committed ncu summary must belong to the committed kernel sources
        assert json.load(open(path)).get("kernel_src_sha256") == a
