"""ctypes binding of libreftr_b200.so (the C ABI declared in include/reftr_b200.h).

The library is built in-tree by ``reftr_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is no CPU or
PyTorch fallback: if the library is missing, every op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REFTR_B200_LIB") or os.path.join(_HERE, "libreftr_b200.so")  # (override: kernel experiments)


class Geom(C.Structure):
    _fields_ = [("mode", C.c_int), ("Wp", C.c_int), ("HpWp", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Rs", C.c_int)]


class Dropout(C.Structure):
    """rb_dropout: device pointer to the uint64 seed, site id, drop probability."""
    _fields_ = [("seed", C.c_void_p), ("site", C.c_int), ("p", C.c_float)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("A", C.c_void_p), ("a_rows", C.c_longlong), ("a_cols", C.c_int), ("lda", C.c_longlong),
        ("B", C.c_void_p), ("b_rows", C.c_longlong), ("b_cols", C.c_int), ("ldb", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("taps", C.c_int),
        ("a_rowoff", C.c_int * 16),
        ("b_koff", C.c_int * 16),
        ("splits", C.c_int),
        ("block_n", C.c_int),
        ("out_row_off", C.c_longlong),
        ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldres", C.c_longlong),
        ("res32", C.c_void_p), ("ldres32", C.c_longlong),
        ("mask_src", C.c_void_p), ("ldmask", C.c_longlong),
        ("out", C.c_void_p), ("ldo", C.c_longlong),
        ("out32", C.c_void_p), ("ldo32", C.c_longlong),
        ("out32_z_stride", C.c_longlong),
        ("relu", C.c_int),
        ("atomic", C.c_int),
        ("geom", Geom),
        ("drop", C.c_void_p),
        ("drop_gshift", C.c_int),
        ("mask_scale", C.c_float),
        ("bias_grad", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("out_scale", C.c_float),
        ("sm_limit", C.c_int),
    ]


class AdamwSegments(C.Structure):
    """rb_adamw_segments"""
    _fields_ = [("nseg", C.c_int), ("end", C.c_longlong * 32), ("group", C.c_int * 32), ("lr", C.c_float * 8), ("weight_decay", C.c_float * 8)]


_lib = None
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "reftr_b200.h")


def header_prototypes(path=HEADER_PATH):
    """Parses `int rb_xxx(...)` prototypes of the C header -> {name: [ctypes argtypes]} (the header is the single
    source of truth for the ABI; tests check that every declared symbol is exported)."""
    import re
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(rb_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    types.append(C.c_void_p)
                elif a.startswith("long long"):
                    types.append(C.c_longlong)
                elif a.startswith("float"):
                    types.append(C.c_float)
                elif a.startswith("int"):
                    types.append(C.c_int)
                else:
                    raise ValueError(f"unhandled C type in header: {a!r}")
        protos[name] = types
    return protos


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(reftr_b200 has no CPU/PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.rb_last_error.restype = C.c_char_p
        for name, argtypes in header_prototypes().items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
    return _lib


LAUNCHES = 0  # number of C-ABI kernel-launching calls made by this process (bench.py reports it)


def call(name, *args):
    """Calls a C-ABI function; tensors are passed as data_ptr() ints or None; raises on a non-zero return."""
    global LAUNCHES
    LAUNCHES += 1
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name}: {lib().rb_last_error().decode()}")


def check(rc, what):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().rb_last_error().decode()}")


PROFILED_SOURCES = ("common.cuh", "host.h", "host.cu", "build.sh", "gemm.cu", "gemm_skinny.cu", "attention_tc.cu", "attention.cu", "stem_pool.cu")


def kernel_source_hash():
    """SHA-256 (first 16 hex digits) over the sources of the kernels whose ncu captures are committed (the GEMM family, the attention
    kernels, the fused stem) and of the headers / build script they share.  The binary's own hash is not reproducible (two builds of
    identical sources differ), so profile files record this one; sources of other kernels (input pipeline, optimizer ...) do not
    invalidate those captures."""
    import hashlib
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
    h = hashlib.sha256()
    for f in PROFILED_SOURCES:
        h.update(f.encode())
        h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()[:16]
