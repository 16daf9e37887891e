"""ctypes binding of libseedstory_b200.so (the C-ABI declared in include/seedstory_b200.h).

There is no fallback: if the shared library is missing, or a call returns non-zero, a RuntimeError is
raised.  Nothing in this module (or anything it imports) touches `oracle/`.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libseedstory_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "seedstory_b200.h")

_lib = None


class SeedStoryError(RuntimeError):
    pass


_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
}


def _parse_header(path):
    """Return {name: (restype, [argtypes])} for every SS_EXPORT prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"SS_EXPORT\s+(const char\s*\*|int)\s+(\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = re.sub(r"\s+\w+$", "", a).replace("const ", "").strip()
                    argtypes.append(_CTYPES[base])
        protos[name] = (restype, argtypes)
    return protos


def declared_symbols():
    return sorted(_parse_header(HEADER_PATH).keys())


def lib():
    """Load the shared library once; raise loudly when it is absent (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeedStoryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(seedstory_b200 has no CPU or PyTorch fallback)")
    cdll = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _parse_header(HEADER_PATH).items():
        fn = getattr(cdll, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = cdll
    return _lib


def check(rc):
    if rc != 0:
        raise SeedStoryError(lib().ss_last_error().decode())


# kernels launched per C-ABI call (everything else launches exactly one); used for bench.py's `gpu_launches`
_KERNELS_PER_CALL = {"ss_groupnorm_nhwc": 2, "ss_fmha_path_counts": 0, "ss_gemm_row_stat_slots": 0, "ss_groupnorm_ws_floats": 0, "ss_last_error": 0, "ss_version": 0,
                     "ss_require_device": 0, "ss_stream_sync": 0}
_launches = 0


def launch_count():
    return _launches


def add_launches(n):
    """CUDA-graph replays re-issue the kernels recorded at capture time without passing through call()."""
    global _launches
    _launches += n


def call(name, *args):
    global _launches
    check(getattr(lib(), name)(*args))
    _launches += _KERNELS_PER_CALL.get(name, 1)
