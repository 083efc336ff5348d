"""ctypes binding of the C ABI (`include/lotus.h` -> `liblotus.so`).

The prototypes are parsed from the header, so the header is the single authority for the Python side too.
There is no fallback: if the library is missing, `load()` raises and every model forward fails loudly.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblotus.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "lotus.h")

_CTYPES = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32, "float": ctypes.c_float, "double": ctypes.c_double,
}


class LotusError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [(argtype, argname), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(lotus_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.+?)\s*(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _ctype_of(t):
    t = t.replace("const ", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t == "char*" else ctypes.c_void_p
    return _CTYPES[t]


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise LotusError(
                f"{LIB_PATH} not found: build it with robot-3dlotus_b200/build.sh (or __graft_entry__.build()). "
                "There is no CPU / PyTorch fallback for the policy hot path.")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)          # AttributeError here == header/library mismatch
            fn.restype = ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if "int64_t" in ret else ctypes.c_int)
            fn.argtypes = [_ctype_of(t) for t, _ in args]
        self.cdll.lotus_last_error.restype = ctypes.c_char_p
        self.launches = 0                            # kernels-launching C calls issued (bench `gpu_launches`)

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        self.launches += 1
        if rc != 0:
            raise LotusError(f"{name} failed (rc={rc}): {self.cdll.lotus_last_error().decode()}")

    def __getattr__(self, name):
        if name.startswith("lotus_"):
            return lambda *a: self.call(name, *a)
        raise AttributeError(name)


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def exported_symbols():
    """names exported by the shared object (via ctypes lookups of the header's declarations)."""
    lib = load()
    return sorted(lib.protos.keys())
