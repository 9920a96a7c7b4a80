"""ctypes binding of libarkb200.so (the C ABI declared in include/arkb200.h).

The argument types are derived from the header itself, so the binding cannot drift from the
declared ABI.  There is NO fallback: if the shared library is missing or a call fails, this module
raises — the product path never silently runs on a CPU or library substitute.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "arkb200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libarkb200.so")

# enums mirrored from the header
F32, BF16 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
EPI_NONE, EPI_GELU, EPI_TANH, EPI_RELU = 0, 1, 2, 3

_SCALARS = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "size_t": ctypes.c_size_t,
}


def parse_header(path: str = HEADER):
    """Returns {name: (restype, [argtypes])} for every `ark_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?m)^\s*(const\s+char\s*\*|int64_t|int|void)\s+(ark_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "void": None}.get(ret.strip(), ctypes.c_char_p)
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class ArkError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m ark_b200.build` (nvcc, sm_100a). "
                "ark_b200 has no CPU or PyTorch fallback.")
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self._dll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if self._dll.ark_abi_version() != 1:
            raise ImportError("libarkb200.so ABI version mismatch; rebuild")

    def call(self, name, *args):
        """Invoke an int-returning entry point; raise ArkError(ark_last_error()) on non-zero."""
        rc = getattr(self._dll, name)(*args)
        if rc != 0:
            msg = self._dll.ark_last_error()
            raise ArkError(f"{name} failed (code {rc}): {msg.decode() if msg else ''}")

    def raw(self, name):
        """The bare ctypes function (for entry points whose int result is a value, not a status)."""
        return getattr(self._dll, name)

    def launch_count(self) -> int:
        return int(self._dll.ark_launch_count())

    def reset_launch_count(self) -> None:
        self._dll.ark_launch_count_reset()


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
