"""ctypes loader for libhcmoco_sm100.so.

Prototypes are parsed from include/hcmoco.h so the Python side can never drift from the header.
There is no fallback: if the library is missing the import of any compute op raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "hcmoco.h")
LIB_PATH = os.path.join(HERE, "libhcmoco_sm100.so")

_SCALARS = {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "double": ctypes.c_double}


class HcmError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> list of (name, restype, [(ctype, argname)])"""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = []
    for m in re.finditer(r"(?:^|\n)\s*(int|long|const char\*)\s+(hcm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                typ, an = mm.group(1).strip(), mm.group(2)
                if typ.endswith("* const*"):
                    ct = ctypes.POINTER(ctypes.c_void_p)
                elif "*" in typ or typ == "cudaStream_t":
                    ct = ctypes.c_void_p
                else:
                    ct = _SCALARS[typ]
                argl.append((ct, an))
        protos.append((name, ctypes.c_char_p if "char" in ret else (ctypes.c_long if ret == "long" else ctypes.c_int), argl))
    return protos


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HcmError("%s not found - build it with `python -m hcmoco_b200.build` "
                       "(there is no CPU / PyTorch fallback for the compute path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, ret, args in parse_header():
        fn = getattr(lib, name)
        fn.restype = ret
        fn.argtypes = [a for a, _ in args]
    if lib.hcm_abi_version() != 1:
        raise HcmError("ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise HcmError("%s failed (%d): %s" % (what, rc, load().hcm_last_error().decode()))
