"""ctypes binding of libstarst3r_b200.so (the C ABI in include/starst3r_b200.h).

The product path has no CPU fallback: if the shared library is missing or a
call fails, a RuntimeError is raised (never a silent PyTorch/NumPy substitute).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# ST3R_B200_LIB overrides the path (development builds with instrumentation); there is still no fallback.
LIB_PATH = os.environ.get("ST3R_B200_LIB") or os.path.join(_HERE, "libstarst3r_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "starst3r_b200.h")

_lib = None


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.startswith("const char*") and decl.endswith("*") and " " not in decl[11:].strip() \
            else ctypes.c_void_p
    base = decl.split()[0] if decl else ""
    if base == "const":
        base = decl.split()[1]
    return {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "float": ctypes.c_float, "double": ctypes.c_double,
            "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64, "unsigned": ctypes.c_uint64,
            "cudaStream_t": ctypes.c_void_p, "void": None}[base]


def parse_header(path: str = HEADER_PATH):
    """Returns {name: (restype, [argtypes])} for every ST3R_API prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"ST3R_API\s+([\w\s\*]+?)\s*\b(st3r_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if ret.replace(" ", "") == "constchar*" else _ctype(ret)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                # drop the parameter name
                mm = re.match(r"(.*?[\*\s])(\w+)$", a)
                typ = mm.group(1).strip() if mm else a
                argtypes.append(ctypes.c_void_p if "*" in typ else _ctype(typ))
        protos[name] = (restype, argtypes)
    return protos


def load():
    """Loads the library (once) and types every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"starst3r_b200: {LIB_PATH} is missing - build it with `python -m starst3r_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback for the CUDA hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch, fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().st3r_last_error().decode(errors="replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"starst3r_b200.{what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("starst3r_b200: tensors must live on a CUDA device (no CPU fallback)")


def require_cuda_device(device, what="starst3r_b200"):
    """The hot paths only exist as sm_100a kernels: any other device is refused loudly (no CPU fallback)."""
    import torch
    d = torch.device(device if device is not None else "cuda")
    if d.type != "cuda":
        raise RuntimeError(f"{what}: device={device!r} - the hot path only runs on CUDA (sm_100a); "
                           "no CPU fallback is provided")
    return d
