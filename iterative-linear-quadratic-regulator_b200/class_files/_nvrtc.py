"""Minimal ctypes binding of NVRTC (libnvrtc.so.12): compile a CUDA translation unit held in memory into a cubin for
sm_100a, in process, with no nvcc and no host compiler.  Used by class_files/codegen.py to build the kernels of a
user-defined System subclass -- the counterpart of the reference's jit at construction (system_base.py:203-251)."""
import ctypes as C
import glob
import os
import sys

_lib = None


def _candidates():
    yield "libnvrtc.so.12"
    for base in sys.path + [os.path.dirname(os.path.dirname(os.__file__))]:
        for pat in ("nvidia/cuda_nvrtc/lib/libnvrtc.so.12", "site-packages/nvidia/cuda_nvrtc/lib/libnvrtc.so.12"):
            yield from glob.glob(os.path.join(base, pat))
    for root in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if root:
            yield os.path.join(root, "lib64", "libnvrtc.so.12")
            yield os.path.join(root, "lib64", "libnvrtc.so")


def lib():
    global _lib
    if _lib is None:
        err = None
        for path in _candidates():
            try:
                _lib = C.CDLL(path)
                break
            except OSError as e:
                err = e
        if _lib is None:
            raise RuntimeError(f"NVRTC (libnvrtc.so.12) not found: a user-defined System needs it to build its kernels ({err})")
        L = _lib
        L.nvrtcGetErrorString.restype = C.c_char_p
        L.nvrtcGetErrorString.argtypes = [C.c_int]
        L.nvrtcCreateProgram.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p),
                                         C.POINTER(C.c_char_p)]
        L.nvrtcDestroyProgram.argtypes = [C.POINTER(C.c_void_p)]
        L.nvrtcAddNameExpression.argtypes = [C.c_void_p, C.c_char_p]
        L.nvrtcCompileProgram.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        L.nvrtcGetProgramLogSize.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.nvrtcGetProgramLog.argtypes = [C.c_void_p, C.c_char_p]
        L.nvrtcGetLoweredName.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
        L.nvrtcGetCUBINSize.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.nvrtcGetCUBIN.argtypes = [C.c_void_p, C.c_char_p]
        L.nvrtcVersion.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _lib


def version():
    a, b = C.c_int(0), C.c_int(0)
    lib().nvrtcVersion(C.byref(a), C.byref(b))
    return a.value, b.value


def compile_cubin(source, headers, name_expressions, options=(), name="ilqr_user.cu"):
    """source: str; headers: {include name: text}; name_expressions: template instantiations to emit (C++ source
    spelling).  Returns (cubin bytes, {expression: lowered (mangled) kernel name})."""
    L = lib()

    def check(rc, what, prog=None):
        if rc != 0:
            log = ""
            if prog is not None:
                n = C.c_size_t(0)
                L.nvrtcGetProgramLogSize(prog, C.byref(n))
                buf = C.create_string_buffer(n.value + 1)
                L.nvrtcGetProgramLog(prog, buf)
                log = buf.value.decode(errors="replace")
            raise RuntimeError(f"NVRTC {what} failed: {L.nvrtcGetErrorString(rc).decode()}\n{log[-6000:]}")

    names = list(headers)
    hdr_src = (C.c_char_p * len(names))(*[headers[k].encode() for k in names])
    hdr_names = (C.c_char_p * len(names))(*[k.encode() for k in names])
    prog = C.c_void_p()
    check(L.nvrtcCreateProgram(C.byref(prog), source.encode(), name.encode(), len(names), hdr_src, hdr_names), "create")
    try:
        for e in name_expressions:
            check(L.nvrtcAddNameExpression(prog, e.encode()), f"name expression {e}")
        opts = [o.encode() for o in options]
        check(L.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts)), "compile", prog)
        lowered = {}
        for e in name_expressions:
            out = C.c_char_p()
            check(L.nvrtcGetLoweredName(prog, e.encode(), C.byref(out)), f"lowered name of {e}")
            lowered[e] = out.value.decode()
        n = C.c_size_t(0)
        check(L.nvrtcGetCUBINSize(prog, C.byref(n)), "cubin size")
        buf = C.create_string_buffer(n.value)
        check(L.nvrtcGetCUBIN(prog, buf), "cubin")
        return buf.raw, lowered
    finally:
        L.nvrtcDestroyProgram(C.byref(prog))
