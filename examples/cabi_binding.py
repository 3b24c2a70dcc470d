"""Binding include/ilqr_b200.h directly -- the stub of INTEGRATION.md section 2 in runnable form, with torch tensors as the
device buffers (a JAX maintainer would pass `array.unsafe_buffer_pointer()` instead of `tensor.data_ptr()`).

Independent of the package's own ctypes layer (class_files/_cabi.py): the declarations are taken from the header itself
with cffi (ABI mode), so this file only works if the header, the library's exported symbols and the documented
batch-innermost layouts agree.  tests/test_gpu_cabi.py runs it against the oracle.
"""
import os
import re

import cffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ilqr_b200.h")
LIB = os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200", "libilqr_b200.so")


def load(lib_path=LIB, header=HEADER):
    """(ffi, lib): every declaration of the header, bound without any knowledge of the Python package"""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                       # comments
    lines = []
    for line in text.splitlines():
        st = line.strip()
        if st.startswith("#define") and re.match(r"#define\s+\w+\s+-?\d+\s*$", st):
            lines.append(st)                                                # integer constants: cffi understands these
        elif st.startswith("#") or st.startswith('extern "C"') or st == "}":
            continue                                                        # include guards, includes, extern "C" block
        else:
            lines.append(line)
    ffi = cffi.FFI()
    ffi.cdef("\n".join(lines))
    return ffi, ffi.dlopen(lib_path)


def declared_functions(header=HEADER):
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ilqr_\w+)\s*\(", text)))


class Solver:
    """ilqr_create / ilqr_destroy around one problem; arrays are torch CUDA tensors in the header's layouts"""

    def __init__(self, ffi, lib, model, integrator, n, m, N, B, dt, Q, R, Qf, x_target, phys, n_alpha=10, maxiter=100,
                 tol=1e-5):
        self.ffi, self.lib = ffi, lib
        p = ffi.new("ilqr_problem_t *")
        p.model, p.integrator, p.dtype = model, integrator, lib.ILQR_F64
        p.n, p.m, p.N, p.B, p.n_alpha, p.maxiter = n, m, N, B, n_alpha, maxiter
        p.dt, p.tol, p.alpha_factor, p.min_alpha = dt, tol, 0.5, 1e-8
        for i, v in enumerate(phys):
            p.phys[i] = float(v)
        for dst, src in ((p.Q, Q), (p.R, R), (p.Qf, Qf), (p.x_target, x_target)):
            for i, v in enumerate(src):
                dst[i] = float(v)
        out = ffi.new("ilqr_handle_t *")
        rc = lib.ilqr_create(p, out)
        if rc != 0:
            raise RuntimeError(ffi.string(lib.ilqr_strerror(rc)).decode())
        self.h, self.p = out[0], p

    def ptr(self, t):
        return self.ffi.cast("void *", t.data_ptr()) if t is not None else self.ffi.NULL

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.ffi.string(self.lib.ilqr_strerror(rc)).decode())

    def close(self):
        if self.h is not None:
            self.lib.ilqr_destroy(self.h)
            self.h = None
