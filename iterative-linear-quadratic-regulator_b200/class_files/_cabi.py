"""ctypes binding of libilqr_b200.so (include/ilqr_b200.h).

This is the only bridge between the Python solver interface and the CUDA kernels.  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present when a handle is
created, a RuntimeError is raised.
"""
import ctypes as C
import os

NMAX, MMAX, MAX_ALPHAS = 12, 4, 16
MODELS = {"pendulum": 0, "double_pendulum": 1, "ua_double_pendulum": 2, "ltv": 3, "user": 4}
INTEGRATORS = {"euler": 0, "midpoint": 1, "rk4": 2, "backward_euler": 3}
DTYPES = {"float64": 0, "float32": 1}
KERNEL_CLASSES = ("linearize", "backward", "rollout", "init_rollout", "other")
STATUS_NAMES = {0: "converged", 1: "ls_failed", 2: "maxiter", 3: "running"}

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("ILQR_B200_LIB") or os.path.join(_PKG_ROOT, "libilqr_b200.so")


class Problem(C.Structure):
    """ilqr_problem_t"""
    _fields_ = [
        ("model", C.c_int32), ("integrator", C.c_int32), ("dtype", C.c_int32),
        ("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32), ("B", C.c_int32),
        ("n_alpha", C.c_int32), ("maxiter", C.c_int32), ("reserved", C.c_int32),
        ("dt", C.c_double), ("tol", C.c_double), ("alpha_factor", C.c_double), ("min_alpha", C.c_double),
        ("phys", C.c_double * 16),
        ("Q", C.c_double * (NMAX * NMAX)), ("R", C.c_double * (MMAX * MMAX)),
        ("Qf", C.c_double * (NMAX * NMAX)), ("x_target", C.c_double * NMAX),
        ("Ac", C.c_double * (NMAX * NMAX)), ("E", C.c_double * (NMAX * NMAX)),
        ("Bc", C.c_double * (NMAX * MMAX)), ("ltv_amp", C.c_double),
        ("reg_init", C.c_double), ("reg_factor", C.c_double), ("reg_min", C.c_double), ("reg_max", C.c_double),
    ]


# name -> (restype, argtypes); every symbol include/ilqr_b200.h declares
_VP, _I32P, _I64P = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
SIGNATURES = {
    "ilqr_create": (C.c_int, [C.POINTER(Problem), C.POINTER(_VP)]),
    "ilqr_destroy": (C.c_int, [_VP]),
    "ilqr_module_load": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(_VP)]),
    "ilqr_module_unload": (C.c_int, [_VP]),
    "ilqr_create_user": (C.c_int, [C.POINTER(Problem), _VP, C.POINTER(_VP)]),
    "ilqr_workspace_bytes": (C.c_size_t, [_VP]),
    "ilqr_step": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP, _VP]),
    "ilqr_linearize": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ilqr_cost_expansion": (C.c_int, [_VP] + [_VP] * 11 + [_VP]),
    "ilqr_backward": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ilqr_backward_pass": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "ilqr_rollout": (C.c_int, [_VP, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ilqr_forward_linesearch": (C.c_int, [_VP] * 12 + [_VP]),
    "ilqr_solve": (C.c_int, [_VP] * 8 + [_VP, _VP, _VP, C.c_size_t, _VP, _I64P]),
    "ilqr_set_trace": (C.c_int, [_VP, _VP, _VP]),
    "ilqr_set_linesearch_waves": (C.c_int, [_VP, C.c_int, _I32P]),
    "ilqr_get_linesearch_waves": (C.c_int, [_VP, _I32P]),
    "ilqr_set_mu_buffer": (C.c_int, [_VP, _VP]),
    "ilqr_set_profiling": (C.c_int, [_VP, C.c_int]),
    "ilqr_get_kernel_times": (C.c_int, [_VP, C.POINTER(C.c_double), _I64P]),
    "ilqr_fp64_peak": (C.c_int, [C.POINTER(C.c_double), _VP, _VP]),
    "ilqr_mpc_shift": (C.c_int, [_VP, _VP, _VP, _VP]),
    "ilqr_launch_count": (C.c_int64, [_VP]),
    "ilqr_last_cuda_error": (C.c_int, [_VP]),
    "ilqr_strerror": (C.c_char_p, [C.c_int]),
    "ilqr_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, handle=None):
    if rc == 0:
        return
    lib = load()
    msg = lib.ilqr_strerror(rc).decode()
    if rc == -2 and handle is not None:
        msg += f" [cudaError {lib.ilqr_last_cuda_error(handle)}]"
    raise RuntimeError(f"libilqr_b200: {msg} (code {rc})")


def fill(arr, values):
    for i, v in enumerate(values):
        arr[i] = float(v)
