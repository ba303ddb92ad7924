"""ctypes binding of libspuq_sg.so (the C ABI of include/spuq_sg.h).

This is the stub a spuq maintainer would add on the reference side: plain
``ctypes`` over ``extern "C"`` functions, PyTorch used only to own device buffers and streams.

The CUDA library is mandatory.  There is no CPU implementation behind this module: if the shared
object is missing or no CUDA device is present, importing the operators works (so the host-side
index/polynomial code stays usable) but the first call that needs the device raises
``SpuqBackendError``.  ``use_library(path, device="cpu")`` is a TEST HOOK that loads the
host-emulated build of the very same sources (tests/_emu/libspuq_sg_emu.so, built by
``make -C spuq_b200/csrc emu``) so the plan builder and kernel index arithmetic can be exercised on
a GPU-less CI box; it is never selected automatically.
"""
import ctypes as C
import os

import torch

__all__ = ["SpuqBackendError", "SpuqError", "lib", "device", "use_library", "check", "ptr", "stream_ptr"]

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libspuq_sg.so")

PATH_AUTO, PATH_CSR, PATH_ELL, PATH_STENCIL = 0, 1, 2, 3
PATH_NAMES = {PATH_CSR: "csr", PATH_ELL: "ell", PATH_STENCIL: "stencil"}
PCG_RUNNING, PCG_CONVERGED, PCG_PRECOND_NOT_PD, PCG_SINGULAR, PCG_NOT_PD, PCG_NOT_CONVERGED = -1, 0, 1, 2, 3, 4
PCG_STATE_DOUBLES = 16


class SpuqBackendError(RuntimeError):
    """The CUDA backend (shared library or device) is not available."""


class SpuqError(RuntimeError):
    """A libspuq_sg call returned an error code."""


_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/spuq_sg.h declares
SIGNATURES = {
    "spuq_sg_last_error": (C.c_char_p, []),
    "spuq_sg_abi_version": (C.c_int, []),
    "spuq_sg_lambda_sort": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "spuq_sg_lambda_neighbours": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp]),
    "spuq_sg_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       C.c_int32, C.c_int32, C.c_int, C.c_int32]),
    "spuq_sg_plan_destroy": (C.c_int, [_vp]),
    "spuq_sg_plan_info": (C.c_int, [_vp, _i32p, _i64p, _i64p, _i64p, _i64p]),
    "spuq_sg_plan_tune": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "spuq_sg_apply": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp]),
    "spuq_sg_apply_part": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int32, _vp]),
    "spuq_sg_plan_last_launches": (C.c_int, [_vp]),
    "spuq_sg_plan_kernel_name": (C.c_char_p, [_vp]),
    "spuq_sg_reduce_scratch_bytes": (C.c_int64, []),
    "spuq_sg_dot": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_dot2": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_axpy": (C.c_int, [C.c_int64, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_axpy2": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_xpay": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_scal": (C.c_int, [C.c_int64, C.c_double, _vp, _vp]),
    "spuq_sg_precond_mean_fd_forward": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp]),
    "spuq_sg_precond_mean_fd_inverse": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, _vp]),
    "spuq_sg_precond_mean_fd_mode_order": (C.c_int, [_vp, _vp]),
    "spuq_sg_spike_correct": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_neighbour_combine": (C.c_int, [C.c_int64, C.c_int32, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp]),
    "spuq_sg_combine": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp]),
    "spuq_sg_precond_identity": (C.c_int, [C.POINTER(_vp), C.c_double]),
    "spuq_sg_precond_from_plan": (C.c_int, [C.POINTER(_vp), _vp]),
    "spuq_sg_precond_mean_fd": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int32, C.c_int32,
                                           C.c_double, C.c_double, C.c_double]),
    "spuq_sg_precond_sparse_lu": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spuq_sg_precond_mean_fd_rows": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int32, C.c_double, C.c_double, C.c_double,
                                                C.c_int32, C.c_int32, _vp]),
    "spuq_sg_precond_strip_info": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "spuq_sg_precond_strip_forward": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, _vp, _vp, _vp]),
    "spuq_sg_precond_strip_assemble": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "spuq_sg_precond_strip_schur": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp]),
    "spuq_sg_precond_strip_final": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, C.c_int64, _vp, _vp]),
    "spuq_sg_precond_destroy": (C.c_int, [_vp]),
    "spuq_sg_precond_work_bytes": (C.c_int64, [_vp, C.c_int64, C.c_int32]),
    "spuq_sg_precond_last_launches": (C.c_int, [_vp]),
    "spuq_sg_precond_apply": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp]),
    "spuq_sg_pcg_check_zeta": (C.c_int, [_vp, C.c_int32, _vp]),
    "spuq_sg_pcg_check_alpha": (C.c_int, [_vp, _vp]),
    "spuq_sg_pcg_rotate": (C.c_int, [_vp, _vp]),
    "spuq_sg_pcg_check_zeta_hist": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "spuq_sg_pcg_accept": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "spuq_sg_set_gate": (C.c_int, [_vp]),
    "spuq_sg_halo_push": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp,
                                     C.c_int64, _vp, _vp]),
    "spuq_sg_halo_wait": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp]),
    "spuq_sg_copy2d_async": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, _vp]),
    "spuq_sg_flag_wait_geq": (C.c_int, [_vp, C.c_int64, _vp]),
    "spuq_sg_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
    "spuq_sg_ipc_export": (C.c_int, [_vp, C.c_char_p, _i64p]),
    "spuq_sg_ipc_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(_vp)]),
    "spuq_sg_ipc_close": (C.c_int, [C.c_int, _vp]),
    "spuq_sg_pcg_work_bytes": (C.c_int64, [_vp, _vp]),
    "spuq_sg_pcg": (C.c_int, [_vp, _vp, _vp, _vp, C.c_double, C.c_int32, C.c_int32, _vp,
                               _f64p, _i32p, _i32p, _vp, _vp]),
}

_lib = None
_device = None
_lib_path = None


def _bind(path):
    handle = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return handle


def load_only(path=DEFAULT_LIB):
    """Load and bind the library without touching a device (used by the CPU-side ABI test)."""
    return _bind(path)


def use_library(path, device="cpu"):
    """TEST HOOK: route all calls to another build of the library (see module docstring)."""
    global _lib, _device, _lib_path
    _lib = _bind(path)
    _device = torch.device(device)
    _lib_path = path
    return _lib


def reset():
    global _lib, _device, _lib_path
    _lib = _device = _lib_path = None


def lib():
    """The bound library; loads libspuq_sg.so on first use and insists on a CUDA device."""
    global _lib, _device, _lib_path
    if _lib is None:
        if not os.path.isfile(DEFAULT_LIB):
            raise SpuqBackendError(
                "libspuq_sg.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C spuq_b200/csrc`; there is no CPU fallback" % DEFAULT_LIB)
        if not torch.cuda.is_available():
            raise SpuqBackendError("no CUDA device: spuq_b200 runs its operators on a B200 only "
                                   "(there is no CPU fallback)")
        _lib = _bind(DEFAULT_LIB)
        _device = torch.device("cuda", torch.cuda.current_device())
        _lib_path = DEFAULT_LIB
    return _lib


_host_lib = None


def host_lib():
    """Library handle for the host-only entry points (K5 integer set-up): needs the shared object
    but no CUDA device."""
    global _host_lib
    if _lib is not None:
        return _lib
    if _host_lib is None:
        if not os.path.isfile(DEFAULT_LIB):
            raise SpuqBackendError("libspuq_sg.so not found at %s (no CPU fallback)" % DEFAULT_LIB)
        _host_lib = _bind(DEFAULT_LIB)
    return _host_lib


def device():
    lib()
    return _device


def is_emulated():
    return _device is not None and _device.type == "cpu"


def library_path():
    return _lib_path


def check(rc, what=""):
    if rc != 0:
        msg = host_lib().spuq_sg_last_error()
        raise SpuqError("%s failed (code %d): %s" % (what or "libspuq_sg call", rc,
                                                     msg.decode() if msg else "?"))


def ptr(t):
    """Raw address of a tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    """Current torch stream as a cudaStream_t (NULL under emulation)."""
    if _device is not None and _device.type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(_device).cuda_stream)
    return None


def device_index():
    d = device()
    return d.index if d.type == "cuda" and d.index is not None else 0
