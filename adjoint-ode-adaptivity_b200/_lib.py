"""ctypes binding of libdgadj.so (include/dgadj.h).  There is no CPU fallback: if the shared
library is missing the import fails, and `dgadj_create` fails without an sm_100 device."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGADJ_LIB") or os.path.join(_HERE, "libdgadj.so")   # DGADJ_LIB: A/B builds only

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
_STATUS = {OK: "OK", ERR_INVALID: "ERR_INVALID", ERR_NO_DEVICE: "ERR_NO_DEVICE", ERR_CUDA: "ERR_CUDA",
           ERR_UNSUPPORTED: "ERR_UNSUPPORTED", ERR_STATE: "ERR_STATE", ERR_NOMEM: "ERR_NOMEM"}

BC = {"inflow": 0, "periodic": 1}
INFLOW = {"zero": 0, "sin_at": 1, "sin_aat": 2, "table": 3}
FUNCTIONAL = {"int_u": 0, "linear": 0, "int_u2": 1}
SCHEME = {"lserk4": 0, "euler": 1}
FD_ODE = {"sin": 0, "linear": 1}
FD_FUNCTIONAL = {"int_u": 0, "u_N": 1, "int_u2": 2}
HM = 5  # row stride of the even/odd blocks (dgadj_kernels.cuh)


class DgadjError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"dgadj {_STATUS.get(code, code)}: {msg}")


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("bc", C.c_int32),
                ("inflow", C.c_int32), ("functional", C.c_int32), ("scheme", C.c_int32),
                ("reserved", C.c_int32), ("alpha", C.c_double)]


class MarchArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("S", C.c_int32), ("reserved", C.c_int32), ("t0", C.c_double),
                ("a", C.c_double), ("dt", C.c_double), ("a_dev", C.c_void_p), ("dt_dev", C.c_void_p)]


class BurgersArgs(C.Structure):
    """dgadj_burgers_args (include/dgadj.h)"""
    _fields_ = [("B", C.c_int64), ("S", C.c_int32), ("limit", C.c_int32), ("indicator", C.c_int32), ("reserved", C.c_int32),
                ("dt", C.c_double), ("dt_dev", C.c_void_p), ("tvb_M", C.c_double),
                ("invV_host", C.c_void_p), ("V_host", C.c_void_p), ("x_host", C.c_void_p), ("jw_host", C.c_void_p),
                ("invVF_host", C.c_void_p), ("VF_host", C.c_void_p), ("xF_host", C.c_void_p), ("jwF_host", C.c_void_p)]


class TdgLoopArgs(C.Structure):
    """dgadj_tdg_loop_args (include/dgadj.h)"""
    _fields_ = [("B", C.c_int64), ("iters", C.c_int32), ("Ks0", C.c_int32), ("Np", C.c_int32), ("nq_march", C.c_int32),
                ("nq_adj", C.c_int32), ("linear", C.c_int32), ("maxit", C.c_int32), ("y0_per_trajectory", C.c_int32),
                ("tol", C.c_double), ("y0_hard", C.c_double), ("times0_host", C.c_void_p),
                ("march_T0_host", C.c_void_p), ("march_T1_host", C.c_void_p), ("adj_T0_host", C.c_void_p),
                ("adj_T1_host", C.c_void_p)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
# name -> (restype, argtypes); exactly the symbols include/dgadj.h declares
PROTOTYPES = {
    "dgadj_version": (C.c_int, []),
    "dgadj_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "dgadj_destroy": (None, [_P]),
    "dgadj_last_error": (C.c_char_p, [_P]),
    "dgadj_set_operators": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dgadj_set_enriched": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "dgadj_set_functional_weights": (C.c_int, [_P, _P, _P]),
    "dgadj_set_inflow_table": (C.c_int, [_P, C.c_int, _P]),
    "dgadj_set_element_orders": (C.c_int, [_P, _P]),
    "dgadj_set_tuning": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "dgadj_forward": (C.c_int, [_P, C.POINTER(MarchArgs), _P, _P, _P, _P, _P]),
    "dgadj_ckpt_bytes": (C.c_int64, [_P, C.c_int64, C.c_int32]),
    "dgadj_adjoint": (C.c_int, [_P, C.POINTER(MarchArgs), _P, _P, _P, _P, _P, _P]),
    "dgadj_fwd_adj": (C.c_int, [_P, C.POINTER(MarchArgs), _P, _P, _P, _P, _P, _P]),
    "dgadj_fwd_adj_windowed": (C.c_int, [_P, C.POINTER(MarchArgs), C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "dgadj_fwd_adj_host": (C.c_int, [_P, C.POINTER(MarchArgs), _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_forward_host": (C.c_int, [_P, C.POINTER(MarchArgs), _P, _P, _P, _P, _P]),
    "dgadj_rhs": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_double, C.c_double, _P, _P, _P]),
    "dgadj_fd_awr": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_tdg_march": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                  _P, _P, _P, _P, _P]),
    "dgadj_tdg_adjoint": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                    _P, _P, _P, _P, _P, _P]),
    "dgadj_tdg_adjoint_rec": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P, _P, _P]),
    "dgadj_tdg_err_contribution": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dgadj_burgers_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, _P, C.c_int32, C.c_double, _P, _P, _P, _P, _P,
                                        _P, _P, _P, _P, _P, _P]),
    "dgadj_burgers_adjoint": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                        _P, _P, _P, _P]),
    "dgadj_tdg_adapt_loop": (C.c_int, [_P, C.POINTER(TdgLoopArgs), _P, _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_fd_adapt_loop": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_tdg_adapt_loop_pt": (C.c_int, [_P, C.POINTER(TdgLoopArgs), _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_fd_adapt_loop_pt": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "dgadj_burgers_fwd_adj": (C.c_int, [_P, C.POINTER(BurgersArgs), _P, _P, _P, _P, _P, _P, _P, _P]),
    "dgadj_burgers_plan": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dgadj_march_status": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "dgadj_ic_indicator": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "dgadj_refine_shared": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "dgadj_rank": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "dgadj_reduce_indicators": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "dgadj_allreduce_indicators": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "dgadj_reduce_indicator_blocks": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "dgadj_allreduce_indicator_blocks": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "dgadj_measure_dfma_peak": (C.c_int, [_P, C.c_double, _D, _D]),
    "dgadj_device_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dgadj_launch_count": (C.c_int64, [_P]),
    "dgadj_plan": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "dgadj_host_eo_operators": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, _D]),
    "dgadj_host_modal_operators": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, _D]),
}

STATUS_NOT_CONVERGED, STATUS_NON_FINITE = 1, 2


def march_status(lib, h, torch, values=None, its=None, maxit=0, stream=None):
    """status[B] (int32 view of the uint32 word) through `dgadj_march_status`: values = a float64 CUDA tensor
    [B, ...] (NaN / Inf check), its = an int32 CUDA tensor [B, ...] of Newton counts (its > maxit = not converged)."""
    ref = values if values is not None else its
    B = ref.shape[0]
    status = torch.empty(B, dtype=torch.int32, device=ref.device)
    v = None if values is None else values.contiguous()
    i = None if its is None else its.contiguous()
    st = C.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream) if stream is None else stream
    rc = lib.dgadj_march_status(h, B, 0 if v is None else v.numel() // B, C.c_void_p(0 if v is None else v.data_ptr()),
                                0 if i is None else i.numel() // B, C.c_void_p(0 if i is None else i.data_ptr()), int(maxit),
                                C.c_void_p(status.data_ptr()), st)
    if rc != OK:
        raise DgadjError(rc, lib.dgadj_last_error(h).decode())
    return status


_lib = None


def load():
    """Load libdgadj.so (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C adjoint-ode-adaptivity_b200/csrc -j8`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
