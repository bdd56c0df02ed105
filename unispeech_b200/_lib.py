"""ctypes binding of the C-ABI library (include/unispeech_b200.h).

The product path has no CPU / PyTorch fallback: if the shared library is missing, or the device is not
sm_100, calls raise.  Functions take torch CUDA tensors and pass raw device pointers + the current stream.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libunispeech_b200.so"
_lib = None


class Epilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("res1", C.c_void_p), ("res1_bs", C.c_longlong), ("res1_ld", C.c_longlong),
        ("res2", C.c_void_p), ("res2_bs", C.c_longlong), ("res2_ld", C.c_longlong),
        ("gelu_aux", C.c_void_p), ("aux_bs", C.c_longlong), ("aux_ld", C.c_longlong),
        ("out_pre", C.c_void_p), ("pre_bs", C.c_longlong), ("pre_ld", C.c_longlong),
        ("colsum", C.c_void_p),
        ("gelu", C.c_int),
        ("dgelu", C.c_int),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m unispeech_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)"
        )
    lib = C.CDLL(str(_LIB_PATH))
    lib.b200s_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load().b200s_last_error()
        raise RuntimeError(f"unispeech_b200 C-ABI call failed ({rc}): {msg.decode() if msg else '?'}")


def check_device():
    _check(load().b200s_check_device())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> C.c_void_p:
    """Raw cudaStream_t of torch's current stream.  `torch.cuda.current_stream()` costs ~13 us per call (device-index and
    availability checks) and is needed once per kernel launch, so the raw C accessor is used when present."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def ll(v) -> C.c_longlong:
    return C.c_longlong(int(v))


def call(name: str, *args):
    fn = getattr(load(), name)
    _check(fn(*args))


def make_epilogue(bias=None, res1=None, res1_bs=0, res1_ld=0, res2=None, res2_bs=0, res2_ld=0,
                  gelu_aux=None, aux_bs=0, aux_ld=0, out_pre=None, pre_bs=0, pre_ld=0, colsum=None,
                  gelu=False, dgelu=False) -> Epilogue:
    e = Epilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.res1 = res1.data_ptr() if res1 is not None else None
    e.res1_bs, e.res1_ld = int(res1_bs), int(res1_ld)
    e.res2 = res2.data_ptr() if res2 is not None else None
    e.res2_bs, e.res2_ld = int(res2_bs), int(res2_ld)
    e.gelu_aux = gelu_aux.data_ptr() if gelu_aux is not None else None
    e.aux_bs, e.aux_ld = int(aux_bs), int(aux_ld)
    e.out_pre = out_pre.data_ptr() if out_pre is not None else None
    e.pre_bs, e.pre_ld = int(pre_bs), int(pre_ld)
    e.colsum = colsum.data_ptr() if colsum is not None else None
    e.gelu = int(gelu)    # 1: out_pre <- pre-activation, 2: out_pre <- gelu'(pre-activation)
    e.dgelu = int(dgelu)  # 1: aux is the pre-activation, 2: aux already is gelu'(.)
    return e
