"""ctypes binding of liba3gc_b200.so (the C ABI declared in include/a3gc_b200.h).

There is no CPU path and no PyTorch-eager fallback: if the shared library is missing, or a
tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liba3gc_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

VARIANT = {"AAGC": 0, "A3GC": 1, "AGC": 2, "GGRU": 3}
ACT = {"linear": 0, "tanh": 1, "relu": 2}
PRECISION = {"fp32": 0, "bf16": 1}
ENGINE = {"auto": 0, "simt": 1, "tc": 2}

_f32p = C.POINTER(C.c_float)
ABI_VERSION = 3          # A3GC_ABI_VERSION of include/a3gc_b200.h this binding was written against


class GcParams(C.Structure):
    _fields_ = [("gcn_kernel", C.c_void_p), ("adj", C.c_void_p), ("gcn_bias", C.c_void_p)]


class CellParams(C.Structure):
    _fields_ = [
        ("gcn_kernel", C.c_void_p * 4),
        ("adjacency", C.c_void_p * 4),
        ("gcn_bias", C.c_void_p * 4),
        ("attention_w", C.c_void_p),
        ("attention_wq", C.c_void_p),
        ("attention_wh", C.c_void_p),
        ("attention_u", C.c_void_p),
        ("attention_bs", C.c_void_p),
        ("attention_bu", C.c_void_p),
        ("g_gcn_kernel", C.c_void_p),
        ("g_adjacency", C.c_void_p),
        ("dense_in_w", C.c_void_p * 3),
        ("dense_in_b", C.c_void_p * 3),
        ("dense_hid_w", C.c_void_p * 3),
    ]


class Tape(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("gates", "u", "c", "hh", "e", "hp", "a", "q", "s")]


class TapeGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dzm", "dep", "dqs", "dqp", "dap", "dzm_hi16", "dzm_lo16")]


class NetParams(C.Structure):
    _fields_ = [("linear_in", GcParams), ("rnn", (CellParams * 2) * 2), ("linear_out", GcParams), ("packed_rnn", C.c_void_p * 2)]


# every symbol include/a3gc_b200.h declares: name -> (restype, argtypes)
_PTR2 = C.c_void_p * 2
SYMBOLS = {
    "a3gc_abi_version": (C.c_int, []),
    "a3gc_last_error": (C.c_char_p, []),
    "a3gc_device_count": (C.c_int, []),
    "a3gc_gc_forward": (C.c_int, [C.POINTER(GcParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "a3gc_layer_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "a3gc_layer_forward": (C.c_int, [C.c_int, C.c_int, C.POINTER(CellParams), C.POINTER(C.c_int),
                                     C.c_void_p, C.c_int64, C.c_int64,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_packed_weights_bytes": (C.c_size_t, [C.c_int] * 6),
    "a3gc_pack_weights": (C.c_int, [C.c_int, C.c_int, C.POINTER(CellParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_net_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "a3gc_net_forward": (C.c_int, [C.c_int, C.POINTER(NetParams), C.c_void_p,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_net_forward_raw": (C.c_int, [C.c_int, C.POINTER(NetParams)] + [C.c_void_p] * 7 +
                             [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                              C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                              C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_layer_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "a3gc_layer_train_forward": (C.c_int, [C.c_int, C.c_int, C.POINTER(CellParams), C.POINTER(C.c_int),
                                           C.c_void_p, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(Tape), C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_layer_backward": (C.c_int, [C.c_int, C.c_int, C.POINTER(CellParams), C.POINTER(C.c_int),
                                      C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(Tape), C.POINTER(TapeGrads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "a3gc_prepare_input": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int, C.c_void_p]),
    "a3gc_concat_stage_input": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "a3gc_reduced_to_full_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "a3gc_train_split_tf32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "a3gc_train_hprev_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_int, C.c_int, C.c_void_p]),
    "a3gc_train_split_mixed": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "a3gc_train_hprev_split_mixed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                               C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "a3gc_train_adjacency_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "a3gc_profile_enable": (C.c_int, [C.c_int]),
    "a3gc_profile_count": (C.c_int, []),
    "a3gc_profile_get": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "a3gc_tc_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "a3gc_tc_mma_bench": (C.c_int, [C.c_int] * 11 + [C.c_void_p, C.c_void_p]),
    "a3gc_tc_stream_bench": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "a3gc_debug_read_tc_trace": (C.c_int, [C.c_void_p]),
    "a3gc_debug_max_active_clusters": (C.c_int, [C.c_int, C.c_int]),
    "a3gc_launch_count": (C.c_int64, []),
    "a3gc_reset_launch_count": (None, []),
}

_lib = None
_lock = threading.Lock()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile liba3gc_b200.so for sm_100a with nvcc (in-tree: a3gc_ip_b200/lib/)."""
    if os.path.exists(LIB_PATH) and not force:
        srcs = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cu", ".cuh"))]
        srcs.append(os.path.join(os.path.dirname(_HERE), "include", "a3gc_b200.h"))
        if all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
            return LIB_PATH
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-8000:])
        print(res.stderr[-8000:])
    if res.returncode != 0:
        raise RuntimeError("building liba3gc_b200.so failed (nvcc, sm_100a); see output above")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                        "a3gc_ip_b200 has no CPU or eager fallback.")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                if l.a3gc_abi_version() != ABI_VERSION:
                    raise RuntimeError("liba3gc_b200.so ABI version mismatch")
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().a3gc_last_error().decode(errors="replace")
        raise RuntimeError(f"{what or 'a3gc'} failed (status {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: a3gc_ip_b200 runs on CUDA devices only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(ts: Optional[Sequence[Optional[torch.Tensor]]], n: int = 2):
    arr = (C.c_void_p * n)()
    if ts is not None:
        for i, t in enumerate(ts):
            arr[i] = ptr(t)
    return arr


class Workspace:
    """Grow-only byte workspace owned by the caller side (torch memory), one per module."""

    def __init__(self) -> None:
        self.buf: Optional[torch.Tensor] = None

    # scratch memory is not module state: copies (copy.deepcopy, pickle / torch.save of a whole module) start empty
    def __deepcopy__(self, memo) -> "Workspace":
        return Workspace()

    def __getstate__(self):
        return {}

    def __setstate__(self, state) -> None:
        self.buf = None

    def release(self) -> None:
        self.buf = None

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.device != device or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf
