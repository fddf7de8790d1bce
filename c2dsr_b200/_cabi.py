"""ctypes binding of libc2dsr_b200.so (include/c2dsr_b200.h).

The library is the only compute path of this package.  If it has not been built, or the
current device is not a Blackwell sm_100 part, every call raises -- there is no CPU or
PyTorch fallback.  PyTorch is used for device memory, streams and (in dist.py) NCCL.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libc2dsr_b200.so")
_lib = None
_lock = threading.Lock()

vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64


class LayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b",
                                  "lin2_w", "lin2_b", "ln1_w", "ln1_b", "ln2_w", "ln2_b")]


class AdamTensor(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("acc", vp), ("m", vp), ("v", vp), ("vmax", vp), ("n", i64)]


MAX_PEERS = 16


class PeerTensor(C.Structure):
    _fields_ = [("t", AdamTensor), ("offset", i64)]


class PeerMap(C.Structure):
    _fields_ = [("world", i32), ("rank", i32), ("grad", vp * MAX_PEERS), ("param", vp * MAX_PEERS), ("grad_mc", vp),
                ("param_mc", vp)]


# name -> (restype, argtypes); mirrors include/c2dsr_b200.h one to one
_DROP = [f32, u64, u64]
_PROTOS = {
    "c2dsr_abi_version": (i32, []),
    "c2dsr_last_error": (C.c_char_p, []),
    "c2dsr_device_check": (i32, []),
    "c2dsr_launch_count": (i64, []),
    "c2dsr_gather_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, f32] + _DROP + [vp]),
    "c2dsr_gather_bwd_workspace_bytes": (i64, [i64, i32, i64, i32]),
    "c2dsr_gather_bwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i64, i32, i64, f32] + _DROP + [vp, i64, vp]),
    "c2dsr_spmm": (i32, [vp, vp, vp, vp, i32, vp, vp, vp, vp, i64, i32, f32, f32, f32, i32] + _DROP + [vp, vp, vp]),
    "c2dsr_mark_rows": (i32, [vp, i64, i64, vp, vp]),
    "c2dsr_spmm_long_row_threshold": (i32, []),
    "c2dsr_gemm_workspace_bytes": (i64, [i64, i64, i64]),
    "c2dsr_gemm": (i32, [i32, i32, i64, i64, i64, f32, vp, i64, vp, i64, f32, vp, i64, vp, i32] + _DROP
                   + [vp, i64, vp]),
    "c2dsr_gemm_tc_workspace_bytes": (i64, [i64, i64, i64]),
    "c2dsr_gemm_tc": (i32, [i32, i32, i64, i64, i64, vp, i64, vp, i64, f32, vp, i64, vp, i32] + _DROP
                      + [i32, vp, i64, vp]),
    "c2dsr_colsum": (i32, [vp, i64, i64, i64, vp, i32, vp, i64, vp]),
    "c2dsr_wsum": (i32, [vp, vp, i64, vp, vp]),
    "c2dsr_encoder_saved_floats": (i64, [i64, i32, i32, i32]),
    "c2dsr_encoder_workspace_bytes": (i64, [i64, i32, i32, i32]),
    "c2dsr_encoder_fwd": (i32, [vp, i32, vp, vp, vp, vp, i64, i32, i32, i32, i64, i32, i32, f32] + _DROP
                          + [vp, vp, vp, i64, vp]),
    "c2dsr_encoder_select_workspace_bytes": (i64, [i64, i32, i32, i32]),
    "c2dsr_encoder_fwd_select": (i32, [vp, i32, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, i32, i32, f32, vp, vp, i64,
                                       vp]),
    "c2dsr_encoder_padkeys_workspace_bytes": (i64, [i64, i32, i32]),
    "c2dsr_encoder_padkeys_prepare": (i32, [vp, i32, vp, i32, i32, f32, vp, vp, i64, vp]),
    "c2dsr_gather_select_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, f32, i64, vp]),
    "c2dsr_encoder_fwd_padkeys": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, i32, i32, f32, vp, vp,
                                        i64, vp]),
    "c2dsr_encoder_bwd": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, i32, i32, f32] + _DROP
                          + [vp, vp, vp, i64, vp]),
    "c2dsr_attention_fwd": (i32, [vp, vp, i64, i32, i32, i32, i64] + _DROP + [vp, vp, vp]),
    "c2dsr_attention_bwd": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, i64] + _DROP + [vp, vp]),
    "c2dsr_add_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, f32] + _DROP + [vp]),
    "c2dsr_ln_bwd": (i32, [vp, vp, vp, vp, vp, i32, i64, i32, vp]),
    "c2dsr_ln_param_grad": (i32, [vp, vp, vp, vp, vp, i64, i32, vp]),
    "c2dsr_infomax_workspace_bytes": (i64, [i64, i32]),
    "c2dsr_infomax_fwd": (i32, [vp] * 11 + [i64, i32, i32, f32, vp, vp, vp, vp, vp, i64, vp]),
    "c2dsr_infomax_bwd": (i32, [vp] * 8 + [i64, i32, i32, f32] + [vp] * 9 + [vp, i64, vp]),
    "c2dsr_score_ldz": (i64, [i64]),
    "c2dsr_score_ce_fwd": (i32, [vp, vp, vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, i64, vp]),
    "c2dsr_score_ce_bwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, i64, vp]),
    "c2dsr_compact_rows": (i32, [vp, i64, i64, vp, vp]),
    "c2dsr_score_ce_tc_workspace_bytes": (i64, [i64, i64, i32, i32]),
    "c2dsr_score_ce_fwd_tc": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, i64, vp]),
    "c2dsr_score_ce_bwd_tc": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
    "c2dsr_score_shard": (i32, [vp, vp, vp, i64, i64, i32, vp, i64, vp, i64, vp]),
    "c2dsr_pick_target": (i32, [vp, i64, vp, i64, i64, i64, vp, vp]),
    "c2dsr_rank_from_scores": (i32, [vp, i64, vp, vp, vp, i64, i64, i64, i64, vp, vp]),
    "c2dsr_split_bf16": (i32, [vp, i64, i32, i64, vp, vp, vp]),
    "c2dsr_score_tc_workspace_bytes": (i64, [i64, i64, i32]),
    "c2dsr_score_target_tc": (i32, [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, vp, vp, vp, i64, vp]),
    "c2dsr_score_count_tc": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, vp, vp, vp, i64, vp, i64, vp]),
    "c2dsr_score_target_full_tc": (i32, [vp, vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, i64, vp]),
    "c2dsr_eval_partition": (i32, [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "c2dsr_eval_ranks": (i32, [vp, vp, vp, vp, i64, vp, vp]),
    "c2dsr_adamw_amsgrad": (i32, [vp, i32, i64, f32, f32, f32, f32, f32, i32, vp]),
    "c2dsr_loss_rows_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                  vp]),
    "c2dsr_loss_rows_bwd_workspace_bytes": (i64, [i64, i32]),
    "c2dsr_loss_rows_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i64, vp, vp, vp, vp, vp, i64, vp]),
    "c2dsr_graph_build_workspace_bytes": (i64, [i64, i64]),
    "c2dsr_graph_build": (i32, [vp, vp, i64, i64, i32, vp, vp, vp, vp, vp, i64, vp]),
    "c2dsr_step_state_bytes": (i32, []),
    "c2dsr_step_state_set": (i32, [vp, i64, f32, vp]),
    "c2dsr_step_state_set_lr": (i32, [vp, f32, vp]),
    "c2dsr_step_begin": (i32, [vp, u64, vp]),
    "c2dsr_adamw_amsgrad_dyn": (i32, [vp, i32, i64, vp, f32, f32, f32, f32, i32, vp]),
    "c2dsr_preprocess_train": (i32, [vp, vp, vp, i64, i64, i64, i32, vp, vp, vp]),
    "c2dsr_preprocess_eval": (i32, [vp, vp, i64, i64, i64, i32, i32, vp, vp, vp, vp]),
    "c2dsr_adamw_amsgrad_peer": (i32, [vp, i32, i64, vp, vp, f32, f32, f32, f32, i32, vp]),
    "c2dsr_axpby": (i32, [vp, vp, vp, i64, f32, f32, vp]),
}
EXPORTS = tuple(_PROTOS)


class C2dsrError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load(check_device: bool = True):
    """Load the shared library (once) and bind every prototype.  Raises if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_LIB_PATH):
                    raise C2dsrError(
                        f"{_LIB_PATH} not found: build it with `python -m c2dsr_b200.build` "
                        "(nvcc, sm_100a).  c2dsr_b200 has no CPU / PyTorch fallback.")
                lib = C.CDLL(_LIB_PATH)
                for name, (res, args) in _PROTOS.items():
                    fn = getattr(lib, name)
                    fn.restype, fn.argtypes = res, args
                if lib.c2dsr_abi_version() != 1:
                    raise C2dsrError("libc2dsr_b200.so ABI version mismatch; rebuild the library")
                _lib = lib
    if check_device:
        _require_device()
    return _lib


_device_ok = False


def _require_device():
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise C2dsrError("c2dsr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    rc = _lib.c2dsr_device_check()
    if rc != 0:
        raise C2dsrError(_lib.c2dsr_last_error().decode())
    _device_ok = True


# Optional per-entry-point device timing (bench.py): PROFILE = {"names": set or None, "events": []}.
# When set, matching calls are bracketed by CUDA events on the current stream.
PROFILE = None


def call(name: str, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    lib = _lib if (_lib is not None and _device_ok) else load()
    prof = PROFILE
    if prof is not None and (prof["names"] is None or name in prof["names"]):
        # inside a stream capture the events become event-record nodes that are re-timed on every replay
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        (prof.setdefault("graph_events", []) if ext else prof["events"]).append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise C2dsrError(f"{name} failed ({rc}): {lib.c2dsr_last_error().decode()}")


# kernel launches replayed from CUDA graphs (the C-side counter only sees launches made through call())
REPLAYED_LAUNCHES = 0


def launch_count() -> int:
    return int(load(check_device=False).c2dsr_launch_count()) + REPLAYED_LAUNCHES


class DynSeed(int):
    """A dropout 'seed' that is the device address of c2dsr_step_state.key (see the header): ops pass it with
    tag | SEED_INDIRECT so that kernels read the per-step key words themselves."""


SEED_INDIRECT = 1 << 63


def seed_tag(seed, tag: int) -> int:
    return (tag | SEED_INDIRECT) if isinstance(seed, DynSeed) else tag


def query(name: str, *args) -> int:
    """Invoke a size-query entry point (no device needed)."""
    return int(getattr(load(check_device=False), name)(*args))


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise C2dsrError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise C2dsrError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise C2dsrError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def stream() -> int:
    """Raw cudaStream_t of torch's current stream (the C-level getter: this runs ~150 times per step)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _Workspace:
    """One growable scratch buffer per (device, stream); kernels on a stream run in order, so reuse is safe."""

    def __init__(self):
        self.buf = {}
        self.pinned = False      # set once a CUDA graph has captured a scratch address: regrown buffers are kept
        self.retired = []

    def get(self, nbytes: int, device) -> torch.Tensor:
        # one buffer per (device, stream): branches running on side streams must not share scratch
        key = ((device if isinstance(device, torch.device) else torch.device(device)).index or 0, stream())
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None and self.pinned:
                self.retired.append(b)
            b = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


workspace = _Workspace()
