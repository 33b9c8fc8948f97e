"""ctypes binding of libmml_b200.so -- the ONLY compute backend of this package.

Every entry point declared in include/mml_b200.h is bound here with explicit
argtypes.  There is no CPU or PyTorch-eager fallback: if the library is missing
or a call returns non-zero, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libmml_b200.so")

_P = c_void_p
ABI_VERSION = 3          # must equal MML_ABI_VERSION of include/mml_b200.h
_SIGNATURES = {
    "mml_abi_version": (ctypes.c_int, []),
    "mml_last_error": (c_char_p, []),
    "mml_launch_count": (c_int64, []),
    "mml_device_error_flags": (ctypes.c_int, [_P, c_int32]),
    "mml_crd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "mml_crd_fused_loss_grad": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, _P, _P, c_int32, _P, _P, c_int64, c_int64,
        c_float, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mml_crd_fused_loss_grad_multipos": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, c_int64, c_int64,
        c_float, _P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mml_crd_relation_diff": (ctypes.c_int, [_P, _P, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, c_int64, _P, _P]),
    "mml_crd_sort_columns_max": (c_int32, []),
    "mml_crd_sort_columns": (ctypes.c_int, [_P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int64, _P, c_int64, _P]),
    "mml_crd_kmeans_max_clusters": (c_int32, []),
    "mml_crd_kmeans_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "mml_crd_kmeans_lloyd": (ctypes.c_int, [_P, c_int64, c_int32, _P, _P, c_int32, c_int32, _P, _P, c_int32, c_int32, _P, _P, _P, _P, _P,
                                            _P, ctypes.c_size_t, _P]),
    "mml_crd_knn_max_positives": (c_int32, []),
    "mml_crd_knn_workspace_bytes": (c_int64, [c_int64, c_int64, c_int32]),
    "mml_crd_knn_inv_norms": (ctypes.c_int, [_P, c_int64, c_int32, _P, c_int64, _P, _P]),
    "mml_crd_knn_positives": (ctypes.c_int, [_P, c_int64, c_int32, _P, _P, c_int32, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P, _P,
                                             _P, ctypes.c_size_t, _P]),
    "mml_crd_scores": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, _P, _P, c_int32, _P, c_int64, c_int64,
        c_float, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mml_crd_weighted_rows": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, c_int32, _P, _P, _P, c_int64, c_int64, _P, _P, _P, c_size_t, _P]),
    "mml_shard_route_strided": (ctypes.c_int, [_P, c_int64, c_int64, c_int32, c_int64, c_int32, _P, _P, _P]),
    "mml_crd_peer_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "mml_crd_fused_loss_grad_peer": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, _P, _P, _P, c_int32, c_int64, c_int32, c_int32, _P,
        c_float, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "mml_crd_scores_peer": (ctypes.c_int, [
        _P, _P, c_int64, c_int32, _P, _P, _P, _P, c_int32, c_int64, c_int32, c_int32, c_float, _P, _P, c_size_t, _P]),
    "mml_symm_push": (ctypes.c_int, [_P, c_int32, _P, _P, _P, c_int32, _P]),
    "mml_symm_pull_reduce": (ctypes.c_int, [_P, c_int32, c_int64, c_int64, c_int64, c_int64, c_int32, _P, _P,
                                            c_int64, c_int32, _P, _P]),
    "mml_shard_count": (ctypes.c_int, [_P, c_int64, c_int64, c_int32, c_int64, c_int32, _P, _P]),
    "mml_shard_scatter": (ctypes.c_int, [_P, c_int64, c_int64, c_int32, c_int64, c_int32, _P, _P, _P]),
    "mml_crd_memory_update": (ctypes.c_int, [
        _P, _P, c_int32, _P, _P, _P, c_int64, c_float, c_int64, c_int64, _P]),
    "mml_instance_sample": (ctypes.c_int, [_P, c_int64, _P, _P, _P, c_int32, c_int64, c_int32, c_int32, c_int32,
                                           ctypes.c_uint64, _P, _P, _P]),
    "mml_l2norm_fwd": (ctypes.c_int, [_P, c_int64, c_int32, _P, _P, _P]),
    "mml_l2norm_bwd": (ctypes.c_int, [_P, _P, _P, c_int64, c_int32, _P, _P]),
    "mml_alias_build_host": (ctypes.c_int, [_P, c_int64, _P, _P]),
    "mml_alias_gather_prob": (ctypes.c_int, [_P, _P, c_int64, _P, _P]),
    "mml_alias_select": (ctypes.c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P]),
    "mml_kron_num_chunks": (c_int64, [c_int32, c_int32, c_int32]),
    "mml_kron_chunk_table_host": (ctypes.c_int, [c_int32, c_int32, c_int32, _P]),
    "mml_kron_packed_floats": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_pack_weight": (ctypes.c_int, [_P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P]),
    "mml_kron_fwd_supported": (ctypes.c_int, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_fwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_linear_fwd": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32,
        _P, _P, c_size_t, _P]),
    "mml_kron_fwd_stat_tiles": (c_int64, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_linear_fwd_stats": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32,
        _P, _P, _P, c_size_t, _P]),
    "mml_bn_relu_fwd": (ctypes.c_int, [_P, c_int64, c_int32, _P, c_int32, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P]),
    "mml_kron_wgrad_supported": (ctypes.c_int, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_linear_wgrad": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32,
        _P, _P, c_size_t, _P]),
    "mml_kron_dgrad_supported": (ctypes.c_int, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_packed_t_floats": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_pack_weight_t": (ctypes.c_int, [_P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P]),
    "mml_kron_dgrad_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "mml_kron_linear_dgrad": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32,
        _P, _P, _P, _P, c_size_t, _P]),
    "mml_kron_linear_fwd_simt": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32, _P, _P]),
    "mml_kron_linear_bwd_simt": (ctypes.c_int, [
        _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, c_int32, c_float, ctypes.c_uint64, _P, c_int32,
        _P, _P, _P, _P, _P]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the C-ABI library; fail loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "This package has no CPU / eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.mml_abi_version() != ABI_VERSION:
            raise RuntimeError("libmml_b200.so ABI version mismatch")
        _lib = _DeviceGuardedLib(handle)
    return _lib


class _DeviceGuardedLib:
    """The C ABI works on the CURRENT CUDA device (launches, cudaFuncSetAttribute, tensor maps, the error word).  Every
    entry point that enqueues work takes the stream `cur_stream(tensor.device)` produced while its arguments are built;
    that call notes the device, and the entry point then runs with that device current when it is not already (a module
    on cuda:1 while the process' current device is cuda:0, as the reference's modules allow)."""

    def __init__(self, handle):
        self._handle = handle
        self._fns = {}

    def __getattr__(self, name):
        fn = self._fns.get(name)
        if fn is None:
            raw = getattr(self._handle, name)

            def fn(*args, _raw=raw):
                dev = getattr(_pending_device, "index", None)
                _pending_device.index = None
                if dev is not None and dev != torch.cuda.current_device():
                    with torch.cuda.device(dev):
                        return _raw(*args)
                return _raw(*args)
            self._fns[name] = fn
        return fn


_pending_device = threading.local()      # .index: device noted by cur_stream() for the call this thread is assembling


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().mml_last_error().decode()}")


def dptr(t: torch.Tensor | None, dtype=None) -> c_void_p | None:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mml_b200 kernels take CUDA tensors only (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return c_void_p(t.data_ptr())


def hptr(t: torch.Tensor) -> c_void_p:
    if t.is_cuda or not t.is_contiguous():
        raise RuntimeError("expected a contiguous host tensor")
    return c_void_p(t.data_ptr())


def cur_stream(device) -> c_void_p:
    device = torch.device(device)
    _pending_device.index = device.index if device.index is not None else torch.cuda.current_device()
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count() -> int:
    return int(lib().mml_launch_count())


DEVERR_CRD_INDEX = 1
DEVERR_SHARD_OWNER = 2


def device_error_flags(reset: bool = True) -> int:
    """Sticky error flags of the current CUDA device (synchronises with it); see include/mml_b200.h MML_DEVERR_*."""
    out = ctypes.c_uint32(0)
    check(lib().mml_device_error_flags(ctypes.byref(out), int(reset)), "mml_device_error_flags")
    return int(out.value)


def check_device_errors() -> None:
    """Raise if a kernel saw an out-of-range row id since the last check (the reference's index_select would have hit a
    device-side assert; these kernels clamp / drop the id and set a flag instead of touching memory out of bounds)."""
    f = device_error_flags(reset=True)
    if f & DEVERR_CRD_INDEX:
        raise IndexError("mml_b200: contrast_idx / idx holds a row id outside [0, n_data) (clamped to row 0 by the kernels)")
    if f & DEVERR_SHARD_OWNER:
        raise IndexError("mml_b200: a contrast_idx entry lies outside the sharded bank (dropped by the routing kernel)")
