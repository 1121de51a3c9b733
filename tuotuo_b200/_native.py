"""ctypes binding of ``libttlda.so`` (C ABI declared in ``include/ttlda.h``).

There is no CPU fallback: if the library is missing or a call fails, this module raises.  torch is used only
for device memory and streams — every function takes ``torch.Tensor`` arguments, checks device/dtype/contiguity
and passes ``data_ptr()`` plus the current CUDA stream to the C entry point.
"""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TTLDA_LIB") or os.path.join(_PKG, "libttlda.so")  # TTLDA_LIB: tuning builds only

NPART = 8
MAX_K = 1024

_i64 = ctypes.c_int64
_f64 = ctypes.c_double
_ptr = ctypes.c_void_p

# name -> (restype, argtypes)   — mirrors include/ttlda.h one to one
SIGNATURES = {
    "ttlda_version": (ctypes.c_int, []),
    "ttlda_last_error": (ctypes.c_char_p, []),
    "ttlda_check_device": (ctypes.c_int, [ctypes.c_int]),
    "ttlda_ld_for": (_i64, [_i64]),
    "ttlda_dirichlet_expectation_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr]),
    "ttlda_dirichlet_expectation_1d_f64": (ctypes.c_int, [_ptr, _f64, _ptr, _i64, _ptr]),
    "ttlda_bow_to_csr_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "ttlda_bow_to_csr": (ctypes.c_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_csr_to_csc_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "ttlda_csr_to_csc": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr,
                                        _i64, _ptr]),
    "ttlda_csr_to_csc_block": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr,
                                              _ptr, _i64, _ptr]),
    "ttlda_widen_i32": (ctypes.c_int, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _ptr]),
    "ttlda_validate_csr": (ctypes.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _i64, _i64, ctypes.c_int, _ptr, _ptr]),
    "ttlda_reduce_workspace_bytes": (_i64, []),
    "ttlda_theta_refresh_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _f64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_estep_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _f64, _ptr, _ptr, _ptr, _ptr,
                                       _ptr, _ptr, _ptr, _i64, _f64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_estep_hot_rows": (_i64, [_i64]),
    "ttlda_build_estream": (ctypes.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ttlda_estep_hot_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                           _ptr, _ptr, _ptr, _i64, _f64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_estep_fixed_point_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _f64, _ptr, _ptr, _ptr,
                                                   _ptr, _ptr, _i64, _f64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_sstats_atomic_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "ttlda_sstats_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "ttlda_sstats_csc_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                            _ptr, _ptr, _i64, _ptr]),
    "ttlda_sstats_windowed_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "ttlda_sstats_csc_windowed_f64": (ctypes.c_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                                     _ptr, _i64, _ptr]),
    "ttlda_sstats_csc_block_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                                  _ptr, ctypes.c_int, _ptr, _i64, _ptr]),
    "ttlda_sstats_fold_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr]),
    "ttlda_mstep_workspace_bytes": (_i64, [_i64, _i64]),
    "ttlda_mstep_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _f64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_mstep_online_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _f64, _f64, _f64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                              _i64, _ptr]),
    "ttlda_beta_refresh_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _f64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64,
                                              _ptr]),
    "ttlda_mstep_slice_lambda_f64": (ctypes.c_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _i64, _f64, _ptr, _ptr, _ptr,
                                                    _ptr, _i64, _ptr]),
    "ttlda_mstep_slice_beta_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _f64, _ptr, _ptr, _i64, _ptr,
                                                  ctypes.c_int, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ttlda_elbo_tokens_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _i64,
                                             _ptr]),
    "ttlda_hyper_stats_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _i64, _ptr]),
    "ttlda_hyper_colstats_workspace_bytes": (_i64, [_i64, _i64]),
    "ttlda_hyper_colstats_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _i64, _ptr]),
    "ttlda_psi_f64": (ctypes.c_int, [_ptr, _i64, _ptr, _ptr]),
    "ttlda_probe_gather_f64": (ctypes.c_int, [_ptr, _i64, _i64, _ptr, _i64, _ptr, _ptr]),
    "ttlda_transpose_pad_f64": (ctypes.c_int, [_ptr, _i64, _i64, _ptr, _i64, _ptr]),
    "ttlda_transpose_unpad_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr]),
}


class TTLDAError(RuntimeError):
    pass


_lib = None
launch_count = 0  # C-ABI calls made so far
kernel_count = 0  # kernels of libttlda launched so far (bench.py reports the per-region delta as gpu_launches)
timing = None     # set to {} by bench.py: name -> [(start_event, end_event), ...] on the current stream

# hand-written kernels launched by each entry point (CUB sort/RLE passes and memsets are not counted)
KERNELS_PER_CALL = {
    "ttlda_dirichlet_expectation_f64": 1, "ttlda_dirichlet_expectation_1d_f64": 1,
    "ttlda_bow_to_csr": 2, "ttlda_csr_to_csc": 4, "ttlda_csr_to_csc_block": 4, "ttlda_widen_i32": 1, "ttlda_validate_csr": 1,
    "ttlda_theta_refresh_f64": 2, "ttlda_estep_f64": 2, "ttlda_estep_hot_f64": 2, "ttlda_build_estream": 1, "ttlda_estep_fixed_point_f64": 2,
    "ttlda_sstats_csc_f64": 2, "ttlda_sstats_csc_windowed_f64": 0, "ttlda_sstats_atomic_f64": 1, "ttlda_sstats_csc_block_f64": 2, "ttlda_sstats_fold_f64": 1,
    "ttlda_mstep_f64": 4, "ttlda_mstep_online_f64": 4,
    "ttlda_beta_refresh_f64": 4, "ttlda_mstep_slice_lambda_f64": 2, "ttlda_mstep_slice_beta_f64": 3, "ttlda_elbo_tokens_f64": 2, "ttlda_hyper_stats_f64": 2, "ttlda_hyper_colstats_f64": 2, "ttlda_psi_f64": 1,
    "ttlda_transpose_pad_f64": 1, "ttlda_transpose_unpad_f64": 1, "ttlda_probe_gather_f64": 1,
}


def load():
    """Load libttlda.so (no fallback).  Safe without a GPU: only binds symbols."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TTLDAError(
            f"{LIB_PATH} not found: build it with `python -m tuotuo_b200.build` — there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, name: str):
    if rc != 0:
        msg = load().ttlda_last_error().decode("utf-8", "replace")
        raise TTLDAError(f"{name} failed ({rc}): {msg}")


class _DevPtr(ctypes.c_void_p):
    """A device pointer that remembers which GPU it lives on (`dev`), so that `_call` can launch there."""
    dev = -1


class _StreamOfCall:
    """Placeholder for the `stream` argument: `_call` substitutes the current stream OF THE DEVICE THE TENSORS LIVE
    ON (not of the process's current device)."""


_STREAM = _StreamOfCall()


def _p(t: torch.Tensor | None, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise TTLDAError("libttlda takes CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise TTLDAError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TTLDAError(f"expected dtype {dtype}, got {t.dtype}")
    p = _DevPtr(t.data_ptr())
    p.dev = t.device.index
    return p


def _stream():
    return _STREAM


def _call(name, *args):
    """One C-ABI call.  The entry points launch on the CURRENT device (kernels, memsets, CUB), so the call is made
    with the tensors' device current and that device's current stream — `LDASmoothed(K, device="cuda:1")` works
    whatever `torch.cuda.current_device()` is.  Tensors on different devices are refused."""
    global launch_count, kernel_count
    devs = {a.dev for a in args if isinstance(a, _DevPtr)}
    if len(devs) > 1:
        raise TTLDAError(f"{name}: tensors on different CUDA devices {sorted(devs)}")
    dev = devs.pop() if devs else torch.cuda.current_device()
    launch_count += 1
    kernel_count += KERNELS_PER_CALL[name]
    fn = getattr(load(), name)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        args = [ctypes.c_void_p(stream.cuda_stream) if a is _STREAM else a for a in args]
        if timing is None:
            _check(fn(*args), name)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = fn(*args)
        e1.record(stream)
        timing.setdefault(name, []).append((e0, e1))
        _check(rc, name)


def ld_for(K: int) -> int:
    return (int(K) + 3) // 4 * 4


def version() -> int:
    return int(load().ttlda_version())


def check_device(device=0) -> None:
    """Raise unless `device` (index or torch.device) is an sm_100 GPU."""
    if isinstance(device, torch.device):
        if device.type != "cuda":
            raise TTLDAError(f"libttlda needs a CUDA device, got {device} (there is no CPU fallback)")
        device = device.index if device.index is not None else torch.cuda.current_device()
    _check(load().ttlda_check_device(int(device)), "ttlda_check_device")


f64, i64, i32 = torch.float64, torch.int64, torch.int32


# ---- thin wrappers ------------------------------------------------------------------------------------------------
def reduce_workspace_bytes() -> int:
    return int(load().ttlda_reduce_workspace_bytes())


def dirichlet_expectation(a, cols, out_elog=None, out_exp=None):
    rows, ld = a.shape
    _call("ttlda_dirichlet_expectation_f64", _p(a, f64), rows, int(cols), ld, _p(out_elog, f64), _p(out_exp, f64),
          _stream())


def dirichlet_expectation_1d(doc_topic, prior, out):
    _call("ttlda_dirichlet_expectation_1d_f64", _p(doc_topic, f64), float(prior), _p(out, f64), doc_topic.numel(),
          _stream())


def bow_to_csr_workspace_bytes(T, D, V) -> int:
    r = int(load().ttlda_bow_to_csr_workspace_bytes(int(T), int(D), int(V)))
    if r < 0:
        _check(r, "ttlda_bow_to_csr_workspace_bytes")
    return r


def bow_to_csr(token_ids, doc_offsets, V, rowptr, colidx, counts, nnz_out, ws):
    T, D = token_ids.numel(), doc_offsets.numel() - 1
    _call("ttlda_bow_to_csr", _p(token_ids, i32), _p(doc_offsets, i64), T, D, int(V), _p(rowptr, i64),
          _p(colidx, i32), _p(counts, i32), _p(nnz_out, i64), _p(ws), ws.numel() * ws.element_size(), _stream())


def csr_to_csc_workspace_bytes(nnz, D, V, nblocks=1) -> int:
    r = int(load().ttlda_csr_to_csc_workspace_bytes(int(nnz), int(D), int(V), int(nblocks)))
    if r < 0:
        _check(r, "ttlda_csr_to_csc_workspace_bytes")
    return r


def csr_to_csc(rowptr, colidx, counts, D, V, nnz, nblocks, block_docs, colptr, docidx, ccounts, ws, csr2csc=None):
    _call("ttlda_csr_to_csc", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), int(D), int(V), int(nnz),
          int(nblocks), int(block_docs), _p(colptr, i64), _p(docidx, i32), _p(ccounts, i32), _p(csr2csc, torch.int32),
          _p(ws), ws.numel() * ws.element_size(), _stream())


BAD_ROWPTR, BAD_COLIDX, BAD_COUNT = 1, 2, 4


def widen_i32(src, dst, bound=0, flags=None):
    """uint8 / int16-or-uint16 bit patterns (unsigned) -> int32, elementwise.  bound > 0: items >= bound are stored as 0
    and reported in flags (int32[1], BAD_COLIDX)."""
    if src.element_size() not in (1, 2) or src.numel() != dst.numel():
        raise TTLDAError("widen_i32: source must hold 1- or 2-byte items, one per destination item")
    _call("ttlda_widen_i32", _p(src), src.element_size(), _p(dst, i32), src.numel(), int(bound), _p(flags, i32),
          _stream())


def validate_csr(rowptr, colidx, counts, V, flags, pos_base=0, sanitize=False, nnz=None):
    """OR the BAD_* bits of a CSR corpus (or, with rowptr None, of a token-id stream) into flags (int32[1], device)."""
    D = rowptr.numel() - 1 if rowptr is not None else 0
    n = int(nnz) if nnz is not None else (colidx.numel() if colidx is not None else 0)
    _call("ttlda_validate_csr", _p(rowptr, i64), D, int(pos_base), _p(colidx, i32), _p(counts, i32), n, int(V),
          int(bool(sanitize)), _p(flags, i32), _stream())


def bad_corpus_message(flags: int) -> str:
    what = [m for b, m in ((BAD_ROWPTR, "rowptr must start at 0, be non-decreasing and end at nnz"),
                           (BAD_COLIDX, "word ids must lie in [0, V)"), (BAD_COUNT, "counts must be >= 0")) if flags & b]
    return "malformed corpus: " + "; ".join(what)


def csr_to_csc_block(rowptr_block, colidx, counts, V, doc_base, pos_base, nnz_block, colptr_block, docidx, ccounts, ws,
                     csr2csc=None):
    """One document block of a blocked mirror; colidx/counts/docidx/ccounts/csr2csc are the corpus-wide arrays."""
    _call("ttlda_csr_to_csc_block", _p(rowptr_block, i64), _p(colidx, i32), _p(counts, i32), rowptr_block.numel() - 1,
          int(V), int(doc_base), int(pos_base), int(nnz_block), _p(colptr_block, i64), _p(docidx, i32),
          _p(ccounts, i32), _p(csr2csc, torch.int32), _p(ws), ws.numel() * ws.element_size(), _stream())


def theta_refresh(gamma, K, alpha_vec, alpha_sum, eT, partial, ws):
    D, ld = gamma.shape
    _call("ttlda_theta_refresh_f64", _p(gamma, f64), D, int(K), ld, _p(alpha_vec, f64), float(alpha_sum), _p(eT, f64),
          _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def estep(rowptr, colidx, counts, K, V, alpha_vec, alpha_sum, eT_in, eB, gamma_in, gamma_out, eT_out,
          max_iter, tol, iters, partial, ws, csr2csc=None, ratio_csc=None):
    D, ld = gamma_out.shape
    _call("ttlda_estep_f64", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), D, int(K), int(V), ld,
          _p(alpha_vec, f64), float(alpha_sum), _p(eT_in, f64), _p(eB, f64), _p(gamma_in, f64), _p(gamma_out, f64),
          _p(eT_out, f64), _p(csr2csc, torch.int32), _p(ratio_csc, f64), int(max_iter), float(tol), _p(iters, i64), _p(partial, f64), _p(ws),
          ws.numel() * ws.element_size(), _stream())


def estep_hot_rows(ld) -> int:
    return max(0, int(load().ttlda_estep_hot_rows(int(ld))))


def build_estream(rowptr, colidx, counts, csr2csc, V, slot_of_word, e_idx, e_cnt, e_pos):
    _call("ttlda_build_estream", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), _p(csr2csc, torch.int32),
          rowptr.numel() - 1, int(V), _p(slot_of_word, i32), _p(e_idx, i32), _p(e_cnt, i32), _p(e_pos, torch.int32),
          _stream())


def estep_hot(rowptr, e_idx, e_cnt, e_pos, hot_words, K, V, alpha_vec, eT_in, eB, gamma_in, gamma_out, ratio_csc,
              max_iter, tol, iters, partial, ws):
    D, ld = gamma_out.shape
    _call("ttlda_estep_hot_f64", _p(rowptr, i64), _p(e_idx, i32), _p(e_cnt, i32), _p(e_pos, torch.int32),
          _p(hot_words, i32), hot_words.numel(), D, int(K), int(V), ld, _p(alpha_vec, f64), _p(eT_in, f64), _p(eB, f64),
          _p(gamma_in, f64), _p(gamma_out, f64), _p(ratio_csc, f64), int(max_iter), float(tol), _p(iters, i64),
          _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def estep_fixed_point(rowptr, colidx, counts, K, V, alpha_vec, alpha_sum, eT_in, eB, gamma_in, gamma_out, eT_out,
                      max_iter, tol, iters, partial, ws):
    D, ld = gamma_out.shape
    _call("ttlda_estep_fixed_point_f64", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), D, int(K), int(V), ld,
          _p(alpha_vec, f64), float(alpha_sum), _p(eT_in, f64), _p(eB, f64), _p(gamma_in, f64), _p(gamma_out, f64),
          _p(eT_out, f64), int(max_iter), float(tol), _p(iters, i64), _p(partial, f64), _p(ws),
          ws.numel() * ws.element_size(), _stream())


def sstats_workspace_bytes(nnz, V, ld, nblocks=1) -> int:
    return int(load().ttlda_sstats_workspace_bytes(int(nnz), int(V), int(ld), int(nblocks)))


def sstats_csc(colptr, docidx, ccounts, V, K, nnz, nblocks, eT, eB, sstats, ws, ratio_csc=None):
    global kernel_count
    ld = eT.shape[1]
    kernel_count += 1 if nblocks > 1 else 0  # + sstats_fold_kernel
    _call("ttlda_sstats_csc_f64", _p(colptr, i64), _p(docidx, i32), _p(ccounts, i32), eT.shape[0], int(V), int(K), ld,
          int(nnz),
          int(nblocks), _p(eT, f64), _p(eB, f64), _p(ratio_csc, f64), _p(sstats, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def sstats_windowed_workspace_bytes(nnz, V, kw, nblocks=1) -> int:
    r = int(load().ttlda_sstats_windowed_workspace_bytes(int(nnz), int(V), int(kw), int(nblocks)))
    if r < 0:
        _check(r, "ttlda_sstats_windowed_workspace_bytes")
    return r


def sstats_csc_windowed(colptr, docidx, V, K, nnz, nblocks, kw, eT, ratio_csc, sstats, ws):
    """Statistics from the E-step's ratios, one topic window of `kw` columns at a time (long K-vectors)."""
    global kernel_count
    ld = eT.shape[1]
    kernel_count += -(-ld // int(kw)) * (3 if nblocks > 1 else 2)  # per window: pass + carry (+ fold)
    _call("ttlda_sstats_csc_windowed_f64", _p(colptr, i64), _p(docidx, i32), eT.shape[0], int(V), int(K), ld, int(nnz),
          int(nblocks), int(kw), _p(eT, f64), _p(ratio_csc, f64), _p(sstats, f64), _p(ws), ws.numel() * ws.element_size(),
          _stream())


def sstats_csc_block(colptr_block, docidx, ccounts, V, K, pos_base, nnz_block, eT, eB, sstats_block, ws, ratio_csc=None,
                     accumulate=False):
    """Statistics of one document block ([V][ld] partial, or added to sstats_block when accumulate); docidx / ccounts /
    ratio_csc / eT are the corpus-wide arrays."""
    ld = eT.shape[1]
    _call("ttlda_sstats_csc_block_f64", _p(colptr_block, i64), _p(docidx, i32), _p(ccounts, i32), eT.shape[0], int(V),
          int(K), ld, int(pos_base), int(nnz_block), _p(eT, f64), _p(eB, f64), _p(ratio_csc, f64), _p(sstats_block, f64),
          int(bool(accumulate)), _p(ws), ws.numel() * ws.element_size(), _stream())


def sstats_fold(sstats_blocks, sstats):
    nb, V, ld = sstats_blocks.shape
    _call("ttlda_sstats_fold_f64", _p(sstats_blocks, f64), nb, V, ld, _p(sstats, f64), _stream())


def sstats_atomic(rowptr, colidx, counts, K, V, eT, eB, sstats):
    D, ld = eT.shape
    _call("ttlda_sstats_atomic_f64", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), D, int(K), int(V), ld,
          _p(eT, f64), _p(eB, f64), _p(sstats, f64), _stream())


def mstep_workspace_bytes(V, ld) -> int:
    return int(load().ttlda_mstep_workspace_bytes(int(V), int(ld)))


def mstep(sstats, K, eta, eB, lam, elogbeta, colsum, partial, ws):
    V, ld = sstats.shape
    _call("ttlda_mstep_f64", _p(sstats, f64), V, int(K), ld, float(eta), _p(eB, f64), _p(lam, f64), _p(elogbeta, f64),
          _p(colsum, f64), _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def mstep_online(sstats, K, eta, scale, rho, eB, lam, elogbeta, colsum, partial, ws):
    V, ld = sstats.shape
    _call("ttlda_mstep_online_f64", _p(sstats, f64), V, int(K), ld, float(eta), float(scale), float(rho), _p(eB, f64),
          _p(lam, f64), _p(elogbeta, f64), _p(colsum, f64), _p(partial, f64), _p(ws), ws.numel() * ws.element_size(),
          _stream())


def _ptr_array(ptrs):
    """Host array of device pointers (ints) for the peer-table arguments; returns (ctypes array, device index or None)."""
    arr = (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    return arr


def mstep_slice_lambda(src_ptrs, src_row0, w0, n, K, eta, eB, lam, colsum_part, ws, multicast=0):
    """Phase A of the sharded M-step on rows [w0, w0+n): src_ptrs = device addresses (ints) of the statistics tables to
    sum in order — peer-mapped tables of every rank, or one locally reduced buffer; multicast = their NVSwitch multicast
    address (0: none), in which case the sum is read with multimem.ld_reduce."""
    ld = eB.shape[1]
    arr = _ptr_array(src_ptrs)
    _call("ttlda_mstep_slice_lambda_f64", ctypes.cast(arr, ctypes.c_void_p), len(src_ptrs),
          ctypes.c_void_p(int(multicast)) if multicast else None, int(src_row0), int(w0),
          int(n), int(K), ld, float(eta), _p(eB, f64), _p(lam, f64), _p(colsum_part, f64), _p(ws),
          ws.numel() * ws.element_size(), _stream())


def mstep_slice_beta(colsum_parts, w0, n, K, eta, lam, dst_ptrs, add_colsum_term, colsum, partial, ws, multicast=0):
    """Phase B: colsum_parts [nparts][ld] (rank order) -> colsum; exp(E[log beta]) rows [w0, w0+n) stored into every table
    of dst_ptrs (device addresses: this rank's table, or all ranks' through peer memory)."""
    nparts, ld = colsum_parts.shape
    arr = _ptr_array(dst_ptrs)
    _call("ttlda_mstep_slice_beta_f64", _p(colsum_parts, f64), nparts, int(w0), int(n), int(K), ld, float(eta),
          _p(lam, f64), ctypes.cast(arr, ctypes.c_void_p), len(dst_ptrs),
          ctypes.c_void_p(int(multicast)) if multicast else None, int(bool(add_colsum_term)), _p(colsum, f64),
          _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def beta_refresh(lam, K, eta, eB, elogbeta, colsum, partial, ws):
    V, ld = lam.shape
    _call("ttlda_beta_refresh_f64", _p(lam, f64), V, int(K), ld, float(eta), _p(eB, f64), _p(elogbeta, f64),
          _p(colsum, f64), _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def elbo_tokens(rowptr, colidx, counts, K, eT, eB, partial, ws):
    D, ld = eT.shape
    _call("ttlda_elbo_tokens_f64", _p(rowptr, i64), _p(colidx, i32), _p(counts, i32), D, int(K), eB.shape[0], ld,
          _p(eT, f64),
          _p(eB, f64), _p(partial, f64), _p(ws), ws.numel() * ws.element_size(), _stream())


def hyper_stats(x, cols, partial, ws):
    rows, ld = x.shape
    _call("ttlda_hyper_stats_f64", _p(x, f64), rows, int(cols), ld, _p(partial, f64), _p(ws),
          ws.numel() * ws.element_size(), _stream())


def hyper_colstats(x, cols, out):
    """out[k] = sum_r (psi(x[r,k]) - psi(sum_j x[r,j])), exact digamma; allocates its own scratch."""
    rows, ld = x.shape
    n = int(load().ttlda_hyper_colstats_workspace_bytes(int(rows), int(cols)))
    if n < 0:
        _check(n, "ttlda_hyper_colstats_workspace_bytes")
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    _call("ttlda_hyper_colstats_f64", _p(x, f64), rows, int(cols), ld, _p(out, f64), _p(ws), n, _stream())


def psi_exact(x, out):
    _call("ttlda_psi_f64", _p(x, f64), x.numel(), _p(out, f64), _stream())


def probe_gather(table, idx, sink):
    """Loads-only gather of table[idx[i], :] with the product traversal (roofline probe, bench.py)."""
    rows, ld = table.shape
    _call("ttlda_probe_gather_f64", _p(table, f64), rows, ld, _p(idx, i32), idx.numel(), _p(sink, f64), _stream())


def transpose_pad(src, dst):
    """dst[c, r] = src[r, c];  src (rows, cols) contiguous, dst (cols, dst_ld>=rows), padding zeroed."""
    rows, cols = src.shape
    assert dst.shape[0] == cols
    _call("ttlda_transpose_pad_f64", _p(src, f64), rows, cols, _p(dst, f64), dst.shape[1], _stream())


def transpose_unpad(src, cols, dst):
    """dst[c, r] = src[r, c] for c < cols;  src (rows, src_ld), dst (cols, rows)."""
    rows, src_ld = src.shape
    assert dst.shape == (cols, rows)
    _call("ttlda_transpose_unpad_f64", _p(src, f64), rows, int(cols), src_ld, _p(dst, f64), _stream())
