"""ctypes binding of libtvc.so (include/tvc.h).  There is no CPU fallback: every compute entry
point raises when the library or an sm_100 device is missing."""
from __future__ import annotations

import ctypes as C
import math
import threading
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libtvc.so"

TVC_OK, TVC_ERR_INVALID, TVC_ERR_CUDA, TVC_ERR_NO_DEVICE, TVC_ERR_UNSUPPORTED, TVC_ERR_OOM = range(6)
TVC_F32, TVC_BF16, TVC_F16 = 0, 1, 2
GALLERY_NORMALIZE, GALLERY_NO_MASTER = 1, 2
SEARCH_NORMALIZE_Q, SEARCH_SKIP_SELF, SEARCH_PREPARED_Q = 1, 2, 4
MAX_K, MAX_VARIANTS, MAX_REFS, NSCORES = 56, 16, 16, 24
FLAG_DET_ADV, FLAG_CC_ADV, FLAG_SIGMA_ADV = 1, 2, 4

SCORE_NAMES = (
    "original_similarity", "text_variant_consistency", "text_variant_std", "text_variant_min",
    "text_variant_var", "retrieval_consistency", "retrieval_std", "generative_consistency",
    "generative_std", "generative_max", "cross_modal_variance", "cross_variant_mean",
    "cross_variant_min", "cross_variant_var", "det_text_variants", "det_sd_reference",
    "det_consistency", "aggregated_score", "overall_score", "threshold", "confidence",
    "n_retrieval", "n_generative", "reference_sigma")
SCORE_INDEX = {n: i for i, n in enumerate(SCORE_NAMES)}


class Scatter(C.Structure):
    """tvc_scatter: where the candidates of a sharded search go (include/tvc.h)."""
    _fields_ = [("n_slices", C.c_int32), ("slot", C.c_int32), ("rows_per_slice", C.c_int64),
                ("val", C.c_void_p * 16), ("idx", C.c_void_p * 16)]


class TvcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libtvc status {status}: {message}")
        self.status = status


class DetectorParams(C.Structure):
    _fields_ = [
        ("n_variants", C.c_int32), ("n_retrieval", C.c_int32), ("n_generative", C.c_int32),
        ("methods", C.c_uint32), ("aggregation", C.c_int32),
        ("w_text_variants", C.c_float), ("w_sd_reference", C.c_float), ("w_consistency", C.c_float),
        ("detection_threshold", C.c_float), ("voting", C.c_int32), ("cc_weights", C.c_float * 4),
        ("cc_base_threshold", C.c_float), ("cc_adaptive", C.c_int32),
        ("dedup_threshold", C.c_float), ("sigma_threshold", C.c_float),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "cc_weights"}
        d["cc_weights"] = tuple(self.cc_weights)
        return d


_EXPORTS = {
    # name: (restype, argtypes)
    "tvc_version": (C.c_int, []),
    "tvc_status_string": (C.c_char_p, [C.c_int]),
    "tvc_detector_params_default": (None, [C.POINTER(DetectorParams)]),
    "tvc_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "tvc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "tvc_last_error": (C.c_char_p, [C.c_void_p]),
    "tvc_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "tvc_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "tvc_ctx_release_workspace": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "tvc_ctx_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "tvc_ctx_last_search_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "tvc_gallery_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64,
                                     C.c_uint32, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tvc_gallery_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "tvc_gallery_truncate": (C.c_int, [C.c_void_p, C.c_int64]),
    "tvc_gallery_move_row": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "tvc_gallery_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]),
    "tvc_gallery_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_void_p)]),
    "tvc_gallery_get_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tvc_gallery_destroy": (C.c_int, [C.c_void_p]),
    "tvc_gallery_wrap_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                       C.POINTER(C.c_void_p)]),
    "tvc_gallery_export_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tvc_gallery_import_ipc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                         C.POINTER(C.c_void_p)]),
    "tvc_gallery_group_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]),
    "tvc_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int32,
                             C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvc_similarity_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32,
                                        C.c_uint32, C.c_void_p, C.c_void_p]),
    "tvc_merge_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvc_consistency_sims": (C.c_int, [C.c_void_p, C.POINTER(DetectorParams), C.c_int64] + [C.c_void_p] * 10),
    "tvc_consistency_emb": (C.c_int, [C.c_void_p, C.POINTER(DetectorParams), C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p,          # img txt var
                                      C.c_void_p, C.c_void_p, C.c_int32,           # ret gallery idx ncand
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,  # gen g_cnt gen_gallery gen_idx ncand
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,  # scores flags sv sr sg
                                      C.c_void_p]),
    "tvc_candidate_width": (C.c_int, [C.c_int32]),
    "tvc_query_row_bytes": (C.c_int, [C.c_int32]),
    "tvc_prepare_queries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_uint32, C.c_int32,
                                      C.POINTER(C.c_void_p), C.c_int64, C.c_void_p]),
    "tvc_search_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int32,
                                        C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvc_rerank_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "tvc_exchange_merge": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.POINTER(C.c_void_p), C.c_void_p]),
    "tvc_exchange_rescore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                       C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "tvc_exchange_finalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvc_peer_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tvc_peer_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "tvc_peer_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tvc_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tvc_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tvc_reference_vector_rule": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tvc_retrieval_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p]),
    "tvc_k_occurrence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_EXPORTS)

_lib = None
_lib_lock = threading.Lock()


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libtvc.so (building it with nvcc when absent) and type every entry point."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            if not build_if_missing:
                raise TvcError(TVC_ERR_NO_DEVICE, f"{LIB_PATH} is missing; run __graft_entry__.build()")
            from .build import build
            build()
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def default_params(**overrides) -> DetectorParams:
    p = DetectorParams()
    load_library().tvc_detector_params_default(C.byref(p))
    for k, v in overrides.items():
        if k == "cc_weights":
            for i, w in enumerate(v):
                p.cc_weights[i] = w
        else:
            setattr(p, k, v)
    return p


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


_TORCH_DTYPES = None


def _dtype_code(x) -> int:
    global _TORCH_DTYPES
    if _is_torch(x):
        import torch
        if _TORCH_DTYPES is None:
            _TORCH_DTYPES = {torch.float32: TVC_F32, torch.bfloat16: TVC_BF16, torch.float16: TVC_F16}
        return _TORCH_DTYPES[x.dtype]
    if x.dtype == np.float32:
        return TVC_F32
    if x.dtype == np.float16:
        return TVC_F16
    raise TypeError(f"unsupported dtype {x.dtype}")


def _rows(x):
    """Contiguous float rows as numpy (host) or torch (host/device) without changing residence."""
    if _is_torch(x):
        import torch
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        return x.contiguous()
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float16):
        a = a.astype(np.float32)
    return np.ascontiguousarray(a)


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def _stream_of(*tensors) -> Optional[int]:
    for t in tensors:
        if t is not None and _is_torch(t) and t.is_cuda:
            import torch
            return torch.cuda.current_stream(t.device).cuda_stream
    return None


class Context:
    """One libtvc context per CUDA device."""

    _by_device = {}
    _lock = threading.Lock()

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.tvc_ctx_create(self.device, C.byref(h))
        if rc != TVC_OK:
            raise TvcError(rc, self.lib.tvc_status_string(rc).decode())
        self.handle = h

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            try:
                import torch
                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            except Exception:
                device = 0
        with cls._lock:
            ctx = cls._by_device.get(device)
            if ctx is None:
                ctx = cls._by_device[device] = Context(device)
            return ctx

    @classmethod
    def release_all_workspaces(cls) -> int:
        """release_workspace() on every context that already exists (never creates one, never raises:
        it is called from the mirrors' clear_cache(), which the reference's pipeline invokes)."""
        with cls._lock:
            live = list(cls._by_device.values())
        freed = 0
        for ctx in live:
            try:
                freed += ctx.release_workspace()
            except TvcError:
                pass
        return freed

    def check(self, rc: int):
        if rc != TVC_OK:
            msg = self.lib.tvc_last_error(self.handle)
            raise TvcError(rc, (msg.decode() if msg else "") or self.lib.tvc_status_string(rc).decode())

    # -- bookkeeping ---------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.tvc_ctx_launch_count(self.handle))

    def set_option(self, name: str, value: int):
        self.check(self.lib.tvc_ctx_set_option(self.handle, name.encode(), int(value)))

    def release_workspace(self) -> int:
        """Frees the grow-only device workspaces (re-grown on demand); returns the bytes released."""
        n = C.c_int64()
        self.check(self.lib.tvc_ctx_release_workspace(self.handle, C.byref(n)))
        return int(n.value)

    def set_timing(self, enabled: bool):
        self.check(self.lib.tvc_ctx_set_timing(self.handle, int(enabled)))

    def search_kernel_ms(self) -> Tuple[float, int]:
        ms, n = C.c_float(), C.c_int64()
        self.check(self.lib.tvc_ctx_last_search_kernel_ms(self.handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # -- kernel (a) helpers ----------------------------------------------------------------
    def merge_topk(self, sims, idx, k: int):
        """sims/idx [m, parts, k] (torch cuda or numpy) -> ([m,k] sims, [m,k] idx)."""
        m, parts = int(sims.shape[0]), int(sims.shape[1])
        if _is_torch(sims):
            import torch
            sims = sims.contiguous().float()
            idx = idx.contiguous().to(torch.int64)
            out_s = torch.empty((m, k), dtype=torch.float32, device=sims.device)
            out_i = torch.empty((m, k), dtype=torch.int64, device=sims.device)
        else:
            sims = np.ascontiguousarray(sims, dtype=np.float32)
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            out_s = np.empty((m, k), np.float32)
            out_i = np.empty((m, k), np.int64)
        self.check(self.lib.tvc_merge_topk(self.handle, _ptr(sims), _ptr(idx), m, parts, k, _ptr(out_s),
                                           _ptr(out_i), _stream_of(sims)))
        return out_s, out_i

    # -- sharded search (peer memory) --------------------------------------------------------
    def peer_alloc(self, nbytes: int) -> Tuple[int, bytes]:
        """Zero-filled device buffer shareable with the other ranks: (device pointer, 64-byte IPC handle)."""
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self.check(self.lib.tvc_peer_alloc(self.handle, int(nbytes), C.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self.check(self.lib.tvc_peer_open(self.handle, C.create_string_buffer(handle, 64), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int):
        self.check(self.lib.tvc_peer_close(self.handle, C.c_void_p(ptr)))

    def peer_free(self, ptr: int):
        self.check(self.lib.tvc_peer_free(self.handle, C.c_void_p(ptr)))

    def peer_copy(self, dst: int, src: int, nbytes: int, stream: Optional[int] = None):
        self.check(self.lib.tvc_peer_copy(self.handle, C.c_void_p(int(dst)), C.c_void_p(int(src)), int(nbytes), stream))

    def exchange_merge(self, m: int, parts: int, kp: int, cand_val: int, cand_idx: int, req_dst, stream):
        arr = (C.c_void_p * len(req_dst))(*[int(p) for p in req_dst])
        self.check(self.lib.tvc_exchange_merge(self.handle, int(m), int(parts), int(kp), C.c_void_p(cand_val),
                                               C.c_void_p(cand_idx), len(req_dst), arr, stream))

    def exchange_rescore(self, shard, q_f32: int, d: int, owners: int, rows_per_slice: int, m_total: int, kp: int,
                         req: int, score_dst, stream):
        arr = (C.c_void_p * len(score_dst))(*[int(p) for p in score_dst])
        self.check(self.lib.tvc_exchange_rescore(self.handle, shard.handle, C.c_void_p(q_f32), int(d), int(owners),
                                                 int(rows_per_slice), int(m_total), int(kp), C.c_void_p(req), arr, stream))

    def exchange_finalize(self, m: int, kp: int, k: int, threshold: float, req: int, score: int, device):
        import torch
        sims = torch.empty((m, k), dtype=torch.float32, device=device)
        idx = torch.empty((m, k), dtype=torch.int64, device=device)
        self.check(self.lib.tvc_exchange_finalize(self.handle, int(m), int(kp), int(k), float(threshold), C.c_void_p(req),
                                                  C.c_void_p(score), _ptr(sims), _ptr(idx),
                                                  torch.cuda.current_stream(device).cuda_stream))
        return sims, idx

    def query_row_bytes(self, d: int) -> int:
        return int(self.lib.tvc_query_row_bytes(int(d)))

    def prepare_queries(self, rows, dst_ptrs, dst_row0: int = 0, normalize: bool = False):
        """rows [m, d] cuda tensor -> bf16 GEMM operand rows written at row dst_row0 of every buffer in
        dst_ptrs (device pointers: own or peer query buffers)."""
        q = _rows(rows)
        arr = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
        self.check(self.lib.tvc_prepare_queries(self.handle, _ptr(q), _dtype_code(q), int(q.shape[0]), int(q.shape[1]),
                                                SEARCH_NORMALIZE_Q if normalize else 0, len(dst_ptrs), arr,
                                                int(dst_row0), _stream_of(q)))

    def candidate_width(self, k: int) -> int:
        return int(self.lib.tvc_candidate_width(int(k)))

    def rerank_candidates(self, gallery, queries, cand_val: int, cand_idx: int, parts: int, kp: int, k: int,
                          threshold: float = -math.inf):
        """Phase 2 of the sharded search: queries [m, d] fp32 cuda; cand_val / cand_idx are device
        pointers of the [parts, m, kp] receive buffers.  Returns (sims [m,k], idx [m,k]) cuda tensors."""
        import torch
        q = queries.contiguous().to(torch.float32)
        m = int(q.shape[0])
        sims = torch.empty((m, k), dtype=torch.float32, device=q.device)
        idx = torch.empty((m, k), dtype=torch.int64, device=q.device)
        self.check(self.lib.tvc_rerank_candidates(self.handle, gallery.handle, _ptr(q), m, int(q.shape[1]), int(parts),
                                                  int(kp), C.c_void_p(cand_val), C.c_void_p(cand_idx), int(k),
                                                  float(threshold), _ptr(sims), _ptr(idx), _stream_of(q)))
        return sims, idx

    def reference_vector_rule(self, img, ret_gallery=None, ret_idx=None, gen=None, sigma_threshold: float = 0.30,
                              want_ref: bool = True):
        """README.md:474-482: img [Q,d]; ret_idx [Q,V,k] into ret_gallery and/or gen [Q,V,m,d].
        Returns (s [Q,V], ref_sim [Q], sigma [Q], flags [Q]); numpy in/out or torch cuda in/out."""
        q, d = int(img.shape[0]), int(img.shape[1])
        torch_mode = _is_torch(img)
        if ret_idx is not None:
            v, k = int(ret_idx.shape[1]), int(ret_idx.shape[2])
        else:
            v, k = int(gen.shape[1]), 0
        m = int(gen.shape[2]) if gen is not None else 0
        if torch_mode:
            import torch
            img = img.contiguous().float()
            ret_idx = ret_idx.contiguous().to(torch.int64) if ret_idx is not None else None
            gen = gen.contiguous().float() if gen is not None else None
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=img.device)  # noqa: E731
            s, ref, sig, fl = mk((q, v), torch.float32), mk((q,), torch.float32), mk((q,), torch.float32), mk((q,), torch.uint8)
        else:
            img = np.ascontiguousarray(img, np.float32)
            ret_idx = np.ascontiguousarray(ret_idx, np.int64) if ret_idx is not None else None
            gen = np.ascontiguousarray(gen, np.float32) if gen is not None else None
            s, ref, sig, fl = np.empty((q, v), np.float32), np.empty(q, np.float32), np.empty(q, np.float32), np.empty(q, np.uint8)
        self.check(self.lib.tvc_reference_vector_rule(
            self.handle, q, d, v, _ptr(img), ret_gallery.handle if ret_gallery is not None else None, _ptr(ret_idx), k,
            _ptr(gen), m, float(sigma_threshold), _ptr(s), _ptr(ref) if want_ref else None, _ptr(sig), _ptr(fl),
            _stream_of(img)))
        return s, (ref if want_ref else None), sig, fl

    # -- retrieval metrics ---------------------------------------------------------------------
    def retrieval_metrics(self, topk_idx, rel_ptr, rel_idx, k_values):
        """Per-query [rr, ap, recall@K.., precision@K.., ndcg@K..] from ranked lists topk_idx [q, k] and a
        CSR relevance set (rel_ptr [q+1], rel_idx).  numpy in -> numpy out, torch cuda in -> torch cuda out."""
        q, k = int(topk_idx.shape[0]), int(topk_idx.shape[1])
        ks = (C.c_int32 * len(k_values))(*[int(x) for x in k_values])
        cols = 2 + 3 * len(k_values)
        if _is_torch(topk_idx):
            import torch
            topk_idx, rel_ptr, rel_idx = (t.contiguous().to(torch.int64) for t in (topk_idx, rel_ptr, rel_idx))
            out = torch.empty((q, cols), dtype=torch.float32, device=topk_idx.device)
            n_rel = int(rel_idx.numel())
        else:
            topk_idx, rel_ptr, rel_idx = (np.ascontiguousarray(t, dtype=np.int64) for t in (topk_idx, rel_ptr, rel_idx))
            out = np.empty((q, cols), np.float32)
            n_rel = int(rel_idx.size)
        self.check(self.lib.tvc_retrieval_metrics(self.handle, _ptr(topk_idx), q, k, _ptr(rel_ptr),
                                                  _ptr(rel_idx) if n_rel else None, n_rel, ks, len(k_values),
                                                  _ptr(out), _stream_of(topk_idx)))
        return out

    # -- kernel (c) ------------------------------------------------------------------------
    def k_occurrence(self, idx, n_bins: int, idx_base: int = 0, counts=None):
        """idx [m, k] int64 (torch cuda or numpy) -> counts [n_bins] int32 (same residence)."""
        k = int(idx.shape[-1]) if idx.ndim > 1 else 1
        m = int(idx.numel() // k) if _is_torch(idx) else int(idx.size // k)
        zero = counts is None
        if _is_torch(idx):
            import torch
            idx = idx.contiguous().to(torch.int64)
            if counts is None:
                counts = torch.empty((n_bins,), dtype=torch.int32, device=idx.device)
        else:
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            if counts is None:
                counts = np.empty((n_bins,), np.int32)
        self.check(self.lib.tvc_k_occurrence(self.handle, _ptr(idx), m, k, int(idx_base), int(n_bins),
                                             _ptr(counts), int(zero), _stream_of(idx, counts)))
        return counts

    # -- kernel (b) ------------------------------------------------------------------------
    def consistency_sims(self, params: DetectorParams, s0, sv=None, sr=None, r_cnt=None, sg=None,
                         g_cnt=None, sxv=None):
        q = int(s0.shape[0])
        torch_mode = _is_torch(s0)

        def prep(x, dt):
            if x is None:
                return None
            if torch_mode:
                import torch
                return x.contiguous().to(torch.float32 if dt == "f" else torch.int32)
            return np.ascontiguousarray(x, dtype=np.float32 if dt == "f" else np.int32)

        s0, sv, sr, sg, sxv = (prep(x, "f") for x in (s0, sv, sr, sg, sxv))
        r_cnt, g_cnt = prep(r_cnt, "i"), prep(g_cnt, "i")
        if torch_mode:
            import torch
            scores = torch.empty((q, NSCORES), dtype=torch.float32, device=s0.device)
            flags = torch.empty((q,), dtype=torch.uint8, device=s0.device)
        else:
            scores = np.empty((q, NSCORES), np.float32)
            flags = np.empty((q,), np.uint8)
        self.check(self.lib.tvc_consistency_sims(self.handle, C.byref(params), q, _ptr(s0), _ptr(sv), _ptr(sr),
                                                 _ptr(r_cnt), _ptr(sg), _ptr(g_cnt), _ptr(sxv), _ptr(scores),
                                                 _ptr(flags), _stream_of(s0)))
        return scores, flags

    def consistency_emb(self, params: DetectorParams, img, txt, var=None, ret_gallery=None, ret_idx=None,
                        gen=None, g_cnt=None, gen_gallery=None, gen_idx=None, return_sims: bool = False):
        q, d = int(img.shape[0]), int(img.shape[1])
        torch_mode = _is_torch(img)

        def prep(x, dt):
            if x is None:
                return None
            if torch_mode:
                import torch
                return x.contiguous().to({"f": torch.float32, "i": torch.int32, "l": torch.int64}[dt])
            return np.ascontiguousarray(x, dtype={"f": np.float32, "i": np.int32, "l": np.int64}[dt])

        img, txt, var, gen = (prep(x, "f") for x in (img, txt, var, gen))
        ret_idx, gen_idx, g_cnt = prep(ret_idx, "l"), prep(gen_idx, "l"), prep(g_cnt, "i")
        n_rc = int(ret_idx.shape[1] * (ret_idx.shape[2] if ret_idx.ndim == 3 else 1)) if ret_idx is not None else 0
        n_gc = int(gen_idx.shape[1] * (gen_idx.shape[2] if gen_idx.ndim == 3 else 1)) if gen_idx is not None else 0
        V, R, G = params.n_variants, params.n_retrieval, params.n_generative
        if var is not None and int(var.shape[1]) != V:
            raise ValueError(f"var has {var.shape[1]} variants, params.n_variants = {V}")
        if gen is not None and int(gen.shape[1]) != G:
            raise ValueError(f"gen has {gen.shape[1]} references, params.n_generative = {G}")

        def empty(shape, dt):
            if torch_mode:
                import torch
                return torch.empty(shape, dtype=dt[0], device=img.device)
            return np.empty(shape, dt[1])

        if torch_mode:
            import torch
            f32, u8 = (torch.float32, np.float32), (torch.uint8, np.uint8)
        else:
            f32, u8 = (None, np.float32), (None, np.uint8)
        scores, flags = empty((q, NSCORES), f32), empty((q,), u8)
        sv = empty((q, V), f32) if return_sims else None
        sr = empty((q, R), f32) if return_sims else None
        sg = empty((q, G), f32) if return_sims else None
        self.check(self.lib.tvc_consistency_emb(
            self.handle, C.byref(params), q, d, _ptr(img), _ptr(txt), _ptr(var),
            ret_gallery.handle if ret_gallery is not None else None, _ptr(ret_idx), n_rc,
            _ptr(gen), _ptr(g_cnt), gen_gallery.handle if gen_gallery is not None else None, _ptr(gen_idx), n_gc,
            _ptr(scores), _ptr(flags), _ptr(sv), _ptr(sr), _ptr(sg), _stream_of(img)))
        if return_sims:
            return scores, flags, (sv, sr, sg)
        return scores, flags


class Gallery:
    """HBM-resident row store searched by kernel (a) (the FAISS IndexFlatIP replacement)."""

    def __init__(self, rows=None, dim: Optional[int] = None, *, normalize: bool = False,
                 keep_master: bool = True, global_row_offset: int = 0, capacity: int = 0,
                 ctx: Optional[Context] = None, device: Optional[int] = None):
        if rows is None and dim is None:
            raise ValueError("need rows or dim")
        if rows is not None:
            rows = _rows(rows)
            if rows.ndim != 2:
                raise ValueError("gallery rows must be [N, d]")
            dim = int(rows.shape[1])
            if _is_torch(rows) and rows.is_cuda and device is None and ctx is None:
                device = rows.device.index
        self.ctx = ctx or Context.get(device)
        self.dim = int(dim)
        self.flags = (GALLERY_NORMALIZE if normalize else 0) | (0 if keep_master else GALLERY_NO_MASTER)
        h = C.c_void_p()
        n = int(rows.shape[0]) if rows is not None else 0
        self.ctx.check(self.ctx.lib.tvc_gallery_create(
            self.ctx.handle, _ptr(rows) if n else None, _dtype_code(rows) if n else TVC_F32, n, self.dim,
            int(global_row_offset), self.flags, int(capacity), _stream_of(rows), C.byref(h)))
        self.handle = h
        self.global_row_offset = int(global_row_offset)

    @classmethod
    def wrap_rows(cls, rows, global_row_offset: int = 0, ctx: Optional[Context] = None) -> "Gallery":
        """Non-owning view over fp32 CUDA rows [n, d] (keeps `rows` alive); not searchable."""
        if not (_is_torch(rows) and rows.is_cuda and rows.dim() == 2):
            raise ValueError("wrap_rows needs a 2-D CUDA tensor")
        import torch
        rows = rows.contiguous().to(torch.float32)
        self = cls.__new__(cls)
        self.ctx = ctx or Context.get(rows.device.index)
        self.dim = int(rows.shape[1])
        self.flags = 0
        self._keepalive = rows
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tvc_gallery_wrap_f32(self.ctx.handle, _ptr(rows), int(rows.shape[0]), self.dim,
                                                         int(global_row_offset), C.byref(h)))
        self.handle = h
        self.global_row_offset = int(global_row_offset)
        return self

    # -- row shards across the GPUs of one box ---------------------------------------------------
    def export_ipc(self) -> bytes:
        """64-byte CUDA IPC handle of the fp32 master (send it to the peer ranks)."""
        buf = C.create_string_buffer(64)
        self.ctx.check(self.ctx.lib.tvc_gallery_export_ipc(self.handle, buf))
        return buf.raw

    @classmethod
    def import_ipc(cls, handle: bytes, n: int, dim: int, global_row_offset: int, ctx: Context) -> "Gallery":
        """View over a PEER rank's fp32 master (read over NVLink by kernel (b)); not searchable."""
        self = cls.__new__(cls)
        self.ctx, self.dim, self.flags = ctx, int(dim), 0
        h = C.c_void_p()
        ctx.check(ctx.lib.tvc_gallery_import_ipc(ctx.handle, C.create_string_buffer(handle, 64), int(n), int(dim),
                                                 int(global_row_offset), C.byref(h)))
        self.handle = h
        self.global_row_offset = int(global_row_offset)
        return self

    @classmethod
    def group(cls, parts) -> "Gallery":
        """Group of row shards with disjoint global index ranges, usable as ret_gallery / gen_gallery."""
        parts = list(parts)
        ctx = parts[0].ctx
        arr = (C.c_void_p * len(parts))(*[p.handle for p in parts])
        self = cls.__new__(cls)
        self.ctx, self.dim, self.flags = ctx, parts[0].dim, 0
        self._keepalive = parts
        h = C.c_void_p()
        ctx.check(ctx.lib.tvc_gallery_group_create(ctx.handle, arr, len(parts), C.byref(h)))
        self.handle = h
        self.global_row_offset = 0
        return self

    @property
    def has_master(self) -> bool:
        """True when the fp32 master rows are resident (candidates can be re-scored in fp32)."""
        return not (self.flags & GALLERY_NO_MASTER)

    def __len__(self) -> int:
        n = C.c_int64()
        self.ctx.check(self.ctx.lib.tvc_gallery_info(self.handle, C.byref(n), None, None, None))
        return int(n.value)

    @property
    def ntotal(self) -> int:
        return len(self)

    def append(self, rows):
        rows = _rows(rows)
        if rows.ndim == 1:
            rows = rows.reshape(1, -1)
        if int(rows.shape[1]) != self.dim:
            raise ValueError(f"row dimension {rows.shape[1]} != gallery dimension {self.dim}")
        self.ctx.check(self.ctx.lib.tvc_gallery_append(self.handle, _ptr(rows), _dtype_code(rows),
                                                       int(rows.shape[0]), _stream_of(rows)))

    def _check_dim(self, q):
        """The reference raises on a query of the wrong width (np.dot / the FAISS assert); so do we, before a
        pointer with the wrong row stride reaches the library."""
        if int(q.shape[-1]) != self.dim:
            raise ValueError(f"query dimension {int(q.shape[-1])} != gallery dimension {self.dim}")

    def truncate(self, n: int):
        self.ctx.check(self.ctx.lib.tvc_gallery_truncate(self.handle, int(n)))

    def move_row(self, src: int, dst: int):
        self.ctx.check(self.ctx.lib.tvc_gallery_move_row(self.handle, int(src), int(dst), None))

    def get_rows(self, idx):
        """fp32 rows for LOCAL indices: numpy in -> numpy out, torch cuda in -> torch cuda out."""
        if _is_torch(idx) and idx.is_cuda:
            import torch
            idx = idx.contiguous().to(torch.int64)
            out = torch.empty((int(idx.shape[0]), self.dim), dtype=torch.float32, device=idx.device)
            self.ctx.check(self.ctx.lib.tvc_gallery_get_rows(self.handle, _ptr(idx), int(idx.shape[0]), _ptr(out),
                                                             _stream_of(idx)))
            return out
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty((idx.shape[0], self.dim), np.float32)
        self.ctx.check(self.ctx.lib.tvc_gallery_get_rows(self.handle, _ptr(idx), int(idx.shape[0]), _ptr(out), None))
        return out

    def search(self, queries, k: int, threshold: float = -math.inf, *, normalize_queries: bool = False,
               skip_self: bool = False):
        """Top-k by inner product, ordered (similarity desc, index asc).  numpy in -> numpy out (host round
        trip); torch cuda in -> torch cuda out.

        How exact: the candidates of a row are its KP best gallery rows by the bf16 tensor-core score
        (KP = 16 / 32 / 64 for k <= 10 / 26 / 56), and those are re-scored from the fp32 masters, so the
        returned similarities are fp32 and the order among the candidates is the fp32 order.  A true top-k
        row can only be missed if bf16 rounding (~1e-3 for unit rows) pushes it below bf16 rank KP, i.e. only
        among rows whose similarities lie within ~1e-3 of the k-th - the band north_star allows index
        disagreement in.  With keep_master=False (TVC_GALLERY_NO_MASTER) the scores are the bf16 ones.
        k > MAX_K (56) is served by a chunked fp32 similarity matrix + top-k (_search_wide) instead of the fused
        epilogue; the C entry point itself returns TVC_ERR_UNSUPPORTED for such k."""
        q = _rows(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        self._check_dim(q)
        lead = None
        if q.ndim > 2:
            lead = tuple(q.shape[:-1])
            q = q.reshape(-1, q.shape[-1])
        m = int(q.shape[0])
        if int(k) > MAX_K:
            sims, idx = self._search_wide(q, int(k), float(threshold), normalize_queries, skip_self)
            if lead is not None:
                sims, idx = sims.reshape(*lead, k), idx.reshape(*lead, k)
            return sims, idx
        flags = (SEARCH_NORMALIZE_Q if normalize_queries else 0) | (SEARCH_SKIP_SELF if skip_self else 0)
        if _is_torch(q) and q.is_cuda:
            import torch
            sims = torch.empty((m, k), dtype=torch.float32, device=q.device)
            idx = torch.empty((m, k), dtype=torch.int64, device=q.device)
        elif _is_torch(q):
            import torch
            sims = torch.empty((m, k), dtype=torch.float32, pin_memory=q.is_pinned())
            idx = torch.empty((m, k), dtype=torch.int64, pin_memory=q.is_pinned())
        else:
            sims = np.empty((m, k), np.float32)
            idx = np.empty((m, k), np.int64)
        self.ctx.check(self.ctx.lib.tvc_search(self.ctx.handle, self.handle, _ptr(q), _dtype_code(q), m, int(q.shape[1]),
                                               int(k), float(threshold), flags, _ptr(sims), _ptr(idx),
                                               _stream_of(q)))
        if lead is not None:
            sims = sims.reshape(*lead, k)
            idx = idx.reshape(*lead, k)
        return sims, idx

    def _search_wide(self, q, k: int, threshold: float, normalize_queries: bool, skip_self: bool):
        """k > MAX_K (FAISS and the reference accept any k; the in-register epilogue stops at 56): the dense
        similarity tile of a chunk of rows (tvc_similarity_matrix, the same tensor-core GEMM), the k + 32 best per
        row by that bf16 score, re-scored from the fp32 masters like tvc_search does, ordered (similarity desc,
        index asc), `>= threshold`, unused slots (-inf, -1).  Rare path: device selection and sort are torch's."""
        import torch
        host_np = not _is_torch(q)
        dev = torch.device("cuda", self.ctx.device)
        qd = (torch.as_tensor(q) if host_np else q).to(dev, torch.float32)
        if normalize_queries:
            qd = torch.nn.functional.normalize(qd, dim=1)
        m, n = int(qd.shape[0]), len(self)
        out_s = torch.full((m, k), -math.inf, dtype=torch.float32, device=dev)
        out_i = torch.full((m, k), -1, dtype=torch.int64, device=dev)
        kk = min(n, k + 32 + (1 if skip_self else 0))
        step = max(1, min(m, (1 << 28) // max(n, 1)))           # <= 1 GiB of fp32 scores per chunk
        for r0 in range(0, m if n > 0 else 0, step):
            qc = qd[r0:r0 + step]
            cand = torch.topk(self.similarity_matrix(qc), kk, dim=1).indices          # local row numbers
            if self.flags & GALLERY_NO_MASTER:
                sc = torch.gather(self.similarity_matrix(qc), 1, cand)
            else:
                rows = self.get_rows(cand.reshape(-1)).view(qc.shape[0], kk, self.dim)
                sc = torch.einsum("mkd,md->mk", rows, qc)
            gi = cand + self.global_row_offset
            if skip_self:
                me = torch.arange(r0, r0 + qc.shape[0], device=dev)[:, None]
                sc = torch.where(gi == me, torch.full_like(sc, -math.inf), sc)
            order = torch.sort(gi, dim=1, stable=True).indices                         # index asc ...
            sc, gi = torch.gather(sc, 1, order), torch.gather(gi, 1, order)
            order = torch.sort(sc, dim=1, descending=True, stable=True).indices        # ... then similarity desc
            sc, gi = torch.gather(sc, 1, order)[:, :k], torch.gather(gi, 1, order)[:, :k]
            keep = (sc >= threshold) & torch.isfinite(sc)
            w = sc.shape[1]
            out_s[r0:r0 + qc.shape[0], :w] = torch.where(keep, sc, torch.full_like(sc, -math.inf))
            out_i[r0:r0 + qc.shape[0], :w] = torch.where(keep, gi, torch.full_like(gi, -1))
        if host_np:
            return out_s.cpu().numpy(), out_i.cpu().numpy()
        if not q.is_cuda:
            return out_s.cpu(), out_i.cpu()
        return out_s, out_i

    def search_candidates(self, queries, k: int, scatter: Optional[Scatter] = None, *, normalize_queries: bool = False,
                          skip_self: bool = False):
        """Phase 1 of the sharded search on this shard (cuda tensors only).  With `scatter` the candidates
        go to the slice owners' receive buffers and nothing is returned; otherwise returns the local
        (cand_val [m, kp] f32, cand_idx [m, kp] i64 global) lists."""
        import torch
        flags = (SEARCH_NORMALIZE_Q if normalize_queries else 0) | (SEARCH_SKIP_SELF if skip_self else 0)
        if isinstance(queries, tuple):
            # (device pointer of a prepared bf16 operand, rows): see Context.prepare_queries
            qp, m = int(queries[0]), int(queries[1])
            flags |= SEARCH_PREPARED_Q
            code, stream, dev = TVC_BF16, torch.cuda.current_stream().cuda_stream, torch.device("cuda", self.ctx.device)
        else:
            q = _rows(queries)
            if q.ndim > 2:
                q = q.reshape(-1, q.shape[-1])
            self._check_dim(q)
            qp, m, code, stream, dev = _ptr(q), int(q.shape[0]), _dtype_code(q), _stream_of(q), q.device
        val = idx = None
        if scatter is None:
            kp = self.ctx.candidate_width(k)
            val = torch.empty((m, kp), dtype=torch.float32, device=dev)
            idx = torch.empty((m, kp), dtype=torch.int64, device=dev)
        self.ctx.check(self.ctx.lib.tvc_search_candidates(
            self.ctx.handle, self.handle, C.c_void_p(qp), code, m, self.dim, int(k), flags,
            C.byref(scatter) if scatter is not None else None, _ptr(val), _ptr(idx), stream))
        return val, idx

    def similarity_matrix(self, queries, *, normalize_queries: bool = False):
        q = _rows(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        self._check_dim(q)
        m, n = int(q.shape[0]), len(self)
        if _is_torch(q) and q.is_cuda:
            import torch
            out = torch.empty((m, n), dtype=torch.float32, device=q.device)
        else:
            out = np.empty((m, n), np.float32)
        self.ctx.check(self.ctx.lib.tvc_similarity_matrix(
            self.ctx.handle, self.handle, _ptr(q), _dtype_code(q), m, int(q.shape[1]),
            SEARCH_NORMALIZE_Q if normalize_queries else 0, _ptr(out), _stream_of(q)))
        return out

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.ctx.lib.tvc_gallery_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
