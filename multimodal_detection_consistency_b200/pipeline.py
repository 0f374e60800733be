"""Batched TVC scoring: the public call that feeds whole [Q, V, d] tiles through kernels (a), (b), (c).

One "TVC-scored query" (BASELINE.json metric; the stack of experiments/defenses/detector.py:117-170):
  1. every text-variant row is searched top-k against the image gallery and the reference bank
     (kernel a: tvc_search),
  2. the per-query candidate lists (variant-major) are greedily de-duplicated into <= R retrieval and
     <= G generative references and the query image is scored against text, variants and references;
     statistics + AdversarialDetector / ConsistencyChecker decisions (kernel b: tvc_consistency_emb),
  3. the gallery hits are accumulated into the k-occurrence hubness histogram (kernel c).

Multi-GPU: the gallery and the bank are row-sharded, one process per GPU.  Every rank
searches all query rows against its shard; the packed (sim, global idx) candidates are exchanged so
that each rank owns the merged global top-k of its slice of the queries (one all-to-all of 12·k
bytes per row per rank), fetches the few gallery rows its slice needs from the owning shards
(index + row all-to-all), runs kernel (b) on its slice, and the histogram is all-reduced.
`torch.distributed` carries the collectives (NCCL on GPUs; gloo + a CPU engine in the tests).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _native as N


class CudaEngine:
    """Kernels through libtvc.so.  (tests substitute an oracle-backed engine to exercise the
    multi-rank host logic on CPU with gloo; the product has no other engine.)"""

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise N.TvcError(N.TVC_ERR_NO_DEVICE, "TVCScorer needs a CUDA device; libtvc has no CPU fallback")
        self.device = device
        self.ctx = N.Context.get(device.index if device.index is not None else torch.cuda.current_device())

    def make_gallery(self, rows, offset, normalize=False):
        return N.Gallery(rows, global_row_offset=offset, normalize=normalize, ctx=self.ctx)

    def wrap_rows(self, rows, offset=0):
        return N.Gallery.wrap_rows(rows, offset, ctx=self.ctx)

    def search(self, gallery, q, k, threshold=-math.inf):
        return gallery.search(q, k, threshold)

    def merge(self, sims, idx, k):
        return self.ctx.merge_topk(sims, idx, k)

    def get_rows(self, gallery, local_idx):
        return gallery.get_rows(local_idx)

    def consistency(self, params, img, txt, var, ret_gallery, ret_idx, gen, g_cnt, gen_gallery, gen_idx):
        return self.ctx.consistency_emb(params, img, txt, var, ret_gallery=ret_gallery, ret_idx=ret_idx, gen=gen,
                                        g_cnt=g_cnt, gen_gallery=gen_gallery, gen_idx=gen_idx)

    def k_occurrence(self, idx, n_bins, counts):
        return self.ctx.k_occurrence(idx, n_bins, 0, counts)

    def make_exchange(self, dist, group, world, rank):
        return PeerExchange(self.ctx, dist, group, world, rank)

    def peer_group(self, gallery, total_rows, dist, group, world, rank):
        """[own shard + CUDA-IPC views of every peer's fp32 master]: kernel (b) then reads rows of other
        shards straight from peer HBM over NVLink instead of staging them with collectives."""
        info = [None] * world
        dist.all_gather_object(info, (gallery.export_ipc(), len(gallery), gallery.global_row_offset), group=group)
        parts = []
        for r, (handle, n, off) in enumerate(info):
            parts.append(gallery if r == rank else N.Gallery.import_ipc(handle, n, gallery.dim, off, self.ctx))
        return N.Gallery.group(parts)


class PeerExchange:
    """Candidate exchange of the sharded search without a data-path collective.  Every buffer that crosses GPUs
    is allocated through libtvc (tvc_peer_alloc), double-buffered and mapped by all peers with CUDA IPC; the
    kernels themselves store into / the copy engines copy into the peers' HBM over NVLink, and stream-ordered
    barriers (a 4-byte all-reduce) separate the steps.  Per batch:
      begin_batch    this rank's slice of the query rows -> bf16 GEMM operand, stored into EVERY rank's operand
                     buffer; barrier
      push_queries   the same rows in fp32 -> every rank's fp32 query area, by copy engine on a side stream,
                     under the GEMMs
      scatter        GEMM + top-KP of ALL rows on this rank's shard; candidates stored into the slice owners'
                     receive areas                                                       (per search); barrier
      merge          owner: P lists per row -> KP best by GEMM score; the index list goes to every shard's
                     request area                                                         (per search); barrier
      rescore        every shard: the requested rows of ITS range re-scored in fp32 against the fp32 query rows;
                     scores stored into the owners' score areas                           (per search); barrier
      finalize       owner: (fp32 score desc, index asc), threshold, top-k                 (per search)
    `collect` is the round-1 phase 2 (the owner pulls KP master rows per query row from peer HBM); it remains for
    shards without an fp32 master."""

    def __init__(self, ctx, dist, group, world, rank):
        self.ctx, self.dist, self.group, self.world, self.rank = ctx, dist, group, world, rank
        self.areas = {}           # (tag, ...) -> dict(bases=[ptr per rank], nbytes=bytes of one of the two buffers, own=ptr)
        self.parity = {}
        self._token = None
        self._operand = None
        self._qf32 = None
        self._push_stream = None
        self._push_ev = None

    def _shared(self, tag, nbytes: int):
        """Double-buffered peer-mapped area `tag` of 2 x (>= nbytes) on every rank.  Grow-only: a request that
        fits reuses the area (every layout inside is addressed by the request's own strides), a larger one
        replaces it - collectively: batch shapes are the same on all ranks, so all take the same path."""
        area = self.areas.get(tag)
        if area is not None and area["nbytes"] >= nbytes:
            return area
        self._retire(tag)
        ptr, handle = self.ctx.peer_alloc(2 * nbytes)
        info = [None] * self.world
        self.dist.all_gather_object(info, handle, group=self.group)
        bases = [ptr if r == self.rank else self.ctx.peer_open(h) for r, h in enumerate(info)]
        area = dict(bases=bases, nbytes=nbytes, own=ptr)
        self.areas[tag] = area
        self.parity[tag] = 0
        return area

    def _flip(self, tag) -> int:
        b = self.parity[tag]
        self.parity[tag] = b ^ 1
        return b

    def _area(self, tag, rows_per_slice: int, kp: int):
        # layout of one rank's candidate area: [val buf0 | val buf1 | idx buf0 | idx buf1], the region sizes
        # fixed when the area was allocated (the same on every rank)
        n = self.world * rows_per_slice * kp
        area = self._shared(tag, n * 12)
        if "vb" not in area:
            area["vb"], area["ib"] = n * 4, n * 8
        return area

    def begin_batch(self, rows_slice, q_total: int, v: int, lo: int, per: Optional[int] = None):
        """Convert this rank's slice of the query rows to the bf16 GEMM operand once and store it into
        every rank's query buffer (own HBM + peers over NVLink): after the barrier each rank holds the
        operand of the WHOLE batch although it only ever saw (or uploaded) its own slice."""
        d = int(rows_slice.shape[1])
        per = -(-q_total // self.world) if per is None else per
        rows_cap = per * self.world * v
        area = self._shared("q", rows_cap * self.ctx.query_row_bytes(d))
        b = self._flip("q")
        if rows_slice.shape[0] > 0:
            self.ctx.prepare_queries(rows_slice, [base + b * area["nbytes"] for base in area["bases"]], lo * v)
        self.barrier(rows_slice.device)
        self._operand = (area["own"] + b * area["nbytes"], q_total * v)

    def push_queries(self, rows_slice, q_total: int, v: int, lo: int, per: Optional[int] = None):
        """The fp32 rows of this rank's slice -> every rank's fp32 query area, on a side stream by the copy
        engines (no SM taken from the GEMMs they run under).  The barrier that follows the scatters waits for
        them, so every shard holds the fp32 rows of the whole batch when it re-scores."""
        import torch
        d = int(rows_slice.shape[1])
        per = -(-q_total // self.world) if per is None else per
        rows_cap = per * self.world * v
        area = self._shared("qf32", rows_cap * d * 4)
        b = self._flip("qf32")
        self._qf32 = area["own"] + b * area["nbytes"]
        if self._push_stream is None:
            self._push_stream = torch.cuda.Stream(rows_slice.device)
        ps = self._push_stream
        ps.wait_stream(torch.cuda.current_stream(rows_slice.device))
        nb = rows_slice.numel() * 4
        if nb > 0:
            off = lo * v * d * 4
            order = [(self.rank + 1 + i) % self.world for i in range(self.world)]    # peers first, staggered
            for r in order:
                self.ctx.peer_copy(area["bases"][r] + b * area["nbytes"] + off, rows_slice.data_ptr(), nb, ps.cuda_stream)
            rows_slice.record_stream(ps)
        self._push_ev = torch.cuda.Event()
        self._push_ev.record(ps)

    def _release(self, area):
        for r, base in enumerate(area["bases"]):
            try:
                if r == self.rank:
                    self.ctx.peer_free(base)
                else:
                    self.ctx.peer_close(base)
            except Exception:  # noqa: BLE001 - teardown order at interpreter exit
                pass

    def _retire(self, tag):
        """A batch of another size needs areas of another shape: the old ones of this tag are released (every
        rank takes this path together - batch sizes are the same on all ranks), after every GPU has drained
        what may still read or write them.  Without it varying batch sizes leak HBM on every rank."""
        old = [k for k in self.areas if k == tag]
        if not old:
            return
        import torch
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)
        for k in old:
            self._release(self.areas.pop(k))
            self.parity.pop(k, None)

    def close(self):
        """Unmap the peers' buffers and free this rank's (call on every rank, after a barrier)."""
        for area in self.areas.values():
            self._release(area)
        self.areas.clear()

    def barrier(self, device):
        import torch
        if self._push_ev is not None:          # peers may read what this rank's copy engines wrote only after it
            torch.cuda.current_stream(device).wait_event(self._push_ev)
            self._push_ev = None
        if self._token is None:
            self._token = torch.zeros(1, dtype=torch.int32, device=device)
        self.dist.all_reduce(self._token, group=self.group)

    def scatter(self, tag, gallery, q_total: int, v: int, k: int, per: Optional[int] = None):
        """Phase 1 on this rank's shard: GEMM + top-KP of ALL rows (the operand begin_batch spread), the
        candidates stored into the slice owners' receive buffers.  Returns the token the later steps need."""
        per = -(-q_total // self.world) if per is None else per
        rows_per_slice = per * v
        kp = self.ctx.candidate_width(k)
        area = self._area(tag, rows_per_slice, kp)
        b = self._flip(tag)
        sc = N.Scatter()
        sc.n_slices, sc.slot, sc.rows_per_slice = self.world, self.rank, rows_per_slice
        for r, base in enumerate(area["bases"]):
            sc.val[r] = base + b * area["vb"]
            sc.idx[r] = base + 2 * area["vb"] + b * area["ib"]
        gallery.search_candidates(self._operand, k, sc)
        return dict(area=area, b=b, kp=kp, tag=tag, rps=rows_per_slice, total=q_total * v)

    def collect(self, token, group_gallery, rows_slice, k: int, threshold: float):
        """Round-1 phase 2 (after a barrier): global top-k of this rank's query slice, (sims [rows, k], idx).
        rows_slice are the slice's fp32 rows; the candidate rows are read from local + peer masters."""
        area, b, kp = token["area"], token["b"], token["kp"]
        if rows_slice.shape[0] == 0:
            import torch
            return (torch.empty((0, k), dtype=torch.float32, device=rows_slice.device),
                    torch.empty((0, k), dtype=torch.int64, device=rows_slice.device))
        own = area["own"]
        return self.ctx.rerank_candidates(group_gallery, rows_slice, own + b * area["vb"],
                                          own + 2 * area["vb"] + b * area["ib"], self.world, kp, k, threshold)

    # -- phase 2 where the rows live ----------------------------------------------------------------
    def merge(self, token, my_rows: int, stream: int):
        area, b, kp, rps = token["area"], token["b"], token["kp"], token["rps"]
        key = token["tag"] + ".req"
        req = self._shared(key, self.world * rps * kp * 8)
        rb = self._flip(key)
        token["req"] = req["own"] + rb * req["nbytes"]
        mine = self.rank * rps * kp * 8
        own = area["own"]
        self.ctx.exchange_merge(my_rows, self.world, kp, own + b * area["vb"], own + 2 * area["vb"] + b * area["ib"],
                                [base + rb * req["nbytes"] + mine for base in req["bases"]], stream)

    def rescore(self, token, shard, d: int, stream: int):
        kp, rps = token["kp"], token["rps"]
        key = token["tag"] + ".score"
        score = self._shared(key, rps * kp * 4)
        sb = self._flip(key)
        token["score"] = score["own"] + sb * score["nbytes"]
        self.ctx.exchange_rescore(shard, self._qf32, d, self.world, rps, token["total"], kp, token["req"],
                                  [base + sb * score["nbytes"] for base in score["bases"]], stream)

    def finalize(self, token, my_rows: int, k: int, threshold: float, device):
        kp, rps = token["kp"], token["rps"]
        return self.ctx.exchange_finalize(my_rows, kp, k, threshold, token["req"] + self.rank * rps * kp * 8,
                                          token["score"], device)


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous row shard of rank `rank`: rows [lo, hi) with ceil(n / world) rows per rank."""
    per = -(-n // world) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def slice_bounds(q: int, world: int, rank: int):
    per = -(-q // world)
    lo = min(q, rank * per)
    return lo, min(q, lo + per)


class TVCScorer:
    """score_batch(img, txt, var) -> scores, decisions, top-k; see the module docstring."""

    def __init__(self, gallery_rows, bank_rows=None, *, k: int = 10, params: Optional[N.DetectorParams] = None,
                 device=None, total_gallery_rows: Optional[int] = None, total_bank_rows: Optional[int] = None,
                 bank_threshold: float = -math.inf, track_hubness: bool = True, process_group=None,
                 engine=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.rank = self.dist.get_rank(process_group) if self.dist else 0
        if engine is None:
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            engine = CudaEngine(dev)
        self.engine = engine
        self.device = engine.device
        self.k = int(k)
        self.params = params if params is not None else N.default_params()
        self.bank_threshold = float(bank_threshold)
        n_local = int(gallery_rows.shape[0])
        self.n_total = int(total_gallery_rows) if total_gallery_rows is not None else n_local
        self.g_lo, self.g_hi = shard_bounds(self.n_total, self.world, self.rank)
        if self.g_hi - self.g_lo != n_local:
            raise ValueError(f"rank {self.rank}: gallery shard has {n_local} rows, expected {self.g_hi - self.g_lo}")
        self.gallery = engine.make_gallery(gallery_rows, self.g_lo)
        self.bank = None
        self.b_total = 0
        if bank_rows is not None:
            b_local = int(bank_rows.shape[0])
            self.b_total = int(total_bank_rows) if total_bank_rows is not None else b_local
            self.b_lo, self.b_hi = shard_bounds(self.b_total, self.world, self.rank)
            if self.b_hi - self.b_lo != b_local:
                raise ValueError(f"rank {self.rank}: bank shard has {b_local} rows, expected {self.b_hi - self.b_lo}")
            self.bank = engine.make_gallery(bank_rows, self.b_lo)
        # multi-rank: rows referenced by kernel (b) live on other shards.  With the CUDA engine they are
        # read in place through CUDA IPC peer mappings; engines without peer memory (the CPU test
        # engine) stage them with index + row all-to-alls (_fetch_rows).
        self._gallery_group = self._bank_group = None
        if self.world > 1 and hasattr(engine, "peer_group"):
            self._gallery_group = engine.peer_group(self.gallery, self.n_total, self.dist, self.group, self.world,
                                                    self.rank)
            if self.bank is not None:
                self._bank_group = engine.peer_group(self.bank, self.b_total, self.dist, self.group, self.world,
                                                     self.rank)
        self._exchange = None
        if self.world > 1 and self._gallery_group is not None and hasattr(engine, "make_exchange"):
            self._exchange = engine.make_exchange(self.dist, self.group, self.world, self.rank)
        # phase 2 of the sharded search: True = re-score where the rows live (PeerExchange.merge / rescore /
        # finalize), False = the round-1 pull of master rows over NVLink (PeerExchange.collect)
        self.rescore_at_shards = True
        self.track_hubness = track_hubness
        self._k_occ = torch.zeros(self.n_total, dtype=torch.int32, device=self.device) if track_hubness else None
        # multi-GPU: the per-step histogram all-reduce runs on its own stream and NCCL communicator, under the
        # next step's upload / operand broadcast; reading `k_occurrence` waits for it
        self._hist_stream = self._hist_ev = self._hist_group = None
        if track_hubness and self.world > 1 and self.device.type == "cuda" and self._exchange is not None:
            self._hist_group = self.dist.new_group(ranks=list(range(self.world))) if process_group is None else None
            self._hist_stream = torch.cuda.Stream(self.device)
        self._host: Dict[str, torch.Tensor] = {}
        self._dev_stage: Dict[str, torch.Tensor] = {}
        self._copy_stream = None
        # pieces a host batch is pipelined in: a count (equal pieces; 1 = off) or relative sizes.  Measured on the
        # bench workload (scripts/e2e_splits.py, ms/step; device-resident 101.4-102.2): unsplit 106.1, 2 equal
        # 103.9, 4 equal 105.0, 8 equal 105.6, (1,7) 103.7, (1,4,3) 103.5, (1,3,3,1) 103.1, (1,6,1) 102.4 - a small
        # first piece hides the upload, a small last piece hides the download, one big launch keeps the GEMM
        # at full-wave efficiency
        self.host_chunks = (1, 6, 1)
        # pageable sources upload at about half the pinned rate and block the host while they do: growing pieces,
        # each uploaded under the search of the one before (a piece's search takes ~3.5x its pageable upload)
        self.host_chunks_pageable = (1, 3, 4)
        self.min_chunk_queries = 1024       # ... when every piece keeps at least this many queries
        self.profile = False            # True: CUDA-event time per phase, read with phase_times()
        self._marks = []

    @property
    def k_occurrence(self):
        """Running k-occurrence histogram over the whole gallery (int32 [N]); on several GPUs the all-reduced one."""
        if self._hist_ev is not None:
            torch.cuda.current_stream(self.device).wait_event(self._hist_ev)
        return self._k_occ

    # ------------------------------------------------------------------ helpers
    def _mark(self, name: str):
        if self.profile and self.device.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def phase_times(self) -> Dict[str, float]:
        """ms per phase accumulated since the last call (profile=True only)."""
        out: Dict[str, float] = {}
        if self._marks:
            torch.cuda.synchronize(self.device)
            for (n0, e0), (n1, e1) in zip(self._marks, self._marks[1:]):
                if n1 != "begin":
                    out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._marks = []
        return out

    def _dev(self, x, dtype=torch.float32):
        if x is None:
            return None
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x)
        return x.to(self.device, dtype=dtype, non_blocking=True).contiguous()

    def _side_upload(self, items):
        """name -> (tensor or None, dtype).  Device tensors pass through (converted on the current stream);
        host tensors are copied on the copy stream so the transfer runs under whatever the current stream
        does next; the caller waits on the returned event before the first kernel that reads them."""
        out, host = {}, {}
        for name, (t, dt) in items.items():
            if t is None:
                out[name] = None
            elif isinstance(t, torch.Tensor) and t.device.type == self.device.type:
                out[name] = t.to(dtype=dt).contiguous()
            elif self.device.type != "cuda":
                out[name] = self._dev(t, dt)
            else:
                host[name] = (t if isinstance(t, torch.Tensor) else torch.as_tensor(t), dt)
        if not host:
            return out, None
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        for name, (t, dt) in host.items():                     # allocated on the consumer's stream
            out[name] = torch.empty(t.shape, dtype=dt, device=self.device)
        cs.wait_stream(main)                                   # the blocks may be recycled from work still in flight
        with torch.cuda.stream(cs):
            for name, (t, dt) in host.items():
                out[name].copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        return out, ev

    def _pinned(self, name: str, like: torch.Tensor) -> torch.Tensor:
        buf = self._host.get(name)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=self.device.type == "cuda")
            self._host[name] = buf
        return buf

    def _a2a(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        """all_to_all_single over dim 0 with per-peer row counts."""
        out = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(out, send.contiguous(), list(map(int, recv_counts)), list(map(int, send_counts)),
                                    group=self.group)
        return out

    def _global_topk(self, gallery, rows_all: torch.Tensor, q_total: int, v: int, threshold: float):
        """Local search of every row, then (multi-rank) exchange + merge so that this rank holds the
        global top-k of its query slice.  Returns (sims [Qs*V, k], idx [Qs*V, k])."""
        sims, idx = self.engine.search(gallery, rows_all, self.k, threshold)
        if self.world == 1:
            return sims, idx
        k = self.k
        per = -(-q_total // self.world)
        send_counts = []
        for r in range(self.world):
            lo, hi = slice_bounds(q_total, self.world, r)
            send_counts.append((hi - lo) * v)
        lo, hi = slice_bounds(q_total, self.world, self.rank)
        mine = (hi - lo) * v
        recv_counts = [mine] * self.world
        del per
        rs = self._a2a(sims, send_counts, recv_counts).view(self.world, mine, k)
        ri = self._a2a(idx, send_counts, recv_counts).view(self.world, mine, k)
        return self.engine.merge(rs.permute(1, 0, 2).contiguous(), ri.permute(1, 0, 2).contiguous(), k)

    def _fetch_rows(self, gallery, total_rows: int, idx: torch.Tensor):
        """Rows of the (sharded) gallery for the global indices in `idx`: returns (view gallery over
        the fetched rows [U, d], remapped idx with the same shape pointing into it)."""
        flat = idx.reshape(-1)
        uniq, inv = torch.unique(flat, return_inverse=True)      # sorted; a leading -1 marks unused slots
        valid = uniq >= 0
        per = -(-total_rows // self.world)
        owner = torch.div(uniq.clamp(min=0), per, rounding_mode="floor")
        owner[~valid] = self.world                                  # never sent
        send_counts = torch.bincount(owner, minlength=self.world + 1)[: self.world]
        recv_counts = torch.empty_like(send_counts)
        self.dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        want = uniq[valid]                                          # sorted by index == sorted by owner
        asked = self._a2a(want, sc, rc)                             # global indices peers want from me
        lo, _ = shard_bounds(total_rows, self.world, self.rank)
        rows = self.engine.get_rows(gallery, asked - lo)            # [sum(rc), d] fp32
        got = self._a2a(rows, rc, sc)                               # rows for `want`, same order
        # remap: position of every idx entry inside `got`; unused slots stay negative
        shift = int((~valid).sum().item())
        remapped = (inv - shift).reshape(idx.shape)
        remapped = torch.where(idx >= 0, remapped, torch.full_like(remapped, -1))
        return self.engine.wrap_rows(got), remapped

    # ------------------------------------------------------------------ the call
    def score_batch(self, img, txt, var, gen=None, g_cnt=None, *, to_host: bool = False, copy: bool = False):
        """img, txt: [Q, d]; var: [Q, V, d] (host or device, fp32).  In multi-rank mode every rank passes
        the SAME full batch and receives the results of its own contiguous slice of the queries.
        Returns a dict: scores [Qs, 24], flags [Qs], topk_idx/topk_sim [Qs, V, k], bank_idx/bank_sim.

        With to_host=True the arrays are the scorer's own PINNED result buffers, BORROWED until the next
        score_batch call on this scorer (which overwrites them - its device-to-host copies land in the same
        memory): consume them before calling again, or pass copy=True to receive private copies.

        Host batches on one GPU are pipelined: the batch is cut into `host_chunks` pieces, piece c+1 is
        uploaded on a copy stream while piece c is searched and scored, and piece c's results go back
        to pinned host memory behind it - only the first upload and the last download are exposed."""
        host_in = not (isinstance(var, torch.Tensor) and var.device.type == "cuda")
        pinned = isinstance(var, torch.Tensor) and var.is_pinned()
        if (host_in and self.world == 1 and self.device.type == "cuda"
                and len(self._piece_bounds(int(var.shape[0]), pinned)) > 1):
            out = self._score_batch_pipelined(img, txt, var, gen, g_cnt, to_host)
        elif (host_in and self.world > 1 and self._exchange is not None and self.device.type == "cuda"
                and len(self._piece_bounds(-(-int(var.shape[0]) // self.world), pinned)) > 1):
            out = self._score_batch_pipelined_multi(img, txt, var, gen, g_cnt, to_host)
        else:
            out = self._score_batch(img, txt, var, gen, g_cnt, to_host=to_host)
        if to_host and copy:
            out = {n: (t.clone() if isinstance(t, torch.Tensor) else t) for n, t in out.items()}
        return out

    def _piece_bounds(self, q_total: int, pinned: bool = True):
        """Query ranges a host batch is pipelined in.  `host_chunks` is a piece count (equal pieces) or a
        sequence of relative piece sizes, e.g. (1, 4, 3): a small first piece shortens the only upload that
        is not hidden behind a search.  One piece (= no pipelining) when a piece would fall under
        `min_chunk_queries`."""
        hc = self.host_chunks if pinned else self.host_chunks_pageable
        cands = [hc, 2]
        if self.world > 1 and pinned and q_total // 4 >= self.min_chunk_queries // 2:
            # a rank's slice on several GPUs: a piece of every rank together is one sharded search, so the first piece
            # may be smaller per rank than on one GPU (512 queries x 8 ranks x 5 variants = 80 query tiles, a full
            # wave); a quarter in front leaves a quarter of the upload exposed instead of the half two halves do
            cands.insert(1, (1, 3))
        for cand in cands:                                  # too small for the configured split: try the next
            weights = [1.0] * int(cand) if isinstance(cand, int) else [float(w) for w in cand]
            if len(weights) < 2 or min(weights) <= 0:
                break
            total, acc, cuts = sum(weights), 0.0, [0]
            for w in weights:
                acc += w
                cuts.append(int(round(q_total * acc / total)))
            cuts[-1] = q_total
            bounds = list(zip(cuts[:-1], cuts[1:]))
            floor = self.min_chunk_queries // 2 if (cand == (1, 3) and self.world > 1) else self.min_chunk_queries
            if min(b - a for a, b in bounds) >= floor:
                return bounds
        return [(0, q_total)]

    def _staging(self, name: str, shape, dtype) -> torch.Tensor:
        buf = self._dev_stage.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._dev_stage[name] = buf
        return buf

    def _score_batch_pipelined_multi(self, img, txt, var, gen, g_cnt, to_host: bool):
        """Host batch on several GPUs (peer-memory path): every rank's slice is cut at the same relative
        positions; piece c of ALL ranks is one run of the sharded protocol, and a rank uploads its rows of piece
        c+1 on the copy stream while piece c is searched, results going back behind it - only the first upload
        and the last download stay exposed (round 1 uploaded the whole slice in front of the first GEMM)."""
        q_total, v, d = int(var.shape[0]), int(var.shape[1]), int(var.shape[2])
        per = -(-q_total // self.world)
        lo, hi = slice_bounds(q_total, self.world, self.rank)
        qs = hi - lo
        pinned = isinstance(var, torch.Tensor) and var.is_pinned()
        bounds = self._piece_bounds(per, pinned=pinned)
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        cs.wait_stream(main)
        srcs = dict(img=(img, torch.float32), txt=(txt, torch.float32), var=(var, torch.float32))
        if gen is not None:
            srcs["gen"] = (gen, torch.float32)
        if g_cnt is not None:
            srcs["g_cnt"] = (g_cnt, torch.int32)
        host = {n: (t if isinstance(t, torch.Tensor) else torch.as_tensor(t)) for n, (t, _) in srcs.items()}
        dev = {n: self._staging("m_" + n, (per,) + tuple(host[n].shape[1:]), dt) for n, (_, dt) in srcs.items()}

        def mine(a, b):                       # this rank's rows of piece [a, b), relative to its slice start
            return min(qs, a), min(qs, b)

        def upload(a, b):
            a, b = mine(a, b)
            with torch.cuda.stream(cs):
                for n in ("var", "img", "txt", "gen", "g_cnt"):
                    if n in dev and b > a:
                        dev[n][a:b].copy_(host[n][lo + a:lo + b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return ev

        pieces, o = [], None
        ev = upload(*bounds[0])
        for i, (a, b) in enumerate(bounds):
            main.wait_event(ev)
            ma, mb = mine(a, b)
            # queries of this piece over all ranks: full slices contribute b - a, the last slice what it has
            q_piece = sum(max(0, min(b, (min(q_total, (r + 1) * per) - r * per)) - a) for r in range(self.world))
            o = self._score_batch(dev["img"][ma:mb], dev["txt"][ma:mb], dev["var"][ma:mb],
                                  dev["gen"][ma:mb] if "gen" in dev else None,
                                  dev["g_cnt"][ma:mb] if "g_cnt" in dev else None, to_host=False,
                                  layout=(q_piece, b - a, self.rank * (b - a)))
            o.pop("slice")
            if to_host:
                for name, t in o.items():
                    buf = self._pinned(name, torch.empty((qs,) + tuple(t.shape[1:]), dtype=t.dtype, device="meta"))
                    buf[ma:mb].copy_(t, non_blocking=True)
            else:
                pieces.append(o)
            if i + 1 < len(bounds):
                ev = upload(*bounds[i + 1])
        if to_host:
            main.synchronize()
            out = {name: self._host[name] for name in o}
            self._mark("to_host")
        else:
            out = {name: torch.cat([pc[name] for pc in pieces], 0) for name in pieces[0]}
        out["slice"] = (lo, hi)
        return out

    def _score_batch_pipelined(self, img, txt, var, gen, g_cnt, to_host: bool):
        q_total = int(var.shape[0])
        bounds = self._piece_bounds(q_total, pinned=isinstance(var, torch.Tensor) and var.is_pinned())
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        cs.wait_stream(main)                     # the staging buffers may still be read by the previous call
        srcs = dict(img=(img, torch.float32), txt=(txt, torch.float32), var=(var, torch.float32))
        if gen is not None:
            srcs["gen"] = (gen, torch.float32)
        if g_cnt is not None:
            srcs["g_cnt"] = (g_cnt, torch.int32)
        host = {n: (t if isinstance(t, torch.Tensor) else torch.as_tensor(t)) for n, (t, _) in srcs.items()}
        dev = {n: self._staging(n, host[n].shape, dt) for n, (_, dt) in srcs.items()}
        def upload(a, b):
            with torch.cuda.stream(cs):
                for n in ("var", "img", "txt", "gen", "g_cnt"):     # var first: the search waits on it
                    if n in dev:
                        dev[n][a:b].copy_(host[n][a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return ev

        # Piece c+1 goes up AFTER piece c's kernels and result copies are queued: a copy from pinned memory is
        # asynchronous either way, but a copy from PAGEABLE memory (what the reference's encoders hand over:
        # NumPy arrays) blocks the host until it is staged - queued in this order it blocks under piece c's
        # search instead of in front of it (B200, bench workload: 131.8 -> see DESIGN.md section 6)
        pieces = []
        ev = upload(*bounds[0])
        for i, (a, b) in enumerate(bounds):
            main.wait_event(ev)
            o = self._score_batch(dev["img"][a:b], dev["txt"][a:b], dev["var"][a:b],
                                  dev["gen"][a:b] if "gen" in dev else None,
                                  dev["g_cnt"][a:b] if "g_cnt" in dev else None, to_host=False)
            o.pop("slice")
            if to_host:
                for name, t in o.items():
                    buf = self._pinned(name, torch.empty((q_total,) + tuple(t.shape[1:]), dtype=t.dtype, device="meta"))
                    buf[a:b].copy_(t, non_blocking=True)
            else:
                pieces.append(o)
            if i + 1 < len(bounds):
                ev = upload(*bounds[i + 1])
        if to_host:
            main.synchronize()
            out = {name: self._host[name] for name in o}
            self._mark("to_host")
        else:
            out = {name: torch.cat([pc[name] for pc in pieces], 0) for name in pieces[0]}
        out["slice"] = (0, q_total)
        return out

    def _score_batch(self, img, txt, var, gen=None, g_cnt=None, *, to_host: bool = False, layout=None):
        """layout=None: the arrays hold the whole batch and this rank takes its contiguous slice of it.
        layout=(q_total, per, lo): peer-memory path only - the arrays hold ONLY this rank's rows of a batch of
        q_total queries cut into slices of `per` queries per rank, this rank's starting at query `lo` (the
        pipelined multi-GPU host path cuts every rank's slice at the same relative positions)."""
        self._mark("begin")
        v, d = int(var.shape[1]), int(var.shape[2])
        if v != self.params.n_variants:
            raise ValueError(f"{v} variants given, params.n_variants = {self.params.n_variants}")
        per = None
        if layout is None:
            q_total = int(var.shape[0])
            lo, hi = slice_bounds(q_total, self.world, self.rank)
        else:
            q_total, per, lo = layout
            hi = lo + int(var.shape[0])
            z = slice(None)                       # the arrays already are this rank's rows
        qs = hi - lo
        if layout is None:
            z = slice(lo, hi)
        k = self.k
        host_var = not (isinstance(var, torch.Tensor) and var.device.type == self.device.type == "cuda")
        b_sim = b_idx = None

        def upload_rest():
            # the rows only kernel (b) reads go up on the copy stream, under the searches - queued BEHIND the
            # variant rows (the copy stream waits for what the current stream holds), which the GEMM waits for
            return self._side_upload(dict(img=(img[z], torch.float32), txt=(txt[z], torch.float32),
                                          gen=(gen[z] if gen is not None else None, torch.float32),
                                          g_cnt=(g_cnt[z] if g_cnt is not None else None, torch.int32)))
        if layout is not None and not (self.world > 1 and self._exchange is not None):
            raise ValueError("layout= serves the peer-memory path only")
        if self.world > 1 and self._exchange is not None:
            # peer-memory path: this rank only touches (uploads) ITS slice of the batch.  The bf16 operand
            # of the whole batch is assembled in every rank's HBM by the prepare kernel's peer stores, the
            # candidates are stored into the slice owners' HBM, and the owners re-rank.
            var_s = self._dev(var[z])
            side, side_ev = upload_rest()
            rows_s = var_s.view(qs * v, d)
            ex = self._exchange
            at_shards = self.rescore_at_shards and self.gallery.has_master and (self.bank is None or self.bank.has_master)
            ex.begin_batch(rows_s, q_total, v, lo, per)
            if at_shards:
                ex.push_queries(rows_s, q_total, v, lo, per)
            # both shard searches first, ONE barrier, then phase 2 of both: a rank whose gallery GEMM finishes
            # early spends the wait in its bank GEMM instead of in a barrier
            tg = ex.scatter("gallery", self.gallery, q_total, v, k, per)
            tb = ex.scatter("bank", self.bank, q_total, v, k, per) if self.bank is not None else None
            ex.barrier(self.device)
            self._mark("search_gemm")
            if at_shards:
                st = torch.cuda.current_stream(self.device).cuda_stream
                toks = [(tg, self.gallery, -math.inf)] + ([(tb, self.bank, self.bank_threshold)] if tb is not None else [])
                for tok, _, _ in toks:
                    ex.merge(tok, qs * v, st)
                ex.barrier(self.device)
                for tok, shard, _ in toks:
                    ex.rescore(tok, shard, d, st)
                ex.barrier(self.device)
                g_sim, g_idx = ex.finalize(tg, qs * v, k, -math.inf, self.device)
                if tb is not None:
                    b_sim, b_idx = ex.finalize(tb, qs * v, k, self.bank_threshold, self.device)
                self._mark("search_phase2")
            else:
                g_sim, g_idx = ex.collect(tg, self._gallery_group, rows_s, k, -math.inf)
                self._mark("search_gallery")
                if tb is not None:
                    b_sim, b_idx = ex.collect(tb, self._bank_group, rows_s, k, self.bank_threshold)
                    self._mark("search_bank")
        else:
            if self.world > 1 and host_var and self.device.type == "cuda":
                # every rank holds the same host batch: upload only this rank's slice over PCIe and
                # all-gather the rest over NVLink (1/world of the host->device traffic per GPU)
                per = -(-q_total // self.world)
                piece = torch.zeros((per, v, d), dtype=torch.float32, device=self.device)
                if qs > 0:
                    piece[:qs].copy_(torch.as_tensor(var[lo:hi]), non_blocking=True)
                full = torch.empty((self.world * per, v, d), dtype=torch.float32, device=self.device)
                self.dist.all_gather_into_tensor(full, piece, group=self.group)
                var = full[:q_total]
            else:
                var = self._dev(var)
            side, side_ev = upload_rest()
            rows_all = var.view(q_total * v, d)
            g_sim, g_idx = self._global_topk(self.gallery, rows_all, q_total, v, -math.inf)
            self._mark("search_gallery")
            if self.bank is not None:
                b_sim, b_idx = self._global_topk(self.bank, rows_all, q_total, v, self.bank_threshold)
                self._mark("search_bank")
            var_s = var[lo:hi]
        if side_ev is not None:
            torch.cuda.current_stream(self.device).wait_event(side_ev)
        img_s, txt_s, gen_s, gcnt_s = side["img"], side["txt"], side["gen"], side["g_cnt"]
        ret_idx = g_idx.view(qs, v * k)
        gen_idx = b_idx.view(qs, v * k) if (b_idx is not None and gen is None) else None
        ret_gal, gen_gal = self.gallery, self.bank
        if self._gallery_group is not None:
            ret_gal, gen_gal = self._gallery_group, self._bank_group
        elif self.world > 1:
            # (ranks with an empty query slice still take part: the collectives must stay matched)
            ret_gal, ret_idx = self._fetch_rows(self.gallery, self.n_total, ret_idx)
            if gen_idx is not None:
                gen_gal, gen_idx = self._fetch_rows(self.bank, self.b_total, gen_idx)
        self._mark("fetch_rows")
        if qs > 0:
            scores, flags = self.engine.consistency(self.params, img_s, txt_s, var_s, ret_gal, ret_idx, gen_s, gcnt_s,
                                                    gen_gal if gen_s is None else None, gen_idx)
        else:
            scores = torch.empty((0, N.NSCORES), dtype=torch.float32, device=self.device)
            flags = torch.empty((0,), dtype=torch.uint8, device=self.device)
        self._mark("consistency")
        if self.track_hubness:
            if self.world > 1:
                local = torch.zeros_like(self._k_occ)
                if qs > 0:
                    self.engine.k_occurrence(g_idx, self.n_total, local)
                if self._hist_stream is not None:
                    hs = self._hist_stream
                    hs.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(hs):
                        self.dist.all_reduce(local, group=self._hist_group if self._hist_group is not None else self.group)
                        self._k_occ += local
                        self._hist_ev = torch.cuda.Event()
                        self._hist_ev.record(hs)
                    local.record_stream(hs)
                else:
                    self.dist.all_reduce(local, group=self.group)
                    self._k_occ += local
            else:
                self.engine.k_occurrence(g_idx, self.n_total, self._k_occ)
        self._mark("hubness")
        out = dict(scores=scores, flags=flags, topk_idx=g_idx.view(qs, v, k), topk_sim=g_sim.view(qs, v, k))
        if b_idx is not None:
            out.update(bank_idx=b_idx.view(qs, v, k), bank_sim=b_sim.view(qs, v, k))
        if to_host:
            host = {}
            for name, t in out.items():
                buf = self._pinned(name, t)
                buf.copy_(t, non_blocking=True)
                host[name] = buf
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
            out = host
            self._mark("to_host")
        out["slice"] = (lo, hi)
        return out

    def close(self):
        """Release the peer-memory buffers of the multi-GPU path (collective: every rank calls it)."""
        if self._exchange is not None:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            if self.dist is not None:
                self.dist.barrier(group=self.group)
            self._exchange.close()
            self._exchange = None

    def reset_hubness(self):
        if self._k_occ is not None:
            self.k_occurrence.zero_()          # (the property orders this behind an all-reduce still in flight)
            if self._hist_stream is not None:
                self._hist_stream.wait_stream(torch.cuda.current_stream(self.device))
