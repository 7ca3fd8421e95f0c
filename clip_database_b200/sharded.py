"""Row-sharded multi-GPU search (no reference equivalent — the
reference is one process, one thread).

The embedding matrix is split into contiguous rowid ranges, one per rank / GPU.
Every rank scans its shard for the (replicated) query and produces a local top-k;
the only exchange step is that every rank's k candidates reach every rank, which then
merges the G lists with the same (distance, rowid) order.  Because shards are
contiguous in rowid order, (distance, shard, position) == (distance, rowid), so the
sharded answer is bit-identical to the unsharded one.  Two ways to do the exchange:

  fused (default)  ``clipdb_search_sharded_device``: the scan kernel's last CTA stores its
                   shard's candidates into every peer GPU's inbox over NVLink (peer memory
                   mapped with CUDA IPC), waits for theirs and merges — ONE launch per rank per
                   query, no collective call.  ``torch.distributed`` is used once, to swap the
                   64-byte IPC handles.
  NCCL             ONE all-gather of a packed k-entry record per rank (8 + 12k + 4 bytes,
                   252 B at k = 20) followed by a merge kernel.  Also used for batched
                   searches (nq x k candidates per rank).

``ShardedIndex`` is the orchestration (record layout, exchange, merge order);
the per-rank work is delegated to a backend.  The product backend is
``CudaShardBackend`` (CUDA kernels through the C ABI); tests drive the same
orchestration over gloo with a CPU stand-in backend.  ``multigpu.MultiGpuIndex`` is the
single-process variant (the ranks are contexts of one process).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np


def shard_bounds(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges [lo, hi) in scan order, sizes differing by at most one."""
    return [(n_total * g // world, n_total * (g + 1) // world) for g in range(world)]


@dataclass(frozen=True)
class RecordLayout:
    """Byte layout of one rank's packed result: nan(int64) | rowids(int64[k]) |
    dist(float32[k]) | count(int32) | pad to 16."""
    k: int

    @property
    def off_nan(self) -> int:
        return 0

    @property
    def off_rowids(self) -> int:
        return 8

    @property
    def off_dist(self) -> int:
        return 8 + 8 * self.k

    @property
    def off_count(self) -> int:
        return 8 + 12 * self.k

    @property
    def nbytes(self) -> int:
        return (8 + 12 * self.k + 4 + 15) // 16 * 16


@dataclass(frozen=True)
class BatchRecordLayout:
    """One rank's packed result for nq queries: nan(int64[nq]) | rowids(int64[nq][k]) |
    dist(float32[nq][k]) | count(int32[nq]) | flags(int32[nq]) | pad to 16."""
    nq: int
    k: int

    @property
    def off_nan(self) -> int:
        return 0

    @property
    def off_rowids(self) -> int:
        return 8 * self.nq

    @property
    def off_dist(self) -> int:
        return 8 * self.nq + 8 * self.nq * self.k

    @property
    def off_count(self) -> int:
        return 8 * self.nq + 12 * self.nq * self.k

    @property
    def off_flags(self) -> int:
        return self.off_count + 4 * self.nq

    @property
    def nbytes(self) -> int:
        return (self.off_flags + 4 * self.nq + 15) // 16 * 16


class CudaShardBackend:
    """Per-rank work on the GPU: local scan + top-k into a record, merge of gathered records."""

    def __init__(self, index):
        import torch
        self.index = index
        self.torch = torch
        self.device = torch.device("cuda", index.device)
        index.use_torch_stream()
        self._h_query = None

    def new_buffer(self, nbytes: int):
        return self.torch.zeros(nbytes, dtype=self.torch.uint8, device=self.device)

    def to_device(self, query: np.ndarray):
        """Host query -> device through a pinned staging buffer (one async H2D)."""
        t = self.torch
        q = np.ascontiguousarray(query, dtype=np.float32).ravel()
        if self._h_query is None or self._h_query.numel() != q.shape[0]:
            self._h_query = t.empty(q.shape[0], dtype=t.float32).pin_memory()
            self._d_query = t.empty(q.shape[0], dtype=t.float32, device=self.device)
        self._h_query.numpy()[:] = q
        self._d_query.copy_(self._h_query, non_blocking=True)
        return self._d_query

    def local_search(self, d_query, k: int, record, lay: RecordLayout, use_mask: bool) -> None:
        self.index.search_into_record(d_query, k, record, lay.off_rowids, lay.off_dist, lay.off_count,
                                      lay.off_nan, use_mask=use_mask)

    # ---- fused exchange: scan + peer-memory exchange + merge in one launch per rank ----------
    def connect_exchange(self, dist, group, world: int, rank: int) -> None:
        """Allocate this rank's inbox, swap CUDA IPC handles with the other ranks (a host-side
        all_gather_object, once) and map their inboxes.  Raises if peer mapping is impossible."""
        handle, _ = self.index.exchange_init(world, rank)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.index.exchange_connect(handles)
        dist.barrier(group=group)

    def fused_search(self, d_query, k: int, out_dist, out_rowids, out_n, out_nan, use_mask: bool) -> None:
        self.index.search_sharded_device(d_query, k, out_rowids, out_dist, out_n, out_nan, use_mask=use_mask)

    def fused_search_batch(self, d_queries, k: int, out_dist, out_rowids, out_n, out_nan, flags) -> None:
        """<= 256 queries per pass; outputs are the merged answers over all shards."""
        nq = d_queries.shape[0]
        for q0 in range(0, nq, 256):
            q1 = min(q0 + 256, nq)
            self.index.search_batch_sharded_device(d_queries[q0:q1], k, out_rowids[q0:q1], out_dist[q0:q1],
                                                   out_n[q0:q1], out_nan[q0:q1], flags[q0:q1])

    @property
    def batch_enabled(self) -> bool:
        return bool(getattr(self.index, "batch_enabled", False))

    def merge(self, gathered, k: int, lay: RecordLayout, out_dist, out_rowids, out_n) -> None:
        self.index.merge_records_device(gathered, k, lay.off_rowids, lay.off_dist, lay.off_count,
                                        out_dist, out_rowids, out_n)

    def local_search_batch(self, d_queries, k: int, record, lay: BatchRecordLayout) -> None:
        """nq queries against this rank's shard into a packed batch record; async.  Uses the
        tensor-core batched path when the index has its bf16 store (256 queries per pass); queries
        it could not answer (candidate overflow, zero norm) are marked in the record's flags and
        re-run through the exact path by ``ShardedIndex.search_batch`` on every rank."""
        nq, dim = d_queries.shape
        base = record.data_ptr()
        qp = d_queries.data_ptr()
        if not getattr(self.index, "batch_enabled", False) or k > 128:
            record[lay.off_flags:lay.off_flags + 4 * nq].zero_()
            self.index.search_ptrs(qp, nq, k, base + lay.off_rowids, base + lay.off_dist, base + lay.off_count,
                                   base + lay.off_nan)
            return
        for q0 in range(0, nq, 256):
            m = min(256, nq - q0)
            self.index.search_batch_ptrs(qp + q0 * dim * 4, m, k, base + lay.off_rowids + q0 * k * 8,
                                         base + lay.off_dist + q0 * k * 4, base + lay.off_count + q0 * 4,
                                         base + lay.off_nan + q0 * 8, base + lay.off_flags + q0 * 4)

    def merge_batch(self, gathered, nq: int, k: int, lay: BatchRecordLayout, out_dist, out_rowids, out_n) -> None:
        self.index.merge_batch_records_device(gathered, nq, k, lay.off_rowids, lay.off_dist, lay.off_count,
                                              out_dist, out_rowids, out_n)

    def new_batch_outputs(self, nq: int, k: int):
        t = self.torch
        return (t.empty((nq, max(k, 1)), dtype=t.float32, device=self.device),
                t.empty((nq, max(k, 1)), dtype=t.int64, device=self.device),
                t.zeros(nq, dtype=t.int32, device=self.device))

    def queries_to_device(self, queries: np.ndarray):
        """Host queries -> device through a pinned staging buffer (one async H2D, no host-side sync)."""
        t = self.torch
        q = np.ascontiguousarray(queries, dtype=np.float32)
        key = q.shape
        if getattr(self, "_hq_key", None) != key:
            self._hq = t.empty(key, dtype=t.float32).pin_memory()
            self._dq = t.empty(key, dtype=t.float32, device=self.device)
            self._hq_key = key
        self._hq.numpy()[:] = q
        self._dq.copy_(self._hq, non_blocking=True)
        return self._dq

    def fetch_batch(self, out_dist, out_rowids, out_n, flags):
        """(rowids, dist, counts, flags) of a batched search as numpy copies: four async D2H copies into pinned
        memory and ONE stream synchronisation."""
        t = self.torch
        src = (out_rowids, out_dist, out_n, flags.contiguous())
        key = tuple((x.shape, x.dtype) for x in src)
        if getattr(self, "_hb_key", None) != key:
            self._hb = [t.empty(x.shape, dtype=x.dtype).pin_memory() for x in src]
            self._hb_key = key
        for h, d in zip(self._hb, src):
            h.copy_(d, non_blocking=True)
        t.cuda.current_stream(self.device).synchronize()
        return tuple(h.numpy().copy() for h in self._hb)

    def new_outputs(self, k: int):
        """(dist[k], rowids[k], n[1]) as views into ONE packed device record, so the host
        fetches a result with a single D2H copy (``fetch``)."""
        t = self.torch
        lay = RecordLayout(max(k, 1))
        self._out_lay = lay
        self._out = t.zeros(lay.nbytes, dtype=t.uint8, device=self.device)
        self._h_out = t.zeros(lay.nbytes, dtype=t.uint8).pin_memory()
        kk = max(k, 1)
        self.out_nan = self._out[lay.off_nan:lay.off_nan + 8].view(t.int64)
        return (self._out[lay.off_dist:lay.off_dist + 4 * kk].view(t.float32),
                self._out[lay.off_rowids:lay.off_rowids + 8 * kk].view(t.int64),
                self._out[lay.off_count:lay.off_count + 4].view(t.int32))

    def fetch(self, k: int):
        """One async D2H of the packed result + stream sync -> (rowids, dist) numpy copies."""
        lay = self._out_lay
        self._h_out.copy_(self._out, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        h = self._h_out.numpy()
        m = int(h[lay.off_count:lay.off_count + 4].view(np.int32)[0])
        if m == -2:
            raise RuntimeError("sharded search: the ranks asked different questions (k differs between ranks)")
        if m < 0:
            raise RuntimeError("sharded search: a peer GPU did not deliver its candidates in time "
                               "(option xchg_timeout_ms); every rank must issue the same searches")
        ids = h[lay.off_rowids:lay.off_rowids + 8 * m].view(np.int64).copy()
        d = h[lay.off_dist:lay.off_dist + 4 * m].view(np.float32).copy()
        return ids, d


class ShardedIndex:
    """One rank's view of a row-sharded store."""

    FUSED_K_MAX = 128

    def __init__(self, backend, group=None, fused="auto"):
        """``fused``: True = single-query searches go through the fused peer-memory exchange
        (one launch per rank, no collective), False = local search + NCCL all-gather + merge
        kernel, "auto" = fused when the backend can map its peers' memory."""
        import torch.distributed as dist
        self.backend = backend
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._k = None
        self.fused = False
        if fused and self.world > 1 and hasattr(backend, "connect_exchange"):
            ok = 1
            try:
                backend.connect_exchange(dist, group, self.world, self.rank)
            except Exception:
                if fused is True:
                    raise
                ok = 0
            # every rank must take the same path
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=group)
            self.fused = all(flags)

    def resync(self) -> None:
        """COLLECTIVE: after a rank failed to enqueue a search, or an exchange timed out, every rank calls this
        to restart the fused exchange from one common sequence number (agreed with an all-reduce MAX) — stale
        inbox records carry older numbers and are ignored.  Ranks still spinning in a kernel are released first."""
        if not (self.fused and self.world > 1):
            return
        index = self.backend.index
        index.exchange_abort(True)
        index.synchronize()
        t = self.backend.torch
        self._issued = getattr(self, "_issued", 0) + 16
        top = t.tensor([self._issued], dtype=t.int64, device=self.backend.device)
        self.dist.all_reduce(top, op=self.dist.ReduceOp.MAX, group=self.group)
        self._issued = int(top[0])
        index.exchange_abort(False)
        index.exchange_set_epoch(self._issued, self._issued)
        self.dist.barrier(group=self.group)

    def _prepare(self, k: int) -> None:
        if self._k == k:
            return
        self.layout = RecordLayout(k)
        self.record = self.backend.new_buffer(self.layout.nbytes)
        self.gathered = self.backend.new_buffer(self.layout.nbytes * self.world).view(self.world, self.layout.nbytes)
        self.out_dist, self.out_rowids, self.out_n = self.backend.new_outputs(k)
        self._k = k

    def search_device(self, d_query, k: int, use_mask: bool = False):
        """Enqueue local scan -> all-gather -> merge.  Returns device tensors
        (dist[k], rowids[k], n[1]) valid after the stream / collective completes;
        every rank holds the same merged answer."""
        self._prepare(k)
        if self.fused and 1 <= k <= self.FUSED_K_MAX:
            self._issued = getattr(self, "_issued", 0) + 1
            self.backend.fused_search(d_query, k, self.out_dist, self.out_rowids, self.out_n, self.backend.out_nan,
                                      use_mask)
            return self.out_dist, self.out_rowids, self.out_n
        self.backend.local_search(d_query, k, self.record, self.layout, use_mask)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered.view(-1), self.record, group=self.group)
            src = self.gathered
        else:
            src = self.record.view(1, -1)
        self.backend.merge(src, k, self.layout, self.out_dist, self.out_rowids, self.out_n)
        return self.out_dist, self.out_rowids, self.out_n

    def search_batch_device(self, d_queries, k: int):
        """Enqueue: local batched search -> ONE all-gather of nq*k candidates per rank -> per-query
        merge.  Returns device tensors (dist [nq, k], rowids [nq, k], n [nq], flags [world, nq]);
        ``flags != 0`` marks queries some shard could not answer through the batched path."""
        nq = d_queries.shape[0]
        if self.fused and self.world > 1 and 1 <= k <= self.FUSED_K_MAX:
            # no collective call: the candidates are exchanged over peer memory by the batched path's
            # last kernel, or (no bf16 store) by each query's scan kernel
            fkey = ("fused", nq, k)
            if getattr(self, "_bkey", None) != fkey:
                t = self.backend.torch
                self._bout = self.backend.new_batch_outputs(nq, k)
                self._bnan = t.zeros(nq, dtype=t.int64, device=self.backend.device)
                self._bflags = t.zeros(nq, dtype=t.int32, device=self.backend.device)
                self._bkey = fkey
            out_dist, out_rowids, out_n = self._bout
            self._issued = getattr(self, "_issued", 0) + max(nq, 1)
            if getattr(self.backend, "batch_enabled", False):
                self.backend.fused_search_batch(d_queries, k, out_dist, out_rowids, out_n, self._bnan, self._bflags)
            else:
                self._bflags.zero_()
                for q in range(nq):
                    self.backend.fused_search(d_queries[q], k, out_dist[q], out_rowids[q], out_n[q:q + 1],
                                              self._bnan[q:q + 1], False)
            return out_dist, out_rowids, out_n, self._bflags
        key = (nq, k)
        if getattr(self, "_bkey", None) != key:
            lay = BatchRecordLayout(nq, k)
            self._blay = lay
            self._brecord = self.backend.new_buffer(lay.nbytes)
            self._bgathered = self.backend.new_buffer(lay.nbytes * self.world).view(self.world, lay.nbytes)
            self._bout = self.backend.new_batch_outputs(nq, k)
            self._bkey = key
        lay = self._blay
        self.backend.local_search_batch(d_queries, k, self._brecord, lay)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self._bgathered.view(-1), self._brecord, group=self.group)
            gathered = self._bgathered
        else:
            gathered = self._brecord.view(1, -1)
        out_dist, out_rowids, out_n = self._bout
        self.backend.merge_batch(gathered, nq, k, lay, out_dist, out_rowids, out_n)
        flags = gathered[:, lay.off_flags:lay.off_flags + 4 * nq]
        return out_dist, out_rowids, out_n, flags

    def search_batch(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """nq host queries -> (rowids [nq, k], distances [nq, k], counts [nq]); synchronous.
        Every rank scans its shard for all queries, ONE all-gather moves nq*k candidates per
        rank, every rank merges per query with the (distance, rowid) order.  Queries flagged by
        any shard are re-run one at a time through the exact sharded search — on every rank,
        since every rank sees the same gathered flags."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        d_q = self.backend.queries_to_device(queries)
        out_dist, out_rowids, out_n, flags = self.search_batch_device(d_q, k)
        if hasattr(self.backend, "fetch_batch"):
            ids, dist, counts, h_flags = self.backend.fetch_batch(out_dist, out_rowids, out_n, flags)
        else:
            ids, dist, counts = (x.cpu().numpy().copy() for x in (out_rowids, out_dist, out_n))
            h_flags = flags.contiguous().cpu().numpy()
        flagged = np.flatnonzero(h_flags.view(np.int32).reshape(-1, nq).any(axis=0))
        if (counts < 0).any():
            raise RuntimeError("sharded batch search: a peer GPU did not deliver in time, or the ranks asked "
                               "different questions (every rank must issue the same searches)")
        for q in flagged.tolist():
            r_ids, r_d = self.search(queries[q], k)
            counts[q] = len(r_ids)
            ids[q, :len(r_ids)] = r_ids
            dist[q, :len(r_d)] = r_d
        unused = np.arange(ids.shape[1])[None, :] >= counts[:, None]     # as GpuIndex.search leaves them
        ids[unused] = -1
        dist[unused] = np.nan
        return ids, dist, counts

    def nan_rows(self) -> int:
        """Admitted rows with NaN distance over all shards for the last search."""
        if self.fused and self._k is not None and 1 <= self._k <= self.FUSED_K_MAX:
            return int(self.backend.out_nan.cpu()[0])
        src = self.gathered if self.world > 1 else self.record.view(1, -1)
        return int(src[:, :8].contiguous().view(-1).cpu().numpy().view(np.int64).sum())

    def search(self, query: np.ndarray, k: int, use_mask: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """Host query in, host (rowids, distances) out; synchronous."""
        if k <= 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
        d_query = self.backend.to_device(query)
        dist, rowids, n = self.search_device(d_query, k, use_mask)
        if hasattr(self.backend, "fetch"):
            return self.backend.fetch(k)
        m = int(n.cpu()[0])
        # copies: the output tensors are reused by the next search
        return rowids[:m].cpu().numpy().copy(), dist[:m].cpu().numpy().copy()
