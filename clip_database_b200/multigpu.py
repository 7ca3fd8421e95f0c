"""Row-sharded search over several GPUs of one box from ONE process (no torchrun).

Same sharding and the same kernels as ``sharded.ShardedIndex`` (contiguous rowid ranges, the fused
peer-memory exchange of ``clipdb_search_sharded_device``), but the ranks are contexts of this
process: their inboxes are wired with raw pointers (``cudaDeviceEnablePeerAccess``) instead of
CUDA IPC handles, and one host thread enqueues one launch per GPU — the kernels wait for each
other on the device, the host only waits for GPU 0's result.  This is what
``ImageDatabase(db_path, devices=[0, 1, ...])`` uses.  No reference equivalent (the reference is
one process, one thread, no GPU on this path).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .index import GpuIndex, SearchResult
from .sharded import RecordLayout, shard_bounds


class MultiGpuIndex:
    FUSED_K_MAX = 128

    def __init__(self, devices: Sequence[int], scan_ctas: Optional[int] = None, timeout_ms: int = 10000):
        """``devices``: CUDA device per shard, in rowid order.  The same device may appear more than
        once (tests on a 1-GPU box); then ``scan_ctas`` must leave room for all shards' kernels to
        be resident together (they wait for each other)."""
        import torch
        self.torch = torch
        self.devices = [int(d) for d in devices]
        self.world = len(self.devices)
        if self.world < 1:
            raise ValueError("at least one device")
        self.shards: List[GpuIndex] = []
        self._streams = []            # one stream per shard: shards on the same device must be able to overlap
        for d in self.devices:
            with torch.cuda.device(d):
                idx = GpuIndex(d)
                stream = torch.cuda.Stream(device=d)
                idx.set_stream(stream.cuda_stream)
                self._streams.append(stream)
            share = self.devices.count(d)
            if scan_ctas:
                idx.set_option("scan_ctas", int(scan_ctas))
            elif share > 1:      # shards sharing a GPU wait for each other: all their CTAs must be resident
                idx.set_option("scan_ctas", max(1, (idx.get_option("sm_count") - 4) // share))
            idx.set_option("xchg_timeout_ms", int(timeout_ms))
            self.shards.append(idx)
        self.bounds: List[Tuple[int, int]] = []
        self.num_rows = 0
        self._connected = False
        self._k = None
        self._masked = False          # every shard holds an admission mask
        self._issued = 0              # upper bound on the exchange sequence numbers used so far

    # ---- store ------------------------------------------------------------------------
    def load(self, rows: np.ndarray, rowids: Optional[np.ndarray] = None) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n = rows.shape[0]
        if n < self.world:
            raise ValueError("fewer rows than shards")
        ids = None if rowids is None else np.ascontiguousarray(rowids, dtype=np.int64)
        self.bounds = shard_bounds(n, self.world)
        for (lo, hi), idx in zip(self.bounds, self.shards):
            idx.load(rows[lo:hi], ids[lo:hi] if ids is not None else np.arange(lo, hi, dtype=np.int64))
        self.num_rows = n
        self._connect()

    def load_shard(self, rank: int, rows: np.ndarray, rowids: np.ndarray) -> None:
        """Load one shard at a time (a store larger than host memory is read range by range,
        ``loader.shard_rowid_range``); call ``finish_load`` after the last one."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.shape[0] < 1:
            raise ValueError("every shard needs at least one row")
        self.shards[rank].load(rows, np.ascontiguousarray(rowids, dtype=np.int64))
        self._shard_rows = getattr(self, "_shard_rows", {})
        self._shard_rows[rank] = rows.shape[0]

    def note_shard_loaded(self, rank: int, n_rows: int) -> None:
        """Shard ``rank`` was filled directly (``shards[rank].reserve`` + ``append``, the streaming loader)."""
        if n_rows < 1:
            raise ValueError("every shard needs at least one row")
        self._shard_rows = getattr(self, "_shard_rows", {})
        self._shard_rows[rank] = int(n_rows)

    def update_row(self, position: int, row) -> None:
        """Overwrite the row at global scan position ``position`` (in-place re-embedding)."""
        for (lo, hi), idx in zip(self.bounds, self.shards):
            if lo <= position < hi:
                idx.update_row(position - lo, row)
                return
        raise IndexError("position out of range")

    def finish_load(self) -> None:
        counts = [self._shard_rows[r] for r in range(self.world)]
        self.bounds, at = [], 0
        for c in counts:
            self.bounds.append((at, at + c))
            at += c
        self.num_rows = at
        self._connect()

    def _connect(self) -> None:
        if self._connected:
            return
        inboxes = [idx.exchange_init(self.world, r)[1] for r, idx in enumerate(self.shards)]
        for idx in self.shards:
            idx.exchange_connect_pointers(inboxes, self.devices)
        self._connected = True

    def append(self, rows, rowids=None) -> None:
        """New rows (larger rowids) go to the last shard: the ranges stay contiguous."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        lo, hi = self.bounds[-1]
        ids = None if rowids is None else np.ascontiguousarray(rowids, dtype=np.int64)
        self.shards[-1].append(rows, ids if ids is not None else np.arange(hi, hi + rows.shape[0], dtype=np.int64))
        self.bounds[-1] = (lo, hi + rows.shape[0])
        self.num_rows += rows.shape[0]
        if self._masked:               # the last shard dropped its mask (sized for the old row count): so do the others
            self.clear_mask()

    def note_appended(self, m: int) -> None:
        """``m`` rows were appended to the LAST shard directly (``shards[-1].append_sqlite``)."""
        lo, hi = self.bounds[-1]
        self.bounds[-1] = (lo, hi + int(m))
        self.num_rows += int(m)
        if self._masked:
            self.clear_mask()

    def set_mask(self, admitted) -> None:
        bits = np.asarray(admitted).astype(bool)
        if bits.shape != (self.num_rows,):
            raise ValueError("mask must have one entry per row")
        for (lo, hi), idx in zip(self.bounds, self.shards):
            idx.set_mask(bits[lo:hi])
        self._masked = True

    def clear_mask(self) -> None:
        for idx in self.shards:
            idx.clear_mask()
        self._masked = False

    def resync(self) -> None:
        """Bring the shards' exchanges back in step after a launch failed on some of them or a wait timed out:
        shards that did launch are released from their in-kernel wait (abort flag), every stream is drained, and
        all shards restart from one common sequence number above anything used so far (stale inbox records
        carry older numbers and are ignored)."""
        if not self._connected:
            return
        for idx in self.shards:
            idx.exchange_abort(True)
        for s in self._streams:
            s.synchronize()
        self._issued += 16
        for idx in self.shards:
            idx.exchange_abort(False)
            idx.exchange_set_epoch(self._issued, self._issued)

    @property
    def dim(self) -> int:
        return self.shards[0].dim

    @property
    def launch_count(self) -> int:
        return sum(s.launch_count for s in self.shards)

    # ---- search -------------------------------------------------------------------------
    def _prepare(self, k: int) -> None:
        if self._k == k:
            return
        t = self.torch
        lay = RecordLayout(max(k, 1))
        self._lay = lay
        self._out, self._q, self._views = [], [], []
        t.cuda.synchronize()
        for d in self.devices:
            dev = t.device("cuda", d)
            out = t.zeros(lay.nbytes, dtype=t.uint8, device=dev)
            self._out.append(out)
            self._q.append(t.empty(self.dim, dtype=t.float32, device=dev))
            kk = max(k, 1)
            self._views.append((out[lay.off_rowids:lay.off_rowids + 8 * kk].view(t.int64),
                                out[lay.off_dist:lay.off_dist + 4 * kk].view(t.float32),
                                out[lay.off_count:lay.off_count + 4].view(t.int32),
                                out[lay.off_nan:lay.off_nan + 8].view(t.int64)))
        self._h_q = t.empty(self.dim, dtype=t.float32).pin_memory()
        self._h_out = t.zeros(lay.nbytes, dtype=t.uint8).pin_memory()
        for d in set(self.devices):
            t.cuda.synchronize(d)      # buffers were created on torch's default streams
        self._k = k

    def search(self, query: np.ndarray, k: int, use_mask: bool = False) -> Tuple[np.ndarray, np.ndarray, int]:
        """(rowids, distances, NaN rows over all shards) for one query; synchronous.  One kernel
        launch per GPU; every GPU ends up with the merged answer, GPU 0's is fetched."""
        # everything that can be refused is refused BEFORE the first launch: a shard that launched while another
        # did not would wait in-kernel for a record that never comes, and the sequence numbers would drift apart
        if not 1 <= k <= self.FUSED_K_MAX:
            raise ValueError(f"k must be 1..{self.FUSED_K_MAX} on the multi-GPU path")
        if use_mask and not self._masked:
            raise ValueError("use_mask set but no mask installed (set_mask after the last append)")
        q_host = np.ascontiguousarray(query, dtype=np.float32).ravel()
        if q_host.shape[0] != self.dim:
            raise ValueError(f"query must be [{self.dim}]")
        t = self.torch
        self._prepare(k)
        self._h_q.numpy()[:] = q_host
        self._issued += 1
        try:
            for r, (d, idx) in enumerate(zip(self.devices, self.shards)):
                with t.cuda.device(d), t.cuda.stream(self._streams[r]):
                    self._q[r].copy_(self._h_q, non_blocking=True)
                    ids, dist, n, nan = self._views[r]
                    idx.search_sharded_device(self._q[r], k, ids, dist, n, nan, use_mask=use_mask)
            with t.cuda.device(self.devices[0]), t.cuda.stream(self._streams[0]):
                self._h_out.copy_(self._out[0], non_blocking=True)
            for s in self._streams:                     # results of GPU 0; every H2D of the pinned query done
                s.synchronize()
        except Exception:
            self.resync()
            raise
        lay = self._lay
        h = self._h_out.numpy()
        m = int(h[lay.off_count:lay.off_count + 4].view(np.int32)[0])
        if m < 0:
            self.resync()
            raise RuntimeError("multi-GPU search: a shard did not deliver its candidates in time"
                               if m == -1 else "multi-GPU search: the shards were asked different questions")
        ids = h[lay.off_rowids:lay.off_rowids + 8 * m].view(np.int64).copy()
        dist = h[lay.off_dist:lay.off_dist + 4 * m].view(np.float32).copy()
        nan = int(h[lay.off_nan:lay.off_nan + 8].view(np.int64)[0])
        return ids, dist, nan

    # ---- batched (bf16 pre-selection on every shard, candidates exchanged by the last kernel) ---
    def enable_batch(self) -> None:
        """bf16 copy on every shard.  The shards must be on DIFFERENT GPUs: the batched path's last
        kernel waits for the other shards while their contractions need whole SMs."""
        if len(set(self.devices)) != len(self.devices):
            raise ValueError("the batched multi-GPU path needs one GPU per shard")
        for idx in self.shards:
            idx.enable_batch()
        self.batch_enabled = True

    def search_batch(self, queries: np.ndarray, k: int) -> SearchResult:
        """nq queries over all shards: one batched pass per 256 queries on every GPU; queries some
        shard could not answer through the batched path are re-run through ``search``."""
        t = self.torch
        if not getattr(self, "batch_enabled", False):
            raise RuntimeError("call enable_batch() first")
        if not 1 <= k <= self.FUSED_K_MAX:
            raise ValueError(f"k must be 1..{self.FUSED_K_MAX} on the multi-GPU path")
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}]")
        if k > min(hi - lo for lo, hi in self.bounds):
            raise ValueError("k exceeds the smallest shard: use search() per query")
        nq = q.shape[0]
        h_q = t.from_numpy(q).pin_memory()
        outs = []
        self._issued += (nq + 255) // 256
        try:
            for r, (d, idx) in enumerate(zip(self.devices, self.shards)):
                with t.cuda.device(d), t.cuda.stream(self._streams[r]):
                    dev = t.device("cuda", d)
                    dq = h_q.to(dev, non_blocking=True)
                    o = (t.empty((nq, k), dtype=t.int64, device=dev), t.empty((nq, k), dtype=t.float32, device=dev),
                         t.zeros(nq, dtype=t.int32, device=dev), t.zeros(nq, dtype=t.int64, device=dev),
                         t.zeros(nq, dtype=t.int32, device=dev))
                    for q0 in range(0, nq, 256):
                        q1 = min(q0 + 256, nq)
                        idx.search_batch_sharded_device(dq[q0:q1], k, o[0][q0:q1], o[1][q0:q1], o[2][q0:q1],
                                                        o[3][q0:q1], o[4][q0:q1])
                    outs.append((dq, o))
            for s in self._streams:
                s.synchronize()
        except Exception:
            self.resync()
            raise
        ids, dist, n, nan, flags = (x.cpu().numpy() for x in outs[0][1])
        if (n < 0).any():
            self.resync()
            raise RuntimeError("multi-GPU batch search: a shard did not deliver in time")
        res = SearchResult(ids.copy(), dist.copy(), n.copy(), nan.copy())
        for qi in np.flatnonzero(flags).tolist():
            r_ids, r_d, r_nan = self.search(q[qi], k)
            res.counts[qi] = len(r_ids)
            res.rowids[qi, :len(r_ids)] = r_ids
            res.distances[qi, :len(r_d)] = r_d
            res.nan_rows[qi] = r_nan
        unused = np.arange(k)[None, :] >= res.counts[:, None]
        res.rowids[unused] = -1
        res.distances[unused] = np.nan
        return res

    def search_any_k(self, query: np.ndarray, k: int, use_mask: bool = False) -> Tuple[np.ndarray, np.ndarray, int]:
        """Any k: the fused path for 1 <= k <= 128, else per-shard searches merged on the host with
        the same (distance, rowid) order (shards are contiguous rowid ranges)."""
        if 1 <= k <= self.FUSED_K_MAX:
            return self.search(query, k, use_mask)
        if k <= 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32), 0
        parts = [idx.search(query, k, use_mask=use_mask) for idx in self.shards]
        ids = np.concatenate([p.row(0)[0] for p in parts])
        dist = np.concatenate([p.row(0)[1] for p in parts])
        shard = np.concatenate([np.full(int(p.counts[0]), r) for r, p in enumerate(parts)])
        pos = np.concatenate([np.arange(int(p.counts[0])) for p in parts])
        order = np.lexsort((pos, shard, dist))[:k]
        return ids[order], dist[order], int(sum(int(p.nan_rows[0]) for p in parts))

    # ---- the subset of GpuIndex that ImageDatabase uses -------------------------------------
    def blend(self, *args, **kwargs):
        return self.shards[0].blend(*args, **kwargs)

    def blend_search(self, e1, k: int, e2=None, weights=(0.5, 0.5), negatives=(), negative_weights=(),
                     metric="cosine", use_mask: bool = False) -> SearchResult:
        """Blend / negatives on GPU 0 (K3), then the sharded scan."""
        query, _flags = self.shards[0].blend(e1, e2, weights, negatives, negative_weights)
        if getattr(self, "prefer_batch", False) and getattr(self, "batch_enabled", False) and not use_mask \
                and 1 <= int(k) <= min(self.FUSED_K_MAX, min(hi - lo for lo, hi in self.bounds)):
            # tiered shards: the exact scan would stream the host tier over PCIe; the bf16 pre-selection + exact
            # re-rank gives the same answer from the resident copy
            return self.search_batch(query[None, :], int(k))
        ids, dist, nan = self.search_any_k(query, int(k), use_mask)
        kc = max(int(k), 0)
        res = SearchResult(np.full((1, kc), -1, dtype=np.int64), np.full((1, kc), np.nan, dtype=np.float32),
                           np.array([len(ids)], dtype=np.int32), np.array([nan], dtype=np.int64))
        res.rowids[0, :len(ids)] = ids
        res.distances[0, :len(ids)] = dist
        return res

    def load_codes(self, *args, **kwargs):
        return self.shards[0].load_codes(*args, **kwargs)

    def set_code_mask(self, *args, **kwargs):
        return self.shards[0].set_code_mask(*args, **kwargs)

    def binary_search(self, *args, **kwargs):
        return self.shards[0].binary_search(*args, **kwargs)

    def close(self) -> None:
        for idx in self.shards:
            idx.close()
        self.shards = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
