"""Row-sharded search over the GPUs of one box: one process per GPU (torch.distributed, NCCL over
NVLink), each rank mirrors a contiguous block of the collection's rows, produces a local top-k with
the single-GPU scan, and ONE all-gather of a packed per-rank record feeds the final merge kernel
(SURVEY.md 8e).  The reference is single-process (no counterpart); results are independent of the
number of ranks (appendix B-15).

torch is plumbing here (device buffers, streams, the process group); every search step runs in
libsyzgy_b200.so.  The per-rank compute is reached through a small adaptor (`CudaShard`) so that
the host logic of this file (partitioning, record layout, the collective, unpacking) can be
exercised on CPU under the gloo backend with a stand-in shard (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _capi


def shard_bounds(nrows: int, world: int) -> list[int]:
    """Contiguous row blocks, sizes differing by at most one: rank r owns [b[r], b[r+1])."""
    base, extra = divmod(int(nrows), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < extra else 0))
    return b


def record_layout(nq: int, k: int) -> tuple[int, int, int, int]:
    """Packed per-rank record exchanged by the all-gather, in 8-byte words:
    [ids nq*k u64][dist nq*k f64][n nq u32, padded to 8 bytes][flags nq u32, padded]
    -> (off_ids, off_dist, off_n, words); the flags start at off_n + (nq + 1) // 2."""
    off_ids, off_dist, off_n = 0, nq * k, 2 * nq * k
    words = off_n + 2 * ((nq + 1) // 2)
    return off_ids, off_dist, off_n, words


def flags_offset(nq: int, k: int) -> int:
    return 2 * nq * k + (nq + 1) // 2


class CudaShard:
    """The rank-local mirror on one B200, driven through the C ABI."""

    def __init__(self, dim: int, quant: int, metric: int, device: int):
        self.index = _capi.Index(dim, quant, metric, device)
        self.device = torch.device("cuda", device)

    def topk_into(self, tq: torch.Tensor, k: int, rec: torch.Tensor, nq: int, mask_id: int = -1, flags: int = 0,
                  batched: bool = False):
        off_ids, off_dist, off_n, _ = record_layout(nq, k)
        base = rec.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = self.index.search_batch_dev if batched else self.index.search_topk_dev
        fn(tq.data_ptr(), nq, k, base + 8 * off_ids, base + 8 * off_dist, base + 8 * off_n,
           stream, mask_id, flags, d_out_flags=base + 8 * flags_offset(nq, k))

    def merge_into(self, gathered: torch.Tensor, world: int, nq: int, k: int, out: torch.Tensor):
        off_ids, off_dist, off_n, words = record_layout(nq, k)
        g, o = gathered.data_ptr(), out.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        off_f = flags_offset(nq, k)
        self.index.merge_topk_dev(g + 8 * off_ids, g + 8 * off_dist, g + 8 * off_n, world, nq, k, o + 8 * off_ids,
                                  o + 8 * off_dist, o + 8 * off_n, stream, rank_stride_bytes=8 * words,
                                  d_g_flags=g + 8 * off_f, d_out_flags=o + 8 * off_f)

    def close(self):
        self.index.close()


class ShardedIndex:
    """One rank's view of a row-sharded collection."""

    def __init__(self, dim: int, quant: int, metric: int, rank: int = 0, world: int = 1, device: int = 0,
                 group=None, shard=None):
        self.dim, self.quant, self.metric = dim, quant, metric
        self.rank, self.world, self.group = rank, world, group
        self.shard = shard if shard is not None else CudaShard(dim, quant, metric, device)
        self.device = self.shard.device
        self._bufs = {}
        self.total_rows = 0
        self.uncertain_total = 0   # queries returned uncertified after the whole escalation ladder (this rank's count)
        self.digits_option = 0     # the caller's SZG_OPT_DIGITS / SZG_OPT_MIN_CANDIDATE_MODE (set them through set_options)
        self.min_mode_option = -1

    def set_options(self, digits: int = 0, min_candidate_mode: int = -1):
        self.digits_option, self.min_mode_option = digits, min_candidate_mode
        self.shard.index.set_option(_capi.OPT_DIGITS, digits)
        self.shard.index.set_option(_capi.OPT_MIN_CANDIDATE_MODE, min_candidate_mode)

    # -- ingest -------------------------------------------------------------------------------
    def fill_synthetic(self, seed: int, nrows: int):
        """Rank r mirrors rows [b[r], b[r+1]) of the synthetic collection `seed` (ids = row index)."""
        b = shard_bounds(nrows, self.world)
        self.shard.index.fill_synthetic(seed, b[self.rank], b[self.rank + 1] - b[self.rank])
        self.total_rows = nrows
        return b[self.rank], b[self.rank + 1]

    def upsert(self, ids, codes):
        """Routes records by id: rank = id mod world (the owner never changes, so replace/remove stay local)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(ids.size, -1)
        mine = (ids % np.uint64(self.world)) == np.uint64(self.rank)
        if mine.any():
            self.shard.index.upsert(ids[mine], codes[mine])

    def remove(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        mine = (ids % np.uint64(self.world)) == np.uint64(self.rank)
        return self.shard.index.remove(ids[mine]) if mine.any() else 0

    # -- search -------------------------------------------------------------------------------
    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            _, _, _, words = record_layout(nq, k)
            rec = torch.zeros(words, dtype=torch.int64, device=self.device)
            gathered = torch.zeros(words * self.world, dtype=torch.int64, device=self.device) if self.world > 1 else rec
            out = torch.zeros(words, dtype=torch.int64, device=self.device) if self.world > 1 else rec
            pin = self.device.type == "cuda"
            h_out = torch.zeros(words, dtype=torch.int64, pin_memory=pin)
            self._bufs[key] = (rec, gathered, out, h_out)
        return self._bufs[key]

    def search_topk_dev(self, tq: torch.Tensor, k: int, mask_id: int = -1, flags: int = 0, batched: bool = False) -> torch.Tensor:
        """Queries resident on the device ([nq, dim] float64).  Enqueues local scan -> all-gather ->
        merge on the current stream and returns the packed result record (device int64 words,
        see record_layout); nothing is synchronised with the host.  batched=True serves the local step with the
        tensor-core contraction (szg_search_batch_dev) instead of nq single-query scans."""
        nq = tq.shape[0]
        rec, gathered, out, _ = self._buffers(nq, k)
        if batched:
            self.shard.topk_into(tq, k, rec, nq, mask_id, flags, batched=True)
        else:
            self.shard.topk_into(tq, k, rec, nq, mask_id, flags)
        if self.world == 1:
            return rec
        dist.all_gather_into_tensor(gathered, rec, group=self.group)
        self.shard.merge_into(gathered, self.world, nq, k, out)
        return out

    def search_topk(self, queries, k: int, mask_id: int = -1, flags: int = 0, batched: bool = False):
        """Host buffers in, host buffers out (the call a user of the C ABI makes).  Returns
        (ids [nq,k] uint64, dist [nq,k] float64, n [nq] uint32)."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ValueError(f"query has {q.shape[1]} dimensions, collection has {self.dim}")
        nq = q.shape[0]
        if self.world == 1 and isinstance(self.shard, CudaShard):
            fn = self.shard.index.search_batch if batched else self.shard.index.search_topk
            ids, dd, n, _ = fn(q, k, mask_id, flags)
            return ids, dd, n
        tq = torch.from_numpy(q)
        if self.device.type == "cuda":
            tq = tq.pin_memory().to(self.device, non_blocking=True)
        out = self.search_topk_dev(tq, k, mask_id, flags, batched=batched)
        h_out = self._buffers(nq, k)[3]
        h_out.copy_(out, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        ids, dd, n = unpack_record(h_out.numpy(), nq, k)
        unc = unpack_flags(h_out.numpy(), nq, k) & 1
        self.uncertain_total = getattr(self, "uncertain_total", 0)
        if unc.any() and isinstance(self.shard, CudaShard) and not (flags & _capi.F_NO_FP64_VERIFY):
            # Some rank could not certify its local candidate set with the fast surrogate.  The merged flags are the OR over
            # the ranks, so every rank sees the same set and walks the same ladder as the single-device host call
            # (collect_and_escalate): the precise 3-digit surrogate first, then candidate sets of 64, 128, 256 rows; what is
            # still uncertified after that is counted and returned as it is.  The handle's options are restored afterwards.
            ix = self.shard.index
            sel = np.nonzero(unc)[0]
            ladder = [(3, -1), (3, 1), (3, 2), (3, 3)]  # (digits, minimum candidate mode)
            try:
                for digits, mode in ladder:
                    if sel.size == 0:
                        break
                    ix.set_option(_capi.OPT_DIGITS, digits)
                    ix.set_option(_capi.OPT_MIN_CANDIDATE_MODE, mode)
                    out2 = self.search_topk_dev(tq[torch.from_numpy(sel).to(self.device)].contiguous(), k, mask_id, flags)
                    h2 = self._buffers(len(sel), k)[3]
                    h2.copy_(out2, non_blocking=True)
                    torch.cuda.current_stream(self.device).synchronize()
                    i2, d2, n2 = unpack_record(h2.numpy(), len(sel), k)
                    ids[sel], dd[sel], n[sel] = i2, d2, n2
                    sel = sel[(unpack_flags(h2.numpy(), len(sel), k) & 1).astype(bool)]
            finally:
                ix.set_option(_capi.OPT_DIGITS, self.digits_option)
                ix.set_option(_capi.OPT_MIN_CANDIDATE_MODE, self.min_mode_option)
            self.uncertain_total += int(sel.size)
        return ids, dd, n

    def close(self):
        self.shard.close()


def unpack_flags(words: np.ndarray, nq: int, k: int) -> np.ndarray:
    w = np.ascontiguousarray(words[:record_layout(nq, k)[3]]).view(np.uint64)
    return w[flags_offset(nq, k):].view(np.uint32)[:nq].copy()


def unpack_record(words: np.ndarray, nq: int, k: int):
    off_ids, off_dist, off_n, total = record_layout(nq, k)
    w = np.ascontiguousarray(words[:total]).view(np.uint64)
    ids = w[off_ids:off_ids + nq * k].reshape(nq, k).copy()
    dd = w[off_dist:off_dist + nq * k].view(np.float64).reshape(nq, k).copy()
    n = w[off_n:off_n + (nq + 1) // 2].view(np.uint32)[:nq].copy()
    return ids, dd, n
