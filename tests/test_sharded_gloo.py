"""world_size-2 CPU test (gloo) of the multi-GPU host logic in syzgydb_b200/sharded.py: row partitioning,
the packed per-rank record, the single all-gather and the unpacking.  The rank-local compute (which is
CUDA-only in the product) is replaced by a stand-in shard built on the oracle; on a GPU box the same
logic runs for real in tests/test_gpu_parity.py (merge kernel) and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as o
from syzgydb_b200.sharded import ShardedIndex, record_layout, shard_bounds, unpack_record


class _FakeIndex:
    def __init__(self, dims, bits):
        self.dims, self.bits = dims, bits
        self.codes = np.zeros((0, o.vector_size(bits, dims)), np.uint8)
        self.ids = np.zeros(0, np.uint64)

    def fill_synthetic(self, seed, row0, nrows):
        self.codes = np.concatenate([self.codes, o.synth_rows(seed, row0, nrows, self.dims, self.bits)])
        self.ids = np.concatenate([self.ids, np.arange(row0, row0 + nrows, dtype=np.uint64)])

    def upsert(self, ids, codes):
        self.codes = np.concatenate([self.codes, codes])
        self.ids = np.concatenate([self.ids, ids])

    def close(self):
        pass


class OracleShard:
    """Stand-in for CudaShard: same adaptor interface, CPU tensors, oracle compute."""

    def __init__(self, dims, bits, metric):
        self.index = _FakeIndex(dims, bits)
        self.device = torch.device("cpu")
        self.dims, self.bits, self.metric = dims, bits, metric

    def topk_into(self, tq, k, rec, nq, mask_id=-1, flags=0, batched=False):
        self.batched_calls = getattr(self, "batched_calls", 0) + int(batched)
        off_ids, off_dist, off_n, _ = record_layout(nq, k)
        w = rec.numpy().view(np.uint64)
        for qi in range(nq):
            ri, rd, _ = o.search_exact(self.index.codes, self.index.ids, self.dims, self.bits, self.metric,
                                       tq[qi].numpy(), k=k)
            w[off_ids + qi * k: off_ids + qi * k + ri.size] = ri
            w[off_dist + qi * k: off_dist + qi * k + rd.size] = rd.view(np.uint64)
            w[off_n:].view(np.uint32)[qi] = ri.size

    def merge_into(self, gathered, world, nq, k, out):
        _, _, _, words = record_layout(nq, k)
        parts = [unpack_record(gathered.numpy()[g * words:(g + 1) * words], nq, k) for g in range(world)]
        off_ids, off_dist, off_n, _ = record_layout(nq, k)
        w = out.numpy().view(np.uint64)
        for qi in range(nq):
            ids = np.concatenate([p[0][qi, :p[2][qi]] for p in parts])
            dd = np.concatenate([p[1][qi, :p[2][qi]] for p in parts])
            order = sorted(range(ids.size), key=lambda j: (dd[j], str(int(ids[j]))))[:k]
            w[off_ids + qi * k: off_ids + qi * k + len(order)] = ids[order]
            w[off_dist + qi * k: off_dist + qi * k + len(order)] = dd[order].view(np.uint64)
            w[off_n:].view(np.uint32)[qi] = len(order)

    def close(self):
        pass


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims, bits, metric, n, k, nq = 24, 8, o.COSINE, 3001, 7, 3
        sh = ShardedIndex(dims, bits, metric, rank, world, shard=OracleShard(dims, bits, metric))
        r0, r1 = sh.fill_synthetic(5, n)
        assert (r0, r1) == tuple(shard_bounds(n, world)[rank:rank + 2])
        # routed upsert: every rank is handed all records, keeps id % world == rank
        extra_ids = np.arange(10_000, 10_040, dtype=np.uint64)
        extra = o.synth_rows(6, 0, 40, dims, bits)
        sh.upsert(extra_ids, extra)
        assert sh.shard.index.ids.size == (r1 - r0) + 20
        qs = o.synth_queries(9, 0, nq, dims)
        ids, dd, cnt = sh.search_topk(qs, k)
        allcodes = np.concatenate([o.synth_rows(5, 0, n, dims, bits), extra])
        allids = np.concatenate([np.arange(n, dtype=np.uint64), extra_ids])
        for qi in range(nq):
            ri, rd, _ = o.search_exact(allcodes, allids, dims, bits, metric, qs[qi], k=k)
            assert cnt[qi] == k and ids[qi].tolist() == ri.tolist() and np.array_equal(dd[qi], rd)
        # batched=True takes the same collective path (local batch step -> all-gather -> merge)
        ids2, dd2, cnt2 = sh.search_topk(qs, k, batched=True)
        assert sh.shard.batched_calls == 1 and np.array_equal(ids2, ids) and np.array_equal(dd2, dd) and np.array_equal(cnt2, cnt)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_sharded_search_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_and_record_layout():
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(10_000_000, 8)[-1] == 10_000_000 and len(set(np.diff(shard_bounds(10_000_000, 8)))) == 1
    off_ids, off_dist, off_n, words = record_layout(5, 10)
    assert (off_ids, off_dist, off_n, words) == (0, 50, 100, 106)
    w = np.zeros(words, dtype=np.int64)
    w.view(np.uint64)[0:50] = np.arange(50)
    w.view(np.float64)[50:100] = np.arange(50) * 0.5
    w[100:103].view(np.uint32)[:5] = 10
    ids, dd, n = unpack_record(w, 5, 10)
    assert ids[4, 9] == 49 and dd[1, 0] == 5.0 and n.tolist() == [10] * 5
