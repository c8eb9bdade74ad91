"""N > 1 host-side logic on CPU (gloo, world size 2): read sharding, the entry partition plan, and the
doc-partitioned orchestration (df all-reduce, doc-id bases, hit-list all_gather, merge) of
biospectra_b200/partition.py, driven with an oracle-backed engine and checked against the single-index oracle."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from biospectra_b200 import partition

TOPN = 10


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """The interface of partition.GpuPartitionEngine over the C oracle (CPU tensors)."""

    def __init__(self, ol, entries, k, qconf):
        self.device = torch.device("cpu")
        self.qconf = qconf
        self.ox = ol.OracleIndex(k, False)
        for s in entries:
            self.ox.add_entry(s.encode())
        self.ox.finish()
        self.df = torch.from_numpy(self.ox.dense_df().astype(np.int32))
        self.doc_base = 0

    def local_stats(self):
        return self.ox.stats()["n_docs"], self.ox.sum_ttf()

    def df_tensor(self):
        return self.df

    def set_global(self, doc_base, global_docs, global_sum_ttf):
        self.doc_base = doc_base
        self.ox.set_global(global_docs, global_sum_ttf, self.df.numpy().astype(np.uint32))

    def classify_records(self, seq_bytes, offsets, device_reads=None):
        reads = [bytes(seq_bytes[offsets[i]:offsets[i + 1]]) for i in range(len(offsets) - 1)]
        r = self.ox.classify(reads, **self.qconf)
        doc = r["top_doc"].copy()
        doc[doc >= 0] += self.doc_base
        # the record layout of pack_lists_kernel: n_top | doc[10] | score bits[10]
        rec = np.concatenate([r["n_top"][:, None], doc, r["top_score"].view(np.int32)], axis=1).astype(np.int32)
        assert rec.shape[1] == partition.REC_WORDS
        return torch.from_numpy(r["status"]), torch.from_numpy(np.ascontiguousarray(rec))

    def gather_buffer(self, world, n):
        return torch.empty(world, n, partition.REC_WORDS, dtype=torch.int32)

    def merge_records(self, n, parts, status, gathered):
        """numpy restatement of merge_records_kernel (collector order: score desc, doc asc; hits equal to the max)"""
        g = gathered.numpy()
        out = {"status": status.numpy().copy(), "n_top": np.zeros(n, np.int32), "n_hits": np.zeros(n, np.int32),
               "label": np.zeros(n, np.int32), "top_doc": np.full((n, TOPN), -1, np.int32), "top_score": np.zeros((n, TOPN), np.float32)}
        for i in range(n):
            cand = []
            for p in range(parts):
                rec = g[p, i]
                sc = rec[1 + TOPN:].view(np.float32)
                for j in range(int(rec[0])):
                    cand.append((-float(sc[j]), int(rec[1 + j])))
            cand.sort()
            cand = cand[:TOPN]
            out["n_top"][i] = len(cand)
            for j, (ns, d) in enumerate(cand):
                out["top_doc"][i, j] = d
                out["top_score"][i, j] = np.float32(-ns)
            nh = sum(1 for ns, _ in cand if ns == cand[0][0]) if cand else 0
            out["n_hits"][i] = nh
            out["label"][i] = 0 if nh == 0 else (2 if nh == 1 else 1)
        return out


def _worker(rank, world, port, ents, reads, k, qconf, expected):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_lib as ol
        plan = partition.plan_entries([len(s) for s in ents], world)
        e0, e1 = plan[rank]
        eng = OracleEngine(ol, ents[e0:e1], k, qconf)
        dp = partition.DocPartitionedClassifier(eng).sync_statistics()
        assert dp.global_docs == expected["n_docs"], (dp.global_docs, expected["n_docs"])
        assert dp.doc_base == expected["doc_base"][rank]
        data = np.frombuffer(b"".join(r.encode() for r in reads), np.uint8)
        off = np.zeros(len(reads) + 1, np.int64)
        np.cumsum([len(r) for r in reads], out=off[1:])
        res = dp.classify(data, off)
        assert dp.collectives == 1          # one all_gather per query batch
        exp = expected["res"]
        assert np.array_equal(res["status"], exp["status"])
        assert np.array_equal(res["n_top"], exp["n_top"]), (res["n_top"], exp["n_top"])
        assert np.array_equal(res["n_hits"], exp["n_hits"])
        assert np.array_equal(res["label"], exp["label"])
        for i in range(len(reads)):
            nt = exp["n_top"][i]
            assert np.array_equal(res["top_doc"][i, :nt], exp["top_doc"][i, :nt]), (i, res["top_doc"][i], exp["top_doc"][i])
            assert res["top_score"][i, :nt].tobytes() == exp["top_score"][i, :nt].tobytes(), i
        # read sharding (replica mode): blocks are contiguous, disjoint and cover the reads
        lo, hi = partition.shard_reads(len(reads), rank, world)
        blocks = [None] * world
        dist.all_gather_object(blocks, (lo, hi))
        assert blocks[0][0] == 0 and blocks[-1][1] == len(reads)
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,alg,sim", [(4, 2, 0), (5, 0, 1), (6, 1, 0)])
def test_doc_partitioned_matches_single_index(oracle_lib, k, alg, sim):
    rng = random.Random(31 + k)
    ents = [common.rand_seq(rng, rng.randint(60, 400)) for _ in range(11)]
    ents[4] = ents[1][:200] + common.rand_seq(rng, 50)            # related entries on both sides of the split
    ents[9] = ents[1][50:260]
    ents[6] = ents[6][:30] + "-" + ents[6][30:]                    # an entry without a reverse doc shifts the doc ids
    reads = common.make_reads(400 + k, ents, n=40, max_len=120)
    qconf = dict(skips=0, algorithm=alg, scoring=sim, msm=0.5, threads=1)
    ox = common.build_oracle(oracle_lib, ents, k, False)
    res = ox.classify([r.encode() for r in reads], **qconf)
    plan = partition.plan_entries([len(s) for s in ents], 2)
    assert plan[0][0] == 0 and plan[0][1] == plan[1][0] and plan[1][1] == len(ents) and 0 < plan[0][1] < len(ents)
    docs_per_entry = [1 if "-" in s else 2 for s in ents]
    expected = {"n_docs": ox.stats()["n_docs"], "doc_base": [0, sum(docs_per_entry[:plan[0][1]])],
                "res": {kk: res[kk] for kk in ("status", "n_top", "n_hits", "label", "top_doc", "top_score")}}
    mp.spawn(_worker, args=(2, free_port(), ents, reads, k, qconf, expected), nprocs=2, join=True)


def test_plan_entries_properties():
    rng = random.Random(3)
    for _ in range(200):
        n = rng.randint(1, 30)
        lengths = [rng.randint(1, 1000) for _ in range(n)]
        world = rng.randint(1, 9)
        plan = partition.plan_entries(lengths, world)
        assert plan[0][0] == 0 and plan[-1][1] == n
        assert all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
        assert all(a <= b for a, b in plan)
    assert partition.plan_entries([10] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert partition.shard_reads(10, 0, 3) == (0, 3) and partition.shard_reads(10, 2, 3) == (6, 10)
