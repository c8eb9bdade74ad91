"""Multi-GPU plumbing (one process per GPU, torch.distributed).

* read-sharded mode: `shard_reads` -- every rank holds a replica of the index and classifies a contiguous block of
  the reads.  No collective on the data path (this is how the reference scales: replicated-index servers,
  classify/ClassifierClient.java:124-143).
* doc-partitioned mode: `DocPartitionedClassifier` -- rank r holds the documents of a contiguous slice of the
  reference entries.  Lucene's docFreq / maxDoc / sumTotalTermFreq are collection-wide
  (lucene!search/IndexSearcher#termStatistics, #collectionStatistics), so after the local build the ranks
  all-reduce the dense per-term df array and exchange (docs, tokens); at query time every rank scores all reads
  against its own docs, ONE all_gather moves the per-rank hit lists (one 84-byte record per read and rank: n_top |
  doc[10] | score[10]), and the merge reads the records where the collective left them and applies the collector order (score desc, global doc id asc), the equal-to-max rule (Classifier.java:358-366) and the label.

The engine behind it is anything with the small interface of `GpuPartitionEngine` (the CUDA library); the CPU
tests drive the same orchestration over gloo with an oracle-backed engine.
"""
from __future__ import annotations

import numpy as np

TOPN = 10


def shard_reads(n_reads: int, rank: int, world: int):
    """contiguous block [lo, hi) of the reads for this rank"""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def plan_entries(lengths, world: int):
    """Contiguous entry ranges balanced by bases: entry e goes to the rank whose base interval contains the midpoint
    of e (the same rule bsp_classifier_open applies to an index file when conf.part_count > 1).  -> [(e0, e1)]"""
    lengths = [int(x) for x in lengths]
    total = sum(lengths)
    n = len(lengths)
    out = []
    for rank in range(world):
        lo, hi = total * rank // world, total * (rank + 1) // world
        run, e0, e1, started = 0, n, n, False
        for e, ln in enumerate(lengths):
            mid = run + ln // 2
            if not started and mid >= lo:
                e0, started = e, True
            if mid >= hi:
                e1 = e
                break
            run += ln
        if not started:
            e0 = n
        if rank == world - 1:
            e1 = n
        out.append((e0, max(e0, e1)))
    return out


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory (no copy, no ownership)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


REC_WORDS = 1 + 2 * TOPN        # int32 words per (read, rank) record: n_top | doc[10] | score bits[10]


class GpuPartitionEngine:
    """Adapter of biospectra_b200.Classifier (a handle opened with conf.part_rank / part_count, or built from this
    rank's slice of the entries) to the interface DocPartitionedClassifier drives.  The reads go up from page-locked
    staging buffers on a side stream; outputs and records live in buffers that are reused from step to step."""

    def __init__(self, classifier, device):
        import torch
        self.clf = classifier
        self.device = torch.device(device)
        self._copy_stream = torch.cuda.Stream(self.device)
        self._pin = {}          # name -> pinned host tensor
        self._dev = {}          # name -> device tensor

    def local_stats(self):
        return self.clf.partition_local_stats()

    def df_tensor(self):
        import torch
        ptr, n = self.clf.partition_df_buffer()
        return torch.as_tensor(_DevArray(ptr, n, "<i4"), device=self.device)      # df < 2^31: int32 view of the u32 array

    def set_global(self, doc_base, global_docs, global_sum_ttf):
        self.clf.partition_set_global(doc_base, global_docs, global_sum_ttf)

    def _buf(self, name, shape, dtype, pinned=False):
        import torch
        pool = self._pin if pinned else self._dev
        t = pool.get(name)
        need = int(np.prod(shape))
        if t is None or t.dtype != dtype or t.numel() < need:
            t = (torch.empty(need, dtype=dtype, pin_memory=True) if pinned
                 else torch.empty(need, dtype=dtype, device=self.device))
            pool[name] = t
        return t[:need].view(*shape)

    def upload(self, seq_bytes, offsets):
        """host reads -> device (page-locked staging, asynchronous copies on a side stream; waited for before use)"""
        import torch
        seq_bytes = np.asarray(seq_bytes, np.uint8)
        offsets = np.asarray(offsets, np.int64)
        n = len(offsets) - 1
        h_seq = self._buf("seq", (len(seq_bytes),), torch.uint8, pinned=True)
        h_off = self._buf("off", (n + 1,), torch.int64, pinned=True)
        h_seq.numpy()[:] = seq_bytes
        h_off.numpy()[:] = offsets
        d_seq = self._buf("seq", (len(seq_bytes),), torch.uint8)
        d_off = self._buf("off", (n + 1,), torch.int64)
        with torch.cuda.stream(self._copy_stream):
            d_seq.copy_(h_seq, non_blocking=True)
            d_off.copy_(h_off, non_blocking=True)
        return d_seq, d_off, int(offsets[-1])

    def _outputs(self, tag, n):
        import torch
        o = {k: self._buf(tag + k, (n,), torch.int32) for k in ("status", "label", "n_hits", "n_top")}
        o["top_doc"] = self._buf(tag + "top_doc", (n, TOPN), torch.int32)
        o["top_score"] = self._buf(tag + "top_score", (n, TOPN), torch.float32)
        return o

    def classify_records(self, seq_bytes, offsets, device_reads=None):
        """reads -> (status[n], records[n, 21]) on the device, doc ids GLOBAL.  device_reads = (d_seq, d_off, total_bytes)
        skips the upload (reads already resident)."""
        import torch
        n = len(offsets) - 1
        d_seq, d_off, total = device_reads if device_reads is not None else self.upload(seq_bytes, offsets)
        o = self._outputs("loc_", n)
        rec = self._buf("rec", (n, REC_WORDS), torch.int32)
        self._copy_stream.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
        self.clf.classify_device(n, d_seq.data_ptr(), d_off.data_ptr(), total, o["status"].data_ptr(), o["label"].data_ptr(),
                                 o["n_hits"].data_ptr(), o["n_top"].data_ptr(), o["top_doc"].data_ptr(), o["top_score"].data_ptr())
        self.clf.pack_topk_device(n, o["n_top"].data_ptr(), o["top_doc"].data_ptr(), o["top_score"].data_ptr(), rec.data_ptr())
        return o["status"], rec

    def gather_buffer(self, world, n):
        import torch
        return self._buf("gathered", (world, n, REC_WORDS), torch.int32)

    def merge_records(self, n, parts, status, gathered):
        import torch
        o = self._outputs("out_", n)
        torch.cuda.current_stream(self.device).synchronize()        # the collective wrote `gathered` on torch's stream
        self.clf.merge_topk_records_device(n, parts, gathered.data_ptr(), status.data_ptr(), o["status"].data_ptr(),
                                           o["label"].data_ptr(), o["n_hits"].data_ptr(), o["n_top"].data_ptr(),
                                           o["top_doc"].data_ptr(), o["top_score"].data_ptr())
        return o


class DocPartitionedClassifier:
    """Orchestrates the collectives of the doc-partitioned mode over `torch.distributed`: at open time the statistics
    (one all_gather of two scalars + one all_reduce of the df array), per query batch exactly ONE all_gather."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.doc_base = 0
        self.global_docs = 0
        self.global_sum_ttf = 0
        self.collectives = 0         # data-path collectives issued by classify() so far

    def sync_statistics(self):
        """after the local index is built: global df (in place), maxDoc, sumTotalTermFreq, this rank's doc id base"""
        import torch
        import torch.distributed as dist
        n_docs, sum_ttf = self.engine.local_stats()
        mine = torch.tensor([n_docs, sum_ttf], dtype=torch.int64, device=self.engine.device)
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.group)
        docs = [int(t[0]) for t in every]
        self.doc_base = sum(docs[:self.rank])
        self.global_docs = sum(docs)
        self.global_sum_ttf = sum(int(t[1]) for t in every)
        df = self.engine.df_tensor()
        dist.all_reduce(df, op=dist.ReduceOp.SUM, group=self.group)
        self.engine.set_global(self.doc_base, self.global_docs, self.global_sum_ttf)
        return self

    def classify(self, seq_bytes, offsets, device_reads=None):
        """every rank passes the SAME reads; returns the merged per-read results (identical on every rank)"""
        import torch.distributed as dist
        n = len(offsets) - 1
        status, rec = self.engine.classify_records(seq_bytes, offsets, device_reads)
        gathered = self.engine.gather_buffer(self.world, n)
        dist.all_gather_into_tensor(gathered.view(self.world * n, REC_WORDS), rec, group=self.group)      # rank-major concatenation
        self.collectives += 1
        # the per-read status depends on the read only (tokens, clause count): identical on every rank
        return self.engine.merge_records(n, self.world, status, gathered)
