"""Doc-partitioned mode on one GPU: two handles hold the two halves of the reference (what two ranks would hold);
the df all-reduce is done by hand on the exposed device buffers, the per-rank records are concatenated like the all_gather does,
and bsp_merge_topk_records_device must reproduce the single-index oracle (which is also what Lucene would return: idf, maxDoc
and avgdl are collection-wide)."""
import random

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,alg,sim,full", [(5, 2, 0, 1), (6, 0, 1, 1), (4, 1, 0, 0), (10, 2, 0, 0)])
def test_two_partitions_merge_to_single_index(oracle_lib, k, alg, sim, full):
    import torch
    from biospectra_b200 import Configuration, Indexer, partition
    rng = random.Random(50 + k)
    ents = [common.rand_seq(rng, rng.randint(80, 500)) for _ in range(13)]
    ents[5] = ents[2][:300] + common.rand_seq(rng, 40)
    ents[11] = ents[2][20:280]
    ents[7] = ents[7][:50] + "*" + ents[7][50:]                      # no reverse doc
    reads = common.make_reads(600 + k, ents, n=64, max_len=130)
    ox = common.build_oracle(oracle_lib, ents, k, False)
    res_o = ox.classify([r.encode() for r in reads], skips=0, algorithm=alg, scoring=sim, msm=0.5, threads=2)
    alg_name = ["NAIVE_KMER", "CHAIN_PROXIMITY", "PAIRED_PROXIMITY"][alg]
    dev = torch.device("cuda", 0)
    plan = partition.plan_entries([len(s) for s in ents], 2)
    engines = []
    for r, (e0, e1) in enumerate(plan):
        ix = Indexer(Configuration(index_path="", kmer_size=k))
        for i in range(e0, e1):
            ix.add_entry("f%d.fa" % i, "h%d" % i, "", ents[i])
        clf = ix.close_to_classifier(Configuration(index_path="", kmer_size=k, kmer_skips=0, query_algorithm=alg_name,
                                                   scoring_algorithm="bm25" if sim else "tfidf", full_topn=full))
        engines.append(partition.GpuPartitionEngine(clf, dev))
    try:
        # what DocPartitionedClassifier.sync_statistics does with all_gather / all_reduce
        stats = [e.local_stats() for e in engines]
        dfs = [e.df_tensor() for e in engines]
        total = dfs[0] + dfs[1]
        for d in dfs:
            d.copy_(total)
        torch.cuda.synchronize()
        base = 0
        for e, (nd, _) in zip(engines, stats):
            e.set_global(base, sum(s[0] for s in stats), sum(s[1] for s in stats))
            base += nd
        data = np.frombuffer(b"".join(r.encode() for r in reads), np.uint8)
        off = np.zeros(len(reads) + 1, np.int64)
        np.cumsum([len(r) for r in reads], out=off[1:])
        # what DocPartitionedClassifier.classify does around its one all_gather_into_tensor: per-rank records, concatenated
        # rank-major, merged in place by bsp_merge_topk_records_device
        parts = []
        for e in engines:
            status, rec = e.classify_records(data, off)
            torch.cuda.synchronize()
            parts.append((status.clone(), rec.clone()))
        gathered = engines[0].gather_buffer(2, len(reads))
        gathered.copy_(torch.stack([p[1] for p in parts], dim=0))
        torch.cuda.synchronize()
        out = engines[0].merge_records(len(reads), 2, parts[0][0], gathered)
        # the three-array form of the same exchange (bsp_merge_topk_device) agrees
        recs = [p[1].cpu().numpy() for p in parts]
        n_top3 = torch.from_numpy(np.stack([r[:, 0] for r in recs], axis=1).copy()).to(dev)
        doc3 = torch.from_numpy(np.stack([r[:, 1:11] for r in recs], axis=1).copy()).to(dev)
        score3 = torch.from_numpy(np.stack([r[:, 11:21] for r in recs], axis=1).copy().view(np.float32)).to(dev)
        o3 = {k_: torch.empty(len(reads), dtype=torch.int32, device=dev) for k_ in ("status", "label", "n_hits", "n_top")}
        o3["top_doc"] = torch.empty(len(reads), 10, dtype=torch.int32, device=dev)
        o3["top_score"] = torch.empty(len(reads), 10, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        engines[0].clf.merge_topk_device(len(reads), 2, n_top3.data_ptr(), doc3.data_ptr(), score3.data_ptr(), parts[0][0].data_ptr(),
                                         o3["status"].data_ptr(), o3["label"].data_ptr(), o3["n_hits"].data_ptr(), o3["n_top"].data_ptr(),
                                         o3["top_doc"].data_ptr(), o3["top_score"].data_ptr())
        torch.cuda.synchronize()
        for key in o3:
            assert torch.equal(o3[key], out[key]), key
        torch.cuda.synchronize()
        res_g = {key: v.cpu().numpy() for key, v in out.items()}
        common.assert_classify_equal(res_g, res_o, reads, "two partitions k%d" % k, hits_only=not full)
        # stored fields are addressed with the ids the merge reports (global): every top hit resolves on the partition
        # that owns it, with the fields of that very document, and is refused by the other one
        doc_entry = []                                                 # global doc id -> (entry, direction)
        for i in range(len(ents)):
            doc_entry.append((i, "forward"))
            if i != 7:
                doc_entry.append((i, "reverse"))
        seen_owner = set()
        for i in range(len(reads)):
            for j in range(int(res_g["n_top"][i])):
                d = int(res_g["top_doc"][i, j])
                owner = 1 if d >= stats[0][0] else 0
                seen_owner.add(owner)
                f = engines[owner].clf.doc_fields(d)
                assert (f[1], f[2]) == ("h%d" % doc_entry[d][0], doc_entry[d][1]), (d, f)
                assert engines[owner].clf.doc_info(d) == ox.doc_info(d)
                with pytest.raises(Exception):
                    engines[1 - owner].clf.doc_fields(d)
        assert seen_owner == {0, 1}
    finally:
        for e in engines:
            e.clf.close()
