"""Shared helpers for the parity tests: seeded corpora / reads and the GPU-vs-oracle comparison."""
from __future__ import annotations

import random

import numpy as np


def rand_seq(rng, n, junk=0.0, junk_chars="NXRYKM acgt-*1"):
    out = []
    for _ in range(n):
        if junk and rng.random() < junk:
            out.append(rng.choice(junk_chars))
        else:
            out.append(rng.choice("ACGT"))
    return "".join(out)


def make_corpus(seed, n_entries=(1, 6), length=(0, 300), junk=0.0, related=True):
    """Adversarial small corpus: shared substrings, homopolymers, periodic sequence, junk chars."""
    rng = random.Random(seed)
    ents = []
    base = rand_seq(rng, rng.randint(30, max(31, length[1])))
    for _ in range(rng.randint(*n_entries)):
        L = rng.randint(*length)
        mode = rng.random()
        if related and mode < 0.3:
            s = base[:L] if L < len(base) else base
        elif mode < 0.45:
            s = rng.choice("ACGT") * L
        elif mode < 0.6:
            s = (rand_seq(rng, rng.randint(1, 4)) * (L + 1))[:L]
        else:
            s = rand_seq(rng, L, junk=junk)
        ents.append(s)
    return ents


def make_reads(seed, ents, n=12, max_len=120, err=0.04, n_frac=0.2):
    rng = random.Random(seed)
    qs = []
    while len(qs) < n:
        src = rng.choice(ents) if (ents and rng.random() < 0.8) else rand_seq(rng, 100)
        if len(src) > 5:
            a = rng.randint(0, len(src) - 5)
            q = src[a:a + rng.randint(1, max_len)]
        else:
            q = rand_seq(rng, rng.randint(1, 40))
        q = list(q)
        alphabet = "ACGTN" if rng.random() < n_frac else "ACGT"
        for i in range(len(q)):
            if rng.random() < err:
                q[i] = rng.choice(alphabet)
        q = "".join(q)
        if q:
            qs.append(q)
    return qs


def build_oracle(ol, ents, k, min_strand):
    ox = ol.OracleIndex(k, min_strand)
    for s in ents:
        ox.add_entry(s.encode())
    return ox.finish()


def build_gpu(ents, k, min_strand, qconf_kw, index_path=""):
    from biospectra_b200 import Configuration, Indexer
    iconf = Configuration(index_path=index_path, kmer_size=k, min_strand_kmer=min_strand)
    ix = Indexer(iconf)
    for i, s in enumerate(ents):
        ix.add_entry("f%d.fa" % i, "h%d" % i, "", s)
    qconf_kw = dict(qconf_kw)
    qconf_kw.setdefault("full_topn", 1)        # compare the whole top-10 list unless a test asks for the result set only
    q = Configuration(index_path=index_path, kmer_size=k, min_strand_kmer=min_strand, **qconf_kw)
    return ix.close_to_classifier(q)


def assert_index_equal(gpu, ox, all_terms=True, sample_terms=None):
    gs, os_ = gpu.stats(), ox.stats()
    assert gs == os_, (gs, os_)
    for d in range(os_["n_docs"]):
        assert gpu.doc_info(d) == ox.doc_info(d), d
    terms = sample_terms if sample_terms is not None else [ox.term_at(i) for i in range(os_["n_terms"])]
    for t in terms:
        od, otf, opos = ox.postings(t)
        gd, gtf, gpos = gpu.postings(t)
        assert np.array_equal(od, gd), t
        assert np.array_equal(otf, gtf), t
        assert np.array_equal(np.concatenate(opos) if opos else np.zeros(0, np.int32), gpos), t


def assert_classify_equal(res_g, res_o, reads=None, what="", hits_only=False):
    """hits_only: the GPU ran in result-set mode (conf.full_topn = 0): it reports exactly the hits equal to the maximum
    (n_top == n_hits), which must equal the head of the oracle's top-10 list."""
    n = len(res_o["status"])
    assert np.array_equal(res_g["status"], res_o["status"]), (what, np.nonzero(res_g["status"] != res_o["status"]))
    if hits_only:
        assert np.array_equal(res_g["n_top"], res_g["n_hits"]), what
        res_o = dict(res_o, n_top=res_o["n_hits"])
    for key in ("n_top", "n_hits", "label"):
        bad = np.nonzero(res_g[key] != res_o[key])[0]
        assert bad.size == 0, (what, key, bad[:5], reads[bad[0]] if reads is not None else None,
                               res_g[key][bad[:5]], res_o[key][bad[:5]])
    for i in range(n):
        nt = res_o["n_top"][i]
        assert np.array_equal(res_g["top_doc"][i, :nt], res_o["top_doc"][i, :nt]), \
            (what, i, reads[i] if reads is not None else None, res_g["top_doc"][i], res_o["top_doc"][i],
             res_g["top_score"][i], res_o["top_score"][i])
        assert res_g["top_score"][i, :nt].tobytes() == res_o["top_score"][i, :nt].tobytes(), \
            (what, i, reads[i] if reads is not None else None, res_g["top_score"][i], res_o["top_score"][i])
