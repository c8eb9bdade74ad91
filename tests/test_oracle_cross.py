"""The fast C oracle must agree bit-for-bit with the literal Python spec oracle on seeded adversarial
corpora (junk characters, homopolymers, periodic sequence, shared substrings, reads with N): index
statistics, postings, positions, norm bytes, and per-read top-10 doc ids / score bit patterns / labels."""
import random

import numpy as np
import pytest

import common
import spec_oracle as so

CASES = []
_rng = random.Random(11)
for _i in range(30):
    CASES.append(dict(seed=100 + _i, k=_rng.choice([1, 2, 3, 3, 4, 5, 6, 8, 10, 12, 16, 31]),
                      min_strand=_rng.random() < 0.4, skips=_rng.choice([0, 0, 0, 1, 2, 5]),
                      alg=_rng.choice([0, 1, 2]), sim=_rng.choice([0, 1]),
                      msm=_rng.choice([0.0, 0.5, 0.5, 0.9, 1.0]), junk=_rng.choice([0, 0, 0.01, 0.05])))


def build_spec(ents, k, ms):
    ix = so.Index(k, ms)
    for i, s in enumerate(ents):
        ix.add_entry("f%d.fa" % i, "h%d" % i, "", s)
    return ix


@pytest.mark.parametrize("case", CASES, ids=lambda c: "s%d-k%d-a%d" % (c["seed"], c["k"], c["alg"]))
def test_c_oracle_matches_spec(case, oracle_lib):
    ents = common.make_corpus(case["seed"], junk=case["junk"], length=(0, 200))
    reads = common.make_reads(2000 + case["seed"], ents, n=16, max_len=90)
    k, ms = case["k"], case["min_strand"]
    spec = build_spec(ents, k, ms)
    ox = common.build_oracle(oracle_lib, ents, k, ms)
    st = ox.stats()
    assert st["n_docs"] == spec.max_doc
    assert st["n_terms"] == len(spec.postings)
    assert st["n_postings"] == sum(len(v) for v in spec.postings.values())
    assert st["n_positions"] == sum(d.ntok for d in spec.docs)
    for d, doc in enumerate(spec.docs):
        assert ox.doc_info(d) == (doc.ntok, doc.norm)
    for t, pl in spec.postings.items():
        docs, tfs, pos = ox.postings(t)
        assert docs.tolist() == sorted(pl)
        for d, p in zip(docs.tolist(), pos):
            assert p.tolist() == pl[d]
    res = ox.classify([r.encode() for r in reads], skips=case["skips"], algorithm=case["alg"], scoring=case["sim"],
                      msm=case["msm"], threads=2)
    for i, r in enumerate(reads):
        s = so.search(spec, r, skips=case["skips"], algorithm=case["alg"], scoring=case["sim"], msm_ratio=case["msm"])
        if s is None:
            assert res["status"][i] == 1, (i, r)
            continue
        hits, nc, msm = s
        assert res["status"][i] == 0
        assert (res["n_clauses"][i], res["msm"][i]) == (nc, msm)
        assert res["n_top"][i] == len(hits)
        assert res["top_doc"][i, :len(hits)].tolist() == [d for d, _ in hits], (i, r)
        assert res["top_score"][i, :len(hits)].tobytes() == np.array([x for _, x in hits], np.float32).tobytes(), (i, r)
        nh = sum(1 for _, x in hits if float(hits[0][1]) - float(x) == 0) if hits else 0
        assert res["n_hits"][i] == nh
        assert res["label"][i] == (0 if nh == 0 else (2 if nh == 1 else 1))


def test_fasta_reader_semantics():
    # fasta-utils FASTAReader#readNext
    txt = ">h1 desc\r\nacgt\nNN ac\n;comment\nGG\n>h2\rTT\r\n  \n"
    ents = so.read_fasta(txt)
    assert ents == [(">h1 desc", "ACGTNN ACGG"), (">h2", "TT")]
    with pytest.raises(ValueError):
        so.read_fasta("ACGT\n")
    with pytest.raises(ValueError):
        so.read_fasta(">a\n>b\nAC\n")
