"""configuration.json knobs, mirroring src/biospectra/Configuration.java:37-44,78-166."""
from __future__ import annotations

import json

from . import _lib

_ALGS = {"NAIVE_KMER": 0, "CHAIN_PROXIMITY": 1, "PAIRED_PROXIMITY": 2}
_ALG_NAMES = {v: k for k, v in _ALGS.items()}
_KEYS = ("index_path", "kmer_size", "kmer_skips", "min_strand_kmer", "query_term_min_should_match",
         "worker_threads", "index_ram_buffer", "query_algorithm", "scoring_algorithm")


class Configuration:
    """Same keys and defaults as the reference.  GPU-only settings (device, partition,
    residency) are NOT configuration.json keys (the reference's Jackson mapper rejects
    unknown properties); they are attributes set in code / by environment."""

    def __init__(self, **kw):
        self.index_path = None
        self.kmer_size = 10                         # DEFAULT_KMERSIZE
        self.kmer_skips = 5                         # DEFAULT_KMERSKIPS
        self.min_strand_kmer = False
        self.query_term_min_should_match = 0.5
        self.worker_threads = 4
        self.index_ram_buffer = 16
        self.query_algorithm = "PAIRED_PROXIMITY"
        self.scoring_algorithm = "default"
        # GPU placement (not serialised)
        self.device = 0
        self.part_rank = 0
        self.part_count = 1
        self.dense_tables = 0                       # 0 auto, 1 on, 2 off
        self.positional = 0
        self.full_topn = 0                          # 1: report the whole top-10 list, not only the hits equal to the max
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown configuration key %r" % k)
            setattr(self, k, v)

    @classmethod
    def from_json(cls, text: str) -> "Configuration":
        if not text:
            raise ValueError("json is empty or null")
        d = json.loads(text)
        c = cls()
        for k, v in d.items():
            if k not in _KEYS:                      # FAIL_ON_UNKNOWN_PROPERTIES (JsonSerializer.java:33-35)
                raise ValueError("Unrecognized field %r" % k)
            setattr(c, k, v)
        if isinstance(c.query_algorithm, str) and c.query_algorithm not in _ALGS:
            raise ValueError("unknown query_algorithm %r" % c.query_algorithm)
        return c

    @classmethod
    def from_file(cls, path: str) -> "Configuration":
        with open(path, "r") as f:
            return cls.from_json(f.read())

    def to_json(self) -> str:
        d = {k: getattr(self, k) for k in _KEYS}
        if not isinstance(d["query_algorithm"], str):
            d["query_algorithm"] = _ALG_NAMES[d["query_algorithm"]]
        return json.dumps(d)

    @property
    def algorithm_id(self) -> int:
        a = self.query_algorithm
        return _ALGS[a] if isinstance(a, str) else int(a)

    @property
    def scoring_id(self) -> int:
        """Configuration.java:169-179: only "bm25" (any case) selects BM25."""
        s = self.scoring_algorithm
        if isinstance(s, int):
            return s
        return _lib.BM25 if (s or "").lower() == "bm25" else _lib.TFIDF

    def to_struct(self) -> _lib.BspConfig:
        c = _lib.BspConfig()
        c.kmer_size = int(self.kmer_size)
        c.kmer_skips = int(self.kmer_skips)
        c.min_strand_kmer = 1 if self.min_strand_kmer else 0
        c.query_algorithm = self.algorithm_id
        c.scoring_algorithm = self.scoring_id
        c.worker_threads = int(self.worker_threads)
        c.index_ram_buffer = int(self.index_ram_buffer)
        c.device = int(self.device)
        c.query_term_min_should_match = float(self.query_term_min_should_match)
        c.part_rank = int(self.part_rank)
        c.part_count = int(self.part_count)
        c.dense_tables = int(self.dense_tables)
        c.positional = int(self.positional)
        c.full_topn = int(self.full_topn)
        return c
