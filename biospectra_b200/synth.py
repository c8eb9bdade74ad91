"""Synthetic workloads of SURVEY.md 8(d): i.i.d. uniform ACGT reference genomes and reads that
follow the reference's own simulator (verification/RandomReadSampler.java:99-149 and
verification/SequencingArtifactsSimulator.java:38-82): uniform start inside one entry, forward
strand only, length R + U{-R/10 .. R/10-1}, exactly floor(len*e) substitutions at distinct
positions, each to one of the 3 other bases uniformly.

Two generators with the same distribution:
  * numpy (PCG64)  -- CPU, used for the small configs and the tests;
  * torch on CUDA  -- used by bench.py for the 10 Gbp config, genomes are produced directly in
    device memory and handed to the indexer as device pointers (nothing that large crosses PCIe).
"""
from __future__ import annotations

import numpy as np

ASCII = np.frombuffer(b"ACGT", np.uint8)


def genome_lengths(n_genomes: int, total_bp: int, seed: int, lo=1.5e6, hi=6.5e6) -> np.ndarray:
    """log-uniform lengths scaled to sum to total_bp (C2/C3/C4); equal lengths if lo == hi."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if lo == hi:
        L = np.full(n_genomes, total_bp // n_genomes, np.int64)
    else:
        raw = np.exp(rng.uniform(np.log(lo), np.log(hi), n_genomes))
        L = np.maximum((raw * (total_bp / raw.sum())).astype(np.int64), 1000)
    return L


def header_of(g: int) -> str:
    return "syn|%d| synthetic genome %d" % (g, g)


# ------------------------------------------------------------------------ numpy
def genome_numpy(seed: int, g: int, length: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64([seed, g]))
    return ASCII[rng.integers(0, 4, length, dtype=np.uint8)]


def reads_from_genome_numpy(rng, genome: np.ndarray, n: int, R: int, err: float):
    """-> list of uint8 arrays"""
    out = []
    L = len(genome)
    for _ in range(n):
        ln = R + int(rng.integers(0, max(R // 5, 1))) - R // 10
        if ln <= 0:
            ln = R
        ln = min(ln, L)
        st = int(rng.integers(0, L - ln + 1))
        rd = genome[st:st + ln].copy()
        ne = int(ln * err)
        if ne:
            pos = rng.permutation(ln)[:ne]
            code = np.searchsorted(ASCII, rd[pos])              # A0 C1 G2 T3 (ASCII order)
            rd[pos] = ASCII[(code + rng.integers(1, 4, ne)) % 4]
        out.append(rd)
    return out


def sample_reads_numpy(genomes, n_reads: int, R: int, err: float, seed: int):
    """Vectorised version of reads_from_genome_numpy over several genomes (uint8 arrays): same distribution
    (uniform genome by length, uniform start, length R + U{-R/10..R/10-1}, exactly floor(len*err) substitutions at
    distinct positions).  -> (list of bytes, source genome index array)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens_g = np.array([len(g) for g in genomes], np.int64)
    per = rng.multinomial(n_reads, lens_g / lens_g.sum())
    maxlen = R + max(R // 5, 1)
    out, src = [], []
    for gi, g in enumerate(genomes):
        n = int(per[gi])
        if not n:
            continue
        L = len(g)
        ln = R + rng.integers(0, max(R // 5, 1), n) - R // 10
        ln = np.minimum(np.where(ln <= 0, R, ln), L)
        st = (rng.random(n) * (L - ln + 1)).astype(np.int64)
        st = np.minimum(st, L - ln)
        cols = np.arange(maxlen)[None, :]
        valid = cols < ln[:, None]
        win = g[np.minimum(st[:, None] + cols, L - 1)]
        ne = np.floor(ln * err).astype(np.int64)
        key = np.where(valid, rng.random((n, maxlen)), 2.0)
        rank = key.argsort(axis=1).argsort(axis=1)
        sub = (rank < ne[:, None]) & valid
        code = np.searchsorted(ASCII, win)
        win = np.where(sub, ASCII[(code + rng.integers(1, 4, (n, maxlen))) % 4], win)
        for i in range(n):
            out.append(win[i, :ln[i]].tobytes())
        src.extend([gi] * n)
    order = rng.permutation(len(out))
    return [out[i] for i in order], np.asarray(src, np.int64)[order]


def make_small_workload(n_genomes=10, genome_len=2_000_000, n_reads=10_000, R=100, err=0.04, seed=1, read_seed=11):
    """C1-style workload entirely on the CPU.  Returns (genomes[list of bytes], reads[list of bytes], src[list])."""
    genomes = [genome_numpy(seed, g, genome_len) for g in range(n_genomes)]
    rng = np.random.Generator(np.random.PCG64(read_seed))
    src = rng.integers(0, n_genomes, n_reads)
    reads = []
    for g in src:
        reads.append(reads_from_genome_numpy(rng, genomes[int(g)], 1, R, err)[0].tobytes())
    return [x.tobytes() for x in genomes], reads, [int(x) for x in src]


# ------------------------------------------------------------------------ torch / CUDA
def build_reference_and_reads_cuda(indexer, lengths, n_reads: int, R: int, err: float, seed: int, read_seed: int,
                                   device, keep_genomes=(), rank_slice=None, n_frac: float = 0.0):
    """Generates every genome on `device`, adds it to `indexer` through the device-pointer entry
    point, and samples reads from it while it is alive.

    Returns dict(seq_bytes uint8[total] (pinned host), offsets int64[n+1], src int32[n] (source genome),
    kept {g: bytes} for genomes listed in keep_genomes)."""
    import torch

    lengths = np.asarray(lengths, np.int64)
    G = len(lengths)
    prng = np.random.Generator(np.random.PCG64(read_seed))
    per = prng.multinomial(n_reads, lengths / lengths.sum()) if n_reads else np.zeros(G, np.int64)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    gen = torch.Generator(device=device)
    chunks, lens_all, src_all = [], [], []
    kept = {}
    maxlen = R + max(R // 5, 1)
    for g in range(G):
        L = int(lengths[g])
        wanted = indexer is not None and (rank_slice is None or rank_slice(g))
        if not wanted and g not in keep_genomes and int(per[g]) == 0:
            continue                                    # nobody needs this genome (another rank's slice, no reads drawn)
        gen.manual_seed(int(seed) * 1_000_003 + g)
        codes = torch.randint(0, 4, (L,), dtype=torch.uint8, device=device, generator=gen)
        seq = lut[codes.long()]
        del codes
        if wanted:
            # the library packs on its own (non-blocking) stream: the genome must be complete first
            torch.cuda.current_stream(device).synchronize()
            indexer.add_entry_device("%d.fna" % g, header_of(g), "", seq.data_ptr(), L)
        if g in keep_genomes:
            kept[g] = seq.cpu().numpy().tobytes()
        n = int(per[g])
        if n:
            gen.manual_seed(int(read_seed) * 1_000_003 + g)
            var = torch.randint(0, max(R // 5, 1), (n,), device=device, generator=gen) - R // 10
            ln = R + var
            ln = torch.where(ln <= 0, torch.full_like(ln, R), ln).clamp(max=L)
            st = (torch.rand(n, device=device, generator=gen, dtype=torch.float64) * (L - ln + 1).double()).long()
            st = torch.minimum(st, L - ln)
            idx = st[:, None] + torch.arange(maxlen, device=device)[None, :]
            valid = torch.arange(maxlen, device=device)[None, :] < ln[:, None]
            win = seq[idx.clamp(max=L - 1)]
            # exactly floor(len*err) substitutions at distinct positions: rank of a random key
            ne = torch.floor(ln.double() * err).long()
            key = torch.rand(n, maxlen, device=device, generator=gen)
            key = torch.where(valid, key, torch.full_like(key, 2.0))
            rank = key.argsort(dim=1).argsort(dim=1)
            sub = (rank < ne[:, None]) & valid
            shift = torch.randint(1, 4, (n, maxlen), device=device, generator=gen)
            code = torch.bucketize(win.long(), lut.long())
            win = torch.where(sub, torch.take(lut, (code + shift) % 4), win)
            chunks.append(win[valid])
            lens_all.append(ln)
            src_all.append(torch.full((n,), g, dtype=torch.int32, device=device))
        del seq
    if not chunks:
        return dict(seq_bytes=np.zeros(0, np.uint8), offsets=np.zeros(1, np.int64), src=np.zeros(0, np.int32), kept=kept)
    lens = torch.cat(lens_all)
    src = torch.cat(src_all)
    flat = torch.cat(chunks)
    del chunks
    # shuffle reads so that consecutive reads do not share a source genome
    n = lens.numel()
    gen.manual_seed(int(read_seed) * 7 + 13)
    perm = torch.randperm(n, device=device, generator=gen)
    off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(lens, 0)
    new_lens = lens[perm]
    new_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    new_off[1:] = torch.cumsum(new_lens, 0)
    # gather bytes: for every output byte, its source index
    read_of_byte = torch.repeat_interleave(torch.arange(n, device=device), new_lens)
    within = torch.arange(flat.numel(), device=device) - new_off[read_of_byte]
    gathered = flat[off[perm][read_of_byte] + within]
    if n_frac > 0:          # a fraction of the reads gets one 'N' (not part of the BASELINE workload: exercises the positional path)
        pick = torch.rand(n, device=device, generator=gen) < n_frac
        at = new_off[:-1] + (torch.rand(n, device=device, generator=gen) * new_lens.double()).long().clamp(max=int(new_lens.max()) - 1)
        at = torch.minimum(at, new_off[1:] - 1)
        gathered[at[pick]] = ord("N")
    seq_host = torch.empty(gathered.numel(), dtype=torch.uint8, pin_memory=True)
    seq_host.copy_(gathered)
    off_host = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
    off_host.copy_(new_off)
    return dict(seq_bytes=seq_host.numpy(), offsets=off_host.numpy(), src=src[perm].cpu().numpy(), kept=kept,
                _pin=(seq_host, off_host), d_seq=gathered, d_off=new_off)
