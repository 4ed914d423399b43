"""Seeded synthetic mpileup2readcounts tables of the shapes named in BASELINE.json.

The reference publishes no data set for this path; the generator
follows SURVEY.md section 8(d): per-sample depth scale, NB depths, per (position, allele)
log-uniform error rate and over-dispersion, planted variants and a few common germline sites.

Layout produced: ``counts`` int32 [P][9][n] with planes in the table's column order
(depth, A, T, C, G, a, t, c, g — bin/needlestack.r:188-198) and ``ref`` uint8 [P] with
0..3 = A, T, C, G.
"""
from dataclasses import dataclass

import numpy as np

BASES = "ATCG"


@dataclass
class Config:
    name: str
    n: int
    P: int
    depth: float
    e_lo: float
    e_hi: float
    s_lo: float
    s_hi: float
    af_lo: float = 0.02
    af_hi: float = 0.5
    af_log: bool = False
    tn_pairs: int = 0
    seed: int = 20240000


CONFIGS = {
    "C1": Config("C1", 20, 10_000, 100, -4, -2, -2, 0, seed=20240001),
    "C2": Config("C2", 100, 1_000_000, 100, -4, -2, -2, 0, seed=20240002),
    "C3": Config("C3", 300, 50_000, 50_000, -4.5, -3, -3, -1.5, af_lo=-3.5, af_hi=-1, af_log=True, seed=20240003),
    "C4": Config("C4", 400, 10_000_000, 100, -4, -2, -2, 0, tn_pairs=200, seed=20240004),
    "C5": Config("C5", 50, 100_000_000, 30, -4, -2, -2, 0, seed=20240005),
}


def generate(cfg: Config, P=None, seed=None, p_variant=1e-3, p_germline=1e-4, p_refswap=0.0):
    """Returns (ref uint8 [P], counts int32 [P][9][n]).  Chunk-wise callers pass P and a seed.
    p_refswap: probability that a position has the reference genome carrying the MINOR allele (the planes of the
    REF base and of one ALT base are exchanged, so that ALT/DP > 0.8 in most samples): the reference-inversion
    branch of needlestack.r:330-337.  Drawn after everything else: 0 leaves the batches of earlier seeds unchanged."""
    P = cfg.P if P is None else int(P)
    n = cfg.n
    rng = np.random.Generator(np.random.PCG64(cfg.seed if seed is None else seed))
    scale = rng.lognormal(0.0, 0.25, size=n)
    mean_dp = cfg.depth * scale[None, :]
    size_dp = 20.0
    lam = rng.gamma(size_dp, mean_dp / size_dp, size=(P, n))
    DP = rng.poisson(lam).astype(np.int64)
    ref = rng.integers(0, 4, size=P).astype(np.uint8)
    # per (position, alt) error counts
    e = 10.0 ** rng.uniform(cfg.e_lo, cfg.e_hi, size=(P, 3))
    sg = 10.0 ** rng.uniform(cfg.s_lo, cfg.s_hi, size=(P, 3))
    y = np.empty((P, 3, n), dtype=np.int64)
    for k in range(3):
        m = e[:, k, None] * DP
        shape = 1.0 / sg[:, k, None]
        lam_k = rng.gamma(shape, np.maximum(m, 1e-300) / shape, size=(P, n))
        y[:, k, :] = rng.poisson(lam_k)
    # planted variants
    nvar = rng.binomial(P, p_variant)
    vpos = rng.choice(P, size=nvar, replace=False) if nvar else np.empty(0, dtype=np.int64)
    for p in vpos:
        k = rng.integers(0, 3)
        ns = min(n, 1 + rng.poisson(0.5))
        who = rng.choice(n, size=ns, replace=False)
        if cfg.af_log:
            af = 10.0 ** rng.uniform(cfg.af_lo, cfg.af_hi, size=ns)
        else:
            af = rng.uniform(cfg.af_lo, cfg.af_hi, size=ns)
        if cfg.tn_pairs:  # somatic: tumours only (samples 0..tn_pairs-1 are the tumours)
            who = who[who < cfg.tn_pairs] if np.any(who < cfg.tn_pairs) else who % cfg.tn_pairs
            af = af[: len(who)]
        y[p, k, who] += rng.binomial(DP[p, who], af)
    ngerm = rng.binomial(P, p_germline)
    gpos = rng.choice(P, size=ngerm, replace=False) if ngerm else np.empty(0, dtype=np.int64)
    for p in gpos:
        k = rng.integers(0, 3)
        carriers = rng.random(n) < 0.3
        if cfg.tn_pairs:  # germline: both members of a pair carry it
            half = carriers[: cfg.tn_pairs]
            carriers = np.concatenate([half, half])[:n]
        af = np.where(rng.random(n) < 0.5, 0.5, 1.0)
        y[p, k, carriers] += rng.binomial(DP[p, carriers], af[carriers])
    # alt counts cannot exceed depth in total: truncate in allele order
    room = DP.copy()
    for k in range(3):
        y[:, k, :] = np.minimum(y[:, k, :], room)
        room -= y[:, k, :]
    refcount = room
    counts = np.zeros((P, 9, n), dtype=np.int32)
    counts[:, 0, :] = DP
    alts = np.array([[b for b in range(4) if b != r] for r in range(4)])  # setdiff order
    fwd = rng.binomial(y, 0.5)
    rev = y - fwd
    rfwd = rng.binomial(refcount, 0.5)
    rrev = refcount - rfwd
    pidx = np.arange(P)
    for k in range(3):
        b = alts[ref, k]
        counts[pidx, 1 + b, :] = fwd[:, k, :]
        counts[pidx, 5 + b, :] = rev[:, k, :]
    counts[pidx, 1 + ref, :] = rfwd
    counts[pidx, 5 + ref, :] = rrev
    if p_refswap > 0:
        for p in np.where(rng.random(P) < p_refswap)[0]:
            b = alts[ref[p], rng.integers(0, 3)]
            r = ref[p]
            counts[p, [1 + r, 1 + b]] = counts[p, [1 + b, 1 + r]]
            counts[p, [5 + r, 5 + b]] = counts[p, [5 + b, 5 + r]]
    return ref, counts


def tn_pairs(cfg: Config):
    """0-based index sets for tumour/normal mode: samples [0, k) tumours, [k, 2k) their normals."""
    k = cfg.tn_pairs
    return dict(tindex=np.arange(k, dtype=np.int32), nindex=np.arange(k, 2 * k, dtype=np.int32),
                only_t=np.empty(0, dtype=np.int32), only_n=np.empty(0, dtype=np.int32))


def random_indel_cells(P, n, seed=0, p_site=0.02):
    """INS / DEL cells for write_table: dict (p, s) -> (ins_cell, del_cell) in the mpileup2readcounts
    spelling 'count:SEQ|count:seq' (upper case = forward strand, lower case = reverse)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    cells = {}
    for p in np.where(rng.random(P) < p_site)[0]:
        pool = ["".join(rng.choice(list("ACGT"), size=rng.integers(1, 5))) for _ in range(rng.integers(1, 4))]
        carriers = np.where(rng.random(n) < rng.choice([0.1, 0.5, 0.9]))[0]
        for s in carriers:
            out = []
            for _ in range(2):
                ent = []
                for sq in pool:
                    if rng.random() < 0.6:
                        if rng.random() < 0.8:
                            ent.append(f"{rng.integers(1, 40)}:{sq}")
                        if rng.random() < 0.8:
                            ent.append(f"{rng.integers(1, 40)}:{sq.lower()}")
                rng.shuffle(ent)
                out.append("|".join(ent) if ent else "NA")
            cells[(int(p), int(s))] = tuple(out)
    return cells


def write_table(path, ref, counts, chrom="chr1", start=1, names=None, indel_cells=None):
    """Write a real mpileup2readcounts TEXT table (3 + 11 n tab-separated columns, header line,
    indel cells NA) so the reference driver could consume it."""
    P, _, n = counts.shape
    names = names or [f"S{i + 1}" for i in range(n)]
    with open(path, "w") as f:
        hdr = ["chr", "loc", "ref"]
        for s in names:
            hdr += [f"{s}_{c}" for c in ("depth", "A", "T", "C", "G", "a", "t", "c", "g", "Insertion", "Deletion")]
        f.write("\t".join(hdr) + "\n")
        for p in range(P):
            cells = [chrom, str(start + p), BASES[ref[p]] if ref[p] < 4 else "N"]
            row = counts[p]
            for i in range(n):
                ins, dele = (indel_cells or {}).get((p, i), ("NA", "NA"))
                cells += [str(int(row[j, i])) for j in range(9)] + [ins, dele]
            f.write("\t".join(cells) + "\n")
