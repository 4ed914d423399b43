"""Region sharding across GPUs (one process per GPU).

The reference's only parallelism is process-level data parallelism over genomic regions
(needlestack.nf:421-503: bed_cut.r splits the BED into --nsplit pieces, one Rscript per piece, the VCFs
are concatenated and sorted at the end).  Here contiguous position ranges are assigned to ranks; there
is no collective on the data path — only the per-rank summaries (and, when asked for, the called
records) are gathered on rank 0 in genomic order.
"""
import numpy as np


def shard_range(n_positions, world_size, rank):
    """contiguous [lo, hi) slice of the positions owned by `rank` (sizes differ by at most one)"""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_positions), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_calls(local, lo, dist=None, dst=0):
    """Gathers the emitted records of every rank on `dst`, re-based to global unit ids and merged in
    genomic order (the reference sorts at the end, needlestack.nf:546-553).  `local` is the dict
    returned by Caller.call_snv for the rank's own slice starting at position `lo`."""
    rec = {k: np.asarray(local[k]) for k in ("emit_unit", "gq", "qval_minaf", "gt", "status", "pooled", "sor", "rvsb",
                                              "fs", "pvalue", "qvalue")}
    rec["emit_unit"] = rec["emit_unit"] + 3 * int(lo)
    summary = dict(n_fitted=int(local["n_fitted"]), n_gated=int(local["n_gated"]), n_emitted=int(local["n_emitted"]))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rec, summary
    world = dist.get_world_size()
    parts = [None] * world if dist.get_rank() == dst else None
    dist.gather_object((rec, summary), parts, dst=dst)
    if dist.get_rank() != dst:
        return None, None
    order = np.argsort([p[0]["emit_unit"][0] if len(p[0]["emit_unit"]) else np.iinfo(np.int64).max for p in parts],
                       kind="stable")
    merged = {k: np.concatenate([parts[i][0][k] for i in order]) for k in rec}
    tot = {k: sum(p[1][k] for p in parts) for k in summary}
    return merged, tot
