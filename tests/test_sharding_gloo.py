"""N > 1 path on CPU: two gloo ranks each own a contiguous region, run the per-unit path on it (here the
serial emulation of the kernel bodies stands in for the GPU) and rank 0 gathers the calls in genomic
order; the result must equal the single-process run.  No collective on the data path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from needlestack_b200 import shard


def test_shard_ranges_cover_exactly():
    for P in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(P, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _emul_result(ref, counts):
    import emul
    import parity
    from needlestack_b200 import _abi
    E = parity.emul_snv_batch(emul, ref, counts, emul.make_params())
    emit = np.where((E["flags"] & _abi.NS_F_GATE2) != 0)[0]
    out = {k: E[k][emit] for k in ("gq", "qval_minaf", "gt", "status", "pooled", "sor", "rvsb", "fs", "pvalue",
                                   "qvalue")}
    out["emit_unit"] = emit.astype(np.int64)
    out.update(n_fitted=int(((E["flags"] & _abi.NS_F_GATED) == 0).sum()),
               n_gated=int(((E["flags"] & _abi.NS_F_GATED) != 0).sum()), n_emitted=len(emit))
    return out


def _worker(rank, world, port, q):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import golden_util
    G = golden_util.load("c1_default")
    lo, hi = shard.shard_range(len(G["ref"]), world, rank)
    local = _emul_result(G["ref"][lo:hi], G["counts"][lo:hi])
    merged, tot = shard.gather_calls(local, lo, dist)
    if rank == 0:
        q.put((merged, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, tot = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import golden_util
    G = golden_util.load("c1_default")
    one = _emul_result(G["ref"], G["counts"])
    single, tot1 = shard.gather_calls(one, 0, None)
    assert tot == tot1
    assert np.array_equal(merged["emit_unit"], single["emit_unit"])
    assert np.array_equal(np.where(G["gate2"])[0], single["emit_unit"])
    for k in ("gq", "qval_minaf", "sor", "rvsb", "fs"):
        assert np.allclose(merged[k], single[k], rtol=0, atol=0, equal_nan=True), k
    assert np.array_equal(merged["gt"], single["gt"]) and np.array_equal(merged["status"], single["status"])
