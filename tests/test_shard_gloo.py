"""The multi-GPU host logic on CPU: world_size 2 over gloo (the GPU box runs the same code over NCCL).
Sharding by bases, reference broadcast, and the gather of per-rank record arrays into global read order."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_results(lo, hi):
    """Deterministic stand-in for Aligner.align_packed on reads [lo, hi): read i has i % 3 records of i % 5 + 1 ops."""
    from vacmap_b200.align import RECORD_DTYPE
    counts = [(i % 3) for i in range(lo, hi)]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    recs = np.zeros(int(off[-1]), dtype=RECORD_DTYPE)
    cig = []
    k = 0
    for i in range(lo, hi):
        for j in range(i % 3):
            n = i % 5 + 1
            recs[k]["contig"] = i
            recs[k]["q_st"] = j
            recs[k]["cigar_off"] = len(cig)
            recs[k]["cigar_len"] = n
            cig += [(i * 16 + t) for t in range(n)]
            k += 1
    return off, recs, np.array(cig, dtype=np.uint32)


def _worker(rank, world, port, lengths, q):
    import torch.distributed as dist
    from vacmap_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref = [("chr1", "ACGT" * 50), ("chr2", "TTGACA" * 7)] if rank == 0 else None
        got_ref = shard.broadcast_reference(ref)
        blocks = shard.partition_by_bases(lengths, world)
        lo, hi = blocks[rank]
        res = shard.gather_records(*_fake_results(lo, hi))
        if rank == 0:
            off, recs, cig = res
            q.put(("ok", got_ref, blocks, off.tolist(), recs["contig"].tolist(), recs["cigar_off"].tolist(), cig.tolist()))
        else:
            assert res is None
            q.put(("ref", got_ref))
    finally:
        dist.destroy_process_group()


def test_partition_by_bases_balances_and_covers():
    from vacmap_b200 import shard
    rng = np.random.default_rng(5)
    for world in (1, 2, 4, 8):
        for lengths in ([15000] * 100, list(rng.integers(100, 60000, 57)), [60_000_000, 1000, 1000, 1000], [5], []):
            blocks = shard.partition_by_bases(lengths, world)
            assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == len(lengths)
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            if len(lengths) >= 4 * world and max(lengths) * 4 * world < sum(lengths):
                per = [sum(lengths[lo:hi]) for lo, hi in blocks]
                assert max(per) - min(per) <= 2 * max(lengths)


@pytest.mark.timeout(120)
def test_broadcast_and_gather_world2():
    lengths = [1000 + 37 * i for i in range(23)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    ref = [("chr1", "ACGT" * 50), ("chr2", "TTGACA" * 7)]
    ok = [m for m in msgs if m[0] == "ok"][0]
    other = [m for m in msgs if m[0] == "ref"][0]
    assert ok[1] == ref and other[1] == ref
    want_off, want_recs, want_cig = _fake_results(0, len(lengths))
    assert ok[3] == want_off.tolist()
    assert ok[4] == want_recs["contig"].tolist()
    assert ok[5] == want_recs["cigar_off"].tolist()
    assert ok[6] == want_cig.tolist()
