"""Multi-GPU plumbing: one process per GPU (torch.distributed), reads sharded, no data-path collective.

Reads are independent units -- the reference itself hands them out one at a time to its worker processes
(`mammap_clrnano.py:24110-24117`) -- so rank r aligns a contiguous block of every super-batch, balanced by
cumulative BASE count (a 15 kb read and a 60 Mb contig are not the same amount of work).  Collectives, both
outside the per-read path: the reference sequence is broadcast once from rank 0 (every rank then builds its own
index replica in its HBM), and the alignment records of a super-batch are gathered on rank 0, which writes the SAM
text.  NCCL on the GPU box; the same code runs over gloo for the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def partition_by_bases(lengths, world):
    """Contiguous blocks [lo, hi) of reads, one per rank, with (nearly) equal cumulative base counts."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        # the boundary read goes to the side that leaves the split closer to the target
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` (length first, then the data as a uint8 tensor)."""
    dev = _device()
    n = torch.tensor([len(payload) if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy())


def broadcast_reference(ref, src=0):
    """ref = [(name, sequence)] on `src` (anything elsewhere) -> the same list on every rank."""
    if dist.get_rank() == src:
        names = "\n".join(n for n, _ in ref).encode()
        lens = np.array([len(s) for _, s in ref], dtype=np.int64).tobytes()
        seqs = "".join(s for _, s in ref).encode()
    else:
        names = lens = seqs = b""
    names = broadcast_bytes(names, src).decode().split("\n")
    lens = np.frombuffer(broadcast_bytes(lens, src), dtype=np.int64)
    seqs = broadcast_bytes(seqs, src).decode()
    out, o = [], 0
    for n, ln in zip(names, lens):
        out.append((n, seqs[o:o + int(ln)]))
        o += int(ln)
    return out


def _gather_array(arr, dst):
    """Variable-length gather of a 1-D contiguous array (as bytes) on `dst`: list of per-rank arrays there, None elsewhere."""
    dev = _device()
    world, rank = dist.get_world_size(), dist.get_rank()
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.size], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if raw.size:
        pad[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return [b[:sz].cpu().numpy().view(arr.dtype) for b, sz in zip(bufs, sizes)]


def gather_records(rec_off, recs, cig, dst=0):
    """Per-rank results of `Aligner.align_packed` for the rank's block of reads -> on `dst`, the results of the whole
    super-batch in global read order (rec_off over all reads, records with cigar_off rebased, one CIGAR arena)."""
    parts = [_gather_array(np.diff(rec_off).astype(np.int64), dst), _gather_array(recs, dst), _gather_array(cig, dst)]
    if dist.get_rank() != dst:
        return None
    counts, rec_parts, cig_parts = parts
    # rebase CIGAR offsets rank by rank
    out_recs, shift = [], 0
    for r, c in zip(rec_parts, cig_parts):
        r = r.copy()
        r["cigar_off"] += shift
        shift += len(c)
        out_recs.append(r)
    all_counts = np.concatenate(counts) if counts else np.zeros(0, np.int64)
    off = np.concatenate([[0], np.cumsum(all_counts)]).astype(np.int64)
    return off, np.concatenate(out_recs), np.concatenate(cig_parts)
