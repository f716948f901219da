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
    """ref = [(name, sequence)] on `src` (anything elsewhere) -> the same list on every rank.  The sequences travel as
    one uint8 tensor (no Python string joins of a multi-gigabase reference); see `broadcast_index` for shipping the
    BUILT index instead."""
    if dist.get_rank() == src:
        names = "\n".join(n for n, _ in ref).encode()
        lens = np.array([len(s) for _, s in ref], dtype=np.int64)
        cat = np.empty(int(lens.sum()), dtype=np.uint8)
        o = 0
        for (_, sq), ln in zip(ref, lens):
            cat[o:o + ln] = np.frombuffer(sq.encode() if isinstance(sq, str) else sq, dtype=np.uint8)
            o += int(ln)
    else:
        names, lens, cat = b"", np.zeros(0, np.int64), np.zeros(0, np.uint8)
    names = broadcast_bytes(names, src).decode().split("\n")
    lens = np.frombuffer(broadcast_bytes(lens.tobytes(), src), dtype=np.int64)
    cat = broadcast_array(cat, src)
    out, o = [], 0
    for n, ln in zip(names, lens):
        out.append((n, cat[o:o + int(ln)].tobytes().decode()))
        o += int(ln)
    return out


def broadcast_array(arr, src=0):
    """Broadcast a 1-D numpy array (dtype known on every rank from `arr.dtype`) from `src`; returns the array."""
    dev = _device()
    rank = dist.get_rank()
    n = torch.tensor([arr.size if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    if rank == src:
        buf = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)
    else:
        buf = torch.empty(int(n.item()) * arr.dtype.itemsize, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=src)
    return arr if rank == src else buf.cpu().numpy().view(arr.dtype)


def gather_records(rec_off, recs, cig, dst=0):
    """Per-rank results of `Aligner.align_packed` for the rank's block of reads -> on `dst`, the results of the whole
    super-batch in global read order (rec_off over all reads, records with cigar_off rebased, one CIGAR arena).
    One all_gather of the three sizes, then every rank's arrays go straight into their slices of the destination's
    buffers (unpadded point-to-point sends; nothing is padded to the largest rank)."""
    dev = _device()
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = np.diff(rec_off).astype(np.int64)
    mine = [np.ascontiguousarray(counts).view(np.uint8).reshape(-1), np.ascontiguousarray(recs).view(np.uint8).reshape(-1),
            np.ascontiguousarray(cig).view(np.uint8).reshape(-1)]
    sizes = torch.zeros((world, 3), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([[a.size for a in mine]], dtype=torch.int64, device=dev))
    sizes = sizes.cpu().numpy()
    if rank != dst:
        ops = []
        for a in mine:
            if a.size:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(a).to(dev, non_blocking=True), dst))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return None
    total = sizes.sum(axis=0)
    bufs = [torch.empty(int(total[k]), dtype=torch.uint8, device=dev) for k in range(3)]
    start = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(sizes, axis=0)])
    ops = []
    for r in range(world):
        for k in range(3):
            n = int(sizes[r, k])
            if n == 0:
                continue
            sl = bufs[k][int(start[r, k]):int(start[r, k]) + n]
            if r == dst:
                sl.copy_(torch.from_numpy(mine[k]).to(dev, non_blocking=True))
            else:
                ops.append(dist.P2POp(dist.irecv, sl, r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    all_counts = bufs[0].cpu().numpy().view(np.int64)
    out_recs = bufs[1].cpu().numpy().view(recs.dtype).copy()
    out_cig = bufs[2].cpu().numpy().view(cig.dtype)
    # rebase the CIGAR offsets rank by rank
    rec_start = start[:, 1] // max(recs.dtype.itemsize, 1)
    cig_start = start[:, 2] // max(cig.dtype.itemsize, 1)
    for r in range(world):
        out_recs["cigar_off"][int(rec_start[r]):int(rec_start[r + 1])] += int(cig_start[r])
    off = np.concatenate([[0], np.cumsum(all_counts)]).astype(np.int64)
    return off, out_recs, out_cig


class _DevView:
    """Device memory owned by someone else, exposed through the CUDA array interface (1-D uint8)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_index(index, src=0, ctx=None, device=None):
    """The BUILT index from `src` to every rank over the backend's broadcast (NCCL: HBM to HBM over NVLink): the five
    device arrays of `Index.arrays()` arrive in torch tensors on every other rank and are adopted there
    (`Index.adopt`) -- nobody but `src` sketches, sorts or hashes anything.  `index` is ignored on the other ranks.
    Needs the NCCL backend (the arrays live in device memory)."""
    from .align import Index
    rank = dist.get_rank()
    dev = _device()
    if dev.type != "cuda":
        raise RuntimeError("broadcast_index moves device arrays: NCCL backend only")
    if rank == src:
        arrs, meta = index.arrays()
        names = "\n".join(index.names).encode()
        lens = np.asarray(index.lens, dtype=np.int64)
        sizes = np.array([b for _, b in arrs], dtype=np.int64)
    else:
        arrs, meta, names, lens, sizes = None, np.zeros(8, np.int64), b"", np.zeros(0, np.int64), np.zeros(5, np.int64)
    names = broadcast_bytes(names, src).decode().split("\n")
    lens = np.frombuffer(broadcast_bytes(lens.tobytes(), src), dtype=np.int64)
    meta = np.frombuffer(broadcast_bytes(meta.tobytes(), src), dtype=np.int64)
    sizes = np.frombuffer(broadcast_bytes(sizes.tobytes(), src), dtype=np.int64)
    tensors = []
    for i in range(5):
        if rank == src:
            # a zero-copy torch view of the library's device array (CUDA array interface): the source of the broadcast
            t = torch.as_tensor(_DevView(arrs[i][0], int(sizes[i])), device=dev) if sizes[i] else torch.empty(0, dtype=torch.uint8, device=dev)
        else:
            t = torch.empty(int(sizes[i]), dtype=torch.uint8, device=dev)
        if sizes[i]:
            dist.broadcast(t, src=src)
        tensors.append(t)
    if rank == src:
        return index
    torch.cuda.synchronize()
    return Index.adopt(names, lens, [t.data_ptr() for t in tensors], sizes, meta, ctx=ctx, device=dev.index if device is None else device,
                       keep=tensors)
