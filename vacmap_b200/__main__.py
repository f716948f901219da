"""`python -m vacmap_b200 -ref R -read Q [Q ...] -mode H|L|S -o out.sam` -- the reference's command line
(`vacmap:95-296`) over the CUDA path, for the flags that reach the per-read alignment path and its SAM text.
Reads stream through in super-batches (one vm_align_submit per batch, the next one submitted before the previous
is collected); records of a batch are written in read order, a read that does not map writes nothing (quirk A2).
`<ref>.w{w}_k{k}.mmi` index files are read and written in minimap2's format (vacmap_b200/mmi.py).
Several GPUs: launch one process per GPU (`python -m torch.distributed.run --nproc-per-node N -m vacmap_b200 ...`,
the counterpart of the reference's `-t` worker processes, vacmap:391-420): rank 0 builds the index and broadcasts the
built tables (NCCL), super-batches go round robin to the ranks, every rank writes its SAM text to a part file and
rank 0 stitches the parts together in input order.
`-read x.bam` (unaligned BAM, as the reference reads it through pysam) and `-o out.bam` / `out.sorted.bam` (the reference
pipes its SAM text into samtools) are handled by vacmap_b200/bam.py.  Not mirrored: `-mode R`."""
import argparse
import os
import sys

import numpy as np

from . import align, bam, sam


RG_FLAGS = ("ID", "SM", "LB", "PL", "DS", "DT", "PU", "PI", "PG", "CN", "FO", "KS", "PM", "BC")


def rg_metadata(args):
    """collect_rg_metadata (vacmap:62-74) + the default group {"ID": "1", "SM": "sample"} (vacmap:214-218)."""
    rg = {}
    for f in RG_FLAGS:
        v = getattr(args, "rg_" + f.lower(), None)
        if v is not None:
            rg[f] = str(v)
    if rg and "ID" not in rg:
        sys.exit("The --rg-id option is required when any other --rg-* option is supplied.")
    return rg or {"ID": "1", "SM": "sample"}


def build_parser():
    p = argparse.ArgumentParser(prog="vacmap_b200", description="VACmap per-read alignment path on B200 (CUDA)")
    p.add_argument("-ref", required=True)
    p.add_argument("-read", required=True, nargs="+")
    p.add_argument("-mode", required=True, choices=["H", "L", "S", "asm"])
    p.add_argument("-workdir", help="accepted for compatibility with the reference's `-mode asm` (nothing is spilled to disk here)")
    p.add_argument("-o", default="-")
    p.add_argument("--force", action="store_true")
    p.add_argument("--nowriteindex", action="store_true", help="do not save the reference index (<ref>.w{w}_k{k}.mmi) for reuse")
    p.add_argument("-t", type=int, default=0, help="host glue threads (0 = all cores)")
    p.add_argument("-k", type=int, default=15)
    p.add_argument("-w", type=int, default=10)
    p.add_argument("-c", type=int, default=100)
    p.add_argument("-maxdivergence", type=float)
    p.add_argument("-globalpenalty", type=float)
    p.add_argument("-localpenalty", type=float)
    p.add_argument("-globalmaxdiff", type=int, default=50)
    p.add_argument("-localmaxdiff", type=int, default=30)
    for f in ("eqx", "MD", "L", "markunbalancetra", "nodiscard", "copycomments", "H", "fakecigar", "Q"):
        p.add_argument("--" + f, action="store_true")
    p.add_argument("--cs", nargs="?", const="short", default=None)
    p.add_argument("--rg-id", dest="rg_id", default=None)
    for f in RG_FLAGS[1:]:      # the other @RG fields of vacmap:134-150 (need --rg-id, vacmap:72-73)
        p.add_argument("--rg-" + f.lower(), dest="rg_" + f.lower(), default=None)
    p.add_argument("--debug", action="store_true", help="per-batch stage times and read counts on stderr (the reference's --debug prints its own trace)")
    p.add_argument("--batch-bases", type=int, default=150_000_000, help="bases per super-batch")
    p.add_argument("--device", type=int, default=0)
    return p


def options_from(args):
    """The `pdict` of vacmap:177-296 (mode defaults, `golbal_` spelling and all)."""
    opt = align.default_option("S" if args.mode == "asm" else args.mode)
    if args.mode == "asm":
        opt["eqx"] = True                  # vacmap:243-244: asm mode forces --eqx
    opt.update({"c": args.c, "eqx": args.eqx or args.mode == "asm", "md": args.MD, "cigar2cg": args.L, "copycomments": args.copycomments, "H": args.H,
                "fakecigar": args.fakecigar, "Q": args.Q, "mode": args.mode, "rg-id": rg_metadata(args)["ID"], "golbal_maxdiff": args.globalmaxdiff,
                "local_maxdiff": args.localmaxdiff, "shortcs": args.cs != "long"})
    if args.maxdivergence is not None:
        opt["maxdivergence"] = args.maxdivergence
    if args.globalpenalty is not None:
        opt["golbal_skipcost"] = args.globalpenalty
    if args.localpenalty is not None:
        opt["local_skipcost"] = args.localpenalty
    if args.markunbalancetra:
        opt["markunbalancetra"] = True
    if args.nodiscard:
        opt["nodiscard"] = True
    return opt


def read_records(path, want_comments=False):
    """Records of one input file: FASTA / FASTQ (.gz) through the reader, `.bam` as the reference takes them from pysam
    (vacmap:439-466)."""
    if path.endswith(".bam"):
        return bam.read_bam(path)
    return align.read_fastx(path, read_comment=want_comments)


class BamSink:
    """File-like front of bam.BamWriter: takes the SAM text (str or bytes, in arbitrary pieces) the emitter writes."""

    def __init__(self, path, header):
        self.w = bam.BamWriter(path, header)
        self.tail = ""
        self.closed = False

    def write(self, text):
        if not isinstance(text, str):
            text = bytes(text).decode()
        lines = (self.tail + text).split("\n")
        self.tail = lines.pop()
        self.w.write_sam_lines(lines)

    def close(self):
        if self.tail:
            self.w.write_sam_lines([self.tail])
            self.tail = ""
        self.w.close()
        self.closed = True


def batches(paths, want_comments, batch_bases):
    """Reads of all input files in order, cut into batches of ~batch_bases; a read name seen before is skipped
    (the reference's `unique_set`, vacmap:428-476)."""
    cur, n = [], 0
    seen = set()
    for path in paths:
        for rec in read_records(path, want_comments):
            if rec[0] in seen:
                continue
            seen.add(rec[0])
            cur.append(rec)
            n += len(rec[1])
            if n >= batch_bases:
                yield cur
                cur, n = [], 0
    if cur:
        yield cur


def stitch_parts(out, part_paths, sizes_per_rank):
    """Rank 0's last step of a multi-GPU run: batch b was written by rank b % world as the (b // world)-th block of
    its part file (`sizes_per_rank[r]` = byte length of each block); copy the blocks to `out` in batch order."""
    world = len(part_paths)
    files = [open(p, "rb") for p in part_paths]
    try:
        b = 0
        while True:
            r, i = b % world, b // world
            if i >= len(sizes_per_rank[r]):
                break
            n = sizes_per_rank[r][i]
            while n > 0:
                buf = files[r].read(min(n, 1 << 24))
                if not buf:
                    raise IOError("part file %s is shorter than its block list" % part_paths[r])
                out.write(buf)
                n -= len(buf)
            b += 1
    finally:
        for f in files:
            f.close()


def main(argv=None):
    args = build_parser().parse_args(argv)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        # one process per GPU (torchrun): this rank's device is its LOCAL_RANK
        import torch
        import torch.distributed as dist
        args.device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(args.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.device))
    if args.o != "-":
        if not args.o.endswith((".sam", ".bam")):
            sys.exit("output path must end in .sam or .bam (sorted.bam: coordinate-sorted) or be '-'")
        if os.path.isfile(args.o) and not args.force:
            sys.exit("output file exists (use --force)")
    opt = options_from(args)
    # vacmap:324-344 -- `<ref>.w{w}_k{k}.mmi` is used when it exists and written (minimap2's format) when it does not;
    # here the index itself is always built on the GPU, the file carries the sequences
    refpath = args.ref
    index_name = "%s.w%d_k%d.mmi" % (refpath, args.w, args.k)
    if not args.nowriteindex and os.path.isfile(index_name):
        refpath = index_name
    if dist is not None:
        from . import shard
        index = shard.broadcast_index(align.Index(refpath, w=args.w, k=args.k, device=args.device) if rank == 0 else None,
                                      device=args.device)
    else:
        index = align.Index(refpath, w=args.w, k=args.k, device=args.device)
    if rank == 0 and not args.nowriteindex and refpath != index_name and not refpath.endswith("mmi"):
        index.write_mmi(index_name)
    ref = [(n, index.seq(n)) for n in index.names]
    al = align.Aligner(index, opt, args.mode, host_threads=args.t)
    # the reference takes contig2seq from index.seq() (vacmap:363): non-ACGT bases are N there
    contig2seq = {n: index.seq(n) for n, _ in ref}
    contig2iloc = {n: i for i, (n, _) in enumerate(ref)}
    header = sam.header_text([(n, len(s)) for n, s in ref], rg=rg_metadata(args),
                             command_line=" ".join(sys.argv if argv is None else ["vacmap_b200"] + list(argv)))
    part_path = None
    if dist is not None:
        # this rank's SAM text goes to a part file next to the output (or in the temp directory for stdout)
        import tempfile
        stem = args.o if args.o != "-" else os.path.join(tempfile.gettempdir(), "vacmap_b200.%s" % os.environ.get("MASTER_PORT", "0"))
        part_path = "%s.part%d" % (stem, rank)
        out = open(part_path, "wb")
    else:
        if args.o.endswith(".bam"):
            out = BamSink(args.o, header)
        else:
            out = sys.stdout.buffer if args.o == "-" else open(args.o, "wb")
            out.write(header.encode())
    block_sizes = []      # multi-GPU: bytes of SAM text per unit (super-batch, or contig in asm mode) this rank owned
    mark = [0]

    def end_block():
        if dist is not None:
            out.flush()
            block_sizes.append(out.tell() - mark[0])
            mark[0] = out.tell()

    try:
        if args.mode == "asm":
            # one contig at a time (the reference's asm workers, vacmap:394-397 -> assembly_get_readmap_DP_test) and the mode's
            # own emitter (iterator_get_bam_dict_str, mammap_asm.py:22757-22941); several GPUs: contigs round robin
            from . import asm
            seen = set()
            unit = 0
            asm_table = None
            for path in args.read:
                for rec in read_records(path):
                    if rec[0] in seen:
                        continue
                    seen.add(rec[0])
                    unit += 1
                    if (unit - 1) % world != rank:
                        continue
                    rows = asm.assembly_align(rec[0], rec[1], index, opt)
                    if rows:
                        # the mode's own emitter (iterator_get_bam_dict_str), from the library's host code: MD / cs over a
                        # 60 Mb contig are not a job for a Python loop
                        if asm_table is None:
                            asm_table = sam.ContigTable(index)
                        rec_off, recs, cig = sam.pack_rows([rows], asm_table.names)
                        sam.batch_text([(rec[0], rec[1].upper()) + tuple(rec[2:])], rec_off, recs, cig, asm_table, opt, md=opt["md"],
                                       shortcs=opt["shortcs"], cigar2cg=opt["cigar2cg"], markunbalancetra=opt["markunbalancetra"],
                                       use_qual=not args.Q, threads=args.t, sink=out, asm=True)
                    end_block()
        else:
            pending = None

            table = sam.ContigTable(index)

            def collect(p):
                # the whole batch's SAM text from the library's host threads (csrc/vm_sam.cu: byte for byte what
                # sam.get_bam_dict_str / get_bam_dict_str_comments write read by read; a read on which the reference's emitter
                # raises writes nothing, like its worker, clrnano:24116-24125)
                handle, recs_in, packed = p
                rec_off, recs, cig = al.wait(handle)
                reads = [(r[0], None) + tuple(r[2:]) for r in recs_in]
                sam.batch_text(reads, rec_off, recs, cig, table, opt, md=opt["md"], shortcs=opt["shortcs"], cigar2cg=opt["cigar2cg"],
                               markunbalancetra=opt["markunbalancetra"], copycomments=args.copycomments, use_qual=not args.Q,
                               threads=args.t, sink=out, packed_seqs=packed)
                if args.debug:
                    top = sorted(((v, k) for k, v in al.last_stage_ms.items() if not k.startswith(("n_", "c_", "cpu_", "t_"))), reverse=True)[:8]
                    sys.stderr.write("[vacmap_b200] batch of %d reads, %d records; ms: %s\n"
                                     % (len(recs_in), len(recs), ", ".join("%s %.1f" % (k, v) for v, k in top)))
                end_block()

            # several GPUs: every rank parses the input (the read-name filter needs all names) and keeps every world-th batch
            for bi, batch in enumerate(batches(args.read, args.copycomments, args.batch_bases)):
                if bi % world != rank:
                    continue
                enc = [r[1].upper().encode() for r in batch]
                off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
                cat = b"".join(enc)
                nxt = (al.submit_packed(cat, off), batch, (cat, off))
                if pending is not None:
                    collect(pending)
                pending = nxt
            if pending is not None:
                collect(pending)
        if dist is not None:
            out.close()
            sizes = [None] * world
            dist.all_gather_object(sizes, block_sizes)      # doubles as the barrier: every part file is complete
            if rank == 0:
                final = sys.stdout.buffer if args.o == "-" else (BamSink(args.o, header) if args.o.endswith(".bam") else open(args.o, "wb"))
                try:
                    if not args.o.endswith(".bam"):
                        final.write(header.encode())
                    stitch_parts(final, ["%s.part%d" % (stem, r) for r in range(world)], sizes)
                finally:
                    if args.o != "-":
                        final.close()
            dist.barrier()
            os.remove(part_path)
    finally:
        if out is not sys.stdout.buffer and not out.closed:
            out.close()
        index.close()
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
