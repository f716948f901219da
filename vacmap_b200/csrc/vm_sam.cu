// SAM text of a batch's records, on the host threads of the library (no CUDA in this file; it needs no device).
//
// C++ twin of vacmap_b200/sam.py::get_bam_dict_str, which mirrors the reference's emitter for the per-read path
// (mammap_clrnano.py): get_bam_dict_str (:20841-21021), get_bam_dict_str_comments (:21022-), reassign_mapq
// (:11661-11707), mergecigar_ (:4773-4796), get_MD_CSshort / get_MD_CSlong (:19012-19112), P_alignmentstring
// (:5391-5424) and output_functions.nm_from_cigar (:300-349).  Same text, same quirks: tag order RG, [CG], SA, NM, MD,
// cs; MD / cs empty unless the CIGAR uses = / X; NM under --H computed at the reference's offsets (H does not advance
// the query); n_cigar counts numbers AND letters; a read whose NM / MD walk runs off a sequence emits nothing (the
// reference raises and its worker swallows the read).  The Python emitter writes ~5 Mbp/s per core; aligned reads
// arrive at ~2 Gbp/s per GPU.
#include "../../include/vacmap_b200.h"
#include "vm_hostpool.hpp"
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

struct vm_text {
    std::string data;
    std::vector<int64_t> off;      // [n_reads + 1]
};

namespace {

using vmp::parallel_for;

struct ReadDropped {};

// Bio.Seq's ambiguous DNA complement (sam.py::_COMP)
struct CompTable {
    unsigned char t[256];
    CompTable()
    {
        for (int i = 0; i < 256; ++i) t[i] = (unsigned char)i;
        const char *a = "ACGTUNRYKMBVDHSWacgtunrykmbvdhsw", *b = "TGCAANYRMKVBHDSWtgcaanyrmkvbhdsw";
        for (int i = 0; a[i]; ++i) t[(unsigned char)a[i]] = (unsigned char)b[i];
    }
};
const CompTable kComp;

inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
inline char lo(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }
inline void put_int(std::string &s, long long v) { s += std::to_string(v); }

const char kOps[] = "MIDNSHP=X";

struct Op { long long n; char op; };

struct Rec {
    int32_t contig, strand;     // strand +1 / -1
    long long q_st, q_en, r_st, r_en;
    int32_t mapq;
    std::vector<Op> ops;        // merged (mergecigar_)
    std::string cigar;          // "".join(oplist)
    long long nm = 0, nm_cigar = 0;
    std::string md, cs, fake;
};

// Python slice s[a:b] of a sequence of length n (a, b >= 0 here)
inline void pyslice(long long n, long long a, long long b, long long &lo_, long long &hi_)
{
    if (a < 0) a = std::max<long long>(0, a + n);
    if (b < 0) b = std::max<long long>(0, b + n);
    lo_ = std::min(a, n);
    hi_ = std::min(b, n);
    if (hi_ < lo_) hi_ = lo_;
}

// output_functions.nm_from_cigar: q / r are the (sliced) query and target
long long nm_from_cigar(const std::vector<Op> &ops, const char *q, long long qn, const char *r, long long rn)
{
    long long nm = 0, qp = 0, rp = 0;
    for (const Op &o : ops) {
        const long long n = o.n;
        switch (o.op) {
        case 'M': {
            if (qp + n > qn || rp + n > rn) throw ReadDropped();        // the reference's per-base loop raises IndexError
            long long d = 0;
            for (long long i = 0; i < n; ++i) d += up(q[qp + i]) != up(r[rp + i]);
            nm += d; qp += n; rp += n;
            break;
        }
        case 'I': nm += n; qp += n; break;
        case 'D': nm += n; rp += n; break;
        case 'N': rp += n; break;
        case 'S': qp += n; break;
        case '=': qp += n; rp += n; break;
        case 'X': nm += n; qp += n; rp += n; break;
        default: break;
        }
    }
    return nm;
}

// get_MD_CSshort / get_MD_CSlong
void md_cs(const std::vector<Op> &ops, const char *t, long long tn, const char *q, long long qn, bool shortcs, std::string &md, std::string &cs)
{
    md.clear(); cs.clear();
    long long refloc = 0, readloc = 0, equal_value = 0;
    char preop = 0;
    auto tslice = [&](long long a, long long b, long long &l, long long &h) { pyslice(tn, a, b, l, h); };
    for (const Op &o : ops) {
        const long long value = o.n;
        const char op = o.op;
        if (op == 'X') {
            if (equal_value > 0) put_int(md, equal_value);
            else if (preop == 'D') md += '0';
            for (long long j = 0; j < value; ++j) {
                // value == 0 still indexes target[refloc] / query[readloc] once in the reference (the j = 0 terms stand before the loop)
                if (refloc + j >= tn || readloc + j >= qn) throw ReadDropped();
                if (j > 0) md += '0';
                md += t[refloc + j];
                cs += '*';
                cs += lo(t[refloc + j]);
                cs += lo(q[readloc + j]);
            }
            if (value <= 0) {
                if (refloc >= tn || readloc >= qn) throw ReadDropped();
                md += t[refloc];
                cs += '*'; cs += lo(t[refloc]); cs += lo(q[readloc]);
            }
            refloc += value; readloc += value; equal_value = 0;
        } else if (op == '=') {
            if (shortcs) { cs += ':'; put_int(cs, value); }
            else {
                long long l, h;
                tslice(refloc, refloc + value, l, h);
                cs += '=';
                for (long long i = l; i < h; ++i) cs += up(t[i]);
            }
            refloc += value; readloc += value; equal_value += value;
        } else if (op == 'D') {
            if (equal_value > 0) put_int(md, equal_value);
            else if (preop == 'X') md += '0';
            long long l, h;
            tslice(refloc, refloc + value, l, h);
            md += '^';
            md.append(t + l, (size_t)(h - l));
            cs += '-';
            for (long long i = l; i < h; ++i) cs += lo(t[i]);
            refloc += value; equal_value = 0;
        } else if (op == 'I') {
            long long l, h;
            pyslice(qn, readloc, readloc + value, l, h);
            cs += '+';
            for (long long i = l; i < h; ++i) cs += lo(q[i]);
            readloc += value;
            continue;                          // preop unchanged
        } else if (op == 'S' || op == 'H') {
            continue;
        } else {
            md.clear(); cs.clear();
            return;                            // an M (or N, P) run: no MD / cs at all
        }
        preop = op;
    }
    if (equal_value > 0) put_int(md, equal_value);
}

void fake_cigar(const Rec &it, long long qlen, char clip, std::string &out)
{
    out.clear();
    if (it.q_st > 0) { put_int(out, it.q_st); out += clip; }
    const long long diff = it.q_en - it.q_st - it.r_en + it.r_st;
    if (diff > 0) { put_int(out, it.r_en - it.r_st); out += 'M'; put_int(out, diff); out += 'I'; }
    else if (diff < 0) { put_int(out, it.q_en - it.q_st); out += 'M'; put_int(out, -diff); out += 'D'; }
    else { put_int(out, it.q_en - it.q_st); out += 'M'; }
    if (qlen - it.q_en > 0) { put_int(out, qlen - it.q_en); out += clip; }
}

// reassign_mapq :11661-11707 on the rows in their original order
void reassign_mapq(std::vector<Rec> &rows)
{
    const int n = (int)rows.size();
    std::vector<int> g(1, 0);
    while (g.back() < n - 1) {
        const int iloc = g.back();
        int test = iloc;
        const Rec &b = rows[(size_t)iloc];
        bool hit = false;
        while (test + 1 < n) {
            ++test;
            const Rec &t = rows[(size_t)test];
            if (t.contig != b.contig) continue;
            const long long refgap = t.strand == 1 ? t.r_st - b.r_en : b.r_st - t.r_en;
            if (std::llabs(refgap) > 100000) continue;
            if (refgap < 10) { g.push_back(test); hit = true; break; }
        }
        if (!hit) g.push_back(iloc + 1);
    }
    std::vector<char> keep((size_t)n, 0);
    for (int i : g)
        if (i >= 0 && i < n) keep[(size_t)i] = 1;
    for (int i = 0; i < n; ++i)
        if (!keep[(size_t)i]) rows[(size_t)i].mapq = 0;
}

struct Ctx {
    const vm_sam_options *opt;
    int32_t n_contigs;
    const char *const *ctg_names;
    const char *const *ctg_seqs;
    const int64_t *ctg_lens;
};

void one_read(const Ctx &C, const vm_record *recs, int64_t n_rec, const uint32_t *cigar, const char *query, long long qlen,
              const char *name, long long name_len, const char *qual, long long qual_len, const char *comment, long long comment_len,
              std::string &out)
{
    out.clear();
    if (n_rec <= 0) return;
    const vm_sam_options &O = *C.opt;
    std::vector<Rec> rows((size_t)n_rec);
    for (int64_t i = 0; i < n_rec; ++i) {
        const vm_record &r = recs[i];
        Rec &x = rows[(size_t)i];
        x.contig = r.contig; x.strand = r.strand;
        x.q_st = r.q_st; x.q_en = r.q_en; x.r_st = r.r_st; x.r_en = r.r_en; x.mapq = r.mapq;
        // mergecigar_: adjacent runs of the same op merged.  asm mode's mergecigar_nm_ (mammap_asm.py:23126-23155) also sums
        // the X / D / I lengths for NM -- only where a run STARTS: a length merged into the run in front is not added
        for (int32_t k = 0; k < r.cigar_len; ++k) {
            const uint32_t c = cigar[r.cigar_off + k];
            const char op = kOps[(c & 15u) < 9u ? (c & 15u) : 0u];
            if (!x.ops.empty() && x.ops.back().op == op) x.ops.back().n += (long long)(c >> 4);
            else {
                x.ops.push_back(Op{(long long)(c >> 4), op});
                if (op == 'X' || op == 'D' || op == 'I') x.nm_cigar += (long long)(c >> 4);
            }
        }
        for (const Op &o : x.ops) { put_int(x.cigar, o.n); x.cigar += o.op; }
    }
    if (O.markunbalancetra) reassign_mapq(rows);
    // longest query span first; among equal spans the later row first (stable ascending sort, then reversed)
    std::stable_sort(rows.begin(), rows.end(), [](const Rec &a, const Rec &b) { return (a.q_en - a.q_st) < (b.q_en - b.q_st); });
    std::reverse(rows.begin(), rows.end());
    std::string rc_query((size_t)qlen, 'N');
    for (long long i = 0; i < qlen; ++i) rc_query[(size_t)i] = (char)kComp.t[(unsigned char)query[qlen - 1 - i]];
    const char clip = O.hardclip ? 'H' : 'S';
    for (Rec &it : rows) {
        const char *oriented = it.strand == 1 ? query : rc_query.data();
        if (it.contig < 0 || it.contig >= C.n_contigs) throw ReadDropped();
        long long tl, th;
        pyslice(C.ctg_lens[it.contig], it.r_st, it.r_en, tl, th);
        const char *target = C.ctg_seqs[it.contig] + tl;
        const long long tn = th - tl;
        if (!O.md) {
            it.nm = O.asm_mode ? it.nm_cigar : nm_from_cigar(it.ops, oriented, qlen, target, tn);
        } else {
            long long ql, qh;
            pyslice(qlen, it.q_st, it.q_en, ql, qh);
            md_cs(it.ops, target, tn, oriented + ql, qh - ql, O.shortcs != 0, it.md, it.cs);
            it.nm = O.asm_mode ? it.nm_cigar : nm_from_cigar(it.ops, oriented + ql, qh - ql, target, tn);
        }
        if (O.fakecigar) fake_cigar(it, qlen, clip, it.fake);
    }
    const bool have_qual = qual != nullptr && qual_len == qlen;
    const size_t m = rows.size();
    // asm mode (mammap_asm.py:22838-22841, 22873-22887): the second-longest record is primary when the longest has MAPQ 1 and
    // it has not; MAPQ is written as 60 (anything non-zero) or 1
    const size_t primary_iloc = (O.asm_mode && m > 1 && rows[0].mapq == 1 && rows[1].mapq != 1) ? 1 : 0;
    auto mq_of = [&](int32_t v) -> long long { return O.asm_mode ? (v != 0 ? 60 : 1) : (long long)v; };
    for (size_t iloc = 0; iloc < m; ++iloc) {
        const Rec &p = rows[iloc];
        std::string &L = out;
        // fixed fields: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL
        L.append(name, (size_t)name_len); L += '\t';
        put_int(L, (iloc == primary_iloc ? 0 : 2048) + (p.strand == 1 ? 0 : 16)); L += '\t';
        L += C.ctg_names[p.contig]; L += '\t';
        put_int(L, p.r_st + 1); L += '\t';
        put_int(L, mq_of(p.mapq)); L += '\t';
        const bool cg = (long long)p.ops.size() * 2 > 65535 && O.cigar2cg;
        if (cg) L += '*'; else L += p.cigar;
        L += "\t*\t0\t0\t";
        const char *seq = p.strand == 1 ? query : rc_query.data();
        long long sl = 0, sh = qlen;
        if (O.hardclip) pyslice(qlen, p.q_st, p.q_en, sl, sh);
        L.append(seq + sl, (size_t)(sh - sl));
        L += '\t';
        if (have_qual) {
            if (p.strand == 1) L.append(qual + sl, (size_t)(sh - sl));
            else
                for (long long i = sl; i < sh; ++i) L += qual[qlen - 1 - i];
        } else L += '*';
        // tags in the reference's dict insertion order
        if (O.rg_id) { L += "\tRG:Z:"; L += O.rg_id; }
        if (cg) { L += "\tCG:Z:"; L += p.cigar; }
        if (m > 1) {
            L += "\tSA:Z:";
            for (size_t t = 0; t < m; ++t) {
                if (t == iloc) continue;
                const Rec &it = rows[t];
                L += C.ctg_names[it.contig]; L += ',';
                put_int(L, it.r_st + 1); L += ',';
                L += it.strand == 1 ? '+' : '-'; L += ',';
                L += O.fakecigar ? it.fake : it.cigar; L += ',';
                put_int(L, mq_of(it.mapq)); L += ',';
                put_int(L, it.nm); L += ';';
            }
        }
        L += "\tNM:i:"; put_int(L, p.nm);
        if (O.md) { L += "\tMD:Z:"; L += p.md; L += "\tcs:Z:"; L += p.cs; }
        if (O.copycomments && comment != nullptr) {
            // P_alignmentstring_comments :20686-20730: well-formed TAG:TYPE:VALUE pieces of the FASTQ comment whose tag is new
            std::vector<std::string> seen = {"SA", "NM", "MD", "cs"};      // (the fixed field names are not two letters long)
            if (O.rg_id) seen.push_back("RG");
            if (cg) seen.push_back("CG");
            long long a = 0;
            while (a <= comment_len) {
                long long b = a;
                while (b < comment_len && comment[b] != '\t') ++b;
                const char *pc = comment + a;
                const long long n = b - a;
                long long c1 = -1, c2 = -1, colons = 0;
                for (long long i = 0; i < n; ++i)
                    if (pc[i] == ':') { ++colons; if (c1 < 0) c1 = i; else if (c2 < 0) c2 = i; }
                if (colons == 2 && c1 == 2 && c2 == c1 + 2 && strchr("AifZHB", pc[c1 + 1]) != nullptr) {
                    const std::string tag(pc, 2);
                    if (std::find(seen.begin(), seen.end(), tag) == seen.end()) {
                        L += '\t';
                        L.append(pc, (size_t)n);
                        seen.push_back(tag);
                    }
                }
                a = b + 1;
            }
        }
        L += '\n';
    }
}

} // namespace

extern "C" {

int vm_sam_batch(const vm_sam_options *opt, int32_t n_contigs, const char *const *contig_names, const char *const *contig_seqs,
                 const int64_t *contig_lens, int64_t n_reads, const int64_t *rec_off, const vm_record *recs, const uint32_t *cigar,
                 const char *seqs, const int64_t *seq_off, const char *names, const int64_t *name_off, const char *quals,
                 const int64_t *qual_off, const char *comments, const int64_t *comment_off, int32_t threads, vm_text **out)
{
    if (!opt || !out || n_reads < 0 || !rec_off || !seq_off || !name_off || n_contigs <= 0 || !contig_names || !contig_seqs || !contig_lens)
        return VM_ERR_ARG;
    *out = nullptr;
    vm_text *T = nullptr;
    try {
        T = new vm_text();
        std::vector<std::string> per((size_t)n_reads);
        Ctx C{opt, n_contigs, contig_names, contig_seqs, contig_lens};
        if (threads <= 0) threads = vmp::HostPool::get().size();
        parallel_for(n_reads, threads, [&](int64_t r) {
            const int64_t lo_ = rec_off[r], hi_ = rec_off[r + 1];
            if (hi_ <= lo_) return;
            try {
                one_read(C, recs + lo_, hi_ - lo_, cigar, seqs + seq_off[r], seq_off[r + 1] - seq_off[r], names + name_off[r],
                         name_off[r + 1] - name_off[r], (quals && qual_off && qual_off[r + 1] > qual_off[r]) ? quals + qual_off[r] : nullptr,
                         (quals && qual_off) ? qual_off[r + 1] - qual_off[r] : 0,
                         (comments && comment_off && comment_off[r + 1] > comment_off[r]) ? comments + comment_off[r] : nullptr,
                         (comments && comment_off) ? comment_off[r + 1] - comment_off[r] : 0, per[(size_t)r]);
            } catch (const ReadDropped &) {
                per[(size_t)r].clear();        // the reference's worker swallows the read (clrnano:24116-24125)
            }
        }, 8);
        T->off.assign((size_t)n_reads + 1, 0);
        for (int64_t r = 0; r < n_reads; ++r) T->off[(size_t)r + 1] = T->off[(size_t)r] + (int64_t)per[(size_t)r].size();
        T->data.resize((size_t)T->off[(size_t)n_reads]);
        parallel_for(n_reads, threads, [&](int64_t r) {
            if (!per[(size_t)r].empty()) memcpy(&T->data[(size_t)T->off[(size_t)r]], per[(size_t)r].data(), per[(size_t)r].size());
        }, 64);
    } catch (const std::bad_alloc &) {
        delete T;
        return VM_ERR_NOMEM;
    } catch (...) {
        delete T;
        return VM_ERR_ARG;
    }
    *out = T;
    return VM_OK;
}

const char *vm_text_data(vm_text *t) { return t ? t->data.data() : nullptr; }
int64_t vm_text_size(vm_text *t) { return t ? (int64_t)t->data.size() : 0; }
const int64_t *vm_text_offsets(vm_text *t) { return t ? t->off.data() : nullptr; }
void vm_text_free(vm_text *t) { delete t; }

} // extern "C"
