// Host-side construction of the reference index and its upload to HBM (see vm_index.cuh).
#include "vm_index.cuh"
#include <algorithm>
#include <cstring>

namespace {
struct HY {
    uint64_t h, y;
};
}

VmIndex *vm_index_build_host(const std::vector<std::string> &names, const std::vector<std::string> &seqs, int w, int k)
{
    VmIndex *ix = new VmIndex();
    ix->w = w;
    ix->k = k;
    ix->names = names;
    int64_t off = 0;
    for (const std::string &s : seqs) {
        ix->ctg_start.push_back(off);
        ix->ctg_len.push_back((int64_t)s.size());
        off += (int64_t)s.size();
    }
    ix->ref.resize((size_t)off);
    {
        size_t p = 0;
        for (const std::string &s : seqs)
            for (char c : s) ix->ref[p++] = "ACGTN"[vm_nt4((unsigned char)c)];   // as mappy's .seq() returns it
    }
    // ---- minimizers of every contig, in global coordinates ----
    std::vector<HY> all;
    all.reserve((size_t)(off / std::max(1, (w + 1) / 2)) + 16);
    for (size_t c = 0; c < seqs.size(); ++c) {
        const int64_t base = ix->ctg_start[c];
        vm_sketch((const unsigned char *)ix->ref.data() + base, ix->ctg_len[c], w, k, [&](uint64_t h, uint64_t y) {
            all.push_back(HY{h, ((y >> 1) + (uint64_t)base) << 1 | (y & 1)});
        });
    }
    std::sort(all.begin(), all.end(), [](const HY &a, const HY &b) { return a.h != b.h ? a.h < b.h : a.y < b.y; });
    int64_t nk = 0;
    for (size_t i = 0; i < all.size(); ++i)
        if (i == 0 || all[i].h != all[i - 1].h) ++nk;
    ix->n_keys = nk;
    uint64_t slots = 1024;
    while (slots < (uint64_t)nk * 2 + 16) slots <<= 1;
    ix->ht.assign((size_t)slots, VmHtSlot{VM_HT_EMPTY, 0, 0});
    ix->occ.resize(all.size());
    std::vector<uint32_t> counts;
    counts.reserve((size_t)nk);
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        while (j < all.size() && all[j].h == all[i].h) { ix->occ[j] = all[j].y; ++j; }
        uint64_t s = vm_ht_hash(all[i].h) & (slots - 1);
        while (ix->ht[s].key != VM_HT_EMPTY) s = (s + 1) & (slots - 1);
        ix->ht[s] = VmHtSlot{all[i].h, (uint32_t)i, (uint32_t)(j - i)};
        counts.push_back((uint32_t)(j - i));
        i = j;
    }
    // default occurrence cap: mm_idx_cal_max_occ(mi, 2e-4) then mm_mapopt_update's clamps [10, 1000000]
    if (nk > 0) {
        int64_t kth = (int64_t)((1. - 2e-4f) * (double)nk);
        if (kth >= nk) kth = nk - 1;
        std::nth_element(counts.begin(), counts.begin() + kth, counts.end());
        int64_t thres = (int64_t)counts[kth] + 1;
        thres = std::max<int64_t>(10, std::min<int64_t>(thres, 1000000));
        ix->mid_occ_default = (int)thres;
    }
    // ---- 9-mer positions of the whole reference, counting sort by 5-letter code ----
    ix->koff.assign(VM_K9_KEYS + 1, 0);
    const int64_t n = off;
    const int64_t allN = VM_K9_KEYS - 1;   // code of NNNNNNNNN: skipped (:23074-23079)
    const int64_t pow8 = 390625;           // 5^8
    auto scan = [&](auto &&visit) {
        int64_t code = 0;
        for (int64_t i = 0; i < n; ++i) {
            code = (code % pow8) * 5 + vm_code5((unsigned char)ix->ref[i]);
            if (i >= VM_K9 - 1 && code != allN) visit(code, i - (VM_K9 - 1));
        }
    };
    scan([&](int64_t code, int64_t) { ix->koff[code + 1]++; });
    for (int64_t c = 0; c < VM_K9_KEYS; ++c) ix->koff[c + 1] += ix->koff[c];
    ix->kpos.resize((size_t)ix->koff[VM_K9_KEYS]);
    {
        std::vector<int64_t> cur(ix->koff.begin(), ix->koff.end() - 1);
        scan([&](int64_t code, int64_t p) { ix->kpos[(size_t)cur[code]++] = (uint32_t)p; });
    }
    return ix;
}

static int up(void **dst, const void *src, size_t bytes, std::string &err)
{
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 16);
    if (e == cudaSuccess && bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("index upload: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

int vm_index_upload(VmIndex *ix, std::string &err)
{
    if (up(&ix->d_ht, ix->ht.data(), ix->ht.size() * sizeof(VmHtSlot), err)) return -1;
    if (up(&ix->d_occ, ix->occ.data(), ix->occ.size() * 8, err)) return -1;
    if (up(&ix->d_kpos, ix->kpos.data(), ix->kpos.size() * 4, err)) return -1;
    if (up(&ix->d_koff, ix->koff.data(), ix->koff.size() * 8, err)) return -1;
    if (up(&ix->d_ref, ix->ref.data(), ix->ref.size(), err)) return -1;
    ix->ht_slots = ix->ht.size();
    ix->n_occ = (int64_t)ix->occ.size();
    ix->n_kpos = (int64_t)ix->kpos.size();
    ix->dev.ht = (const VmHtSlot *)ix->d_ht;
    ix->dev.ht_mask = ix->ht.size() - 1;
    ix->dev.occ = (const uint64_t *)ix->d_occ;
    ix->dev.kpos = (const uint32_t *)ix->d_kpos;
    ix->dev.koff = (const int64_t *)ix->d_koff;
    ix->dev.ref = (const uint8_t *)ix->d_ref;
    ix->dev.ref_len = (int64_t)ix->ref.size();
    ix->dev.w = ix->w;
    ix->dev.k = ix->k;
    ix->dev.mid_occ = ix->mid_occ_default;
    return vm_index_build_buckets(ix, err);
}

void vm_index_free(VmIndex *ix)
{
    if (!ix) return;
    void *p[] = {ix->d_ht, ix->d_occ, ix->d_kpos, ix->d_koff, ix->d_ref, ix->d_ukeys, ix->d_ucnt};
    if (!ix->borrowed)
        for (void *q : p)
            if (q) cudaFree(q);
    if (ix->d_krow) cudaFree(ix->d_krow);
    if (ix->d_kbk) cudaFree(ix->d_kbk);
    delete ix;
}
