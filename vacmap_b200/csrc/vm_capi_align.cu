// vm_index_* / vm_align_* / vm_pairs_* C ABI over the CUDA backend (vm_backend_cuda.cuh).
#include "vm_backend_cuda.cuh"

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct vm_result {
    std::vector<int64_t> rec_off;        // per read
    std::vector<vm_record> recs;
    std::vector<uint32_t> cigar;
    std::vector<int32_t> status;         // per read (ReadStatus)
    std::vector<double> stage_ms;
    std::vector<std::string> stage_names;
    std::string stage_text;
};

extern "C" {

int vm_index_create(vm_ctx *c, int32_t n_contigs, const char *const *names, const char *const *seqs, const int64_t *lens,
                    int32_t w, int32_t k, vm_index_handle **out)
{
    if (!c) return VM_ERR_ARG;
    if (!out || n_contigs <= 0 || !names || !seqs || !lens || k < 1 || k > 28 || w < 1 || w > 255) {
        c->err = "bad argument";
        return VM_ERR_ARG;
    }
    cudaSetDevice(c->device);
    int64_t total = 0;
    for (int i = 0; i < n_contigs; ++i) total += lens[i];
    if (total >= (1LL << 32) - 64) { c->err = "reference longer than 2^32 bases is not supported (global coordinates are 32-bit)"; return VM_ERR_ARG; }
    vm_index_handle *h = nullptr;
    try {
        h = new vm_index_handle();
        std::string err;
        if (getenv("VM_INDEX_HOST")) {
            // A/B switch: the single-threaded host build of round 1 (the device build must give the same index)
            std::vector<std::string> nm, sq;
            for (int i = 0; i < n_contigs; ++i) {
                nm.emplace_back(names[i]);
                sq.emplace_back(seqs[i], (size_t)lens[i]);
            }
            h->ix = vm_index_build_host(nm, sq, w, k);
            if (vm_index_upload(h->ix, err)) throw std::runtime_error(err);
        } else {
            VmIndex *ix = new VmIndex();
            h->ix = ix;
            ix->w = w;
            ix->k = k;
            ix->ref.resize((size_t)total);
            int64_t off = 0;
            for (int i = 0; i < n_contigs; ++i) {
                ix->names.emplace_back(names[i]);
                ix->ctg_start.push_back(off);
                ix->ctg_len.push_back(lens[i]);
                if (lens[i]) memcpy(&ix->ref[(size_t)off], seqs[i], (size_t)lens[i]);
                off += lens[i];
            }
            if (vm_index_build_device(ix, err)) throw std::runtime_error(err);
        }
    } catch (const std::bad_alloc &) {
        c->err = "out of host memory while building the index";
        if (h) { vm_index_free(h->ix); delete h; }
        return VM_ERR_NOMEM;
    } catch (const std::exception &e) {
        c->err = e.what();
        if (h) { vm_index_free(h->ix); delete h; }
        return VM_ERR_CUDA;
    }
    h->ctg.names = h->ix->names;
    h->ctg.start = h->ix->ctg_start;
    h->ctg.len = h->ix->ctg_len;
    h->ctg.seq = h->ix->ref.data();
    h->ctg.total = (int64_t)h->ix->ref.size();
    *out = h;
    return VM_OK;
}

// The built index as five device arrays (reference, hash table, occurrences, 9-mer positions, 9-mer offsets) plus
// eight scalars: what another rank needs to use the index without building it (vm_index_adopt).
int vm_index_arrays(vm_index_handle *h, const void **ptrs, int64_t *bytes, int64_t *meta)
{
    if (!h || !ptrs || !bytes || !meta) return VM_ERR_ARG;
    VmIndex *ix = h->ix;
    ptrs[0] = ix->d_ref;  bytes[0] = (int64_t)ix->ref.size();
    ptrs[1] = ix->d_ht;   bytes[1] = (int64_t)(ix->ht_slots * sizeof(VmHtSlot));
    ptrs[2] = ix->d_occ;  bytes[2] = ix->n_occ * 8;
    ptrs[3] = ix->d_kpos; bytes[3] = ix->n_kpos * 4;
    ptrs[4] = ix->d_koff; bytes[4] = ((int64_t)VM_K9_KEYS + 1) * 8;
    meta[0] = ix->n_keys; meta[1] = ix->n_occ; meta[2] = ix->n_kpos; meta[3] = (int64_t)ix->ht_slots; meta[4] = ix->mid_occ_default;
    meta[5] = ix->w; meta[6] = ix->k; meta[7] = (int64_t)ix->ref.size();
    return VM_OK;
}

// An index over device arrays the CALLER owns (e.g. received by an NCCL broadcast from the rank that built them): the
// handle only borrows them; they must outlive it.  The normalised reference is copied back to the host for the glue.
int vm_index_adopt(vm_ctx *c, int32_t n_contigs, const char *const *names, const int64_t *lens, const void *const *ptrs,
                   const int64_t *bytes, const int64_t *meta, vm_index_handle **out)
{
    if (!c) return VM_ERR_ARG;
    if (!out || n_contigs <= 0 || !names || !lens || !ptrs || !bytes || !meta) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    vm_index_handle *h = nullptr;
    try {
        h = new vm_index_handle();
        VmIndex *ix = new VmIndex();
        h->ix = ix;
        ix->borrowed = true;
        ix->w = (int)meta[5];
        ix->k = (int)meta[6];
        int64_t off = 0;
        for (int i = 0; i < n_contigs; ++i) {
            ix->names.emplace_back(names[i]);
            ix->ctg_start.push_back(off);
            ix->ctg_len.push_back(lens[i]);
            off += lens[i];
        }
        if (off != meta[7] || bytes[0] != off) throw std::runtime_error("vm_index_adopt: contig lengths do not add up to the reference array");
        ix->d_ref = (void *)ptrs[0]; ix->d_ht = (void *)ptrs[1]; ix->d_occ = (void *)ptrs[2]; ix->d_kpos = (void *)ptrs[3]; ix->d_koff = (void *)ptrs[4];
        ix->n_keys = meta[0]; ix->n_occ = meta[1]; ix->n_kpos = meta[2]; ix->ht_slots = (uint64_t)meta[3]; ix->mid_occ_default = (int)meta[4];
        ix->ref.resize((size_t)off);
        if (off && cudaMemcpy(&ix->ref[0], ix->d_ref, (size_t)off, cudaMemcpyDeviceToHost) != cudaSuccess)
            throw std::runtime_error("vm_index_adopt: cannot read the reference array");
        ix->dev.ht = (const VmHtSlot *)ix->d_ht;
        ix->dev.ht_mask = ix->ht_slots - 1;
        ix->dev.occ = (const uint64_t *)ix->d_occ;
        ix->dev.kpos = (const uint32_t *)ix->d_kpos;
        ix->dev.koff = (const int64_t *)ix->d_koff;
        ix->dev.ref = (const uint8_t *)ix->d_ref;
        ix->dev.ref_len = off;
        ix->dev.w = ix->w;
        ix->dev.k = ix->k;
        ix->dev.mid_occ = ix->mid_occ_default;
        std::string berr;
        if (vm_index_build_buckets(ix, berr)) throw std::runtime_error(berr);
    } catch (const std::exception &e) {
        c->err = e.what();
        if (h) { vm_index_free(h->ix); delete h; }
        return VM_ERR_ARG;
    }
    h->ctg.names = h->ix->names;
    h->ctg.start = h->ix->ctg_start;
    h->ctg.len = h->ix->ctg_len;
    h->ctg.seq = h->ix->ref.data();
    h->ctg.total = (int64_t)h->ix->ref.size();
    *out = h;
    return VM_OK;
}

// numba's argsort (np.argsort inside njit code: numba/misc/quicksort.py) of int64 keys, ties and all -- the host mirrors of
// the reference's Python need it wherever the reference sorts inside njit functions (mammap_asm.py:22754-22755, clrnano:23103)
int vm_argsort_i64(const int64_t *keys, int64_t n, int64_t *order)
{
    if (n < 0 || (n > 0 && (!keys || !order))) return VM_ERR_ARG;
    try {
        std::vector<int64_t> R;
        vmg::argsort_replay<int64_t>(keys, n, R);
        if (n) memcpy(order, R.data(), (size_t)n * 8);
    } catch (const std::exception &) { return VM_ERR_NOMEM; }
    return VM_OK;
}

// The minimizer side of the index for the .mmi writer: the n_keys distinct hashes (ascending), their occurrence counts,
// and the n_minimizers occurrences (global last-base position << 1 | strand) key after key, ascending inside a key.
int vm_index_minimizers(vm_index_handle *h, uint64_t *keys, int32_t *counts, uint64_t *occ)
{
    if (!h || !keys || !counts || !occ) return VM_ERR_ARG;
    VmIndex *ix = h->ix;
    if (ix->d_ukeys && ix->d_ucnt) {
        if (cudaMemcpy(keys, ix->d_ukeys, (size_t)ix->n_keys * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(counts, ix->d_ucnt, (size_t)ix->n_keys * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(occ, ix->d_occ, (size_t)ix->n_occ * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return VM_ERR_CUDA;
        return VM_OK;
    }
    if (ix->ht.empty()) return VM_ERR_STATE;         // an adopted index keeps no key list
    std::vector<std::pair<uint64_t, std::pair<uint32_t, uint32_t>>> ks;
    for (const VmHtSlot &sl : ix->ht)
        if (sl.key != VM_HT_EMPTY) ks.push_back({sl.key, {sl.start, sl.count}});
    std::sort(ks.begin(), ks.end());
    for (size_t i = 0; i < ks.size(); ++i) { keys[i] = ks[i].first; counts[i] = (int32_t)ks[i].second.second; }
    memcpy(occ, ix->occ.data(), ix->occ.size() * 8);
    return VM_OK;
}

void vm_index_destroy(vm_index_handle *h)
{
    if (!h) return;
    vm_index_free(h->ix);
    delete h;
}

int vm_index_info(vm_index_handle *h, int32_t *k, int32_t *w, int32_t *n_contigs, int64_t *n_minimizers, int64_t *n_keys,
                  int32_t *mid_occ)
{
    if (!h) return VM_ERR_ARG;
    if (k) *k = h->ix->k;
    if (w) *w = h->ix->w;
    if (n_contigs) *n_contigs = (int32_t)h->ix->names.size();
    if (n_minimizers) *n_minimizers = h->ix->n_occ;
    if (n_keys) *n_keys = h->ix->n_keys;
    if (mid_occ) *mid_occ = h->ix->mid_occ_default;
    return VM_OK;
}

int vm_index_contig(vm_index_handle *h, int32_t i, const char **name, int64_t *start, int64_t *len, const char **seq)
{
    if (!h || i < 0 || i >= (int32_t)h->ix->names.size()) return VM_ERR_ARG;
    if (name) *name = h->ix->names[i].c_str();
    if (start) *start = h->ix->ctg_start[i];
    if (len) *len = h->ix->ctg_len[i];
    if (seq) *seq = h->ix->ref.data() + h->ix->ctg_start[i];
    return VM_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------
// Batches as jobs: a batch is cut into equal sub-batches ("chunks"), every chunk is one task for the context's
// pool of worker threads.  A worker owns a CUDA stream, device arenas and a backend, so the host glue of one chunk
// overlaps the kernels of the others, and the chunks of a batch submitted while another one is still draining
// start at once (vm_align_submit / vm_align_wait): the pipeline does not run empty between batches.
// ---------------------------------------------------------------------------
struct vm_job {
    vm_ctx *c = nullptr;
    vm_index_handle *h = nullptr;
    vmg::Options opt;
    int threads = 1, resident = 0, workers = 1;
    int64_t n_reads = 0;
    const char *seqs = nullptr;
    const int64_t *seq_off = nullptr;
    std::vector<int64_t> bounds;
    BatchResult br;
    // chunks whose records came back as flat arrays (device-resident extension stage): copies owned by the job
    struct FlatChunk { bool used = false; std::vector<int64_t> rec_off; std::vector<vm_record> recs; std::vector<uint32_t> cigar; };
    std::vector<FlatChunk> flat;
    std::mutex mu;
    std::condition_variable cv;
    int64_t chunks_left = 0;
    std::string err;
    StageTimer timer;
    double fill_cells = 0, fill_bases = 0, fill_jobs = 0, ed_cells = 0, ed_upper = 0, reseed_hits = 0, chain_anchors = 0, band_jobs = 0,
           band_redo = 0, dir_bytes = 0, chain_opcount = 0;
    int64_t launches = 0;
    std::chrono::steady_clock::time_point t0, t_done;
};

namespace {

void job_absorb(vm_job *job, CudaBackend &wb, int64_t launches, const std::string &err)
{
    std::lock_guard<std::mutex> lk(job->mu);
    for (auto &kv : wb.timer.ms) job->timer.add(kv.first.c_str(), kv.second);
    job->fill_cells += wb.fill_cells_; job->fill_bases += wb.fill_bases_; job->fill_jobs += wb.fill_jobs_;
    job->ed_cells += wb.ed_cells_; job->ed_upper += wb.ed_upper_jobs_; job->reseed_hits += wb.reseed_hits_;
    job->chain_anchors += wb.chain_anchors_;
    job->chain_opcount += wb.chain_opcount_;
    job->band_jobs += wb.fill_band_jobs_;
    job->band_redo += wb.fill_band_redo_;
    job->dir_bytes += wb.fill_dir_bytes_;
    job->launches += launches;
    if (!err.empty() && job->err.empty()) job->err = err;
    if (--job->chunks_left == 0) {
        job->t_done = std::chrono::steady_clock::now();
        job->cv.notify_all();
    }
}

// one chunk of a job on backend `wb` (context `wc`); parent = backend holding the resident reads
void run_chunk(vm_job *job, int64_t ci, vm_ctx *wc, CudaBackend &wb, CudaBackend *parent)
{
    std::string err;
    const int64_t l0 = wc->launches;
    const VmSyncStats sync0 = vm_sync_stats();
    try {
        cudaSetDevice(job->c->device);
        wb.set_index(job->h);
        wb.reset_counters();
        wb.host_threads = job->threads;
        Driver drv(wb, job->h->ctg, job->opt, job->h->ix->k, job->threads);
        drv.on_time = [&wb](const char *nm, double ms) {
            wb.timer.add(nm, ms);
            const auto t1 = std::chrono::steady_clock::now();
            Timeline::get().add(&wb, nm, t1 - std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                                                  std::chrono::duration<double, std::milli>(ms)), t1);
        };
        drv.on_cpu = [&wb](const char *nm, double ms) { wb.timer.add((std::string("cpu_") + nm).c_str(), ms); };
        const int64_t r0 = job->bounds[(size_t)ci], nr = job->bounds[(size_t)ci + 1] - r0;
        ReadBatch sb;
        sb.n = nr;
        sb.seq = job->seqs;
        sb.off = job->seq_off + r0;
        if (parent != &wb) {
            if (job->resident) wb.copy_reads_from(*parent, r0, nr);
            else wb.reads_resident = false;
        }
        BatchResult sr;
        try {
            drv.align_batch(sb, sr);
            if (sr.flat) {
                // the backend's page-locked result buffers are reused by its next chunk: copy out now
                vm_job::FlatChunk &fc = job->flat[(size_t)ci];
                fc.used = true;
                fc.rec_off.swap(sr.fr.rec_off);
                static_assert(sizeof(vm_record) == sizeof(vmd::Rec), "record layout");
                fc.recs.resize((size_t)sr.fr.n_rec);
                fc.cigar.resize((size_t)sr.fr.n_ops);
                if (sr.fr.n_rec) memcpy(fc.recs.data(), sr.fr.recs, (size_t)sr.fr.n_rec * sizeof(vm_record));
                if (sr.fr.n_ops) {
                    // a few large memcpys: shared between the pool's threads
                    const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(8, sr.fr.n_ops >> 20));
                    parallel_for(parts, job->threads, [&](int64_t k) {
                        const int64_t lo = sr.fr.n_ops * k / parts, hi = sr.fr.n_ops * (k + 1) / parts;
                        memcpy(fc.cigar.data() + lo, sr.fr.cigar + lo, (size_t)(hi - lo) * 4);
                    }, 1);
                }
                for (int64_t i = 0; i < nr; ++i) job->br.status[(size_t)(r0 + i)] = sr.status[(size_t)i];
            } else
            for (int64_t i = 0; i < nr; ++i) {
                job->br.records[(size_t)(r0 + i)].swap(sr.records[(size_t)i]);
                job->br.status[(size_t)(r0 + i)] = sr.status[(size_t)i];
            }
            for (int k = 0; k < BC_COUNT; ++k) wb.timer.add(kBranchName[k], (double)sr.branch[k]);
        } catch (const std::exception &e) {
            // Something in this chunk could not be processed (e.g. a read beyond a kernel's size limits).  The reference
            // loses only the offending read (`except Exception: continue`, clrnano:24116-24125): redo the chunk one
            // read at a time and leave the reads that fail again without records.
            if (nr <= 1 || cudaGetLastError() != cudaSuccess) throw;
            int64_t dropped = 0;
            for (int64_t i = 0; i < nr; ++i) {
                ReadBatch one;
                one.n = 1;
                one.seq = job->seqs;
                one.off = job->seq_off + r0 + i;
                wb.reads_resident = false;
                try {
                    BatchResult s1;
                    drv.align_batch(one, s1);
                    if (s1.flat) {
                        // one read: back to per-read records
                        std::vector<vmg::Record> &dst = job->br.records[(size_t)(r0 + i)];
                        dst.clear();
                        for (int64_t q = 0; q < s1.fr.n_rec; ++q) {
                            const vmd::Rec &x = s1.fr.recs[q];
                            vmg::Record rec;
                            rec.contig = x.contig; rec.strand = x.strand; rec.q_st = x.q_st; rec.q_en = x.q_en; rec.r_st = x.r_st;
                            rec.r_en = x.r_en; rec.mapq = x.mapq;
                            rec.cigar.assign(s1.fr.cigar + x.cigar_off, s1.fr.cigar + x.cigar_off + x.cigar_len);
                            dst.push_back(std::move(rec));
                        }
                    } else
                    job->br.records[(size_t)(r0 + i)].swap(s1.records[0]);
                    job->br.status[(size_t)(r0 + i)] = s1.status[0];
                    for (int k = 0; k < BC_COUNT; ++k) wb.timer.add(kBranchName[k], (double)s1.branch[k]);
                } catch (const std::exception &) {
                    if (cudaGetLastError() != cudaSuccess) throw;
                    job->br.status[(size_t)(r0 + i)] = RS_FAILED;
                    ++dropped;
                }
            }
            wb.timer.add("n_reads_dropped_on_error", (double)dropped);
        }
    } catch (const std::exception &e) {
        err = e.what();
        if (err.empty()) err = "error";
    }
    wb.timer.add("w_sync_wait", 1e-6 * (double)(vm_sync_stats().ns - sync0.ns));      // this worker thread inside stream waits
    wb.timer.add("n_syncs", (double)(vm_sync_stats().n - sync0.n));
    job_absorb(job, wb, wc->launches - l0, err);
}

struct AlignPool {
    vm_ctx *c;
    std::vector<std::thread> threads;
    std::deque<std::pair<vm_job *, int64_t>> queue;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;

    explicit AlignPool(vm_ctx *c_) : c(c_) {}
    ~AlignPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (std::thread &t : threads) t.join();
    }
    // called from the submitting thread: worker contexts and backends are created here, not in the workers
    bool ensure(int workers, vm_index_handle *h)
    {
        while ((int)threads.size() < workers) {
            const int w = (int)threads.size();
            vm_ctx *wc = vm_ctx_worker(c, w);
            if (!wc) return false;
            if (!wc->backend) {
                wc->backend = new CudaBackend(wc, h);
                wc->backend_free = [](void *q) { delete (CudaBackend *)q; };
            }
            threads.emplace_back([this, w, wc] { loop(w, wc); });
        }
        for (int w = 0; w < (int)threads.size(); ++w) vm_ctx_worker(c, w);   // refresh the table aliases
        return true;
    }
    void push(vm_job *job)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            for (int64_t ci = 0; ci + 1 < (int64_t)job->bounds.size(); ++ci) queue.emplace_back(job, ci);
        }
        cv.notify_all();
    }
    void loop(int w, vm_ctx *wc)
    {
        (void)w;
        for (;;) {
            std::pair<vm_job *, int64_t> task;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !queue.empty(); });
                if (stop) return;
                task = queue.front();
                queue.pop_front();
            }
            run_chunk(task.first, task.second, wc, *(CudaBackend *)wc->backend, (CudaBackend *)c->backend);
        }
    }
};

} // namespace

extern "C" {

int vm_align_submit(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                    const int64_t *seq_off, int32_t resident, vm_job **out)
{
    if (!c) return VM_ERR_ARG;
    if (!h || !p || !out || n_reads < 0 || !seq_off || (n_reads > 0 && !seqs)) { c->err = "bad argument"; return VM_ERR_ARG; }
    if (c->n_extra == 0) { c->err = "vm_set_tables must be called first"; return VM_ERR_STATE; }
    cudaSetDevice(c->device);
    vm_job *job = new vm_job();
    job->c = c;
    job->h = h;
    vmg::Options &opt = job->opt;
    opt.global_skipcost = p->global_skipcost;
    opt.local_skipcost = p->local_skipcost;
    opt.maxdivergence = p->maxdivergence;
    opt.global_maxdiff = p->global_maxdiff;
    opt.local_maxdiff = p->local_maxdiff;
    opt.check_num = p->check_num;
    opt.eqx = p->eqx != 0;
    opt.hardclip = p->hardclip != 0;
    opt.nodiscard = p->nodiscard != 0;
    opt.mode = vmg::ModeConst{p->accept_score, p->max_guides, p->local_maxgap, p->clamp40 != 0};
    job->threads = p->host_threads > 0 ? p->host_threads : HostPool::get().size();
    job->resident = resident != 0;
    job->n_reads = n_reads;
    job->seqs = seqs;
    job->seq_off = seq_off;
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *q) { delete (CudaBackend *)q; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.set_index(h);
        be.reads_resident = resident != 0;
        // `workers` = size of the pool; a job is cut into equal chunks of ~chunk_reads (default: half of the job, at
        // least 512 reads -- every chunk pays the same fixed chain of launch and synchronisation latencies, and with the
        // extension stage's glue on the device the host no longer needs many small chunks to keep its cores busy:
        // measured 93 ms per 10k-read step with 3 workers x 5000 reads, 104 ms with 8 x 1667).  With more workers than
        // chunks per job, the chunks of the next jobs run beside them: six workers over four jobs in flight measured
        // 68.7-69.6 ms per step against 72-73 ms with four workers over three jobs (tests/gpu/ab_share.sh).
        int workers = p->workers > 0 ? p->workers : 6;
        const int64_t chunk = p->chunk_reads > 0 ? p->chunk_reads : std::max<int64_t>(512, (n_reads + 1) / 2);
        if (n_reads <= chunk || workers == 1) workers = 1;
        job->bounds.assign(1, 0);
        if (workers > 1) {
            const int64_t parts = (n_reads + chunk - 1) / chunk;
            for (int64_t i = 1; i <= parts; ++i) {
                const int64_t e = n_reads * i / parts;
                if (e > job->bounds.back()) job->bounds.push_back(e);
            }
        } else job->bounds.push_back(n_reads);
        job->workers = workers;
        job->chunks_left = (int64_t)job->bounds.size() - 1;
        job->br.records.assign((size_t)n_reads, {});
        job->br.status.assign((size_t)n_reads, RS_OK);
        job->flat.assign(job->bounds.size() - 1, vm_job::FlatChunk());
        job->t0 = std::chrono::steady_clock::now();
        if (workers <= 1) {
            // lock-step: the whole batch on the context's own backend, in the calling thread
            run_chunk(job, 0, c, be, &be);
        } else {
            if (!c->pool) {
                c->pool = new AlignPool(c);
                c->pool_free = [](void *q) { delete (AlignPool *)q; };
            }
            AlignPool &pool = *(AlignPool *)c->pool;
            if (!pool.ensure(workers, h)) throw std::runtime_error("cannot create a worker context");
            pool.push(job);
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        delete job;
        return VM_ERR_CUDA;
    }
    *out = job;
    return VM_OK;
}

int vm_align_wait(vm_job *job, vm_result **out)
{
    if (!job || !out) return VM_ERR_ARG;
    vm_ctx *c = job->c;
    {
        std::unique_lock<std::mutex> lk(job->mu);
        job->cv.wait(lk, [&] { return job->chunks_left == 0; });
    }
    if (!job->err.empty()) {
        c->err = job->err;
        delete job;
        return VM_ERR_CUDA;
    }
    cudaSetDevice(c->device);
    const int64_t n_reads = job->n_reads;
    BatchResult &br = job->br;
    vm_result *res = new vm_result();
    const double total = std::chrono::duration<double, std::milli>(job->t_done - job->t0).count();
    auto t1 = std::chrono::steady_clock::now();
    // result arena: offsets by prefix sums, records and CIGAR ops copied in parallel.  A read's records are either in
    // its chunk's flat arrays (device-resident extension stage) or, per read, in br.records (host glue, single-read redo)
    const int64_t n_chunks = (int64_t)job->bounds.size() - 1;
    std::vector<int32_t> chunk_of((size_t)n_reads, 0);
    for (int64_t ci = 0; ci < n_chunks; ++ci)
        for (int64_t r = job->bounds[(size_t)ci]; r < job->bounds[(size_t)ci + 1]; ++r) chunk_of[(size_t)r] = (int32_t)ci;
    auto flat_of = [&](int64_t r, int64_t &lo, int64_t &hi) -> const vm_job::FlatChunk * {
        const vm_job::FlatChunk &fc = job->flat[(size_t)chunk_of[(size_t)r]];
        if (!fc.used || !br.records[(size_t)r].empty()) return nullptr;
        const int64_t i = r - job->bounds[(size_t)chunk_of[(size_t)r]];
        lo = fc.rec_off[(size_t)i]; hi = fc.rec_off[(size_t)i + 1];
        return &fc;
    };
    res->rec_off.assign((size_t)n_reads + 1, 0);
    std::vector<int64_t> cig_off((size_t)n_reads + 1, 0);
    for (int64_t r = 0; r < n_reads; ++r) {
        int64_t ops = 0, nrec = 0, lo = 0, hi = 0;
        if (const vm_job::FlatChunk *fc = flat_of(r, lo, hi)) {
            nrec = hi - lo;
            if (nrec > 0) ops = fc->recs[(size_t)hi - 1].cigar_off + fc->recs[(size_t)hi - 1].cigar_len - fc->recs[(size_t)lo].cigar_off;
        } else {
            for (const vmg::Record &rec : br.records[r]) ops += (int64_t)rec.cigar.size();
            nrec = (int64_t)br.records[r].size();
        }
        res->rec_off[r + 1] = res->rec_off[r] + nrec;
        cig_off[r + 1] = cig_off[r] + ops;
    }
    res->recs.resize((size_t)res->rec_off[n_reads]);
    res->cigar.resize((size_t)cig_off[n_reads]);
    res->status.swap(br.status);
    parallel_for(n_reads, job->threads, [&](int64_t r) {
        int64_t ri = res->rec_off[r], co = cig_off[r], lo = 0, hi = 0;
        if (const vm_job::FlatChunk *fc = flat_of(r, lo, hi)) {
            if (hi <= lo) return;
            // a read's records and CIGARs are contiguous in its chunk's arrays
            const int64_t c0 = fc->recs[(size_t)lo].cigar_off;
            for (int64_t q = lo; q < hi; ++q) {
                vm_record o = fc->recs[(size_t)q];
                o.cigar_off = co + (o.cigar_off - c0);
                res->recs[(size_t)ri++] = o;
            }
            memcpy(res->cigar.data() + co, fc->cigar.data() + c0, (size_t)(cig_off[r + 1] - co) * 4);
            return;
        }
        for (const vmg::Record &rec : br.records[r]) {
            vm_record &o = res->recs[(size_t)ri++];
            o.contig = rec.contig;
            o.strand = rec.strand;
            o.q_st = rec.q_st; o.q_en = rec.q_en; o.r_st = rec.r_st; o.r_en = rec.r_en;
            o.mapq = rec.mapq;
            o.cigar_off = co;
            o.cigar_len = (int32_t)rec.cigar.size();
            std::copy(rec.cigar.begin(), rec.cigar.end(), res->cigar.begin() + co);
            co += (int64_t)rec.cigar.size();
        }
    }, 64);
    Timeline::get().flush();
    StageTimer &tm = job->timer;
    tm.add("n_workers", job->workers);
    tm.add("total", total);
    tm.add("g_result_arena", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    tm.add("n_fill_cells", job->fill_cells);
    tm.add("n_fill_bases", job->fill_bases);
    tm.add("n_fill_jobs", job->fill_jobs);
    tm.add("n_fill_band_jobs", job->band_jobs);
    tm.add("n_fill_band_redo", job->band_redo);
    tm.add("n_fill_dir_bytes", job->dir_bytes);
    tm.add("n_ed_cells", job->ed_cells);
    tm.add("n_ed_upper_jobs", job->ed_upper);
    tm.add("n_reseed_hits", job->reseed_hits);
    tm.add("n_chain_anchors", job->chain_anchors);
    tm.add("n_chain_opcount", job->chain_opcount);
    for (auto &kv : tm.ms) {
        res->stage_names.push_back(kv.first);
        res->stage_ms.push_back(kv.second);
        res->stage_text += kv.first + "=" + std::to_string(kv.second) + ";";
    }
    if (job->workers > 1) c->launches += job->launches;   // lock-step jobs counted on the context directly
    delete job;
    *out = res;
    return VM_OK;
}

static int vm_align_impl(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                         const int64_t *seq_off, int resident, vm_result **out)
{
    if (!out) { if (c) c->err = "bad argument"; return VM_ERR_ARG; }
    vm_job *job = nullptr;
    const int rc = vm_align_submit(c, h, p, n_reads, seqs, seq_off, resident, &job);
    if (rc != VM_OK) return rc;
    return vm_align_wait(job, out);
}

int vm_align_batch(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                   const int64_t *seq_off, vm_result **out)
{
    return vm_align_impl(c, h, p, n_reads, seqs, seq_off, 0, out);
}

// Stage-level entry point for the base-level kernels on raw sequence pairs (parity tests).
// kind 0: global edit distance -> out0[j]; kind 1: z-drop edge extension -> out0 = q_e, out1 = t_e;
// kind 3: as kind 0 inside the Ukkonen band |row - column| <= out1[j] (an INPUT here): out0[j] is the exact
//         distance when it is <= out1[j], otherwise some value > out1[j] (what the divergence filter needs);
// kind 2: global fill -> CIGAR ops of pair j at cigar[cig_off[j] .. cig_off[j] + out0[j]),
//         cig_off[j] = sum over i < j of (tlen_i + qlen_i + 2).
int vm_pairs_batch(vm_ctx *c, int32_t kind, int32_t eqx, int64_t n_pairs, const char *targets, const int64_t *t_off,
                   const char *queries, const int64_t *q_off, int64_t *out0, int64_t *out1, uint32_t *cigar)
{
    if (!c) return VM_ERR_ARG;
    if (n_pairs < 0 || !t_off || !q_off || !out0 || kind < 0 || kind > 3 || (kind == 3 && !out1)) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, nullptr);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        // one pseudo-read holding every target followed by every query
        const int64_t tt = t_off[n_pairs], tq = q_off[n_pairs];
        std::string cat((size_t)(tt + tq), 'N');
        if (tt) memcpy(&cat[0], targets, (size_t)tt);
        if (tq) memcpy(&cat[(size_t)tt], queries, (size_t)tq);
        for (char &ch : cat) ch = "ACGTN"[vm_nt4((unsigned char)ch)];
        const int64_t off[2] = {0, tt + tq};
        ReadBatch b;
        b.n = 1; b.seq = cat.data(); b.off = off;
        be.reads_resident = false;
        be.upload_reads(b);
        be.reset_counters();
        auto ref_of = [&](int64_t lo, int64_t hi) { vmg::SeqRef s; s.src = 1; s.lo = lo; s.hi = hi; return s; };
        if (kind == 0 || kind == 3) {
            std::vector<EdJob> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].a = ref_of(tt + q_off[j], tt + q_off[j + 1]);
                jobs[j].b = ref_of(t_off[j], t_off[j + 1]);
                jobs[j].band = kind == 3 ? out1[j] : -1;
            }
            be.edit_distance(b, jobs, nullptr, 0);
            for (int64_t j = 0; j < n_pairs; ++j) out0[j] = jobs[j].dist;
        } else if (kind == 1) {
            std::vector<ExtJobRef> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].job.target = ref_of(t_off[j], t_off[j + 1]);
                jobs[j].job.query = ref_of(tt + q_off[j], tt + q_off[j + 1]);
            }
            be.extend(b, jobs);
            for (int64_t j = 0; j < n_pairs; ++j) { out0[j] = jobs[j].job.q_e; if (out1) out1[j] = jobs[j].job.t_e; }
        } else {
            std::vector<FillJobRef> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].job.target = ref_of(t_off[j], t_off[j + 1]);
                jobs[j].job.query = ref_of(tt + q_off[j], tt + q_off[j + 1]);
            }
            const uint32_t *ops = be.fill(b, eqx != 0, jobs);
            int64_t co = 0;
            for (int64_t j = 0; j < n_pairs; ++j) {
                out0[j] = (int64_t)jobs[j].cig_len;
                if (cigar && jobs[j].cig_len > 0) memcpy(cigar + co, ops + jobs[j].cig_off, (size_t)jobs[j].cig_len * 4);
                co += (t_off[j + 1] - t_off[j]) + (q_off[j + 1] - q_off[j]) + 2;
            }
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

// Stage-level local re-seeding: the scan + same-diagonal merge of get_localmap_..._guide_1 (clrnano:23138-23344)
// for jobs whose windows and guide chain the caller built (window construction :23095-23136 is host glue).
// Job j: read job_read[j] (given already oriented: its testseq), read positions [readstart[j], readend[j]),
// reference windows win_lo/win_hi[win_off[j] .. win_off[j+1]) (GLOBAL [lo, hi), insertion order), guide points
// gx/gy[g_off[j] .. g_off[j+1]) sorted by read position.  Output: the anchors in the reference's emission order
// as int64 rows in rows[row_off[j] .. row_off[j+1]); VM_ERR_NOMEM (row_off filled) when `cap` rows are too few.
int vm_local_reseed_batch(vm_ctx *c, vm_index_handle *h, int64_t n_reads, const char *seqs, const int64_t *seq_off, int64_t n_jobs,
                          const int32_t *job_read, const int32_t *readstart, const int32_t *readend, const int64_t *win_off,
                          const int64_t *win_lo, const int64_t *win_hi, const int64_t *g_off, const int32_t *gx, const int64_t *gy,
                          int64_t *rows, int64_t cap, int64_t *row_off)
{
    if (!c) return VM_ERR_ARG;
    if (!h || n_reads < 0 || n_jobs < 0 || !seq_off || !row_off || (n_jobs > 0 && (!job_read || !readstart || !readend || !win_off || !g_off))) {
        c->err = "bad argument";
        return VM_ERR_ARG;
    }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.set_index(h);
        be.reads_resident = false;
        ReadBatch b;
        b.n = n_reads; b.seq = seqs; b.off = seq_off;
        std::vector<GuideJobRef> jobs((size_t)n_jobs);
        for (int64_t j = 0; j < n_jobs; ++j) {
            if (job_read[j] < 0 || job_read[j] >= n_reads) { c->err = "bad job_read"; return VM_ERR_ARG; }
            jobs[(size_t)j].read = job_read[j];
            vmg::GuideJob &g = jobs[(size_t)j].job;
            g.readstart = readstart[j];
            g.readend = readend[j];
            g.win_lo.assign(win_lo + win_off[j], win_lo + win_off[j + 1]);
            g.win_hi.assign(win_hi + win_off[j], win_hi + win_off[j + 1]);
            g.gx.assign(gx + g_off[j], gx + g_off[j + 1]);
            g.gy.assign(gy + g_off[j], gy + g_off[j + 1]);
        }
        std::vector<char> need_reverse((size_t)n_reads, 0);
        std::vector<VmAnchor> flat;
        std::vector<int64_t> job_off;
        be.reseed_only(b, need_reverse, jobs, flat, job_off);
        for (int64_t j = 0; j <= n_jobs; ++j) row_off[j] = job_off[(size_t)j];
        if (row_off[n_jobs] > cap) { c->err = "row buffer too small"; return VM_ERR_NOMEM; }
        for (size_t t = 0; t < flat.size(); ++t) {
            int64_t *o = rows + 4 * t;
            o[0] = flat[t].x; o[1] = (int64_t)flat[t].y; o[2] = flat[t].s; o[3] = flat[t].l;
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

// Stage-level local chaining: the reference's get_optimal_chain_..._fine_list (variant 1, clrnano:27305-27528),
// _fine_list_mismatch (variant 2, :28250-28476) or their _fast twins (force_fast) on anchors given as int64 rows
// (readpos, refpos, strand, len); presorted != 0: rows are already ordered by read end as the functions expect,
// else the library sorts them as the caller does (np.argsort(x + len), :28585).  Per read r: score[r] (g_max_scores)
// and the chain -- trimmed like :27508-27527, ASCENDING read order (the reference returns it descending) -- in
// path[path_off[r] .. path_off[r+1]) as int64 rows; used_fast[r] = 1 when a _fast variant produced it.
int vm_chain_local_batch(vm_ctx *c, const vm_chain_params *prm, int32_t presorted, int32_t force_fast, int64_t n_reads,
                         const int64_t *anchors, const int64_t *off, const int32_t *read_len, double *score, int64_t *path,
                         int64_t *path_off, int32_t *used_fast)
{
    if (!c) return VM_ERR_ARG;
    if (!prm || n_reads < 0 || !off || !score || !path_off || (n_reads > 0 && !read_len) || (prm->variant != 1 && prm->variant != 2)) {
        c->err = "bad argument";
        return VM_ERR_ARG;
    }
    if (c->n_extra == 0) { c->err = "vm_set_tables must be called first"; return VM_ERR_STATE; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, nullptr);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        ChainOut out;
        std::vector<double> sc;
        std::vector<int32_t> uf;
        be.chain_local_stage(*prm, presorted != 0, force_fast != 0, n_reads, anchors, off, read_len, out, sc, uf);
        path_off[0] = 0;
        for (int64_t r = 0; r < n_reads; ++r) {
            const ExtractRec &x = out.rec[r];
            score[r] = x.n_anc > 0 ? sc[(size_t)r] : 0.0;
            if (used_fast) used_fast[r] = uf[(size_t)r];
            for (int32_t t = 0; t < x.n_anc && path; ++t) {
                const Anc32 &a = out.x_anc[x.anc_off + t];
                int64_t *o = path + 4 * (path_off[r] + t);
                o[0] = a.x; o[1] = (int64_t)a.y; o[2] = a.s; o[3] = a.l;
            }
            path_off[r + 1] = path_off[r] + x.n_anc;
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

// Stage-level seeding: what `index_object.map(seq, check_num, mid_occ=-1)` followed by
// get_reversed_chain_numpy_rough returns per read.  rows: int64[cap][4]; row_off[n_reads+1] receives
// the ragged offsets; returns VM_ERR_NOMEM (with row_off filled) when cap is too small.
int vm_seed_batch_rows(vm_ctx *c, vm_index_handle *h, int32_t check_num, int64_t n_reads, const char *seqs, const int64_t *seq_off,
                       int64_t *rows, int64_t cap, int64_t *row_off, int32_t *need_reverse)
{
    if (!c) return VM_ERR_ARG;
    if (!h || n_reads < 0 || !seq_off || !row_off || (n_reads > 0 && !seqs)) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.set_index(h);
        be.reads_resident = false;
        ReadBatch b;
        b.n = n_reads; b.seq = seqs; b.off = seq_off;
        std::vector<VmAnchor> flat;
        std::vector<int64_t> a_off;
        std::vector<int32_t> n_out, nrev;
        be.seed_only(b, check_num, flat, a_off, n_out, nrev);
        row_off[0] = 0;
        for (int64_t r = 0; r < n_reads; ++r) row_off[r + 1] = row_off[r] + n_out[r];
        if (row_off[n_reads] > cap) { c->err = "row buffer too small"; return VM_ERR_NOMEM; }
        for (int64_t r = 0; r < n_reads; ++r) {
            if (need_reverse) need_reverse[r] = nrev[r];
            for (int32_t t = 0; t < n_out[r]; ++t) {
                const VmAnchor &a = flat[(size_t)a_off[r] + t];
                int64_t *o = rows + (row_off[r] + t) * 4;
                o[0] = a.x; o[1] = (int64_t)a.y; o[2] = a.s; o[3] = a.l;
            }
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

int vm_reads_upload(vm_ctx *c, vm_index_handle *h, int64_t n_reads, const char *seqs, const int64_t *seq_off)
{
    if (!c) return VM_ERR_ARG;
    if (!h || n_reads < 0 || !seq_off || (n_reads > 0 && !seqs)) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.reads_resident = false;
        ReadBatch b;
        b.n = n_reads; b.seq = seqs; b.off = seq_off;
        be.upload_reads(b);
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

int vm_align_resident(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                      const int64_t *seq_off, vm_result **out)
{
    return vm_align_impl(c, h, p, n_reads, seqs, seq_off, 1, out);
}

int64_t vm_result_num_records(vm_result *r) { return r ? (int64_t)r->recs.size() : 0; }
int64_t vm_result_num_cigar_ops(vm_result *r) { return r ? (int64_t)r->cigar.size() : 0; }
const int64_t *vm_result_read_offsets(vm_result *r) { return r ? r->rec_off.data() : nullptr; }
const vm_record *vm_result_records(vm_result *r) { return r ? r->recs.data() : nullptr; }
const uint32_t *vm_result_cigar(vm_result *r) { return r ? r->cigar.data() : nullptr; }
const char *vm_result_stage_times(vm_result *r) { return r ? r->stage_text.c_str() : ""; }
const int32_t *vm_result_read_status(vm_result *r) { return r ? r->status.data() : nullptr; }
void vm_result_free(vm_result *r) { delete r; }

} // extern "C"
