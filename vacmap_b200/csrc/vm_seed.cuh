// Seeding stage buffers and entry point (see vm_seed.cu).
#pragma once
#include "vm_ctx.cuh"
#include "vm_index.cuh"

struct VmSeedBufs {
    VmDevBuf mz_hash, mz_posz, mz_start, mz_cnt, mz_aoff, n_mz, n_anchor, n_out, need_rev, a_off, t_off, raw, out, table,
        compact, chunk_off, chunk_cnt;
    VmPinnedBuf pin;      // bounce buffer of the per-read counts read back between launches
    void release()
    {
        VmDevBuf *b[] = {&mz_hash, &mz_posz, &mz_start, &mz_cnt, &mz_aoff, &n_mz, &n_anchor, &n_out, &need_rev, &a_off,
                         &t_off, &raw, &out, &table, &compact, &chunk_off, &chunk_cnt};
        for (VmDevBuf *x : b) x->release();
        pin.release();
    }
};

// After the call: B.out holds the filtered / flipped anchors of read r at [a_off_host[r], a_off_host[r] + n_out[r]).
int vm_seed_batch(VmSeedBufs &B, const VmIndexDev &ix, const uint8_t *reads_dev, const int64_t *off_dev,
                  const std::vector<int64_t> &off_host, int check_num, int mid_occ, cudaStream_t stream,
                  std::vector<int32_t> &n_out, std::vector<int32_t> &need_rev, std::vector<int64_t> &a_off_host,
                  int64_t *launches, std::string &err);
