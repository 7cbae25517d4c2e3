#include "unc_selfalign.cuh"
#include "unc_selfalign_host.hpp"
// unc_selfalign_host.inl -- C-ABI of `self_align` (included by unc_abi.cu).  Replaces the reference's
// self_align(bwa_prefix, sample_dist) (src/self_align_ref.cpp:34-91; bound to Python at
// src/pybinder.cpp:59 and called by uncalled/index.py:82).  Like the reference's, it loads the FM index
// and the packed reference itself -- at index time there is no .uncl yet, so unc_index_load cannot be used.

__global__ void __launch_bounds__(128) k_selfalign_count(DevIndex ix, DevSelfAlign A) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) unc_selfalign_count(ix, A, i);
}
__global__ void __launch_bounds__(128) k_selfalign_write(DevIndex ix, DevSelfAlign A) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) unc_selfalign_write(ix, A, i);
}

namespace {
struct SelfAlignBuffers {            // frees whatever was allocated, on every exit path
    void *bwt = nullptr, *pac = nullptr, *pos = nullptr, *lim = nullptr, *count = nullptr, *stage = nullptr,
         *offsets = nullptr, *values = nullptr;
    uint64_t *h_offsets = nullptr, *h_values = nullptr;
    ~SelfAlignBuffers() {
        cudaFree(bwt); cudaFree(pac); cudaFree(pos); cudaFree(lim); cudaFree(count); cudaFree(stage);
        cudaFree(offsets); cudaFree(values);
        free(h_offsets); free(h_values);
    }
};
}  // namespace

extern "C" {

void unc_free(void *p) { free(p); }

int unc_self_align(const char *bwa_prefix, uint32_t sample_dist, uint64_t *n_paths, uint64_t **offsets, uint64_t **values) {
    if (!bwa_prefix || !sample_dist || !n_paths || !offsets || !values) return fail(UNC_E_ARG, "null argument or sample_dist == 0");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(UNC_E_NO_DEVICE, "no CUDA device (the product has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(g_device));
    HostIndex h;
    if (!hix_load_fm(h, bwa_prefix)) return fail(UNC_E_IO, h.error);
    if (h.seq_len >= 0xFFFFFF00ull) return fail(UNC_E_TOO_LARGE, "FM index longer than 2^32 rows is not supported by the u32 device image");
    std::vector<char> pac;
    if (!hix_read_file(std::string(bwa_prefix) + ".pac", pac)) return fail(UNC_E_IO, std::string("cannot read ") + bwa_prefix + ".pac");
    uint64_t total = 0;
    for (uint32_t l : h.lens) total += l;
    if (pac.size() * 4 < total) return fail(UNC_E_IO, "truncated .pac");

    std::vector<u32> pos, lim;
    unc_selfalign_sample(h.lens, sample_dist, pos, lim);
    const size_t n = pos.size();
    SelfAlignBuffers b;
    b.h_offsets = (uint64_t *) malloc((n + 1) * sizeof(uint64_t));
    if (!b.h_offsets) return fail(UNC_E_NOMEM, "out of host memory");
    b.h_offsets[0] = 0;
    if (n) {
        size_t dummy = 0;
        int rc;
        if ((rc = upload(&b.bwt, h.bwt.data(), h.bwt.size() * 4, 64, &dummy)) != UNC_OK) return rc;
        if ((rc = upload(&b.pac, pac.data(), pac.size(), 16, &dummy)) != UNC_OK) return rc;
        if ((rc = upload(&b.pos, pos.data(), n * 4, 0, &dummy)) != UNC_OK) return rc;
        if ((rc = upload(&b.lim, lim.data(), n * 4, 0, &dummy)) != UNC_OK) return rc;
        CUDA_TRY(cudaMalloc(&b.count, n * 4));
        CUDA_TRY(cudaMalloc(&b.stage, n * 4 * (size_t) UNC_SA_STAGE));
        CUDA_TRY(cudaMalloc(&b.offsets, (n + 1) * 8));
        DevIndex ix{};
        ix.bwt = (const uint4 *) b.bwt;
        ix.primary = (u32) h.primary;
        ix.seq_len = (u32) h.seq_len;
        for (int i = 0; i < 5; i++) ix.L2[i] = (u32) h.L2[i];
        DevSelfAlign A{};
        A.pac = (const u8 *) b.pac; A.pos = (const u32 *) b.pos; A.lim = (const u32 *) b.lim; A.n = (u32) n;
        A.count = (u32 *) b.count; A.stage = (u32 *) b.stage;
        const u32 grid = (u32) ((n + 127) / 128);
        k_selfalign_count<<<grid, 128>>>(ix, A);
        std::vector<u32> cnt(n);
        CUDA_TRY(cudaMemcpy(cnt.data(), b.count, n * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) b.h_offsets[i + 1] = b.h_offsets[i] + cnt[i];
        const uint64_t nv = b.h_offsets[n];
        b.h_values = (uint64_t *) malloc((nv ? nv : 1) * sizeof(uint64_t));
        if (!b.h_values) return fail(UNC_E_NOMEM, "out of host memory");
        if (nv) {
            CUDA_TRY(cudaMalloc(&b.values, nv * 8));
            CUDA_TRY(cudaMemcpy(b.offsets, b.h_offsets, (n + 1) * 8, cudaMemcpyHostToDevice));
            A.offsets = (const u64 *) b.offsets; A.values = (u64 *) b.values;
            k_selfalign_write<<<grid, 128>>>(ix, A);
            CUDA_TRY(cudaMemcpy(b.h_values, b.values, nv * 8, cudaMemcpyDeviceToHost));
        }
    } else {
        b.h_values = (uint64_t *) malloc(sizeof(uint64_t));
        if (!b.h_values) return fail(UNC_E_NOMEM, "out of host memory");
    }
    *n_paths = n;
    *offsets = b.h_offsets; *values = b.h_values;
    b.h_offsets = nullptr; b.h_values = nullptr;               // ownership passes to the caller (unc_free)
    return UNC_OK;
}

}  // extern "C"
