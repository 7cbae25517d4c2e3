#include "unc_dtw.cuh"
// unc_dtw_host.inl -- C-ABI of the DTW classes (included by unc_abi.cu).  Replaces DTWr94p / DTWr94d
// (reference src/dtw.hpp:188-232 over DTW<float, u16, Func> :31-183; bound to Python at src/pybinder.cpp:75-91).

__global__ void __launch_bounds__(256) k_dtw(DevDtw D) {
    __shared__ u32 s_pi;
    for (;;) {
        if (threadIdx.x == 0) s_pi = atomicAdd(D.queue, 1u);
        __syncthreads();
        const u32 pi = s_pi;
        __syncthreads();
        if (pi >= D.n_prob) break;
        unc_dtw_problem(D, pi);
    }
}

namespace {
// One device workspace, kept between calls and grown on demand (cudaMalloc / cudaFree of the breadcrumb matrix cost more than
// the sweep); released by unc_shutdown.  Calls are serialised by g_dtw_mutex.
struct DtwWorkspace {
    void *p = nullptr;
    size_t cap = 0;
    int device = -1;
    float last_kernel_ms = 0;
};
DtwWorkspace g_dtw_ws;
std::mutex g_dtw_mutex;
size_t dtw_align(size_t x) { return (x + 255) & ~(size_t) 255; }
}  // namespace

extern "C" int unc_dtw_batch(const float *model_means_stdvs, int cost_kind, const unc_dtw_params *prm, uint32_t n_problems,
                             const float *means, const uint64_t *mean_off, const uint16_t *kmers, const uint64_t *kmer_off,
                             uint64_t *path, const uint64_t *path_off, uint64_t *path_len, float *score) {
    if (!model_means_stdvs || !prm || !means || !mean_off || !kmers || !kmer_off || !path || !path_off || !path_len || !score)
        return fail(UNC_E_ARG, "null argument");
    if (cost_kind < 0 || cost_kind > 1 || prm->subseq < 0 || prm->subseq > 2) return fail(UNC_E_ARG, "cost_kind is 0 or 1, subseq 0, 1 or 2");
    if (n_problems == 0) return UNC_OK;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(UNC_E_NO_DEVICE, "no CUDA device (the product has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(g_device));
    // the template model's tables: PoreModel(means_stdvs, false) (reference src/pore_model.hpp:58-62,77-103)
    std::vector<float> model(3 * 1024);
    for (uint32_t k = 0; k < 1024; k++) {
        const float mean = model_means_stdvs[2 * k], stdv = model_means_stdvs[2 * k + 1];
        model[k] = mean;
        model[1024 + k] = 2 * stdv * stdv;
        model[2048 + k] = (float) std::log(std::sqrt(M_PI * model[1024 + k]));
    }
    std::vector<DevDtwProblem> prob(n_problems);
    uint64_t bc_total = 0, diag_total = 0, edge_total = 0;
    for (uint32_t i = 0; i < n_problems; i++) {
        const uint64_t nc = mean_off[i + 1] - mean_off[i], nr = kmer_off[i + 1] - kmer_off[i];
        if (mean_off[i + 1] < mean_off[i] || kmer_off[i + 1] < kmer_off[i] || nc == 0 || nr == 0)
            return fail(UNC_E_ARG, "every problem needs at least one event mean and one k-mer (offsets ascending)");
        if (nc >= 0x7FFFFFFFull || nr >= 0x7FFFFFFFull) return fail(UNC_E_TOO_LARGE, "a DTW problem with 2^31 rows or columns");
        if (path_off[i + 1] - path_off[i] < nr + nc) return fail(UNC_E_ARG, "path_off must leave rows + columns pairs per problem");
        DevDtwProblem &P = prob[i];
        P.mean_off = mean_off[i]; P.kmer_off = kmer_off[i]; P.n_cols = (u32) nc; P.n_rows = (u32) nr;
        P.bc_off = bc_total; P.diag_off = diag_total; P.edge_off = edge_total; P.path_off = path_off[i];
        bc_total += nr * nc; diag_total += UNC_DTW_WORK_FLOATS(nr, nc); edge_total += nr + nc;
    }
    for (uint64_t k = kmer_off[0]; k < kmer_off[n_problems]; k++)
        if (kmers[k] >= 1024) return fail(UNC_E_ARG, "k-mer code out of range (5-mers: 0..1023)");
    const uint64_t n_means = mean_off[n_problems], n_kmers = kmer_off[n_problems], n_path = path_off[n_problems];
    std::lock_guard<std::mutex> lock(g_dtw_mutex);
    // carve the workspace
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += dtw_align(bytes); return o; };
    const size_t o_model = take(model.size() * 4), o_means = take(n_means * 4), o_kmers = take(n_kmers * 2),
                 o_prob = take(prob.size() * sizeof(DevDtwProblem)), o_bc = take(bc_total), o_diag = take(diag_total * 4),
                 o_edge = take(edge_total * 4), o_path = take(n_path * 16), o_plen = take((size_t) n_problems * 8),
                 o_score = take((size_t) n_problems * 4), o_queue = take(4);
    if (g_dtw_ws.cap < off || g_dtw_ws.device != g_device) {
        if (g_dtw_ws.p) { cudaFree(g_dtw_ws.p); g_dtw_ws.p = nullptr; g_dtw_ws.cap = 0; }
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        if (off + (64u << 20) > free_b)
            return fail(UNC_E_NOMEM, "the DTW matrices of this batch do not fit the device memory: pass fewer problems per call");
        CUDA_TRY(cudaMalloc(&g_dtw_ws.p, off));
        g_dtw_ws.cap = off; g_dtw_ws.device = g_device;
    }
    char *w = (char *) g_dtw_ws.p;
    CUDA_TRY(cudaMemcpy(w + o_model, model.data(), model.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w + o_means, means, n_means * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w + o_kmers, kmers, n_kmers * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w + o_prob, prob.data(), prob.size() * sizeof(DevDtwProblem), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(w + o_queue, 0, 4));
    DevDtw D;
    D.model = (const float *) (w + o_model); D.means = (const float *) (w + o_means); D.kmers = (const u16 *) (w + o_kmers);
    D.prob = (const DevDtwProblem *) (w + o_prob); D.n_prob = n_problems;
    D.bc = (unsigned char *) (w + o_bc); D.diag = (float *) (w + o_diag); D.edge = (float *) (w + o_edge);
    D.path = (u64 *) (w + o_path); D.path_len = (u64 *) (w + o_plen); D.score = (float *) (w + o_score);
    D.cost_kind = cost_kind; D.subseq = prm->subseq; D.dw = prm->dw; D.hw = prm->hw; D.vw = prm->vw;
    D.queue = (u32 *) (w + o_queue);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, g_device));
    const uint32_t grid = std::min<uint32_t>(n_problems, (uint32_t) prop.multiProcessorCount * 4u);   // persistent CTAs, problems from a queue
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    cudaEventRecord(e0);
    k_dtw<<<grid, 256>>>(D);
    cudaEventRecord(e1);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaEventSynchronize(e1);
    if (le == cudaSuccess) cudaEventElapsedTime(&g_dtw_ws.last_kernel_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(le);
    CUDA_TRY(cudaMemcpy(path_len, w + o_plen, (size_t) n_problems * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(score, w + o_score, (size_t) n_problems * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(path + 2 * path_off[0], (u64 *) (w + o_path) + 2 * path_off[0], (n_path - path_off[0]) * 16, cudaMemcpyDeviceToHost));
    return UNC_OK;
}

// wall-clock-free timing of the last call's sweep (CUDA events around k_dtw), for tools/bench_dtw.py
extern "C" float unc_dtw_last_kernel_ms(void) { return g_dtw_ws.last_kernel_ms; }
extern "C" void unc_dtw_release(void) {
    std::lock_guard<std::mutex> lock(g_dtw_mutex);
    if (g_dtw_ws.p) cudaFree(g_dtw_ws.p);
    g_dtw_ws.p = nullptr; g_dtw_ws.cap = 0;
}
