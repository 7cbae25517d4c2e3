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
struct DtwBuffers {                  // frees whatever was allocated, on every exit path
    void *model = nullptr, *means = nullptr, *kmers = nullptr, *prob = nullptr, *bc = nullptr, *diag = nullptr, *edge = nullptr,
         *path = nullptr, *path_len = nullptr, *score = nullptr, *queue = nullptr;
    ~DtwBuffers() {
        cudaFree(model); cudaFree(means); cudaFree(kmers); cudaFree(prob); cudaFree(bc); cudaFree(diag); cudaFree(edge);
        cudaFree(path); cudaFree(path_len); cudaFree(score); cudaFree(queue);
    }
};
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
        bc_total += nr * nc; diag_total += 3 * nr; edge_total += nr + nc;
    }
    for (uint64_t k = kmer_off[0]; k < kmer_off[n_problems]; k++)
        if (kmers[k] >= 1024) return fail(UNC_E_ARG, "k-mer code out of range (5-mers: 0..1023)");
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t n_means = mean_off[n_problems], n_kmers = kmer_off[n_problems], n_path = path_off[n_problems];
    if (bc_total + 4 * (diag_total + edge_total + n_means) + 2 * n_kmers + 16 * n_path + (64u << 20) > free_b)
        return fail(UNC_E_NOMEM, "the DTW matrices of this batch do not fit the device memory: pass fewer problems per call");
    DtwBuffers b;
    size_t dummy = 0;
    int rc;
    if ((rc = upload(&b.model, model.data(), model.size() * 4, 0, &dummy)) != UNC_OK) return rc;
    if ((rc = upload(&b.means, means, n_means * 4, 0, &dummy)) != UNC_OK) return rc;
    if ((rc = upload(&b.kmers, kmers, n_kmers * 2, 0, &dummy)) != UNC_OK) return rc;
    if ((rc = upload(&b.prob, prob.data(), prob.size() * sizeof(DevDtwProblem), 0, &dummy)) != UNC_OK) return rc;
    CUDA_TRY(cudaMalloc(&b.bc, bc_total));
    CUDA_TRY(cudaMalloc(&b.diag, diag_total * 4));
    CUDA_TRY(cudaMalloc(&b.edge, edge_total * 4));
    CUDA_TRY(cudaMalloc(&b.path, n_path * 16));
    CUDA_TRY(cudaMalloc(&b.path_len, (size_t) n_problems * 8));
    CUDA_TRY(cudaMalloc(&b.score, (size_t) n_problems * 4));
    CUDA_TRY(cudaMalloc(&b.queue, 4));
    CUDA_TRY(cudaMemset(b.queue, 0, 4));
    DevDtw D;
    D.model = (const float *) b.model; D.means = (const float *) b.means; D.kmers = (const u16 *) b.kmers;
    D.prob = (const DevDtwProblem *) b.prob; D.n_prob = n_problems;
    D.bc = (unsigned char *) b.bc; D.diag = (float *) b.diag; D.edge = (float *) b.edge;
    D.path = (u64 *) b.path; D.path_len = (u64 *) b.path_len; D.score = (float *) b.score;
    D.cost_kind = cost_kind; D.subseq = prm->subseq; D.dw = prm->dw; D.hw = prm->hw; D.vw = prm->vw;
    D.queue = (u32 *) b.queue;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, g_device));
    const uint32_t grid = std::min<uint32_t>(n_problems, (uint32_t) prop.multiProcessorCount * 4u);   // persistent CTAs, problems from a queue
    k_dtw<<<grid, 256>>>(D);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(path_len, b.path_len, (size_t) n_problems * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(score, b.score, (size_t) n_problems * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(path + 2 * path_off[0], (u64 *) b.path + 2 * path_off[0], (n_path - path_off[0]) * 16, cudaMemcpyDeviceToHost));
    return UNC_OK;
}
