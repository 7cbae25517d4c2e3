#include "unc_stream_logic.hpp"
// unc_stream_host.inl -- C-ABI of the streaming path (included by unc_abi.cu): per-channel persistent
// state on the device, one call per round of chunks.  Replaces, for one flow cell's channels,
// RealtimePool::add_chunk / try_add_chunk + MapperThread::run -> Mapper::process_chunk / map_chunk
// (reference src/realtime_pool.cpp:74-139,316-360, src/mapper.cpp:281-431).

__global__ void __launch_bounds__(128) k_stream_chunks(DevBatch B, DevParams p, DevStream S, const u32 *new_read) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B.n_reads) return;
    unc_stream_chunk(B, p, S, r, new_read[r]);
    B.k1_flags[r] = S.sig[B.chan[r]].evdt.total_events;      // EventDetector::total_events_ so far (reporting)
}

__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS)
k2_map_stream(DevIndex ix, DevParams p, DevBatch B, DevWork W0, DevWorkStrides S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unc_k2_cta_main_stream(ix, p, B, W0, S, (K2Shared *) smem_raw);
}

// the exact-ties instantiation (unc_stream_set_tie_order): the reference's unstable child sort reproduced, unc_pdqsort.cuh
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS_V1)
k2_map_stream_exact(DevIndex ix, DevParams p, DevBatch B, DevWork W0, DevWorkStrides S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unc_k2_cta_main_stream<true>(ix, p, B, W0, S, (K2Shared *) smem_raw);
}

struct unc_stream {
    const unc_index *idx = nullptr;
    unc_params prm;
    DevParams dp;
    uint32_t n_channels = 0, max_chunk_len = 0, max_chunks = 0, ev_stride = 0;
    std::vector<HostChan> ch;
    cudaStream_t stream = nullptr;
    // persistent device state
    DevStream S{};
    DevWork W{};
    DevWorkStrides strides{};
    size_t smem = 0;
    uint32_t grid = 0;
    int tie_order = 0;           // unc_stream_set_tie_order
    float chunk_timeout_ms = 3.0e38f;   // unc_stream_set_chunk_timeout (Mapper::PRMS.chunk_timeout; off by default, as `uncalled map` / sim set FLT_MAX)
    float last_step_ms = 0.0f;
    // per-call buffers (capacity n_channels items)
    void *d_samples = nullptr;
    DevReadDesc *d_reads = nullptr, *h_reads = nullptr;
    float *d_events = nullptr, *d_scale = nullptr, *d_shift = nullptr, *d_mel = nullptr;
    u32 *d_n_events = nullptr, *d_queue = nullptr, *d_flags = nullptr, *d_chan = nullptr, *d_new = nullptr;
    u32 *h_chan = nullptr, *h_new = nullptr, *h_flags = nullptr;
    DevRec *d_out = nullptr;
    unc_paf_rec *h_out = nullptr;
};

extern "C" {

void unc_stream_free(unc_stream *T) {
    if (!T) return;
    cudaFree(T->S.sig); cudaFree(T->S.norm_sig); cudaFree(T->S.map);
    cudaFree(T->W.paths); cudaFree(T->W.hist); cudaFree(T->W.wlist); cudaFree(T->W.ckey); cudaFree(T->W.cks);
    cudaFree(T->W.elist); cudaFree(T->W.order); cudaFree(T->W.rlist); cudaFree(T->W.clu); cudaFree(T->W.dir);
    cudaFree(T->d_samples); cudaFree(T->d_reads); cudaFreeHost(T->h_reads);
    cudaFree(T->d_events); cudaFree(T->d_scale); cudaFree(T->d_shift); cudaFree(T->d_mel);
    cudaFree(T->d_n_events); cudaFree(T->d_queue); cudaFree(T->d_flags); cudaFree(T->d_chan); cudaFree(T->d_new);
    cudaFreeHost(T->h_chan); cudaFreeHost(T->h_new); cudaFreeHost(T->h_flags);
    cudaFree(T->d_out); cudaFreeHost(T->h_out);
    if (T->stream) cudaStreamDestroy(T->stream);
    delete T;
}

int unc_stream_set_tie_order(unc_stream *T, int mode) {
    if (!T || (mode != 0 && mode != 1)) return fail(UNC_E_ARG, "bad argument");
    if (mode == 1) {
        CUDA_TRY(cudaSetDevice(T->idx->device));
        CUDA_TRY(raise_dyn_smem(k2_map_stream_exact, T->smem));
    }
    T->tie_order = mode;
    return UNC_OK;
}

// Mapper::PRMS.chunk_timeout (reference src/mapper.cpp:40,384-390): a read whose chunk has been with the mapper for longer
// than this fails and is marked ended.  A chunk is "with the mapper" for the duration of the step that maps it (the step
// returns when all of its chunks' events are mapped; the reference's per-call evt_timeout only yields a thread and has
// no effect on results), so the test is made once per step, on the step's wall-clock time.
int unc_stream_set_chunk_timeout(unc_stream *T, float ms) {
    if (!T || !(ms > 0)) return fail(UNC_E_ARG, "bad argument");
    T->chunk_timeout_ms = ms;
    return UNC_OK;
}
float unc_stream_last_step_ms(const unc_stream *T) { return T ? T->last_step_ms : 0.0f; }

int unc_stream_create(const unc_index *idx, const unc_params *prm, uint32_t n_channels, uint32_t max_chunk_len,
                      uint32_t max_chunks, unc_stream **out) {
    if (!idx || !prm || !out || n_channels == 0 || max_chunk_len == 0 || max_chunks == 0)
        return fail(UNC_E_ARG, "null/zero argument");
    // a chunk must hold fewer events than the normaliser ring (events are >= 2 samples apart)
    if (max_chunk_len > 8000) return fail(UNC_E_ARG, "max_chunk_len above 8000 samples is not supported by the streaming image");
    std::string err;
    if (unc_check_params(*prm, err)) return fail(UNC_E_ARG, err);
    CUDA_TRY(cudaSetDevice(idx->device));
    unc_stream *T = new unc_stream();
    T->idx = idx; T->prm = *prm; T->dp = unc_make_dev_params(*prm, idx->h);
    T->n_channels = n_channels; T->max_chunk_len = max_chunk_len; T->max_chunks = max_chunks;
    T->ev_stride = (max_chunk_len + 3u) & ~3u;
    T->ch.resize(n_channels);
    auto bail = [&](int code, const std::string &m) { unc_stream_free(T); return fail(code, m); };
#define ST(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return bail(UNC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } while (0)
    ST(cudaStreamCreateWithFlags(&T->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    ST(cudaGetDeviceProperties(&prop, idx->device));
    T->smem = K2_SMEM_BYTES(prm->max_paths);
    ST(raise_dyn_smem(k2_map_stream, T->smem));
    int per_sm = 0;
    ST(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_map_stream, K2_THREADS, T->smem));
    if (per_sm < 1) return bail(UNC_E_CUDA, "k2_map_stream does not fit on an SM");
    T->grid = (uint32_t) prop.multiProcessorCount * (uint32_t) per_sm;
    // one workspace slot per CHANNEL (the path buffers, history and seed clusters of its read in progress)
    const size_t maxp = prm->max_paths, nchmax = (maxp + 31) / 32;
    DevWorkStrides &S = T->strides;
    S.paths = 2 * (nchmax * 160 + maxp) * 2;
    S.hist = 24 * (nchmax * 160 + maxp);
    S.ckey = 2 * maxp;
    S.cks = nchmax * 160;
    S.elist = nchmax * 32;
    S.order = 2 * maxp;
    const size_t rl_cap = 64 * 1024;
    uint64_t ev_cap = std::min<uint64_t>(prm->max_events, (uint64_t) max_chunks * (max_chunk_len / 2 + 1));
    uint64_t mb = std::min<uint64_t>(std::max<uint64_t>(1024, ev_cap * 2), 1u << 17);
    S.rlist = 2 * rl_cap;
    S.clu = (size_t) mb * UNC_BLK * 2;
    S.dir = (size_t) mb + 1;
    T->W.rl_cap = (u32) rl_cap;
    T->W.max_blocks = (u32) mb;
    const size_t n = n_channels;
    size_t per_slot = (S.paths + S.ckey + 2 * S.cks + S.elist) * 16 + S.hist * 8 + S.order * 4 + S.rlist * 8 + S.clu * 16 + S.dir * 16;
    size_t free_b = 0, total_b = 0;
    ST(cudaMemGetInfo(&free_b, &total_b));
    if ((double) n * (double) per_slot > 0.9 * (double) free_b) return bail(UNC_E_NOMEM, "not enough device memory for the per-channel workspaces");
    ST(cudaMalloc(&T->W.paths, n * S.paths * 16));
    ST(cudaMalloc(&T->W.ckey, n * S.ckey * 16));
    ST(cudaMalloc(&T->W.hist, n * S.hist * 8));
    ST(cudaMemset(T->W.hist, 0, n * S.hist * 8));
    ST(cudaMalloc(&T->W.wlist, n * S.cks * 16));
    ST(cudaMalloc(&T->W.cks, n * S.cks * 16));
    ST(cudaMalloc(&T->W.elist, n * S.elist * 16));
    ST(cudaMalloc(&T->W.order, n * S.order * 4));
    ST(cudaMalloc(&T->W.rlist, n * S.rlist * 8));
    ST(cudaMalloc(&T->W.clu, n * S.clu * 16));
    ST(cudaMalloc(&T->W.dir, n * S.dir * 16));
    ST(cudaMalloc(&T->S.sig, n * sizeof(DevChanSig)));
    ST(cudaMemset(T->S.sig, 0, n * sizeof(DevChanSig)));
    ST(cudaMalloc(&T->S.norm_sig, n * UNC_NORM_LEN * 4));
    ST(cudaMemset(T->S.norm_sig, 0, n * UNC_NORM_LEN * 4));
    ST(cudaMalloc(&T->S.map, n * sizeof(DevMapState)));
    ST(cudaMemset(T->S.map, 0, n * sizeof(DevMapState)));
    ST(cudaMalloc(&T->d_samples, n * (size_t) max_chunk_len * 4 + 64));
    ST(cudaMalloc(&T->d_reads, n * sizeof(DevReadDesc)));
    ST(cudaMallocHost(&T->h_reads, n * sizeof(DevReadDesc)));
    ST(cudaMalloc(&T->d_events, n * (size_t) T->ev_stride * 4));
    ST(cudaMalloc(&T->d_scale, n * 4)); ST(cudaMalloc(&T->d_shift, n * 4)); ST(cudaMalloc(&T->d_mel, n * 4));
    ST(cudaMalloc(&T->d_n_events, n * 4)); ST(cudaMalloc(&T->d_queue, 32)); ST(cudaMalloc(&T->d_flags, n * 4));
    ST(cudaMalloc(&T->d_chan, n * 4)); ST(cudaMalloc(&T->d_new, n * 4));
    ST(cudaMallocHost(&T->h_chan, n * 4)); ST(cudaMallocHost(&T->h_new, n * 4)); ST(cudaMallocHost(&T->h_flags, n * 4));
    ST(cudaMalloc(&T->d_out, n * sizeof(DevRec)));
    ST(cudaMallocHost(&T->h_out, n * sizeof(unc_paf_rec)));
#undef ST
    *out = T;
    return UNC_OK;
}

// One round of chunks: at most one entry per channel.  n_samples == 0 means "no more signal for the
// read in progress" (RealtimePool::try_add_chunk with an empty chunk -> Mapper::request_reset).
int unc_stream_step(unc_stream *T, const unc_chunk_desc *chunks, uint32_t n, const void *samples,
                    unc_stream_result *out) {
    if (!T || !chunks || !out || (n && !samples)) return fail(UNC_E_ARG, "null argument");
    if (n > T->n_channels) return fail(UNC_E_ARG, "more chunks than channels");
    CUDA_TRY(cudaSetDevice(T->idx->device));
    const float bp_per_samp = T->prm.bp_per_sec / T->prm.sample_rate;
    std::vector<int> item_of(n, -1);
    std::vector<char> seen(T->n_channels, 0);
    uint32_t m = 0;
    uint64_t hi = 0;
    uint32_t dtype = 0xFFFFFFFFu;
    // validate every descriptor before any channel's bookkeeping changes (an error leaves the stream as it was)
    for (uint32_t i = 0; i < n; i++) {
        const unc_chunk_desc &c = chunks[i];
        if (c.channel >= T->n_channels) return fail(UNC_E_ARG, "channel out of range");
        if (seen[c.channel]) return fail(UNC_E_ARG, "two chunks for one channel in one step");
        seen[c.channel] = 1;
        if (c.n_samples > T->max_chunk_len) return fail(UNC_E_TOO_LARGE, "chunk longer than max_chunk_len");
        if (c.dtype > 1) return fail(UNC_E_ARG, "unknown dtype");
        if (c.n_samples || c.new_read) {
            if (dtype == 0xFFFFFFFFu) dtype = c.dtype;
            if (c.dtype != dtype) return fail(UNC_E_ARG, "mixed dtype in a step");
            hi = std::max<uint64_t>(hi, c.offset + c.n_samples);
        }
    }
    if (hi > (uint64_t) T->n_channels * T->max_chunk_len) return fail(UNC_E_TOO_LARGE, "samples exceed the staging buffer");
    if (dtype == 0xFFFFFFFFu) dtype = 0;
    hi = 0;
    const auto t_step0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < n; i++) {
        const unc_chunk_desc &c = chunks[i];
        if (!stream_admit(T->ch[c.channel], c, T->max_chunks)) continue;
        DevReadDesc &d = T->h_reads[m];
        d.offset = c.offset; d.n_samples = c.n_samples; d.dtype = c.dtype;
        d.cal_range = c.cal_range; d.cal_offset = c.cal_offset; d.cal_digit = c.cal_digit; d.pad = 0;
        hi = std::max<uint64_t>(hi, c.offset + c.n_samples);
        T->h_chan[m] = c.channel; T->h_new[m] = c.new_read ? 1u : 0u;
        item_of[i] = (int) m++;
    }
    if (m) {
        cudaStream_t s = T->stream;
        const uint64_t span = hi * (dtype == UNC_DTYPE_F32 ? 4 : 2);
        CUDA_TRY(cudaMemcpyAsync(T->d_samples, samples, span, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(T->d_reads, T->h_reads, (size_t) m * sizeof(DevReadDesc), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(T->d_chan, T->h_chan, (size_t) m * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(T->d_new, T->h_new, (size_t) m * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemsetAsync(T->d_queue, 0, 32, s));
        DevBatch B;
        memset(&B, 0, sizeof(B));
        B.samples = T->d_samples; B.samples_bytes = span; B.reads = T->d_reads; B.n_reads = m;
        B.events = T->d_events; B.normed = nullptr; B.ev_stride = T->ev_stride; B.n_events = T->d_n_events;
        B.scale = T->d_scale; B.shift = T->d_shift; B.mean_event_len = T->d_mel;
        B.queue = T->d_queue; B.k1_queue = T->d_queue + 1; B.k1_flags = T->d_flags; B.k1_stats = nullptr;
        B.out = T->d_out; B.dbg = nullptr;
        B.seq_offsets = (const u64 *) T->idx->d_seq_off; B.seq_lens = (const u32 *) T->idx->d_seq_len;
        B.n_seqs = (u32) T->idx->h.names.size(); B.l_pac = (u64) T->idx->h.l_pac;
        B.mstate = T->S.map; B.chan = T->d_chan;
        k_stream_chunks<<<(m + 127) / 128, 128, 0, s>>>(B, T->dp, T->S, T->d_new);
        CUDA_TRY(cudaGetLastError());
        if (T->tie_order)
            k2_map_stream_exact<<<std::min<uint32_t>(T->grid, m), K2_THREADS, T->smem, s>>>(T->idx->ix, T->dp, B, T->W, T->strides);
        else
            k2_map_stream<<<std::min<uint32_t>(T->grid, m), K2_THREADS, T->smem, s>>>(T->idx->ix, T->dp, B, T->W, T->strides);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(T->h_out, T->d_out, (size_t) m * sizeof(DevRec), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(T->h_flags, T->d_flags, (size_t) m * 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    T->last_step_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_step0).count();
    const bool timed_out = T->last_step_ms > T->chunk_timeout_ms;
    int worst = UNC_OK;
    for (uint32_t i = 0; i < n; i++) {
        HostChan &h = T->ch[chunks[i].channel];
        if (item_of[i] >= 0) {
            const unc_paf_rec &r = T->h_out[item_of[i]];
            stream_settle(h, r, T->h_flags[item_of[i]], T->prm.max_events, T->max_chunks);
            if (r.status != 0) worst = UNC_E_OVERFLOW;
            if (timed_out && h.state == UNC_STREAM_MAPPING) { h.state = UNC_STREAM_FAILURE; h.ended = 1; }   // map_chunk's first test
        }
        stream_result(h, bp_per_samp, &out[i]);
    }
    if (worst) return fail(worst, "a channel overflowed its seed-cluster workspace");
    return UNC_OK;
}

}  // extern "C"
