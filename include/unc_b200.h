/* include/unc_b200.h -- C-ABI of the B200-native `uncalled map` hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no
 * exceptions or abort() across it, every call returns 0 or a negative unc_status.
 * The library (uncalled_b200/libunc_b200.so) contains ONLY the CUDA path; there is no
 * CPU fallback -- if no CUDA device is usable every compute entry point fails with
 * UNC_E_CUDA / UNC_E_NO_DEVICE.
 *
 * Each entry point names the reference interface it replaces (paths under the reference
 * tree, skovaka/UNCALLED v2.3.0):
 *
 *   unc_index_load      Mapper::load_static            src/mapper.cpp:109-159
 *                       BwaIndex::load_index           src/bwa_index.hpp:116-135
 *   unc_index_build     BwaIndex::create -> bwa_idx_build
 *                                                      src/bwa_index.hpp:92-101
 *   unc_params_default  Mapper::PRMS and sub-structs   src/mapper.cpp:29-52,
 *                       seed_tracker.cpp:28-32, event_detector.cpp:17-26,
 *                       read_buffer.cpp:26-32          (what Conf binds, src/conf.hpp:57-95)
 *   unc_pool_create     MapPool::MapPool(Conf&)        src/map_pool.cpp:28-43
 *   unc_map_batch       MapPool::update -> MapperThread::run -> Mapper::new_read/map_read
 *                                                      src/map_pool.cpp:45-69,130-158,
 *                                                      src/mapper.cpp:188-207
 *   unc_map_batch_ordered  one MapperThread's loop with its long-lived Mapper (`-t 1`)
 *                                                      src/map_pool.cpp:104-158, src/mapper.cpp:88,216-246
 *   unc_pool_set_tie_order / unc_stream_set_tie_order
 *                       pdqsort(next_paths_...) as it is  src/mapper.cpp:531,866-871, submods/pdqsort/pdqsort.h
 *   unc_events_batch    EventDetector::get_means + Normalizer::set_signal/pop
 *                                                      src/event_detector.cpp:133-145,
 *                                                      src/normalizer.cpp:31-44,114-129
 *   unc_pool_free       MapPool::stop                  src/map_pool.cpp:71-82
 */
#ifndef UNC_B200_H
#define UNC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    UNC_OK = 0,
    UNC_E_ARG = -1,          /* bad argument */
    UNC_E_IO = -2,           /* index file missing / unreadable / inconsistent */
    UNC_E_CUDA = -3,         /* CUDA runtime error (see unc_last_error) */
    UNC_E_NO_DEVICE = -4,    /* no usable CUDA device */
    UNC_E_TOO_LARGE = -5,    /* index or batch exceeds what the device image supports */
    UNC_E_NOMEM = -6,
    UNC_E_OVERFLOW = -7      /* a per-read device workspace overflowed (record status says which) */
} unc_status;

/* Mirror of the reference's process-global parameter structs (Conf binds references to
 * them, src/conf.hpp:66-72); snapshot by value at unc_pool_create. */
typedef struct {
    /* Mapper::PRMS, src/mapper.cpp:29-52 */
    uint32_t seed_len;        /* 22; the device image supports seed_len == 22 only */
    uint32_t min_rep_len;     /* 0 */
    uint32_t max_rep_copy;    /* 50 */
    uint32_t max_paths;       /* 10000 (<= 32767) */
    uint32_t max_consec_stay; /* 8 */
    uint32_t max_events;      /* 30000 */
    float max_stay_frac;      /* 0.5 */
    float min_seed_prob;      /* -3.75 */
    /* SeedTracker::PRMS_DEF, src/seed_tracker.cpp:28-32 */
    uint32_t min_map_len;     /* 25 */
    float min_mean_conf;      /* 6.00 */
    float min_top_conf;       /* 1.85 */
    /* EventDetector::PRMS_DEF, src/event_detector.cpp:17-26 (window lengths 3/6 are fixed
     * in the device image) */
    uint32_t window_length1, window_length2;
    float threshold1, threshold2, peak_height, min_mean, max_mean;
    /* ReadBuffer::PRMS, src/read_buffer.cpp:26-32 */
    float bp_per_sec, sample_rate;
} unc_params;

/* One read of a batch.  `offset` counts samples (not bytes) from the start of `samples`. */
typedef struct {
    uint64_t offset;
    uint32_t n_samples;
    uint32_t dtype; /* UNC_DTYPE_F32: calibrated pA; UNC_DTYPE_I16: raw DAC values calibrated on
                       the device as src/read_buffer.cpp:239-242 does (u16 reinterpretation) */
    float cal_range, cal_offset, cal_digit; /* used for UNC_DTYPE_I16 only */
} unc_read_desc;

#define UNC_DTYPE_F32 0
#define UNC_DTYPE_I16 1

/* What Paf::set_mapped / set_read_len receive (src/read_buffer.cpp:133-155) plus counters. */
typedef struct {
    int32_t mapped;       /* Paf::is_mapped_ */
    int32_t fwd;          /* '+' / '-' */
    int32_t rid;          /* index into the .ann sequences, -1 when unmapped / untranslatable */
    int32_t status;       /* 0, or UNC_E_OVERFLOW when a device workspace overflowed for this read */
    uint32_t n_events;    /* valid events detected over the whole (possibly truncated) signal */
    uint32_t events_used; /* Mapper::event_i_ when mapping stopped */
    uint32_t matches;
    uint32_t n_clusters;
    uint64_t rd_len, rd_st, rd_en, rf_st, rf_en, rf_len;
    /* exact work counters (algorithmic bytes of the roofline) */
    uint64_t n_children, n_sources, n_occ_blocks, n_sa_steps, n_seeds;
} unc_paf_rec;

typedef struct {
    float h2d_ms, k1_ms, k2_ms, d2h_ms, total_ms; /* CUDA events on the pool's stream */
    uint32_t kernel_launches;                      /* launches of this library's kernels */
    uint64_t h2d_bytes, d2h_bytes;
    float k1_events_ms;                            /* the event-detection kernel alone (k1_ms also covers the
                                                      serial-redo and normaliser-statistics launches) */
    float pad_;
} unc_timing;

typedef struct {
    uint64_t n_rows;  /* FM-index length (fwd + revcomp) */
    int32_t n_seqs;
    int32_t device;
    uint64_t device_bytes;
    uint32_t n_kmer_groups;
} unc_index_info;

typedef struct unc_index unc_index;
typedef struct unc_pool unc_pool;

const char *unc_strerror(int status);
const char *unc_last_error(void); /* thread-local detail of the last failure */
int unc_device_count(void);
/* Select the CUDA device used by subsequent unc_index_load / unc_pool_create calls of
 * this process (one process per GPU). */
int unc_init(int device);
/* End of use in this process: waits for the selected device to go idle and forgets the device selection and
 * error text.  Objects the caller still holds (indexes, pools, streams) stay valid until freed; the CUDA context
 * is NOT reset, because the host application (e.g. PyTorch in bench.py) may share it. */
int unc_shutdown(void);

int unc_params_default(unc_params *p);

/* Parses <prefix>.bwt/.sa/.ann/.amb and the <preset> line of <prefix>.uncl on the host,
 * builds the device image (Occ blocks, u32 sampled SA, 1024 k-mer FM ranges, pore model,
 * thresholds) and uploads it.  model_path == NULL selects the built-in r9.4 5-mer model
 * table file given by model_table_path (1024 x {mean, stdv} f32, template order). */
int unc_index_load(const char *bwa_prefix, const char *preset, const char *model_table_path,
                   unc_index **out);
int unc_index_get_info(const unc_index *idx, unc_index_info *info);
int unc_index_seq(const unc_index *idx, int rid, const char **name, uint64_t *len);
/* Host-side copies of device tables, for parity tests. */
int unc_index_kmer_range(const unc_index *idx, uint32_t kmer, uint64_t *start, uint64_t *end);
int unc_index_thresholds(const unc_index *idx, float out[64]);
void unc_index_free(unc_index *idx);

/* bwa-compatible FM-index construction (.pac .ann .amb .bwt .sa), host C++. */
int unc_index_build(const char *fasta_path, const char *prefix);

/* Device workspace for batches of at most max_reads reads / max_samples samples in total. */
int unc_pool_create(const unc_index *idx, const unc_params *prm, uint32_t max_reads,
                    uint64_t max_samples, unc_pool **out);
void unc_pool_free(unc_pool *pool);

/* Whole path over a batch held in HOST memory (pinned or pageable): H2D of the samples,
 * event detection + normalisation kernel, mapping kernel, D2H of the records.  Synchronous. */
int unc_map_batch(unc_pool *pool, const unc_read_desc *reads, uint32_t n_reads, const void *samples,
                  unc_paf_rec *out);
/* Same, with the samples already resident in device memory (`d_samples` is a device
 * pointer on the pool's device); descriptors and results stay on the host. */
int unc_map_batch_device(unc_pool *pool, const unc_read_desc *reads, uint32_t n_reads,
                         const void *d_samples, unc_paf_rec *out);

/* Order of children that compare equal in the per-event child sort (reference src/mapper.cpp:531, pdqsort under
 * operator< :866-871; only the LAST child of a run of equal FM ranges survives, :569-572).  mode 0 (default): emission
 * order -- a stable parallel sort; what the reference computes with its sort made stable, and what it computes itself on
 * 99.7 % of reads.  mode 1: the reference's own pdqsort reproduced step by step (one thread per CTA sorts the event's
 * keys serially, kernel k2_map_exact), so that equal children land exactly where the reference's land: results are then
 * those of the unmodified reference on every read, several times slower on reads that run at the max_paths cap.
 * Applies to the batch calls of this pool; unc_stream_set_tie_order is the streaming path's switch. */
int unc_pool_set_tie_order(unc_pool *pool, int mode);

/* The batch mapped as ONE long-lived Mapper maps its reads one after the other -- what `uncalled map -t 1` prints
 * for a multi-read input.  Replaces a MapPool thread's loop over its reads (reference src/map_pool.cpp:104-158):
 * Mapper::reset (src/mapper.cpp:216-246) keeps the sources_added_ flags (:88), so a read starts with the flags its
 * predecessor left set.  unc_map_batch gives every read a clear set (the reference's result for a read that is
 * first on its thread); this call resolves the chain instead: all reads are mapped in one launch, then only those
 * whose predecessor ended with other flags than they were mapped from are mapped again, until nothing changes
 * (uncalled_b200/csrc/unc_ordered_logic.hpp).  carry[32] (k-mer k = bit k&31 of word k>>5): in, the flags before
 * reads[0] (all zero for a new Mapper); out, the flags after the last read -- pass it on to the next batch.
 * n_remapped / n_rounds (optional): reads mapped a second time, extra launches. */
int unc_map_batch_ordered(unc_pool *pool, const unc_read_desc *reads, uint32_t n_reads, const void *samples,
                          int samples_on_device, uint32_t carry[32], unc_paf_rec *out, uint32_t *n_remapped,
                          uint32_t *n_rounds);

/* The same work as two calls, so that batches on DIFFERENT pools overlap: submit queues the copies and both
 * kernels on the pool's stream and returns; wait blocks until they are done and hands the records over.  Reads
 * differ widely in cost (one that never maps takes ~6x one that does), so the last reads of a batch leave CTAs
 * idle; with two pools used alternately the next batch's CTAs fill the SMs the previous batch's tail frees -- what
 * the reference's thread pool gets by handing every idle thread the next read (src/map_pool.cpp:45-69).
 * `samples` (and, for host samples, the buffer they live in) must stay untouched until the matching wait; a pool
 * holds at most one submitted batch. */
int unc_map_batch_submit(unc_pool *pool, const unc_read_desc *reads, uint32_t n_reads, const void *samples,
                         int samples_on_device);
int unc_map_batch_wait(unc_pool *pool, unc_paf_rec *out);
/* Device-side timing of work spread over several pools: record marks event `slot` (0 or 1) at the current end of
 * the pool's stream; elapsed waits for both events and returns the milliseconds from the first to the second (they
 * may belong to different pools of the same device). */
int unc_pool_record(unc_pool *pool, int slot);
int unc_pool_elapsed(unc_pool *from, int from_slot, unc_pool *to, int to_slot, float *ms);

/* Event detection + normalisation alone.  events/normed hold `stride` floats per read
 * (stride >= the longest read's n_samples); n_events, mean_event_len one entry per read. */
int unc_events_batch(unc_pool *pool, const unc_read_desc *reads, uint32_t n_reads, const void *samples,
                     uint32_t stride, float *events, float *normed, uint32_t *n_events,
                     float *mean_event_len);

/* pore-model log-probabilities of one normalised event against all 1024 k-mers, computed
 * on the device with the mapper's own routine (parity probe). */
int unc_match_probs(const unc_index *idx, float event, float out[1024]);
/* FM probes computed on the device (parity probes): one backward step per entry, and SA. */
int unc_fm_neighbors(const unc_index *idx, uint32_t n, const uint64_t *start, const uint64_t *end,
                     const uint8_t *base, uint64_t *ostart, uint64_t *oend);
int unc_fm_sa(const unc_index *idx, uint32_t n, const uint64_t *rows, uint64_t *out);

int unc_pool_last_timing(const unc_pool *pool, unc_timing *t);
/* Counters of the event-detection kernel over the last batch: tiles processed, speculative-FSM
 * re-run rounds, lanes re-run, reads redone by the serial routine (exactness condition failed). */
int unc_pool_k1_stats(const unc_pool *pool, uint32_t out[4]);

/* ---- streaming path (chunks of many channels, persistent per-channel state on the device) ----------
 *   unc_stream_create   RealtimePool::RealtimePool(Conf&): one Mapper per channel  src/realtime_pool.cpp:38-60
 *   unc_stream_step     RealtimePool::add_chunk / try_add_chunk + MapperThread::run ->
 *                       Mapper::new_read(Chunk&) / add_chunk / process_chunk / map_chunk
 *                                                      src/realtime_pool.cpp:74-139,316-360, src/mapper.cpp:210-431
 *   unc_stream_free     RealtimePool::stop_all         src/realtime_pool.cpp:286-297
 * One step hands each listed channel its next chunk and maps ALL of the chunk's events (the reference maps
 * them evt_batch_size at a time and accepts the next chunk only when the previous one is fully mapped, so
 * the results are the same; its wall-clock limits evt_timeout / chunk_timeout do not exist here). */
typedef struct {
    uint32_t channel;     /* 0-based channel index */
    uint32_t new_read;    /* 1: this chunk starts a new read on the channel (Mapper::new_read) */
    uint64_t offset;      /* in samples from `samples` */
    uint32_t n_samples;   /* 0 (with new_read == 0): no more signal for the read in progress */
    uint32_t dtype;       /* UNC_DTYPE_F32 / UNC_DTYPE_I16 */
    float cal_range, cal_offset, cal_digit;
} unc_chunk_desc;

#define UNC_STREAM_INACTIVE 0
#define UNC_STREAM_MAPPING 1
#define UNC_STREAM_SUCCESS 2   /* rec holds the mapping */
#define UNC_STREAM_FAILURE 3   /* gave up: max_events / max_chunks / no more signal (ended = 1 like Paf::set_ended) */

typedef struct {
    int32_t state, ended;
    uint32_t chunks;      /* chunks of the read consumed so far */
    uint32_t pad_;
    unc_paf_rec rec;      /* n_events = events detected so far, events_used = Mapper::event_i_ */
} unc_stream_result;

typedef struct unc_stream unc_stream;
int unc_stream_create(const unc_index *idx, const unc_params *prm, uint32_t n_channels, uint32_t max_chunk_len,
                      uint32_t max_chunks, unc_stream **out);
int unc_stream_step(unc_stream *st, const unc_chunk_desc *chunks, uint32_t n, const void *samples,
                    unc_stream_result *out);
/* Tie order of the per-event child sort for this stream's steps: see unc_pool_set_tie_order (1 = the reference's pdqsort
 * reproduced; applies from the next step on, the per-channel state is unaffected). */
int unc_stream_set_tie_order(unc_stream *st, int mode);
/* Mapper::PRMS.chunk_timeout (src/mapper.cpp:40,384-390; default there 4000 ms, FLT_MAX for `uncalled map` / sim,
 * conf.hpp:89-90): a read whose chunk has been with the mapper for longer fails and is marked ended.  Here a chunk is with
 * the mapper for the duration of the step that maps it, so the test is made once per step on the step's wall-clock time
 * (off until set).  The reference's evt_timeout only makes a thread yield and has no counterpart.  unc_stream_last_step_ms:
 * wall-clock time of the last step, copies included (the per-chunk decision latency a ReadUntil client sees). */
int unc_stream_set_chunk_timeout(unc_stream *st, float ms);
float unc_stream_last_step_ms(const unc_stream *st);
void unc_stream_free(unc_stream *st);

/* ---- `uncalled index` after the BWA build ----------------------------------------------------------
 *   unc_self_align      self_align(bwa_prefix, sample_dist) -> vector<vector<u64>>
 *                                                      src/self_align_ref.cpp:34-91 (Python: src/pybinder.cpp:59,
 *                                                      called by IndexParameterizer, uncalled/index.py:82)
 * For every sampled reference position (glibc srand(0); rand() % sample_dist == 0, one draw per position) the
 * FM range lengths of the search that follows the reference until the range is unique.  Loads <prefix>.bwt /
 * .sa / .ann / .pac itself (no .uncl exists yet).  Result in CSR form: path i = (*values)[(*offsets)[i] ..
 * (*offsets)[i+1]), (*offsets) has *n_paths + 1 entries.  Both arrays are allocated by the library; release
 * each with unc_free. */
int unc_self_align(const char *bwa_prefix, uint32_t sample_dist, uint64_t *n_paths, uint64_t **offsets,
                   uint64_t **values);
void unc_free(void *p);

/* ---- DTW of event means against reference k-mers (analysis) ----------------------
 *   unc_dtw_batch       DTWr94p(means, kmers, prms) / DTWr94d(...) + get_path() / score()
 *                                                      src/dtw.hpp:31-183 (DTW<float, u16, Func>), :188-232 (the two
 *                                                      instances), bound at src/pybinder.cpp:75-91
 * Rows = k-mers, columns = event means.  cost_kind 0 = DTWr94p (cost = -match_prob of the r9.4 TEMPLATE model), 1 = DTWr94d
 * (cost = abs(e - mean_k) as the reference compiles it: int abs(int), the difference truncated towards zero first).
 * model_means_stdvs = the 1024 (mean, stdv) pairs of src/model_r94.inl (uncalled_b200/data/r94_5mer_template.f32).
 * unc_dtw_params = DTWParams (subseq: DTWSubSeq NONE 0 / ROW 1 / COL 2; dw, hw, vw).  Problem p: event means
 * means[mean_off[p] .. mean_off[p+1]), k-mers kmers[kmer_off[p] .. kmer_off[p+1]).  Out, per problem: get_path() as
 * (column, row) u64 pairs from the end cell back to the start at path[2*path_off[p] ...] (path_off[p+1] - path_off[p] >=
 * rows + columns), their number in path_len[p], score() in score[p]; mean_score() = score / path_len.  The matrices of
 * one call (one byte per cell) must fit the device memory: UNC_E_NOMEM otherwise. */
typedef struct { int32_t subseq; float dw, hw, vw; } unc_dtw_params;
int unc_dtw_batch(const float *model_means_stdvs, int cost_kind, const unc_dtw_params *prm, uint32_t n_problems,
                  const float *means, const uint64_t *mean_off, const uint16_t *kmers, const uint64_t *kmer_off,
                  uint64_t *path, const uint64_t *path_off, uint64_t *path_len, float *score);
/* The device workspace of unc_dtw_batch is kept between calls and grown on demand; unc_dtw_release (also called by
 * unc_shutdown) frees it.  unc_dtw_last_kernel_ms: CUDA-event time of the last call's sweep kernel. */
void unc_dtw_release(void);
float unc_dtw_last_kernel_ms(void);

/* ---- fast5 input (host; no libhdf5 needed) -------------------------------------------------------------
 *   unc_fast5_open      Fast5Reader::open_next: format detection and the list of reads
 *                                                      src/fast5_reader.cpp:134-177
 *   unc_fast5_info      ReadBuffer(hdf5_tools::File&, raw_path, ch_path): attributes as the reference parses
 *                       them (text -> atoi / atof)      src/read_buffer.cpp:198-225,
 *                                                      submods/fast5/include/fast5/hdf5_tools.hpp:1013-1141
 *   unc_fast5_load      the same constructor's `file.read(raw_path + "/Signal", int_data)` + truncation to
 *                       max_chunks * chunk_len samples  src/read_buffer.cpp:227-236
 *                       for a range of reads, decoded by several host threads into one staging buffer
 * The signal stays int16: unc_map_batch calibrates it on the device (unc_read_desc dtype 1) exactly as
 * src/read_buffer.cpp:239-242 does (u16 reinterpretation included). */
typedef struct {
    const char *read_id;        /* attribute read_id; valid until the next info/load call or close */
    int32_t number;             /* atoi(read_number) */
    int32_t start_sample;       /* atoi(start_time): 32 bit, like the reference */
    int32_t channel;            /* atoi(channel_number) -- what Paf prints as ch:i: */
    float cal_digitisation, cal_range, cal_offset;
    uint64_t n_samples;         /* after truncation, for unc_fast5_load */
    uint64_t sample_offset;     /* unc_fast5_load: where this read's samples start in dst */
} unc_fast5_read;

typedef struct unc_fast5 unc_fast5;
int unc_fast5_open(const char *path, unc_fast5 **out);
int unc_fast5_count(const unc_fast5 *f, uint32_t *n_reads, int *single_read_format);
int unc_fast5_info(unc_fast5 *f, uint32_t i, unc_fast5_read *info);
/* max_samples_per_read 0 = whole signals; threads 0 = all hardware threads.  UNC_E_TOO_LARGE when the signals
 * need more than `capacity` samples. */
int unc_fast5_load(unc_fast5 *f, uint32_t first, uint32_t n, uint64_t max_samples_per_read, int16_t *dst,
                   uint64_t capacity, unc_fast5_read *info, int threads);
void unc_fast5_close(unc_fast5 *f);
const char *unc_fast5_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
