// unc_stream.cuh -- the STREAMING front end: what Mapper::process_chunk does with one chunk of one
// channel (reference src/mapper.cpp:301-363): EventDetector::add_sample per sample, EventProfiler
// (src/event_profiler.hpp:71-104) per event, streaming Normalizer::push (src/normalizer.cpp:46-75) per
// unmasked event -- all serial state machines, carried across chunks in a per-channel DevChanSig.
//
// A chunk is a few hundred samples (chunk_time 0.1125 s = 450 samples), so the unit of parallelism is
// the CHANNEL: one thread walks one channel's chunk (512 channels per flow cell).  The arithmetic is the
// reference's, operation by operation (no FMA contraction: the translation unit is built with
// -fmad=false and uses the explicit-rounding primitives of unc_warp.cuh).
//
// Because RealtimePool::try_add_chunk (src/realtime_pool.cpp:108-139) hands a channel its next chunk
// only once the previous one is fully mapped, every event of a chunk is pushed before any is popped,
// and Normalizer::at (:114-118) evaluates scale/shift from the statistics at pop time: all events of a
// chunk are normalised with the statistics AFTER the chunk's last push.  The kernel therefore emits the
// chunk's unmasked event means plus one (scale, shift) pair, and the mapper kernel applies them.
#pragma once
#include "unc_device.cuh"

#define UNC_EVP_WIN 25u          /* EventProfiler::PRMS_DEF.win_len (reference src/event_profiler.cpp:4-10) */
#define UNC_EVP_STDV_MIN 5.0f    /* win_stdv_min */
#define UNC_NORM_LEN 6000u       /* Normalizer::PRMS_DEF.len (reference src/normalizer.cpp:4-8) */

struct DevRing {                 // Normalizer statistics + ring cursor (reference src/normalizer.hpp:72-79)
    double mean, varsum;
    u32 n, rd, wr, is_full;
};

struct DevChanSig {              // signal-side state of one channel, persistent across chunks and reads
    DevEvdt evdt;
    DevRing win;                 // EventProfiler::window_ (length 25)
    float win_sig[UNC_EVP_WIN];
    float q[UNC_EVP_WIN + 1];    // EventProfiler::events_ (only the means are consumed)
    u32 q_head, q_size;
    float next_mean;
    u32 p_is_full, to_mask;
    DevRing norm;                // Mapper::norm_ statistics; its ring lives in DevStream::norm_sig
    u32 pad;
};

// Normalizer::push (reference src/normalizer.cpp:46-75): Welford while filling, rolling once wrapped
UNC_DEV bool unc_ring_push(DevRing &z, float *sig, u32 len, float newevt) {
    if (z.is_full) return false;
    double oldevt = (double) sig[z.wr];
    sig[z.wr] = newevt;
    const double nv = (double) newevt;
    if (z.n == len) {
        double oldmean = z.mean;
        z.mean = d_add(z.mean, d_div(d_sub(nv, oldevt), (double) len));
        z.varsum = d_add(z.varsum, d_mul(d_sub(d_sub(d_add(nv, oldevt), oldmean), z.mean), d_sub(nv, oldevt)));
    } else {
        z.n++;
        double dt1 = d_sub(nv, z.mean);
        z.mean = d_add(z.mean, d_div(dt1, (double) z.n));
        double dt2 = d_sub(nv, z.mean);
        z.varsum = d_add(z.varsum, d_mul(dt1, dt2));
    }
    z.wr = (z.wr + 1u) % len;
    z.is_full = z.wr == z.rd ? 1u : 0u;
    return true;
}
// Normalizer::unread_size (reference src/normalizer.cpp:131-134) -- n_, not the ring length
UNC_DEV u32 unc_ring_unread(const DevRing &z) {
    if (z.rd < z.wr) return z.wr - z.rd;
    return (z.n - z.rd) + z.wr;
}

UNC_DEV void unc_ring_reset(DevRing &z, float *sig) {
    z.n = z.rd = z.wr = 0; z.mean = 0.0; z.varsum = 0.0; z.is_full = 0; sig[0] = 0.0f;
}

// EventProfiler::reset (reference src/event_profiler.hpp:54-65)
UNC_DEV void unc_evprof_reset(DevChanSig &c) {
    unc_ring_reset(c.win, c.win_sig);
    c.q_head = c.q_size = 0; c.next_mean = 0.0f; c.p_is_full = 0; c.to_mask = 0;
}

// EventProfiler::add_event (reference src/event_profiler.hpp:71-104); returns event_ready()
UNC_DEV bool unc_evprof_add(DevChanSig &c, float mean) {
    unc_ring_push(c.win, c.win_sig, UNC_EVP_WIN, mean);
    c.q[(c.q_head + c.q_size) % (UNC_EVP_WIN + 1u)] = mean; c.q_size++;
    if (unc_ring_unread(c.win) <= UNC_EVP_WIN / 2u) return false;
    float win_stdv = (float) d_sqrt(d_div(c.win.varsum, (double) c.win.n));     // Normalizer::get_stdv
    if (win_stdv < UNC_EVP_STDV_MIN) c.to_mask = UNC_EVP_WIN - 1u;
    else if (c.to_mask > 0) c.to_mask--;
    if (c.win.is_full) {
        c.next_mean = c.q[c.q_head];
        c.q_head = (c.q_head + 1u) % (UNC_EVP_WIN + 1u); c.q_size--;
        c.win.rd = (c.win.rd + 1u) % UNC_EVP_WIN;                                // Normalizer::pop (its value is unused)
        c.win.is_full = 0;
        c.p_is_full = 1;
    }
    return c.p_is_full && c.to_mask == 0;
}

struct DevStream {
    DevChanSig *sig;             // n_channels
    float *norm_sig;             // n_channels x UNC_NORM_LEN: the streaming normaliser's ring
    DevMapState *map;            // n_channels
};

// One chunk of one channel (item r of the batch).  `new_read` starts a read on the channel
// (Mapper::new_read(Chunk&) -> reset(), reference src/mapper.cpp:210-246): detector and profiler are
// reset, the normaliser keeps its statistics (norm_.skip_unread() only drops unread events -- there are
// none, see the file header), the mapper state is marked for a restart.
// Writes the chunk's unmasked event means, their count, scale/shift and mean_event_len of item r.
UNC_DEV void unc_stream_chunk(const DevBatch &B, const DevParams &p, const DevStream &S, u32 r, u32 new_read) {
    const u32 ch = B.chan[r];
    DevChanSig &c = S.sig[ch];
    float *nsig = S.norm_sig + (size_t) ch * UNC_NORM_LEN;
    if (new_read) {
        unc_evdt_reset(c.evdt);
        unc_evprof_reset(c);
        S.map[ch].started = 0;
    }
    const DevReadDesc rd = B.reads[r];
    float *ev = B.events + (size_t) r * B.ev_stride;
    u32 ne = 0;
    for (u32 i = 0; i < rd.n_samples; i++) {
        float mean;
        if (!unc_evdt_add(c.evdt, p, unc_sample(B.samples, rd, i), &mean)) continue;
        if (!unc_evprof_add(c, mean)) continue;
        // Normalizer::push cannot fail here: the ring is empty at the start of a chunk and a chunk holds
        // fewer than UNC_NORM_LEN events (unc_stream_create bounds the chunk length)
        unc_ring_push(c.norm, nsig, UNC_NORM_LEN, c.next_mean);
        ev[ne++] = c.next_mean;
    }
    c.norm.rd = c.norm.wr;                                         // the mapper pops every event of the chunk
    c.norm.is_full = 0;
    B.n_events[r] = ne;
    B.mean_event_len[r] = f_div(c.evdt.len_sum, (float) c.evdt.total_events);
    float scale = 0.0f, shift = 0.0f;
    if (ne > 0) {                                                  // Normalizer::at (reference src/normalizer.cpp:114-118)
        scale = (float) d_div((double) p.tgt_stdv, d_sqrt(d_div(c.norm.varsum, (double) c.norm.n)));
        shift = (float) d_sub((double) p.tgt_mean, d_mul((double) scale, c.norm.mean));
    }
    B.scale[r] = scale;
    B.shift[r] = shift;
}
