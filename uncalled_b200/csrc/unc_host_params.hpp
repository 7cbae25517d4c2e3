// unc_host_params.hpp -- unc_params (C-ABI) -> DevParams (device).  Requires unc_device.cuh,
// unc_host_index.hpp and include/unc_b200.h to be included first.
#pragma once

static inline int unc_check_params(const unc_params &p, std::string &err) {
    if (p.seed_len != 22) { err = "seed_len must be 22 (device path-record layout)"; return -1; }
    if (p.window_length1 != 3 || p.window_length2 != 6) { err = "event window lengths must be 3/6"; return -1; }
    if (p.max_paths == 0 || p.max_paths > 32767) { err = "max_paths must be in 1..32767"; return -1; }
    return 0;
}

static inline DevParams unc_make_dev_params(const unc_params &p, const HostIndex &h) {
    DevParams d;
    d.max_rep_copy = p.max_rep_copy; d.max_paths = p.max_paths; d.max_consec_stay = p.max_consec_stay;
    d.max_events = p.max_events; d.min_rep_len = p.min_rep_len;
    d.max_stay_frac = p.max_stay_frac; d.min_seed_prob = p.min_seed_prob;
    d.min_map_len = p.min_map_len; d.min_mean_conf = p.min_mean_conf; d.min_top_conf = p.min_top_conf;
    d.threshold1 = p.threshold1; d.threshold2 = p.threshold2; d.peak_height = p.peak_height;
    d.min_mean = p.min_mean; d.max_mean = p.max_mean;
    d.bp_per_sec = p.bp_per_sec; d.sample_rate = p.sample_rate;
    d.tgt_mean = h.model_mean; d.tgt_stdv = h.model_stdv;
    return d;
}

static inline void unc_fill_default_params(unc_params *p) {
    p->seed_len = 22; p->min_rep_len = 0; p->max_rep_copy = 50; p->max_paths = 10000;
    p->max_consec_stay = 8; p->max_events = 30000; p->max_stay_frac = 0.5f; p->min_seed_prob = -3.75f;
    p->min_map_len = 25; p->min_mean_conf = 6.00f; p->min_top_conf = 1.85f;
    p->window_length1 = 3; p->window_length2 = 6; p->threshold1 = 1.4f; p->threshold2 = 9.0f;
    p->peak_height = 0.2f; p->min_mean = 0; p->max_mean = 400; p->bp_per_sec = 450; p->sample_rate = 4000;
}
