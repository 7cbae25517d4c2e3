// Ordered mapping: the reads of a batch as ONE long-lived Mapper maps them one after the other, which is
// what `uncalled map -t 1` computes (reference src/map_pool.cpp:104-158: a MapPool thread keeps its Mapper;
// Mapper::reset, src/mapper.cpp:216-246, clears everything except the sources_added_ flags, :88).  The flags a
// read ends with are the only state its successor inherits, so the chain is resolved by fixed-point iteration
// over whole batches instead of serialising the reads:
//   round 0   every read is mapped from clear flags (read 0 from `carry`, what the previous batch ended with);
//   round k   the reads whose predecessor's final flags differ from the flags they were last mapped from are
//             mapped again from exactly those flags.
// Read 0 is right after round 0, and once read i-1 is right read i is mapped from the right flags in the next
// round, so the first inconsistent index moves forward every round and the loop ends with every read mapped from
// its predecessor's true final flags.  In practice (4.7 Mb index, 4000-sample reads) 4 % of the reads end with a
// flag set and chains are one or two reads long: round 1 re-maps those 4 %, round 2 a handful, round 3 none.
// A read without events never touches the flags (Mapper::map_read does not enter map_next): it passes its
// predecessor's flags on and is never re-mapped.
//
// Host logic only, shared by the product (unc_abi.cu: map_subset launches the kernels on the subset) and the
// emulator harness (tests/emul/emul_main.cpp), like unc_stream_logic.hpp.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/unc_b200.h"

// map_subset(ids, n_ids, flags_in /* n_ids x 32 */, flags_out /* n_ids x 32 */, recs /* n_ids */) -> UNC_OK or an error.
// A read's record status != 0 (workspace overflow) is reported by the caller; its flags are still passed on.
template <class MapSubset>
int unc_ordered_map(uint32_t n, uint32_t carry[32], unc_paf_rec *out, uint32_t *n_remapped, uint32_t *n_rounds,
                    MapSubset map_subset) {
    if (n_remapped) *n_remapped = 0;
    if (n_rounds) *n_rounds = 0;
    if (n == 0) return UNC_OK;
    std::vector<uint32_t> fin((size_t) n * 32, 0u), fout((size_t) n * 32, 0u), ids(n);
    memcpy(fin.data(), carry, 128);
    for (uint32_t i = 0; i < n; i++) ids[i] = i;
    int rc = map_subset(ids.data(), n, fin.data(), fout.data(), out);
    if (rc != UNC_OK && rc != UNC_E_OVERFLOW) return rc;
    int worst = rc;
    std::vector<uint32_t> sub_in, sub_out;
    std::vector<unc_paf_rec> sub_recs;
    uint32_t remapped = 0, rounds = 0;
    for (;;) {
        ids.clear();
        sub_in.clear();
        for (uint32_t i = 0; i < n; i++) {
            if (out[i].n_events == 0) {                     // no map_next call: flags pass through untouched
                const uint32_t *before = i ? &fout[(size_t) (i - 1) * 32] : carry;
                memcpy(&fin[(size_t) i * 32], before, 128);
                memcpy(&fout[(size_t) i * 32], before, 128);
                continue;
            }
            if (i == 0) continue;
            const uint32_t *want = &fout[(size_t) (i - 1) * 32];
            if (memcmp(&fin[(size_t) i * 32], want, 128) != 0) {
                ids.push_back(i);
                sub_in.insert(sub_in.end(), want, want + 32);
            }
        }
        if (ids.empty()) break;
        const uint32_t m = (uint32_t) ids.size();
        sub_out.assign((size_t) m * 32, 0u);
        sub_recs.resize(m);
        rc = map_subset(ids.data(), m, sub_in.data(), sub_out.data(), sub_recs.data());
        if (rc != UNC_OK && rc != UNC_E_OVERFLOW) return rc;
        if (rc) worst = rc;
        for (uint32_t j = 0; j < m; j++) {
            const uint32_t i = ids[j];
            memcpy(&fin[(size_t) i * 32], &sub_in[(size_t) j * 32], 128);
            memcpy(&fout[(size_t) i * 32], &sub_out[(size_t) j * 32], 128);
            out[i] = sub_recs[j];
        }
        remapped += m;
        rounds++;
    }
    memcpy(carry, &fout[(size_t) (n - 1) * 32], 128);
    if (n_remapped) *n_remapped = remapped;
    if (n_rounds) *n_rounds = rounds;
    return worst;
}
