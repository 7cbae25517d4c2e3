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
// flags equivalent to its predecessor's true final ones.
// Which reads have to be mapped again: a read's first event has no children, so its initial flags act only in that
// event's fresh-source walk (reference src/mapper.cpp:605-624), which skips flagged k-mers and clears every flag it
// visits.  With cand = the k-mers that pass the walk's other tests (round 0 also returns this 1024-bit mask per read,
// unc_event0_cand) and fewer candidates than max_paths (the walk then visits all k-mers), two initial flag sets
// that agree on cand give the same sources, the same cleared flags and hence the same mapping and final flags.  So a
// read is re-mapped only if its predecessor's final flags differ from the ones it was mapped from ON A CANDIDATE (or
// the walk could stop early: cand >= max_paths).  In practice (4.7 Mb index, 4000-sample reads) 4 % of the reads end
// with a flag set, and a fraction of a percent of the reads see one of those flags on a candidate.
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

// map_subset(ids, n_ids, flags_in /* n_ids x 32 */, flags_out /* n_ids x 32 */, recs /* n_ids */,
//            cand /* n_ids x 32, or NULL when not wanted */) -> UNC_OK or an error.
// A read's record status != 0 (workspace overflow) is reported by the caller; its flags are still passed on.
template <class MapSubset>
int unc_ordered_map(uint32_t n, uint32_t max_paths, uint32_t carry[32], unc_paf_rec *out, uint32_t *n_remapped,
                    uint32_t *n_rounds, MapSubset map_subset) {
    if (n_remapped) *n_remapped = 0;
    if (n_rounds) *n_rounds = 0;
    if (n == 0) return UNC_OK;
    std::vector<uint32_t> fin((size_t) n * 32, 0u), fout((size_t) n * 32, 0u), cand((size_t) n * 32, 0u), ids(n);
    memcpy(fin.data(), carry, 128);
    for (uint32_t i = 0; i < n; i++) ids[i] = i;
    int rc = map_subset(ids.data(), n, fin.data(), fout.data(), out, cand.data());
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
            const uint32_t *want = &fout[(size_t) (i - 1) * 32], *cd = &cand[(size_t) i * 32];
            uint32_t *have = &fin[(size_t) i * 32];
            if (memcmp(have, want, 128) == 0) continue;
            uint32_t on_cand = 0, n_cand = 0;
            for (int w = 0; w < 32; w++) { on_cand |= (have[w] ^ want[w]) & cd[w]; n_cand += (uint32_t) __builtin_popcount(cd[w]); }
            if (on_cand || n_cand >= max_paths) {
                ids.push_back(i);
                sub_in.insert(sub_in.end(), want, want + 32);
            } else {
                memcpy(have, want, 128);                    // same mapping, same final flags: nothing to redo
            }
        }
        if (ids.empty()) break;
        const uint32_t m = (uint32_t) ids.size();
        sub_out.assign((size_t) m * 32, 0u);
        sub_recs.resize(m);
        rc = map_subset(ids.data(), m, sub_in.data(), sub_out.data(), sub_recs.data(), (uint32_t *) nullptr);
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
