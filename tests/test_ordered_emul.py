"""CPU tier: ordered mapping (`uncalled map -t 1`: one long-lived Mapper, sources_added_ flags carried from read
to read) -- the product's host logic (uncalled_b200/csrc/unc_ordered_logic.hpp) driving the emulated kernels,
against the oracle's one-Mapper chain (pinned to the reference's own long-lived Mapper in test_oracle_pinned.py)."""
import numpy as np
import pytest

import emulib
import orclib
import synth
import synthdata


@pytest.fixture(scope="module")
def setup():
    prefix, g = synthdata.get_index("g200k")
    E, O = emulib.Emu(prefix), orclib.Oracle(prefix)
    # a small path buffer fills on most events, so reads end with flags set and their successors see them
    E.params.max_paths = O.params.max_paths = 300
    sig, _ = synth.reads(g, 40, 2000, seed=21, frac_random=0.4)
    # 5 ends (mapped) with flags set and 6 then extends differently; 7 and 1 never map; 0, 1, 25 leave flags behind
    sigs = [np.ascontiguousarray(sig[i], np.float32) for i in (5, 6, 7, 0, 1, 25, 26)]
    sigs.insert(5, np.full(30, 90.0, np.float32))       # a read without events passes the flags on untouched
    return E, O, sigs


def _chain(O, sigs):
    flat = np.concatenate(sigs)
    lens = np.array([len(s) for s in sigs], np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    return O.map_reads_one_mapper(flat, offs, lens)


def _counts(r):
    return (r.n_children, r.n_sources, r.n_seeds, r.n_clusters)


def test_final_flags_of_every_read_match_the_oracle(setup):
    """flags_out of the kernel, from clear flags and from a predecessor's: the same words the oracle's Mapper holds when
    map_read returns -- including reads that end by mapping, where the kernel has run one event ahead."""
    E, O, sigs = setup
    prev = np.zeros(32, np.uint32)
    some_set = 0
    for s in sigs:
        want_rec, want_flags = O.map_read_flags(s, prev)
        recs, carry, _, _ = E.map_ordered([s], carry=prev)
        assert emulib.paf_tuple(recs[0]) == orclib.paf_tuple(want_rec)
        assert np.array_equal(carry, want_flags)
        some_set += int(want_flags.any())
        prev = want_flags
    assert some_set >= 3


def test_ordered_batch_equals_one_long_lived_mapper(setup):
    E, O, sigs = setup
    want = _chain(O, sigs)
    fresh = [O.map_read(s) for s in sigs]
    differ = [i for i in range(len(sigs)) if (orclib.paf_tuple(want[i]), _counts(want[i])) != (orclib.paf_tuple(fresh[i]), _counts(fresh[i]))]
    assert differ, "the read set must contain reads that depend on their predecessor"
    recs, carry, n_remapped, n_rounds = E.map_ordered(sigs)
    for i in range(len(sigs)):
        assert emulib.paf_tuple(recs[i]) == orclib.paf_tuple(want[i]), i
        assert _counts(recs[i]) == _counts(want[i]), i
    assert n_remapped >= len(differ) and 1 <= n_rounds <= len(sigs)


def test_carry_links_consecutive_batches(setup):
    E, O, sigs = setup
    sigs = sigs[:2]
    want = _chain(O, sigs)
    a, carry, _, _ = E.map_ordered(sigs[:1])
    b, carry2, _, _ = E.map_ordered(sigs[1:], carry=carry)
    got = a + b
    for i in range(len(sigs)):
        assert (emulib.paf_tuple(got[i]), _counts(got[i])) == (orclib.paf_tuple(want[i]), _counts(want[i])), i
    prev = np.zeros(32, np.uint32)
    for s in sigs:
        _, prev = O.map_read_flags(s, prev)
    assert np.array_equal(carry2, prev)


def test_first_event_fills_the_buffer(setup):
    """max_paths below the number of first-event candidates: the fresh-source walk stops early, flags behind the cut
    survive the first event, and the candidate-mask shortcut must not be taken."""
    E, O, sigs = setup
    sigs = [s[:1200] for s in sigs[:4]]
    old = E.params.max_paths
    E.params.max_paths = O.params.max_paths = 40
    try:
        want = _chain(O, sigs)
        recs, carry, n_remapped, _ = E.map_ordered(sigs)
        for i in range(len(sigs)):
            assert (emulib.paf_tuple(recs[i]), _counts(recs[i])) == (orclib.paf_tuple(want[i]), _counts(want[i])), i
        prev = np.zeros(32, np.uint32)
        for s in sigs:
            _, prev = O.map_read_flags(s, prev)
        assert np.array_equal(carry, prev) and prev.any() and n_remapped >= 1
    finally:
        E.params.max_paths = O.params.max_paths = old
