// TEST INFRASTRUCTURE.  Stand-in for the reference's vendored pdqsort (submods/pdqsort, an UNSTABLE sort) used only
// to build oracle/_ref/libuncalled_ref_stable.so: the same reference sources with a STABLE child sort.  The oracle
// and the CUDA path keep children that compare equal under (fm_range, seed_prob) in emission order (DESIGN.md
// section 2, "tie order"); if the sort's instability is the only thing that separates them from the reference, this
// build must agree with them on every read -- tests/test_oracle_pinned.py checks exactly that.
#pragma once
#include <algorithm>

template <class It>
inline void pdqsort(It first, It last) { std::stable_sort(first, last); }
template <class It, class Compare>
inline void pdqsort(It first, It last, Compare comp) { std::stable_sort(first, last, comp); }
