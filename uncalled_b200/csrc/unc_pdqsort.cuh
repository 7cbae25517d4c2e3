// unc_pdqsort.cuh -- the reference's child sort run as it is: pattern-defeating quicksort (the vendored
// submods/pdqsort/pdqsort.h, Orson Peters; the non-branchless variant, PathBuffer not being arithmetic) under
// PathBuffer's operator< (reference src/mapper.cpp:866-871), called at src/mapper.cpp:531.
//
// pdqsort is unstable: children that compare equal -- same FM range, same seed_prob -- end up in an order that is a
// function of the whole array's input order, and the reference then keeps the LAST of each run of equal ranges
// (:569-572).  The default kernel keeps equal children in emission order (a parallel stable sort); 0.3 % of reads
// end a seed differently for it.  The opt-in exact-ties kernel (k2_map_exact) instead has one
// thread run this serial sort over the event's sort keys in emission order, which is the array the reference sorts:
// the algorithm is deterministic and sees the same comparisons, so the keys land exactly where the reference's
// structs land.  Every step follows the header: insertion sorts (:76-121), the bounded partial insertion sort
// (:123-148), partition_right / partition_left (:340-408), the loop with its median selection, pattern-breaking swaps
// and heapsort fallback (:411-504, libstdc++'s make_heap / sort_heap).  The recursion on the left part is an explicit
// stack in global memory (the right part is pushed, the left one continued, so the order of work is the header's).
// ---------------------------------------------------------------------------------------------------------------
// This file is an ALTERED version of pdqsort.h (index-based device code for 16-byte sort keys, explicit stack instead
// of recursion, the non-branchless variant only); it is not the original software.  The original's notice, which may
// not be removed from any source distribution:
//
//     pdqsort.h - Pattern-defeating quicksort.
//
//     Copyright (c) 2015 Orson Peters
//
//     This software is provided 'as-is', without any express or implied warranty. In no event will the
//     authors be held liable for any damages arising from the use of this software.
//
//     Permission is granted to anyone to use this software for any purpose, including commercial
//     applications, and to alter it and redistribute it freely, subject to the following restrictions:
//
//     1. The origin of this software must not be misrepresented; you must not claim that you wrote the
//        original software. If you use this software in a product, an acknowledgment in the product
//        documentation would be appreciated but is not required.
//
//     2. Altered source versions must be plainly marked as such, and must not be misrepresented as
//        being the original software.
//
//     3. This notice may not be removed or altered from any source distribution.
// ---------------------------------------------------------------------------------------------------------------
#pragma once

#ifdef UNC_EMUL
static unsigned long g_emu_pdq_heapsorts;
static int g_emu_pdq_min_bad = 1 << 30;       // smallest budget of unbalanced partitions seen (test search objective)
#endif
// operator< on sort keys: x = fm_start, y = fm_end, z = seed_prob bits (src/mapper.cpp:866-871, Range::operator<)
UNC_DEV bool pq_less(const uint4 &a, const uint4 &b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return u2f(a.z) < u2f(b.z);
}
UNC_DEV void pq_swap(uint4 *a, int i, int j) { const uint4 t = a[i]; a[i] = a[j]; a[j] = t; }
UNC_DEV void pq_sort2(uint4 *a, int i, int j) { if (pq_less(a[j], a[i])) pq_swap(a, i, j); }
UNC_DEV void pq_sort3(uint4 *a, int i, int j, int k) { pq_sort2(a, i, j); pq_sort2(a, j, k); pq_sort2(a, i, j); }

// pdqsort.h:76-97 (guarded) / :99-121 (unguarded: a[b - 1] is known not to exceed anything in [b, e))
UNC_DEV void pq_insertion(uint4 *a, int b, int e, bool guarded) {
    if (b == e) return;
    for (int cur = b + 1; cur != e; cur++) {
        int sift = cur, sift_1 = cur - 1;
        if (pq_less(a[sift], a[sift_1])) {
            const uint4 tmp = a[sift];
            do { a[sift--] = a[sift_1]; } while ((!guarded || sift != b) && pq_less(tmp, a[--sift_1]));
            a[sift] = tmp;
        }
    }
}

// pdqsort.h:123-148: false once more than 8 element moves were needed
UNC_DEV bool pq_partial_insertion(uint4 *a, int b, int e) {
    if (b == e) return true;
    int limit = 0;
    for (int cur = b + 1; cur != e; cur++) {
        if (limit > 8) return false;
        int sift = cur, sift_1 = cur - 1;
        if (pq_less(a[sift], a[sift_1])) {
            const uint4 tmp = a[sift];
            do { a[sift--] = a[sift_1]; } while (sift != b && pq_less(tmp, a[--sift_1]));
            a[sift] = tmp;
            limit += cur - sift;
        }
    }
    return true;
}

// pdqsort.h:340-381: keys equal to the pivot go right; *already = [b, e) was partitioned on entry
UNC_DEV int pq_partition_right(uint4 *a, int b, int e, bool *already) {
    const uint4 pivot = a[b];
    int first = b, last = e;
    while (pq_less(a[++first], pivot)) {}
    if (first - 1 == b) { while (first < last && !pq_less(a[--last], pivot)) {} }
    else { while (!pq_less(a[--last], pivot)) {} }
    *already = first >= last;
    while (first < last) {
        pq_swap(a, first, last);
        while (pq_less(a[++first], pivot)) {}
        while (!pq_less(a[--last], pivot)) {}
    }
    const int pivot_pos = first - 1;
    a[b] = a[pivot_pos];
    a[pivot_pos] = pivot;
    return pivot_pos;
}

// pdqsort.h:384-408: keys equal to the pivot go left
UNC_DEV int pq_partition_left(uint4 *a, int b, int e) {
    const uint4 pivot = a[b];
    int first = b, last = e;
    while (pq_less(pivot, a[--last])) {}
    if (last + 1 == e) { while (first < last && !pq_less(pivot, a[++first])) {} }
    else { while (!pq_less(pivot, a[++first])) {} }
    while (first < last) {
        pq_swap(a, first, last);
        while (pq_less(pivot, a[--last])) {}
        while (!pq_less(pivot, a[++first])) {}
    }
    const int pivot_pos = last;
    a[b] = a[pivot_pos];
    a[pivot_pos] = pivot;
    return pivot_pos;
}

// libstdc++ bits/stl_heap.h __adjust_heap (+ __push_heap) on a[f .. f + len)
UNC_DEV void pq_heap_adjust(uint4 *a, int f, int hole, int len, const uint4 value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (pq_less(a[f + child], a[f + child - 1])) child--;
        a[f + hole] = a[f + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[f + hole] = a[f + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && pq_less(a[f + parent], value)) {
        a[f + hole] = a[f + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[f + hole] = value;
}
// std::make_heap + std::sort_heap: pdqsort's fallback after log2(n) highly unbalanced partitions (pdqsort.h:464-468)
UNC_DEV void pq_heapsort(uint4 *a, int b, int e) {
#ifdef UNC_EMUL
    g_emu_pdq_heapsorts++;       // test statistics (emulator builds only): was the fallback exercised?
#endif
    const int len = e - b;
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; parent--) {
        pq_heap_adjust(a, b, parent, len, a[b + parent]);
        if (parent == 0) break;
    }
    int last = e;
    while (last - b > 1) {
        --last;
        const uint4 value = a[last];
        a[last] = a[b];
        pq_heap_adjust(a, b, 0, last - b, value);
    }
}

// pdqsort(begin, end) over a[0 .. n), n < 2^31.  stack: cap uint4 entries of scratch (begin, end, bad_allowed, leftmost).
// Returns false if the scratch stack was too small (nothing sensible can be returned then; the caller reports it).
UNC_DEV_NOINLINE bool unc_pdq_sort(uint4 *a, u32 n, uint4 *stack, u32 cap) {
    if (n == 0) return true;
    int lg = 0;
    for (u32 m = n; m >>= 1;) lg++;
    u32 sp = 0;
    int begin = 0, end = (int) n, bad_allowed = lg;
    bool leftmost = true;
    for (;;) {
        // ---- pdqsort_loop (pdqsort.h:411-504) on [begin, end)
        bool done = false;
        const int size = end - begin;
        if (size < 24) {
            pq_insertion(a, begin, end, leftmost);
            done = true;
        } else {
            const int s2 = size / 2;
            if (size > 128) {
                pq_sort3(a, begin, begin + s2, end - 1);
                pq_sort3(a, begin + 1, begin + (s2 - 1), end - 2);
                pq_sort3(a, begin + 2, begin + (s2 + 1), end - 3);
                pq_sort3(a, begin + (s2 - 1), begin + s2, begin + (s2 + 1));
                pq_swap(a, begin, begin + s2);
            } else {
                pq_sort3(a, begin + s2, begin, end - 1);
            }
            if (!leftmost && !pq_less(a[begin - 1], a[begin])) {
                begin = pq_partition_left(a, begin, end) + 1;
                continue;
            }
            bool already;
            const int pivot_pos = pq_partition_right(a, begin, end, &already);
            const int l_size = pivot_pos - begin, r_size = end - (pivot_pos + 1);
            const bool highly_unbalanced = l_size < size / 8 || r_size < size / 8;
            if (highly_unbalanced) {
#ifdef UNC_EMUL
                if (bad_allowed - 1 < g_emu_pdq_min_bad) g_emu_pdq_min_bad = bad_allowed - 1;
#endif
                if (--bad_allowed == 0) {
                    pq_heapsort(a, begin, end);
                    done = true;
                } else {
                    if (l_size >= 24) {
                        pq_swap(a, begin, begin + l_size / 4);
                        pq_swap(a, pivot_pos - 1, pivot_pos - l_size / 4);
                        if (l_size > 128) {
                            pq_swap(a, begin + 1, begin + (l_size / 4 + 1));
                            pq_swap(a, begin + 2, begin + (l_size / 4 + 2));
                            pq_swap(a, pivot_pos - 2, pivot_pos - (l_size / 4 + 1));
                            pq_swap(a, pivot_pos - 3, pivot_pos - (l_size / 4 + 2));
                        }
                    }
                    if (r_size >= 24) {
                        pq_swap(a, pivot_pos + 1, pivot_pos + (1 + r_size / 4));
                        pq_swap(a, end - 1, end - r_size / 4);
                        if (r_size > 128) {
                            pq_swap(a, pivot_pos + 2, pivot_pos + (2 + r_size / 4));
                            pq_swap(a, pivot_pos + 3, pivot_pos + (3 + r_size / 4));
                            pq_swap(a, end - 2, end - (1 + r_size / 4));
                            pq_swap(a, end - 3, end - (2 + r_size / 4));
                        }
                    }
                }
            } else if (already && pq_partial_insertion(a, begin, pivot_pos) && pq_partial_insertion(a, pivot_pos + 1, end)) {
                done = true;
            }
            if (!done) {
                // the header recurses into the left part and then loops on the right one: park the right part
                if (sp == cap) return false;
                stack[sp++] = make_uint4((u32) (pivot_pos + 1), (u32) end, (u32) bad_allowed, 0u);
                end = pivot_pos;
                continue;
            }
        }
        if (sp == 0) return true;
        const uint4 t = stack[--sp];
        begin = (int) t.x; end = (int) t.y; bad_allowed = (int) t.z; leftmost = t.w != 0;
    }
}
