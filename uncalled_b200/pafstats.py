"""`uncalled pafstats` (reference uncalled/pafstats.py:125-207) for the PAF this package writes: mapped fraction,
accuracy against a reference PAF (true / false positives and negatives with the reference's location test: same
reference sequence and overlapping reference ranges once both alignments are extended by 1.5x their unaligned read
ends), and the speed summary from the `mt` tags.  Host-side analysis of the mapper's output, vectorised with numpy;
`summary()` returns the numbers, `run()` prints them in the reference's layout.  (DTW, reference src/dtw.hpp, is a
signal-level debugging aid of the reference and is not provided.)"""
import sys

import numpy as np


def parse(lines):
    """PAF lines -> dict of numpy arrays (one row per line; unmapped rows have rf_name '' and zero ranges)."""
    names, rf = [], []
    rows, mt = [], []
    for l in lines:
        if not l.strip() or l[0] == "#":
            continue
        t = l.split()
        names.append(t[0])
        mapped = t[4] != "*"
        if mapped:
            rows.append((int(t[1]), 1, int(t[2]), int(t[3]), t[4] == "+", int(t[6]), int(t[7]), int(t[8])))
            rf.append(t[5])
        else:
            rows.append((int(t[1]), 0, 1, int(t[1]), False, 0, 0, 0))
            rf.append("")
        m = np.nan
        for tag in t[12:]:
            if tag.startswith("mt:f:"):
                m = float(tag[5:])
        mt.append(m)
    a = np.array(rows, dtype=np.int64).reshape(-1, 8)
    return {"name": np.array(names), "qr_len": a[:, 0], "mapped": a[:, 1].astype(bool), "qr_st": a[:, 2], "qr_en": a[:, 3],
            "fwd": a[:, 4].astype(bool), "rf_len": a[:, 5], "rf_st": a[:, 6], "rf_en": a[:, 7], "rf_name": np.array(rf),
            "mt": np.array(mt)}


def _ext_ref(p, ext):
    """PafEntry.ext_ref (reference uncalled/pafstats.py:70-80)."""
    st_shift = (p["qr_st"] * ext).astype(np.int64)
    en_shift = ((p["qr_len"] - p["qr_en"]) * ext).astype(np.int64)
    lo = np.where(p["fwd"], p["rf_st"] - st_shift, p["rf_st"] - en_shift)
    hi = np.where(p["fwd"], p["rf_en"] + en_shift, p["rf_en"] + st_shift)
    return np.maximum(1, lo), np.minimum(p["rf_len"], hi)


def compare(qry, ref, ext=1.5):
    """paf_ref_compare (reference uncalled/pafstats.py:125-163) with one reference alignment per read: counts of
    (tp, tn, fp, fn, na) -- na = mapped here, unmapped or absent in the reference PAF."""
    pos = {n: i for i, n in enumerate(ref["name"])}
    j = np.array([pos.get(n, -1) for n in qry["name"]])
    have = j >= 0
    r_mapped = np.zeros(len(j), bool)
    r_mapped[have] = ref["mapped"][j[have]]
    qlo, qhi = _ext_ref(qry, ext)
    rlo_a, rhi_a = _ext_ref(ref, ext)
    rlo, rhi = np.zeros(len(j), np.int64), np.zeros(len(j), np.int64)
    rlo[have], rhi[have] = rlo_a[j[have]], rhi_a[j[have]]
    same = np.zeros(len(j), bool)
    same[have] = np.char.startswith(qry["rf_name"][have], ref["rf_name"][j[have]])
    overlap = qry["mapped"] & r_mapped & same & (np.maximum(qlo, rlo) <= np.minimum(qhi, rhi))
    q = qry["mapped"]
    return {"tp": int((q & r_mapped & overlap).sum()), "fp": int((q & r_mapped & ~overlap).sum()), "na": int((q & ~r_mapped).sum()),
            "tn": int((~q & ~r_mapped).sum()), "fn": int((~q & r_mapped).sum()), "n": int(len(j))}


def summary(qry, ref=None):
    out = {"reads": int(len(qry["name"])), "mapped": int(qry["mapped"].sum())}
    if ref is not None:
        out["vs_reference"] = compare(qry, ref)
    m = qry["mapped"] & np.isfinite(qry["mt"])
    if m.any():
        ms, bp = qry["mt"][m], qry["qr_en"][m].astype(float)
        out["speed"] = {"bp_per_sec": (float(np.mean(1000 * bp / ms)), float(np.median(1000 * bp / ms))),
                        "bp_mapped": (float(bp.mean()), float(np.median(bp))), "ms_to_map": (float(ms.mean()), float(np.median(ms)))}
    return out


def run(infile, ref_paf=None, max_reads=None, out=sys.stdout):
    lines = list(open(infile))
    qry = parse(lines if max_reads is None else [l for l in lines if l[:1] != "#"][:max_reads])
    s = summary(qry, parse(open(ref_paf)) if ref_paf else None)
    out.write("Summary: %d reads, %d mapped (%.2f%%)\n\n" % (s["reads"], s["mapped"], 100 * s["mapped"] / max(1, s["reads"])))
    if "vs_reference" in s:
        c = s["vs_reference"]
        n = max(1, c["n"])
        out.write("Comparing to reference PAF\n     P     N\nT %6.2f %5.2f\nF %6.2f %5.2f\nNA: %.2f\n\n"
                  % (100 * c["tp"] / n, 100 * c["tn"] / n, 100 * c["fp"] / n, 100 * c["fn"] / n, 100 * c["na"] / n))
    if "speed" in s:
        sp = s["speed"]
        out.write("Speed            Mean    Median\nBP per sec: %9.2f %9.2f\nBP mapped:  %9.2f %9.2f\nMS to map:  %9.2f %9.2f\n"
                  % (sp["bp_per_sec"] + sp["bp_mapped"] + sp["ms_to_map"]))
    return s
