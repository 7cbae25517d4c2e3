"""configs[3] of BASELINE.json at a size one GPU call can carry: a chr1-sized (default 230 Mb, with masked `N` runs)
synthetic index -- FM rows far beyond the 126 MB L2, so the Occ look-ups go to HBM -- long reads, `max_events` 30000.

Everything is made on the box: genome + FASTA, the bwa-compatible index (unc_index_build, own SA-IS), the `.uncl`
thresholds (`uncalled index`'s second half: unc_self_align on the GPU + the parameter search), then the reads are
mapped through unc_map_batch.  Reported: build / self-align / map times, reads/s (device-resident and end to end),
the mapper kernel's algorithmic bytes and roofline fraction, parity of the first reads against the oracle (CPU
restatement, test infrastructure) and the reference CPU arm (oracle/_ref, all usable host threads) on a subsample.

    python tools/bench_chr1.py [--mb 230] [--reads 2048] [--samples 32000] [--parity-reads 64] [--cpu-reads 128]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def make_genome(n, seed, n_runs, run_len):
    """ACGT codes of the sequence the reads come from, and the same with masked stretches marked (code 4 -> 'N')."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, n, dtype=np.uint8)
    masked = g.copy()
    starts = np.sort(rng.integers(0, n - run_len, n_runs))
    for s in starts:
        masked[s:s + run_len] = 4
    return g, masked, starts


def write_fasta(path, codes, name="synthetic_chr1"):
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    s = lut[codes]
    with open(path, "wb") as f:
        f.write((">" + name + "\n").encode())
        n = len(s)
        full = (n // 60) * 60
        body = np.empty((full // 60, 61), np.uint8)
        body[:, :60] = s[:full].reshape(-1, 60)
        body[:, 60] = 10
        f.write(body.tobytes())
        if full < n:
            f.write(s[full:].tobytes() + b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=230.0)
    ap.add_argument("--reads", type=int, default=2048)
    ap.add_argument("--samples", type=int, default=32000)
    ap.add_argument("--parity-reads", type=int, default=64)
    ap.add_argument("--cpu-reads", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--dir", default=os.path.join(ROOT, "bench_data"))
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--reuse", action="store_true", help="use the index files of an earlier --keep run if they are there")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle parity and the CPU arm (profiling runs)")
    args = ap.parse_args()

    import torch
    import synth
    import uncalled_b200 as U
    from uncalled_b200 import index as UI
    import uncalled_b200._native as N
    import bench

    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    n_bases = int(args.mb * 1e6)
    os.makedirs(args.dir, exist_ok=True)
    prefix = os.path.join(args.dir, "chr1_%dm" % int(args.mb))
    info = {}
    t = time.time()
    g, masked, runs = make_genome(n_bases, 4242, n_runs=max(1, n_bases // 1_000_000), run_len=5000)
    have = args.reuse and all(os.path.exists(prefix + e) for e in (".bwt", ".sa", ".pac", ".ann", ".amb", ".uncl"))
    if not have:
        write_fasta(prefix + ".fa", masked)
        info["genome_fasta_s"] = time.time() - t
        t = time.time()
        N.check(N.lib().unc_index_build((prefix + ".fa").encode(), prefix.encode()))
        info["fm_index_build_s"] = time.time() - t
        t = time.time()
        off, val = UI.self_align_csr(prefix, UI.IP.sample_distance(UI.IP.reference_length(prefix), UI.IP.DEFAULTS["max_sample_dist"],
                                                                UI.IP.DEFAULTS["min_samples"], UI.IP.DEFAULTS["max_samples"]))
        info["self_align_gpu_s"] = time.time() - t           # unc_self_align: index load + the two kernels + copy back
        info["self_align_paths"] = int(len(off) - 1)
        t = time.time()
        with open(prefix + UI.UNCL_SUFF, "w") as f:
            f.write(UI.IP.uncl_text(off, val))                # the parameter search of `uncalled index` (host, numpy)
        info["uncl_param_search_s"] = time.time() - t
    info["fm_index_files_mb"] = sum(os.path.getsize(prefix + e) for e in (".bwt", ".sa", ".pac")) / 1e6

    # reads: starts outside the masked stretches (bwa fills N with random bases the generator does not know)
    rng = np.random.default_rng(99)
    L = args.samples
    sig, truth = synth.reads(g, args.reads, L, seed=77, noise_mult=bench.NOISE_MULT)
    t = time.time()
    idx = U.Index(prefix, device=0)
    info["index_load_s"] = time.time() - t
    p = U.default_params()
    p.max_events = 30000
    bm = U.BatchMapper(idx, params=p, max_reads=args.reads, max_samples=args.reads * L)
    descs = U.make_descs([L] * args.reads)
    host = torch.from_numpy(sig.reshape(-1)).pin_memory()
    dev = host.cuda()
    torch.cuda.synchronize()
    bm.map_device(dev.data_ptr(), descs)                    # warm-up
    tms = []
    for _ in range(args.steps):
        out = bm.map_device(dev.data_ptr(), descs)
        tms.append(bm.timing())
    e2e = []
    for _ in range(args.steps):
        out_h = bm.map(host.numpy(), descs)
        e2e.append(bm.timing())
    assert np.array_equal(out, out_h)
    ms = float(np.mean([t_["total_ms"] for t_ in tms]))
    k2_ms = float(np.mean([t_["k2_ms"] for t_ in tms]))
    e2e_ms = float(np.mean([t_["total_ms"] for t_ in e2e]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k2_bytes = 64.0 * float(out["n_occ_blocks"].sum()) + 8.0 * float(out["n_seeds"].sum()) + \
        2 * 56.0 * float(out["n_children"].sum() + out["n_sources"].sum())

    # parity: the first reads against the oracle (CPU restatement), PAF fields and counters
    import orclib
    if args.no_cpu:
        args.parity_reads = args.cpu_reads = 0
    npar = min(args.parity_reads, args.reads)
    cpus = bench.host_cpus()
    want, O = [], None
    if npar or args.cpu_reads:
        O = orclib.Oracle(prefix)
        O.params.max_events = 30000
    if npar:
        t = time.time()
        offs = np.arange(npar, dtype=np.uint64) * L
        want = O.map_batch(np.ascontiguousarray(sig[:npar]).ravel(), offs, np.full(npar, L, np.uint32), cpus["usable"])
        info["oracle_parity_s"] = time.time() - t
    bad = []
    for i in range(npar):
        a, b = orclib.paf_tuple(want[i]), U.paf_key(out[i])
        ca = (want[i].n_children, want[i].n_sources, want[i].n_seeds, want[i].n_clusters)
        cb = (int(out[i]["n_children"]), int(out[i]["n_sources"]), int(out[i]["n_seeds"]), int(out[i]["n_clusters"]))
        if a != b or ca != cb or int(out[i]["status"]) != 0:
            bad.append(i)

    # the reference CPU arm on a subsample (oracle/_ref when built, else the port)
    ncpu = min(args.cpu_reads, args.reads)
    cpu = None
    if ncpu > 0:
        flat = np.ascontiguousarray(sig[:ncpu]).ravel()
        lens = np.full(ncpu, L, np.uint32)
        offs = np.arange(ncpu, dtype=np.uint64) * L
        if orclib.ref_available():
            R = orclib.ref()
            R.ref_load(prefix.encode(), b"default")
            R.ref_set_max_events(30000)
            recs = (orclib.RefPaf * ncpu)()
            t = time.time()
            R.ref_map_batch_mt(orclib.fp(flat), offs.ctypes.data_as(orclib.u64p), lens.ctypes.data_as(orclib.u32p), ncpu, cpus["usable"], recs)
            dt = time.time() - t
            kind = "reference"
        else:
            t = time.time()
            O.map_batch(flat, offs, lens, cpus["usable"])
            dt = time.time() - t
            kind = "port"
        cpu = {"value": ncpu / dt, "unit": "reads/s", "cores": cpus["usable"], "kind": kind, "host_cpus": cpus,
               "sample": "first %d reads, %d threads, %.1f s" % (ncpu, cpus["usable"], dt)}

    line = {"metric": bench.METRIC, "workload": "configs[3]-like: %.0f Mb synthetic index with %d masked N runs (%d FM rows), %d reads x %d "
                                                "samples, max_events 30000, noise %.1f x level stdv" % (args.mb, len(runs), 2 * n_bases, args.reads, L, bench.NOISE_MULT),
            "value": args.reads / (ms / 1e3), "unit": "reads/s", "n_gpus": 1, "ms_per_step": ms, "steps": args.steps,
            "e2e": {"value": args.reads / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(e2e[-1]["h2d_bytes"]),
                    "d2h_bytes_per_step": int(e2e[-1]["d2h_bytes"])},
            "events_per_read": float(out["n_events"].mean()), "events_used_per_read": float(out["events_used"].mean()),
            "mapped_fraction": float(out["mapped"].mean()),
            "roofline": {"kernel": "k2_map", "bound": "hbm", "achieved": k2_bytes / (k2_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": k2_bytes / (k2_ms / 1e3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": k2_bytes,
                         "launch_ms": k2_ms},
            "parity": {"reads": npar, "identical_to_oracle": npar - len(bad), "differing_ids": bad[:16]},
            "cpu_baseline": cpu, "index": info, "device_index_bytes": int(idx.device_bytes)}
    print(json.dumps(line), flush=True)
    if not args.keep:
        for e in (".fa", ".bwt", ".sa", ".pac", ".ann", ".amb", ".uncl"):
            try:
                os.remove(prefix + e)
            except OSError:
                pass
    if bad:
        raise SystemExit("parity failed on reads %s" % bad[:16])


if __name__ == "__main__":
    main()
