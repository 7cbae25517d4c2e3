"""Times K2 for every library under uncalled_b200/variants/ on the bench workload (one
subprocess per variant) and checks that all variants produce identical PAF records."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]

def one(lib, n_reads):
    import numpy as np, torch
    import uncalled_b200._native as N
    N.LIB_PATH = lib
    import uncalled_b200 as U
    import synth, synthdata
    prefix, g = synthdata.get_index("g4m7")
    cache = "/tmp/unc_variants_sig_%d.npy" % n_reads          # the same signals for every variant: generate once per box
    if os.path.exists(cache):
        sig = np.load(cache)
    else:
        sig, _ = synth.reads(g, n_reads, 4000, seed=7)
        np.save(cache, sig)
    idx = U.Index(prefix, device=0)
    bm = U.BatchMapper(idx, max_reads=n_reads, max_samples=n_reads * 4000)
    d = U.make_descs([4000] * n_reads)
    dev = torch.from_numpy(sig.reshape(-1)).cuda()
    ts = []
    for it in range(3):
        out = bm.map_device(dev.data_ptr(), d)
        ts.append(bm.timing())
    keys = ["mapped", "fwd", "rid", "status", "n_events", "events_used", "matches", "rd_len", "rd_st", "rd_en", "rf_st", "rf_en", "n_children", "n_sources", "n_seeds"]
    h = hashlib.sha1(b"".join(np.ascontiguousarray(out[k]).tobytes() for k in keys)).hexdigest()[:12]
    print(json.dumps({"lib": os.path.basename(lib), "k1_ms": [round(t["k1_ms"], 3) for t in ts], "k2_ms": [round(t["k2_ms"], 1) for t in ts],
                      "reads_per_s": round(n_reads / (min(t["total_ms"] for t in ts) / 1e3), 1), "hash": h, "bad_status": int((out["status"] != 0).sum())}), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
        vdir = os.path.join(ROOT, "uncalled_b200", "variants")
        for f in sorted(os.listdir(vdir)):
            if f.endswith(".so"):
                # name__ENV=VAL__ENV2=VAL2.so sets environment knobs for that run
                env = dict(os.environ)
                for kv in f[:-3].split("__")[1:]:
                    k, v = kv.split("=")
                    env[k] = v
                try:
                    r = subprocess.run([sys.executable, __file__, "--one", os.path.join(vdir, f), str(n)], capture_output=True, text=True, timeout=300, env=env)
                    print(r.stdout.strip() or ("FAILED %s: %s" % (f, r.stderr[-400:])), flush=True)
                except subprocess.TimeoutExpired:
                    print("TIMEOUT %s" % f, flush=True)
