#!/usr/bin/env python
"""Generate tests/golden/ fixtures by running the REAL reference (built from /root/reference).

Needs the full reference Python extension (`_uncalled`, with HDF5) built per SURVEY.md 8(c)
into $UNC_REF_BUILD (default /tmp/oracle/ref).  This script is the provenance of every file
under tests/golden/; it cannot run on the GPU box (no /root/reference there), which is why
its outputs are committed.

Outputs
  example_read.npz       raw pA signal (f32) of the example fast5 as Fast5Reader/ReadBuffer
                         calibrates it (reference src/read_buffer.cpp:198-246), read id,
                         channel, start sample; EventDetector.get_events() events;
                         Normalizer.set_signal/pop normalised means; mean_event_len
  example_model.npz      pmodel_r94_complement.match_prob at fixed (event, kmer) points
  example_index.npz      FM-index known answers over the shipped example index:
                         size, k-mer counts, sa(i) for all i
  example_index_files.npz  bytes of the shipped example index (.bwt .sa .ann .amb .pac .uncl)
                         and of example_ref.fa (fixtures for the GPU box)
  example_paf.json       PAF fields of `uncalled map` (MapPool, 1 thread) on the example read
                         for: default; -c 1 (first 4000 samples); -e 100
"""
import json, os, sys, time
import numpy as np

REF_BUILD = os.environ.get("UNC_REF_BUILD", "/tmp/oracle/ref")
sys.path.insert(0, REF_BUILD)
import _uncalled as u  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
EX = "/root/reference/example"
FAST5 = [os.path.join(EX, f) for f in os.listdir(EX) if f.endswith(".fast5")][0]
PREFIX = os.path.join(EX, "index", "example_ref")


def read_example():
    fr = u.Fast5Reader()
    fr.add_fast5(FAST5)
    fr.fill_buffer()
    rb = fr.pop_read()
    return rb


def run_map(**conf_kw):
    """Drive MapPool exactly as scripts/uncalled map_cmd does (reference scripts/uncalled:127-167)."""
    import subprocess, textwrap
    code = textwrap.dedent(f"""
        import sys, time, os
        sys.path.insert(0, {REF_BUILD!r})
        import _uncalled as u
        conf = u.Conf()
        conf.bwa_prefix = {PREFIX!r}
        conf.threads = 1
        for k, v in {conf_kw!r}.items(): setattr(conf, k, v)
        pool = u.MapPool(conf)
        pool.add_fast5({FAST5!r})
        while pool.running():
            for p in pool.update():
                p.print_paf()
            time.sleep(0.01)
        pool.stop()
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.split("\n") if l and not l.startswith("#")]
    assert len(lines) == 1, out
    f = lines[0].split("\t")
    return {"line": "\t".join(x for x in f if not x.startswith("mt:f")), "fields": f[:12]}


def main():
    os.makedirs(OUT, exist_ok=True)
    rb = read_example()
    raw = np.array(rb.raw, dtype=np.float32)
    ed = u.EventDetector()
    evs = ed.get_events(list(raw))
    ev_mean = np.array([e.mean for e in evs], dtype=np.float32)
    ev_start = np.array([e.start for e in evs], dtype=np.uint32)
    ev_len = np.array([e.length for e in evs], dtype=np.uint32)
    mel = np.float32(ed.mean_event_len())
    means2 = np.array(ed.get_means(list(raw)), dtype=np.float32)
    assert (means2 == ev_mean).all()
    model = u.pmodel_r94_complement
    norm = u.Normalizer(model.get_means_mean(), model.get_means_stdv())
    norm.set_signal(list(ev_mean))
    normed = np.array([norm.pop() for _ in range(len(ev_mean))], dtype=np.float32)
    # 4000-sample windows (config 2 read length)
    win_counts = {}
    for off in (0, 4000, 8000, 20000):
        win_counts[str(off)] = len(u.EventDetector().get_means(list(raw[off:off + 4000])))
    np.savez_compressed(os.path.join(OUT, "example_read.npz"), raw=raw, read_id=np.array(rb.id),
                        channel=np.int32(rb.channel), start=np.int64(rb.start),
                        ev_mean=ev_mean, ev_start=ev_start, ev_len=ev_len, mean_event_len=mel,
                        normed=normed, model_mean=np.float32(model.get_means_mean()),
                        model_stdv=np.float32(model.get_means_stdv()),
                        win_offsets=np.array([0, 4000, 8000, 20000]),
                        win_counts=np.array([win_counts[k] for k in ("0", "4000", "8000", "20000")]))

    rng = np.random.default_rng(5)
    ev_pts = np.concatenate([np.float32([90.0, 75.5, 110.25, 90.2083511352539]),
                             rng.uniform(50, 140, 60).astype(np.float32)])
    probs = np.array([[model.match_prob(float(e), k) for k in range(1024)] for e in ev_pts], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "example_model.npz"), events=ev_pts, probs=probs)

    idx = u.BwaIndex(PREFIX, False)
    n = idx.size()
    kc = np.array([idx.get_kmer_count(k) for k in range(1024)], dtype=np.uint64)
    sa = np.array([idx.sa(i) for i in range(1, n + 1)], dtype=np.uint64)  # rows 1..n
    np.savez_compressed(os.path.join(OUT, "example_index.npz"), size=np.uint64(n), kmer_count=kc, sa_1_to_n=sa)

    # the shipped example index + its FASTA, byte for byte, so GPU-box tests can materialise
    # them (no /root/reference there) and so the product's own index builder can be checked
    # for byte identity against `bwa index` output.
    files = {}
    for ext in ("bwt", "sa", "ann", "amb", "pac", "uncl"):
        files[ext] = np.frombuffer(open(PREFIX + "." + ext, "rb").read(), dtype=np.uint8)
    files["fasta"] = np.frombuffer(open(os.path.join(EX, "example_ref.fa"), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "example_index_files.npz"), **files)

    paf = {
        "default": run_map(),
        "max_chunks_1": run_map(max_chunks=1),
        "max_events_100": run_map(max_events=100),
    }
    json.dump(paf, open(os.path.join(OUT, "example_paf.json"), "w"), indent=1)
    print(json.dumps(paf, indent=1))
    print("events", len(ev_mean), "mean_event_len", mel, "windows", win_counts)


if __name__ == "__main__":
    main()
