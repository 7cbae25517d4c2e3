#!/usr/bin/env python
"""Run the REAL reference `uncalled index` (built under $UNC_REF_BUILD from /root/reference, see
SURVEY.md 8(c)) on the deterministic synthetic genomes of tools/synth.py and record the
resulting .uncl parameter lines in tests/golden/synth_uncl.json.

The FM-index files themselves are rebuilt anywhere by the product's own index builder
(byte-identical to bwa's, tests/test_index_build.py); only the threshold line -- which comes
from the reference's Python IndexParameterizer (uncalled/index.py:53-209) -- needs a fixture.
"""
import json, os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402
REF_BUILD = os.environ.get("UNC_REF_BUILD", "/tmp/oracle/ref")
GENOMES = {"g200k": (200_000, 1234), "g4m7": (4_700_000, 1234), "g1m": (1_000_000, 99)}
out = {}
d = tempfile.mkdtemp()
for name, (n, seed) in GENOMES.items():
    fa = os.path.join(d, name + ".fa")
    synth.write_fasta(fa, synth.genome(n, seed))
    subprocess.run([sys.executable, os.path.join(REF_BUILD, "scripts", "uncalled"), "index", "-o",
                    os.path.join(d, name), fa], check=True, env=dict(os.environ, PYTHONPATH=REF_BUILD),
                   capture_output=True)
    out[name] = {"size": n, "seed": seed, "uncl": open(os.path.join(d, name + ".uncl")).read()}
    print(name, out[name]["uncl"].strip())
json.dump(out, open(os.path.join(HERE, "..", "tests", "golden", "synth_uncl.json"), "w"), indent=1)
