#!/bin/bash
# Regenerates tests/golden/fast5/*.fast5 and tests/golden/fast5/golden.json.
#   1. builds the reference's vendored libhdf5 (submods/hdf5, 1.8.21) under $H5 (default /tmp/h5) -- a copy, the
#      reference tree is read-only;
#   2. make_fixtures.c writes the fixture files with that library;
#   3. ref_dump.cpp = the reference's own Fast5Reader + ReadBuffer (compiled from /root/reference) reads them
#      back (and the example fast5 the reference ships) and prints what a mapper would be handed.
# Needs /root/reference; the outputs are committed, so tests do not.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF=${REF:-/root/reference}
H5=${H5:-/tmp/h5}
OUT="$ROOT/tests/golden/fast5"
if [ ! -f "$H5/inst/lib/libhdf5.a" ]; then
    mkdir -p "$H5" && cp -r "$REF/submods/hdf5" "$H5/src" && chmod -R u+w "$H5/src"
    (cd "$H5/src" && ./configure --disable-hl --prefix="$H5/inst" --enable-shared=no --with-pic=yes >/dev/null && make -j16 >/dev/null && make install >/dev/null)
fi
mkdir -p "$OUT" "$H5/bin"
gcc -O1 -I"$H5/inst/include" "$HERE/make_fixtures.c" "$H5/inst/lib/libhdf5.a" -lz -ldl -lm -o "$H5/bin/make_fixtures"
g++ -std=c++11 -O2 -w -include array -I"$REF/src" -I"$REF/submods" -I"$REF/submods/fast5/include" -I"$H5/inst/include" \
    "$HERE/ref_dump.cpp" "$REF/src/fast5_reader.cpp" "$REF/src/read_buffer.cpp" "$REF/src/chunk.cpp" \
    "$H5/inst/lib/libhdf5.a" -lz -ldl -lm -pthread -o "$H5/bin/ref_dump"
"$H5/bin/make_fixtures" "$OUT"
cp "$REF"/example/*.fast5 "$OUT/example_single.fast5"
python - "$H5/bin/ref_dump" "$OUT" <<'PY'
import json, subprocess, sys, glob, os
tool, out = sys.argv[1], sys.argv[2]
files = sorted(glob.glob(os.path.join(out, "*.fast5")))
rows = []
for mc in ("1000000", "2"):
    rows += json.loads(subprocess.run([tool, mc] + files, check=True, capture_output=True, text=True).stdout)
json.dump(rows, open(os.path.join(out, "golden.json"), "w"), indent=0)
print(len(rows), "golden rows from", len(files), "files")
PY
ls -la "$OUT"
