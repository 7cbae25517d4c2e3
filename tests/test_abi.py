"""CPU tier: the C-ABI library loads, exports every symbol include/unc_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    from uncalled_b200 import _native as N
    L = C.CDLL(N.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "unc_b200.h")).read()
    names = sorted(set(re.findall(r"\b(unc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(N.EXPORTS) <= set(names)


def test_struct_layouts_match_header():
    from uncalled_b200 import _native as N
    assert C.sizeof(N.ReadDesc) == 32 and C.sizeof(N.PafRec) == 120 and C.sizeof(N.Params) == 80


def test_header_is_plain_c_and_stream_structs_match(tmp_path):
    """include/unc_b200.h must be usable from C (it is the drop-in boundary), and the ctypes mirrors of the
    streaming structs must have the layout the header declares."""
    import subprocess
    from uncalled_b200 import stream as S
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "unc_b200.h"\nint main(void) { printf("%zu %zu %zu\\n", '
                   'sizeof(unc_chunk_desc), sizeof(unc_stream_result), sizeof(unc_timing)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-o", str(exe), str(src)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from uncalled_b200 import _native as N
    assert sizes == [C.sizeof(S.ChunkDesc), C.sizeof(S.StreamResult), C.sizeof(N.Timing)]


def test_defaults_mirror_reference_params():
    from uncalled_b200 import _native as N
    p = N.default_params()
    assert (p.seed_len, p.max_paths, p.max_events, p.max_consec_stay, p.max_rep_copy) == (22, 10000, 30000, 8, 50)
    assert abs(p.min_seed_prob + 3.75) < 1e-7 and p.min_map_len == 25 and abs(p.min_top_conf - 1.85) < 1e-6
    assert (p.window_length1, p.window_length2) == (3, 6) and abs(p.threshold2 - 9.0) < 1e-7


def test_no_cpu_fallback(example_prefix):
    """Without a CUDA device every compute entry point must fail (never silently use a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import uncalled_b200 as U
    with pytest.raises(U.UncError):
        U.Index(example_prefix)
