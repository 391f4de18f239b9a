"""The C-ABI library loads on a CPU-only box and exports every symbol include/paule_b200.h declares.
No compute call is made here (no GPU)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "paule_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"PAULE_API\s+[\w\s\*]+?\b(paule_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("paule_plan_step", "paule_plan_forward", "paule_lstm_seq_fwd_f32", "paule_lstm_seq_bwd_f32",
                 "paule_plan_loss_f32", "paule_adam_clamp_f32", "paule_linear_f32", "paule_tc_gemm_img", "paule_tc_gemm_pack",
                 "paule_tc_lstm_seq_fwd", "paule_tc_lstm_seq_bwd", "paule_upsample_smooth_f32"):
        assert must in syms
    assert len(syms) >= 24


def test_library_exports_every_declared_symbol():
    from paule_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run ./build.sh (or __graft_entry__.build()) first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"
    # the ctypes signature table covers the whole header, and nothing else
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_struct_layout_matches_header_field_order():
    """ctypes mirror of struct paule_plan / paule_lstm_layer: same field names in the same order."""
    from paule_b200 import _lib
    src = open(HEADER).read()

    def fields(struct_name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct_name, struct_name), src, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[\w]+", part)[-1])
        return names

    assert fields("paule_lstm_layer") == [f[0] for f in _lib.LstmLayer._fields_]
    assert fields("paule_plan") == [f[0] for f in _lib.Plan._fields_]


def test_error_strings_and_no_device_is_loud():
    from paule_b200 import _lib
    lib = _lib.load()
    assert lib.paule_version() >= 100
    assert b"invalid argument" in lib.paule_error_string(1)
    import torch
    if not torch.cuda.is_available():
        # no GPU here: the product path must fail loudly, never fall back
        assert lib.paule_device_check() == 4
        with pytest.raises(_lib.PauleB200Error):
            _lib.require_device()


@pytest.mark.parametrize("backward", [0, 1])
def test_pass_plan_of_the_recurrent_kernels(backward):
    """paule_tc_rnn_pass_plan (host-side only): the passes cover every word exactly once, respect the launch capacity of their
    layout (6 / 5 word groups of 16 nq words) and never cost more summed step time than balanced passes of one layout."""
    from paule_b200 import _lib
    lib = _lib.load()
    groups = 5 if backward else 6
    cost = ([0, 2.87, 3.67, 5.00, 6.45] if backward else [0, 2.23, 3.64, 4.71, 5.79])
    nq, words = (ctypes.c_int32 * 96)(), (ctypes.c_int32 * 96)()
    assert lib.paule_tc_rnn_pass_plan(0, backward, nq, words, 96) == -1
    for B in (1, 16, 64, 96, 97, 256, 320, 321, 384, 385, 400, 512, 1000, 1024, 2048, 5000):
        n = lib.paule_tc_rnn_pass_plan(B, backward, nq, words, 96)
        assert 1 <= n <= 96
        assert sum(words[i] for i in range(n)) == B
        for i in range(n):
            assert 1 <= nq[i] <= 4 and 0 < words[i] <= groups * 16 * nq[i]
            if i + 1 < n:
                assert words[i] % 16 == 0, "only the last pass may end inside a word quarter"
        if B <= groups * 64:
            assert n == 1, "one launch whenever the batch fits one"
        else:
            full = groups * 64                       # balanced passes of the four-quarter layout (round 1's schedule)
            n_bal = -(-B // full)
            per = -(-B // n_bal)
            nq_bal = min(4, -(-per // (groups * 16)))
            assert sum(cost[nq[i]] for i in range(n)) <= n_bal * cost[nq_bal] + 1e-6
