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
