"""The C-ABI library loads and exports every symbol include/sleqp_b200.h declares; compute entry
points fail loudly (no CPU fallback) when there is no device."""
import os
import re

import numpy as np
import pytest
import torch

from sleqp_b200 import B200Error, Fact, Mat, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "sleqp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[a-z_]+)\s*\(", hdr)))
    L = _lib.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(_lib.SYMBOLS)


def test_stats_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "sleqp_b200.h")).read()
    body = hdr[hdr.index("typedef struct b200_stats"): hdr.index("} b200_stats;")]
    fields = re.findall(r"^\s*(?:u?int32_t|u?int64_t|double)\s+([a-z_A-Z0-9]+);", body, flags=re.M)
    assert fields == [f for f, _ in _lib.Stats._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    assert _lib.lib().b200_device_count() == 0
    with pytest.raises(B200Error) as e:
        Fact()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(B200Error):
        Mat()


def test_host_glue_is_c11_and_names_the_reference_interface():
    src = open(os.path.join(ROOT, "sleqp_b200", "host", "fact", "fact_b200.c")).read()
    for needle in ("sleqp_fact_create_default", "SLEQP_FACT_FLAGS_LOWER", ".set_matrix", ".solve", ".solution", ".condition", ".free", "sleqp_vec_set_from_raw"):
        assert needle in src
