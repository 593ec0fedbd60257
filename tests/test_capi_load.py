"""CPU-side checks of the C-ABI boundary: the library loads without a GPU, exports every symbol
include/ntn_b200.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import os
import re

import pytest

from neural_tensor_network_for_kb_completion_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "ntn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ntn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_phase_names(lib):
    assert b"sm_100a" in lib.ntn_version()
    names = [lib.ntn_phase_name(i).decode() for i in range(_capi.NUM_PHASES)]
    assert names[0] == "entity_build" and names[-1] == "finalize" and len(set(names)) == _capi.NUM_PHASES


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.NTNError) as ei:
        _capi.Handle(6, 3, 2, 10, 10, 5, 1e-4)
    assert ei.value.code == _capi.NTN_ECUDA and "no CPU fallback" in str(ei.value)


def test_bad_shapes_rejected_before_touching_the_device(lib):
    with pytest.raises(_capi.NTNError) as ei:
        _capi.Handle(300, 3, 2, 10, 10, 5, 1e-4)
    assert ei.value.code == _capi.NTN_EUNSUP
    with pytest.raises(_capi.NTNError) as ei:
        _capi.Handle(10, 3, 0, 10, 10, 5, 1e-4)
    assert ei.value.code == _capi.NTN_EINVAL
