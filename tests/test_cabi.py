"""The C-ABI library loads without a GPU and exports every symbol include/lpopc_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lpopc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpb_[a-z0-9_]+)\s*\(", src)))


def test_cabi_exports_every_declared_symbol():
    from lpopc_b200 import nlp
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(nlp.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes binding covers the same set
    assert sorted(nlp.SIGNATURES) == names


def test_functor_registry_lists_reference_examples():
    from lpopc_b200 import nlp
    lib = nlp.load_library()
    names = [lib.lpb_functor_name(i).decode() for i in range(lib.lpb_num_functors())]
    for ref_example in ("hypersensitive", "bryson_denham", "launch"):  # Lpopc/example/*
        assert ref_example in names
    assert lib.lpb_functor_name(len(names)) is None


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a usable device lpb_create returns LPB_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        return
    from lpopc_b200 import examples, nlp
    try:
        nlp.TranscribedNLP(examples.hypersensitive())
    except nlp.LpopcError as e:
        assert e.code == -3 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("TranscribedNLP must not construct without a CUDA device")
