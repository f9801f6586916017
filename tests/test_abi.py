"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "heybuddy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("hb_mel_f32", "hb_embed_clips", "hb_embed_windows", "hb_augment_clips_f32", "hb_mlp_forward", "hb_mlp_train_step"):
        assert must in syms


def test_library_exports_every_declared_symbol(native_lib):
    for name in declared_symbols():
        assert hasattr(native_lib, name), f"{name} declared in include/heybuddy_b200.h but not exported"
    assert native_lib.hb_abi_version() == 1
    assert native_lib.hb_mel_frames(23040) == 141
    assert native_lib.hb_mel_frames(17280) == 105
    assert native_lib.hb_mel_frames(100) == 0
    assert native_lib.hb_embed_num_params() == 274440
    assert native_lib.hb_mlp_num_params() == 256417


def test_workspace_size_queries_are_host_only_and_consistent(native_lib):
    """Size queries need no device: a training workspace holds what the backward pass reads, so it exceeds the inference one; both grow
    with the batch; the multi-model one with the model count; bad arguments are errors, not sizes."""
    inf = [native_lib.hb_mlp_workspace_bytes(b, 0) for b in (1, 64, 4096)]
    trn = [native_lib.hb_mlp_workspace_bytes(b, 1) for b in (1, 64, 4096)]
    assert all(a > 0 for a in inf) and inf == sorted(inf) and trn == sorted(trn)
    assert all(t > i for t, i in zip(trn, inf))
    assert native_lib.hb_mlp_workspace_bytes(-1, 0) < 0
    multi = [native_lib.hb_mlp_multi_workspace_bytes(m, 256) for m in (1, 7, 64, 65)]
    assert all(a > 0 for a in multi) and multi == sorted(multi)


def test_binding_covers_header(native_lib):
    from heybuddy_b200 import _native

    bound = {n for n in declared_symbols() if getattr(native_lib, n).argtypes is not None}
    assert bound == set(declared_symbols())


def test_library_is_sm100a_only():
    from heybuddy_b200 import _native

    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", _native.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from heybuddy_b200 import _native
    from heybuddy_b200.embeddings import SpeechEmbeddings

    with pytest.raises(_native.NativeError):
        SpeechEmbeddings()(torch.zeros(23040))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under heybuddy_b200/ may import it."""
    pkg = os.path.join(ROOT, "heybuddy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
