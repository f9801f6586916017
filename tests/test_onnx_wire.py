"""CPU: the re-pin hook (heybuddy_b200/util/onnx_wire.py) reads ONNX files by walking the protobuf wire format."""
import os

import numpy as np
import pytest

from heybuddy_b200 import spec
from heybuddy_b200.util import onnx_wire


def _varint(v):
    out = b""
    while True:
        b = v & 0x7F
        v >>= 7
        out += bytes([b | (0x80 if v else 0)])
        if not v:
            return out


def _ld(num, payload):          # length-delimited field
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _tensor(name, arr):
    body = b"".join(_varint((1 << 3) | 0) + _varint(d) for d in arr.shape) + _varint((2 << 3) | 0) + _varint(1)
    return body + _ld(8, name.encode()) + _ld(9, np.ascontiguousarray(arr, dtype="<f4").tobytes())


def _conv_node(i, w_name, b_name):
    body = _ld(1, f"x{i}".encode()) + _ld(1, w_name.encode()) + _ld(1, b_name.encode()) + _ld(2, f"x{i + 1}".encode())
    attr = _ld(1, b"kernel_shape") + _ld(8, _varint(1) + _varint(3))
    return body + _ld(3, f"conv_{i}".encode()) + _ld(4, b"Conv") + _ld(5, attr)


def _model(layers, weights):
    graph = b""
    for i, (name, kh, kw, cin, cout, *_r) in enumerate(layers):
        w = weights[f"{name}.weight"].transpose(3, 2, 0, 1)          # HWIO -> ONNX OIHW
        graph += _ld(1, _conv_node(i, f"W{i}", f"B{i}")) + _ld(1, _ld(4, b"LeakyRelu"))
        graph += _ld(5, _tensor(f"W{i}", w)) + _ld(5, _tensor(f"B{i}", weights[f"{name}.bias"]))
    return _varint((1 << 3) | 0) + _varint(8) + _ld(7, graph)


def test_speech_embedding_onnx_round_trip(tmp_path):
    weights = spec.init_embedding_weights(seed=5)
    path = tmp_path / "speech-embedding.onnx"
    path.write_bytes(_model(spec.EMBEDDING_LAYERS, weights))
    with pytest.raises(ValueError, match="sha256"):
        onnx_wire.speech_embedding_weights_from_onnx(str(path))                       # not the reference's pinned artefact
    got = onnx_wire.speech_embedding_weights_from_onnx(str(path), check_sha256=False)
    assert sorted(got) == sorted(weights)
    for k in weights:
        np.testing.assert_array_equal(got[k], weights[k])
    model = onnx_wire.read_onnx(str(path))
    assert [n["op_type"] for n in model["nodes"]][:2] == ["Conv", "LeakyRelu"] and model["nodes"][0]["attrs"]["kernel_shape"] == [1, 3]
    # a file whose conv stack differs from the proposed table says what the table should be
    other = [list(l) for l in spec.EMBEDDING_LAYERS]
    other[4][4] = 64
    other[5][3] = 64
    w2 = dict(weights)
    w2["conv2d_4.weight"] = np.zeros((1, 3, 24, 64), np.float32)
    w2["conv2d_4.bias"] = np.zeros(64, np.float32)
    w2["conv2d_5.weight"] = np.zeros((3, 1, 64, 48), np.float32)
    path.write_bytes(_model([tuple(l) for l in other], w2))
    with pytest.raises(ValueError, match=r"conv  4: file \(kh, kw, cin, cout\) = \(1, 3, 24, 64\)"):
        onnx_wire.speech_embedding_weights_from_onnx(str(path), check_sha256=False)
    # the CLI writes an .npz the product's weights= argument takes
    path.write_bytes(_model(spec.EMBEDDING_LAYERS, weights))
    out = tmp_path / "w.npz"
    assert onnx_wire._main(["embed", str(path), str(out), "--any-sha"]) == 0
    with np.load(out) as z:
        np.testing.assert_array_equal(z["conv2d_19.weight"], weights["conv2d_19.weight"])


def test_mel_tables_report(tmp_path):
    fb = spec.mel_filterbank()
    win = spec.hann_window_padded().astype(np.float64)
    k = np.arange(257)[:, None] * np.arange(512)[None, :]
    basis = np.concatenate([np.cos(2 * np.pi * k / 512) * win, -np.sin(2 * np.pi * k / 512) * win]).astype(np.float32)
    graph = _ld(5, _tensor("mel", fb.T.copy())) + _ld(5, _tensor("dft", basis[:, None, :]))
    path = tmp_path / "mel.onnx"
    path.write_bytes(_ld(7, graph))
    rep = onnx_wire.mel_tables_from_onnx(str(path), check_sha256=False)
    assert rep["sha256_ok"] is False and rep["mel_matrix_max_abs_diff"] == 0.0 and rep["dft_basis_max_abs_diff"] < 1e-6


def test_reads_the_reference_classifier_files(golden_dir):
    """The same walker on a real ONNX export: the reference's in-repo hey-buddy.onnx == the fixture made from it (build container only)."""
    path = "/root/reference/src/ts/models/hey-buddy.onnx"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    model = onnx_wire.read_onnx(path)
    g = np.load(os.path.join(golden_dir, "classifier_hey_buddy.npz"))
    names = [n for n, _ in spec.classifier_param_shapes()]
    assert set(names) <= set(model["initializers"])
    for n in names:
        np.testing.assert_array_equal(model["initializers"][n], g[f"param::{n}"])
    assert any(n["op_type"] in ("Gemm", "MatMul") for n in model["nodes"])
