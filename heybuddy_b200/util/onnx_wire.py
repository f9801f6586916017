"""
Re-pin hook: read the reference's two model artefacts -- ``speech-embedding.onnx``
(reference embeddings.py:29-30) and ``mel-spectrogram.onnx`` (spectrogram.py:20-21) -- WITHOUT the ``onnx`` package, by walking
the protobuf wire format, and turn them into what this package loads.

Neither file exists offline, so the embedding CNN's interior (``spec.EMBEDDING_LAYERS``) and the mel tables are restated /
proposed and every parity claim about their VALUES is "unpinned".  The day the files are at hand:

    python -m heybuddy_b200.util.onnx_wire embed speech-embedding.onnx weights.npz     # -> SpeechEmbeddings(weights="weights.npz")
    python -m heybuddy_b200.util.onnx_wire mel   mel-spectrogram.onnx                  # compares its tables with spec's
    python -m heybuddy_b200.util.onnx_wire show  any.onnx                              # nodes + initializer shapes

``embed`` checks the file's sha256 against the reference's pin, walks the graph's Conv nodes in order, converts every kernel
from ONNX's OIHW to the HWIO layout ``hb_embed_create`` packs, and compares the resulting layer table (kernel size, channels)
with ``spec.EMBEDDING_LAYERS``: equal -> an ``.npz`` the product loads; different -> it raises with the real table printed, which
is the edit ``spec.py`` needs (the kernels are generated from that table's shapes; csrc/embed_common.cuh mirrors it).

Field numbers (onnx.proto3): ModelProto.graph = 7; GraphProto.node = 1, .initializer = 5, .input = 11, .output = 12;
NodeProto.input = 1, .output = 2, .name = 3, .op_type = 4, .attribute = 5; AttributeProto.name = 1, .i = 3, .ints = 8, .s = 4;
TensorProto.dims = 1, .data_type = 2 (1 = float), .float_data = 4, .name = 8, .raw_data = 9.
"""
from __future__ import annotations

import hashlib
import struct
import sys
from typing import Any, Dict, Iterator, List, Tuple

import numpy as np

from heybuddy_b200 import spec

__all__ = ["read_onnx", "speech_embedding_weights_from_onnx", "mel_tables_from_onnx", "SPEECH_EMBEDDING_SHA256", "MEL_SPECTROGRAM_SHA256"]

SPEECH_EMBEDDING_SHA256 = "70d164290c1d095d1d4ee149bc5e00543250a7316b59f31d056cff7bd3075c1f"   # embeddings.py:30
MEL_SPECTROGRAM_SHA256 = "ba2b0e0f8b7b875369a2c89cb13360ff53bac436f2895cced9f479fa65eb176f"    # spectrogram.py:21


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, Any]]:
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _ints(wt: int, val: Any) -> List[int]:
    if wt == 0:
        return [val]
    out, p = [], 0
    while p < len(val):
        d, p = _varint(val, p)
        out.append(d)
    return out


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    name, raw, dtype_id, floats = "", None, None, []
    for num, wt, val in _fields(buf):
        if num == 1:
            dims += _ints(wt, val)
        elif num == 2:
            dtype_id = val
        elif num == 4:
            floats += list(struct.unpack(f"<{len(val) // 4}f", val)) if wt == 2 else [struct.unpack("<f", val)[0]]
        elif num == 8:
            name = val.decode()
        elif num == 9:
            raw = val
    if dtype_id != 1:
        return name, np.zeros(0, np.float32)
    arr = np.frombuffer(raw, dtype="<f4") if raw is not None else np.asarray(floats, dtype=np.float32)
    return name, arr.reshape(dims).astype(np.float32).copy()


def _node(buf: bytes) -> Dict[str, Any]:
    node: Dict[str, Any] = {"input": [], "output": [], "name": "", "op_type": "", "attrs": {}}
    for num, wt, val in _fields(buf):
        if num == 1:
            node["input"].append(val.decode())
        elif num == 2:
            node["output"].append(val.decode())
        elif num == 3:
            node["name"] = val.decode()
        elif num == 4:
            node["op_type"] = val.decode()
        elif num == 5:
            aname, ints, single, text = "", [], None, None
            for anum, awt, aval in _fields(val):
                if anum == 1:
                    aname = aval.decode()
                elif anum == 3:
                    single = aval
                elif anum == 8:
                    ints += _ints(awt, aval)
                elif anum == 4:
                    text = aval.decode(errors="replace")
            node["attrs"][aname] = ints if ints else (single if single is not None else text)
    return node


def read_onnx(path: str) -> Dict[str, Any]:
    """``{"initializers": {name: f32 array}, "nodes": [{op_type, name, input, output, attrs}], "sha256": hex}``."""
    with open(path, "rb") as fh:
        model = fh.read()
    out: Dict[str, Any] = {"initializers": {}, "nodes": [], "sha256": hashlib.sha256(model).hexdigest()}
    for num, wt, graph in _fields(model):
        if num != 7 or wt != 2:
            continue
        for gnum, gwt, val in _fields(graph):
            if gnum == 5 and gwt == 2:
                name, arr = _tensor(val)
                if arr.size:
                    out["initializers"][name] = arr
            elif gnum == 1 and gwt == 2:
                out["nodes"].append(_node(val))
    return out


def speech_embedding_weights_from_onnx(path: str, check_sha256: bool = True) -> Dict[str, np.ndarray]:
    """
    ``speech-embedding.onnx`` -> ``{"conv2d.weight": HWIO, "conv2d.bias": ..., ...}`` in ``spec.EMBEDDING_LAYERS`` naming, when the
    file's conv stack has the table's shapes; otherwise raises ``ValueError`` listing the file's real (kh, kw, cin, cout) per conv.
    """
    model = read_onnx(path)
    if check_sha256 and model["sha256"] != SPEECH_EMBEDDING_SHA256:
        raise ValueError(f"{path}: sha256 {model['sha256']} is not the reference's pinned artefact ({SPEECH_EMBEDDING_SHA256}); pass "
                         "check_sha256=False to convert it anyway")
    convs = [n for n in model["nodes"] if n["op_type"] == "Conv"]
    found: List[Tuple[int, int, int, int]] = []
    weights: Dict[str, np.ndarray] = {}
    for i, node in enumerate(convs):
        w = model["initializers"].get(node["input"][1])
        if w is None or w.ndim != 4:
            raise ValueError(f"conv {i} ({node['name']}): weight initializer {node['input'][1]!r} not found / not 4-D")
        cout, cin, kh, kw = w.shape
        found.append((kh, kw, cin, cout))
        name = spec.EMBEDDING_LAYERS[i][0] if i < len(spec.EMBEDDING_LAYERS) else f"conv2d_{i}"
        weights[f"{name}.weight"] = np.ascontiguousarray(w.transpose(2, 3, 1, 0))            # OIHW -> HWIO
        b = model["initializers"].get(node["input"][2]) if len(node["input"]) > 2 else None
        weights[f"{name}.bias"] = (b if b is not None else np.zeros(cout, np.float32)).astype(np.float32)
    want = [(kh, kw, cin, cout) for (_, kh, kw, cin, cout, *_r) in spec.EMBEDDING_LAYERS]
    if found != want:
        lines = [f"  conv {i:2d}: file (kh, kw, cin, cout) = {f}   spec = {want[i] if i < len(want) else None}" for i, f in enumerate(found)]
        raise ValueError(f"{path}: the file's conv stack ({len(found)} convs) differs from spec.EMBEDDING_LAYERS ({len(want)}); the table to "
                         "put into heybuddy_b200/spec.py (and csrc/embed_common.cuh) is:\n" + "\n".join(lines))
    return weights


def mel_tables_from_onnx(path: str, check_sha256: bool = True) -> Dict[str, Any]:
    """
    ``mel-spectrogram.onnx`` -> a report comparing its constant tensors with the restated front end: a ``[*, 32]`` / ``[32, *]``
    matrix against ``spec.mel_filterbank()`` and ``[*, 1, 512]``-like DFT-basis kernels against the windowed DFT of
    ``spec.hann_window_padded()``.  Returns ``{"sha256_ok", "tensors": {name: shape}, "mel_matrix_max_abs_diff", "dft_basis_max_abs_diff"}``
    (a diff is None when no tensor of that shape exists).
    """
    model = read_onnx(path)
    report: Dict[str, Any] = {"sha256_ok": model["sha256"] == MEL_SPECTROGRAM_SHA256, "tensors": {k: v.shape for k, v in model["initializers"].items()},
                              "mel_matrix_max_abs_diff": None, "dft_basis_max_abs_diff": None}
    if check_sha256 and not report["sha256_ok"]:
        raise ValueError(f"{path}: sha256 {model['sha256']} is not the reference's pinned artefact ({MEL_SPECTROGRAM_SHA256})")
    fb = spec.mel_filterbank()                                           # [257, 32]
    win = spec.hann_window_padded().astype(np.float64)
    k = np.arange(spec.N_FREQ)[:, None] * np.arange(spec.N_FFT)[None, :]
    basis = np.concatenate([np.cos(2 * np.pi * k / spec.N_FFT) * win, -np.sin(2 * np.pi * k / spec.N_FFT) * win])   # [514, 512]
    for arr in model["initializers"].values():
        a = np.squeeze(arr)
        if a.shape == fb.shape:
            report["mel_matrix_max_abs_diff"] = float(np.abs(a - fb).max())
        elif a.shape == fb.T.shape:
            report["mel_matrix_max_abs_diff"] = float(np.abs(a.T - fb).max())
        elif a.ndim == 2 and a.shape[1] == spec.N_FFT and a.shape[0] in (spec.N_FREQ, 2 * spec.N_FREQ):
            ref = basis[:a.shape[0]]
            report["dft_basis_max_abs_diff"] = float(min(np.abs(a - ref).max(), np.abs(np.abs(a) - np.abs(ref)).max()))
    return report


def _main(argv: List[str]) -> int:
    if len(argv) < 2 or argv[0] not in ("embed", "mel", "show"):
        print(__doc__)
        return 2
    if argv[0] == "embed":
        weights = speech_embedding_weights_from_onnx(argv[1], check_sha256="--any-sha" not in argv)
        out = argv[2] if len(argv) > 2 and not argv[2].startswith("--") else "speech_embedding_weights.npz"
        np.savez(out, **weights)
        print(f"{out}: {len(weights) // 2} conv layers; load with SpeechEmbeddings(weights={out!r}) or HEYBUDDY_B200_EMBED_WEIGHTS={out}")
    elif argv[0] == "mel":
        print(mel_tables_from_onnx(argv[1], check_sha256="--any-sha" not in argv))
    else:
        model = read_onnx(argv[1])
        print("sha256", model["sha256"])
        for n in model["nodes"]:
            print(n["op_type"], n["name"], n["input"], "->", n["output"], n["attrs"])
        for k, v in model["initializers"].items():
            print("init", k, v.shape)
    return 0


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
