"""
Oracle: wake-word gated-MLP classifier (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``WakeWordMLPModel.forward``
(/root/reference/src/python/heybuddy/wakeword.py:334-348) with
``GatedMultiLayerPerceptron.forward`` (modules/multi_layer_perceptron.py:115-124):

    x[B,16,96] -> flatten -> LN(1536) -> gated(1536->64->96)
      -> 2 x [ LN(96) -> gated(96->64->96) ]        (no residual)
      -> LN(96) -> gated(96->64->1) -> sigmoid
    gated(x) = W_out( silu(W_hidden x + b_h) * (W_gate x + b_g) ) + b_out

and the training-step loss of ``WakeWordTrainer.train_epoch``
(trainer.py:405-462): high-loss selection, weighted BCE (mean over the selected
rows), skip below 128 selected rows.  Dropout(0.1) on the input is train-only
and disabled for parity.

PINNED: ``tests/golden/classifier_*.npz`` hold outputs of the reference's own class
(run from a scratch copy) on in-repo trained weights (src/ts/models/*.onnx).

``forward`` is numpy (float64 by default).  ``forward_backward_torch`` builds the same
graph from torch CPU primitives and lets autograd produce the gradient oracle.
"""
from __future__ import annotations

import struct
from typing import Dict, Optional, Tuple

import numpy as np

from heybuddy_b200 import spec


# ----------------------------------------------------------------------------
# forward (numpy)
# ----------------------------------------------------------------------------
def _layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + spec.LN_EPS) * w + b


def _gated(x: np.ndarray, p: Dict[str, np.ndarray], prefix: str) -> np.ndarray:
    h = x @ p[f"{prefix}.hidden.weight"].T + p[f"{prefix}.hidden.bias"]
    g = x @ p[f"{prefix}.gate.weight"].T + p[f"{prefix}.gate.bias"]
    a = h / (1.0 + np.exp(-h)) * g
    return a @ p[f"{prefix}.output.weight"].T + p[f"{prefix}.output.bias"]


def forward(x: np.ndarray, params: Dict[str, np.ndarray], dtype=np.float64, return_logits: bool = False) -> np.ndarray:
    """``x [B,16,96] -> p [B,1]`` (or the pre-sigmoid logit)."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    s = np.asarray(x, dtype=dtype).reshape(x.shape[0], -1)
    s = _gated(_layer_norm(s, p["norm_in.weight"], p["norm_in.bias"]), p, "mlp_in")
    l = 0
    while f"layers.{l}.0.weight" in p:
        s = _gated(_layer_norm(s, p[f"layers.{l}.0.weight"], p[f"layers.{l}.0.bias"]), p, f"layers.{l}.1")
        l += 1
    z = _gated(_layer_norm(s, p["norm_out.weight"], p["norm_out.bias"]), p, "mlp_out")
    if return_logits:
        return z
    return 1.0 / (1.0 + np.exp(-z))


# ----------------------------------------------------------------------------
# training-step loss + gradients (torch autograd as the checker)
# ----------------------------------------------------------------------------
def forward_backward_torch(
    x: np.ndarray,
    y: np.ndarray,
    params: Dict[str, np.ndarray],
    negative_weight: float = 1.0,
    high_loss_threshold: float = 1e-4,
    dtype=None,
):
    """
    One reference training step's loss and parameter gradients (trainer.py:405-462):
    returns (probabilities[B,1], loss, n_selected, {name: grad}).  The loss is the
    weighted BCE *mean over the selected rows*; rows failing the high-loss test
    contribute nothing.
    """
    import torch
    import torch.nn.functional as F

    dtype = dtype or torch.float64
    P = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in params.items()}

    def gated(t, prefix):
        h = F.linear(t, P[f"{prefix}.hidden.weight"], P[f"{prefix}.hidden.bias"])
        g = F.linear(t, P[f"{prefix}.gate.weight"], P[f"{prefix}.gate.bias"])
        return F.linear(F.silu(h) * g, P[f"{prefix}.output.weight"], P[f"{prefix}.output.bias"])

    xt = torch.tensor(np.asarray(x), dtype=dtype).reshape(x.shape[0], -1)
    s = gated(F.layer_norm(xt, (xt.shape[1],), P["norm_in.weight"], P["norm_in.bias"], spec.LN_EPS), "mlp_in")
    l = 0
    while f"layers.{l}.0.weight" in P:
        s = gated(F.layer_norm(s, (s.shape[1],), P[f"layers.{l}.0.weight"], P[f"layers.{l}.0.bias"], spec.LN_EPS), f"layers.{l}.1")
        l += 1
    z = gated(F.layer_norm(s, (s.shape[1],), P["norm_out.weight"], P["norm_out.bias"], spec.LN_EPS), "mlp_out")
    prob = torch.sigmoid(z)

    yt = torch.tensor(np.asarray(y)).to(torch.int64)
    ps = prob.squeeze(1)
    sel = ((yt == 0) & (ps >= high_loss_threshold)) | ((yt == 1) & (ps < 1 - high_loss_threshold))
    n_sel = int(sel.sum())
    grads = {k: np.zeros(v.shape, dtype=np.float64) for k, v in P.items()}
    loss_val = 0.0
    if n_sel > 0:
        yp = prob[sel]
        yy = yt[sel].to(dtype).unsqueeze(1)
        w = torch.where(yy == 1, torch.ones_like(yy), torch.full_like(yy, negative_weight))
        loss = F.binary_cross_entropy(yp, yy, w)
        loss.backward()
        loss_val = float(loss.detach())
        grads = {k: v.grad.detach().numpy().astype(np.float64) for k, v in P.items()}
    return prob.detach().numpy(), loss_val, n_sel, grads


def learning_rate(step, warmup_steps=0, hold_steps=0, total_steps=0, target_learning_rate=1e-3) -> float:
    """trainer.py:127-156 (cosine decay with linear warm-up and hold)."""
    lr = 0.5 * target_learning_rate * (1 + np.cos(np.pi * (step - warmup_steps - hold_steps) / float(total_steps - warmup_steps - hold_steps)))
    warm = target_learning_rate * (step / warmup_steps) if warmup_steps > 0 else 0.0
    if hold_steps > 0:
        lr = lr if step > warmup_steps + hold_steps else target_learning_rate
    return float(warm if step < warmup_steps else lr)


# ----------------------------------------------------------------------------
# ONNX initializer reader (protobuf wire walk; no ``onnx`` package in this image)
# ----------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield num, wt, val


def read_onnx_initializers(path: str) -> Dict[str, np.ndarray]:
    """
    ModelProto.graph = 7; GraphProto.initializer = 5; TensorProto: dims = 1,
    data_type = 2 (1 = float), float_data = 4, name = 8, raw_data = 9.
    """
    with open(path, "rb") as fh:
        model = fh.read()
    out: Dict[str, np.ndarray] = {}
    for num, wt, graph in _fields(model):
        if num != 7 or wt != 2:
            continue
        for gnum, gwt, tensor in _fields(graph):
            if gnum != 5 or gwt != 2:
                continue
            dims, name, raw, dtype_id, floats = [], None, None, None, []
            for tnum, twt, val in _fields(tensor):
                if tnum == 1:
                    if twt == 0:
                        dims.append(val)
                    else:  # packed
                        p = 0
                        while p < len(val):
                            d, p = _varint(val, p)
                            dims.append(d)
                elif tnum == 2:
                    dtype_id = val
                elif tnum == 4:
                    if twt == 2:
                        floats.extend(struct.unpack(f"<{len(val) // 4}f", val))
                    else:
                        floats.append(struct.unpack("<f", val)[0])
                elif tnum == 8:
                    name = val.decode()
                elif tnum == 9:
                    raw = val
            if name is None or dtype_id != 1:
                continue
            arr = np.frombuffer(raw, dtype="<f4") if raw is not None else np.asarray(floats, dtype=np.float32)
            out[name] = arr.reshape(dims).astype(np.float32).copy()
    return out
