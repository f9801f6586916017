"""
``WakeWordMLPModel`` -- B200 replacement of the reference's default classifier
(``heybuddy/wakeword.py:171-348``; ``DEFAULT_ARCHITECTURE = "perceptron"``, constants.py:73).

Same constructor defaults, ``from_file`` / ``state_dict`` / ``load_state_dict`` with the reference's
parameter names (== the ONNX initializer names of ``src/ts/models/*.onnx``), ``__call__(x[B,16,96]) ->
[B,1]`` probabilities, ``predict(audio)`` through ``SpeechEmbeddings``.  Forward, backward and Adam run
in the hand-written kernels of ``csrc/classifier.cu`` through ``hb_mlp_*``; there is no torch.nn graph.

Only the default architecture (layer_dim 96, 2 layers, gating, no half layers) is built: the
transformer classifier and the half-layer variant are outside the north-star path.
Dropout(0.1) on the input is train-only in the reference (wakeword.py:197,338): ``train_step(dropout=p)`` / ``apply_dropout`` draw
it on the device and ``WakeWordTrainer`` uses it by default; ``dropout=0`` is the parity configuration.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from heybuddy_b200 import _native, spec
from heybuddy_b200.constants import *  # noqa: F401,F403
from heybuddy_b200.util import audio_to_bct_tensor

__all__ = ["WakeWordMLPModel", "MultiWakeWordModel", "pack_classifier_params", "unpack_classifier_params"]


def pack_classifier_params(params: Dict[str, Any]) -> np.ndarray:
    parts = []
    for name, shape in spec.classifier_param_shapes():
        v = params[name]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        assert tuple(v.shape) == shape, (name, v.shape, shape)
        parts.append(np.ascontiguousarray(v, dtype=np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts))


def unpack_classifier_params(flat: np.ndarray) -> Dict[str, np.ndarray]:
    out, off = {}, 0
    for name, shape in spec.classifier_param_shapes():
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape).copy()
        off += n
    assert off == flat.size
    return out


class WakeWordMLPModel:
    """The default wake-word classifier: (batch, 16, 96) speech embeddings -> probability."""

    def __init__(
        self,
        input_shape: Tuple[int, int] = (16, 96),
        layer_dim: int = DEFAULT_LAYER_DIM,
        num_layers: int = DEFAULT_LAYERS,
        use_gating: bool = DEFAULT_USE_GATING,
        use_half_layers: bool = DEFAULT_USE_HALF_LAYERS,
        dropout: float = 0.1,
        activation: Optional[str] = "silu",
        device_id: Optional[int] = None,
        seed: int = 5002,
    ) -> None:
        if (tuple(input_shape), layer_dim, num_layers, use_gating, use_half_layers, activation) != ((16, 96), 96, 2, True, False, "silu"):
            raise NotImplementedError("only the default perceptron architecture (16x96 -> 96, 2 layers, gated, silu) is built")
        self.input_shape = tuple(input_shape)
        self.input_features = input_shape[0] * input_shape[1]
        self.dropout = dropout
        self.device_id = device_id
        self.training = False
        self._handle = None
        self._workspace = None
        self._host_params = pack_classifier_params(spec.init_classifier_weights(seed))

    # -- device plumbing ---------------------------------------------------------------------------
    @property
    def device(self):
        return _native.require_cuda(self.device_id)

    def _ensure(self):
        if self._handle is None:
            import torch

            lib = _native.load()
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _native.check(lib.hb_mlp_create(ctypes.byref(h), self._host_params.ctypes.data, self._host_params.size), "hb_mlp_create")
            self._handle = h
        return self._handle

    def _ws(self, nbytes: int):
        import torch

        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._workspace

    def __del__(self):
        try:
            if self._handle is not None:
                _native.load().hb_mlp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # -- torch.nn.Module-like surface ------------------------------------------------------------------
    def to(self, device=None, **_: Any) -> "WakeWordMLPModel":
        import torch

        if device is not None:
            d = torch.device(device)
            if d.type != "cuda":
                raise _native.NativeError("heybuddy_b200 has no CPU path")
            self.device_id = d.index
        return self

    def best(self) -> "WakeWordMLPModel":
        return self

    def eval(self) -> "WakeWordMLPModel":
        self.training = False
        return self

    def train(self, mode: bool = True) -> "WakeWordMLPModel":
        self.training = mode
        return self

    def _flat_params(self) -> np.ndarray:
        if self._handle is None:
            return self._host_params.copy()
        import torch

        flat = np.empty_like(self._host_params)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _native.check(_native.load().hb_mlp_get_params(self._handle, flat.ctypes.data, flat.size), "hb_mlp_get_params")
        return flat

    def state_dict(self) -> Dict[str, Any]:
        import torch

        return {k: torch.from_numpy(v) for k, v in unpack_classifier_params(self._flat_params()).items()}

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True) -> None:
        names = [n for n, _ in spec.classifier_param_shapes()]
        if strict and set(state_dict.keys()) != set(names):
            raise KeyError(f"state_dict keys differ: missing {set(names) - set(state_dict)}, unexpected {set(state_dict) - set(names)}")
        cur = unpack_classifier_params(self._flat_params())
        cur.update({k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state_dict.items() if k in cur})
        self._host_params = pack_classifier_params(cur)
        if self._handle is not None:
            import torch

            with torch.cuda.device(self.device):
                _native.check(_native.load().hb_mlp_set_params(self._handle, self._host_params.ctypes.data, self._host_params.size),
                              "hb_mlp_set_params")

    def gradients(self) -> Dict[str, np.ndarray]:
        """Parameter gradients of the last ``train_step`` (parity hook)."""
        import torch

        flat = np.empty_like(self._host_params)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _native.check(_native.load().hb_mlp_get_grads(self._ensure(), flat.ctypes.data, flat.size), "hb_mlp_get_grads")
        return unpack_classifier_params(flat)

    @classmethod
    def from_file(cls, path: str, device: Optional[Any] = None) -> "WakeWordMLPModel":
        """Loads a reference ``.pt`` state dict (torch.save(state_dict), trainer.py:186-198) or an ``.npz``."""
        if path.endswith(".npz"):
            with np.load(path) as z:
                sd = {k: z[k] for k in z.files}
        else:
            import torch

            sd = torch.load(path, weights_only=True, map_location="cpu")
        model = cls()
        model.load_state_dict(sd)
        if device is not None:
            model.to(device)
        return model

    def save_onnx(self, path: str, opset_version: int = 19) -> None:
        raise NotImplementedError("ONNX export belongs to the browser runtime, outside the B200 hot path (SURVEY.md 2 #21)")

    # -- forward -----------------------------------------------------------------------------------------
    def forward_device(self, x, out=None):
        """cuda f32 ``[B,16,96]`` (or ``[B,1536]``) -> cuda f32 ``[B,1]`` probabilities."""
        import torch

        x = x.reshape(x.shape[0], -1).contiguous()
        assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == self.input_features
        b = x.shape[0]
        if out is None:
            out = torch.empty((b, 1), dtype=torch.float32, device=x.device)
        lib = _native.load()
        with torch.cuda.device(x.device):
            nbytes = lib.hb_mlp_workspace_bytes(b, 0)
            ws = self._ws(nbytes)
            _native.check(lib.hb_mlp_forward(self._ensure(), x.data_ptr(), out.data_ptr(), b, ws.data_ptr(), ws.numel(),
                                             _native.stream_ptr(x.device)), "hb_mlp_forward")
        return out

    def __call__(self, x):
        import torch

        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return self.forward_device(x.to(self.device, dtype=torch.float32))

    forward = __call__

    def set_loss_scale(self, scale: float) -> None:
        """Loss / gradient multiplier of the following training steps (the reference's ``1 / accumulation_steps``, trainer.py:441)."""
        if getattr(self, "_loss_scale", 1.0) != float(scale):
            import torch

            with torch.cuda.device(self.device):
                _native.check(_native.load().hb_mlp_set_loss_scale(self._ensure(), float(scale)), "hb_mlp_set_loss_scale")
            self._loss_scale = float(scale)

    def apply_dropout(self, x, p: float, seed: int = 0):
        """
        Train-mode ``nn.Dropout(p)`` on the classifier input (wakeword.py:197,338): a dropped copy of ``x`` drawn on the device
        (``hb_mlp_dropout``: Philox4x32-10, key = seed, one counter block per four elements, call index = number of calls so far).
        ``p == 0`` returns ``x`` itself.
        """
        import torch

        if not p:
            return x
        x = x.reshape(x.shape[0], -1).contiguous()
        out = torch.empty_like(x)
        call = getattr(self, "_dropout_calls", 0)
        self._dropout_calls = call + 1
        with torch.cuda.device(x.device):
            _native.check(_native.load().hb_mlp_dropout(x.data_ptr(), out.data_ptr(), x.numel(), float(p), int(seed) & (2 ** 64 - 1), call,
                                                        _native.stream_ptr(x.device)), "hb_mlp_dropout")
        return out

    def optimizer_state_dict(self, lr: float = DEFAULT_LEARNING_RATE) -> Dict[str, Any]:
        """The fused step's Adam state as a ``torch.optim.Adam(model.parameters()).state_dict()`` (what trainer.py:197-198 saves)."""
        import torch

        n = self._host_params.size
        m, v, step = np.empty(n, np.float32), np.empty(n, np.float32), ctypes.c_int(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _native.check(_native.load().hb_mlp_get_adam(self._ensure(), m.ctypes.data, v.ctypes.data, ctypes.byref(step), n), "hb_mlp_get_adam")
        ms, vs = unpack_classifier_params(m), unpack_classifier_params(v)
        names = [name for name, _ in spec.classifier_param_shapes()]
        state = {} if step.value == 0 else {
            i: {"step": torch.tensor(float(step.value)), "exp_avg": torch.from_numpy(ms[k]), "exp_avg_sq": torch.from_numpy(vs[k])}
            for i, k in enumerate(names)}
        group = {"lr": float(lr), "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: Dict[str, Any]) -> None:
        """Inverse of :meth:`optimizer_state_dict`; accepts the reference's ``<name>_optimizer.pt`` (parameters in registration order)."""
        import torch

        names = [name for name, _ in spec.classifier_param_shapes()]
        state = sd.get("state", {})
        zeros = unpack_classifier_params(np.zeros_like(self._host_params))
        ms, vs, step = dict(zeros), {k: a.copy() for k, a in zeros.items()}, 0
        for i, k in enumerate(names):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            ms[k] = st["exp_avg"].detach().cpu().numpy().astype(np.float32)
            vs[k] = st["exp_avg_sq"].detach().cpu().numpy().astype(np.float32)
            step = max(step, int(float(st["step"])))
        m, v = pack_classifier_params(ms), pack_classifier_params(vs)
        with torch.cuda.device(self.device):
            _native.check(_native.load().hb_mlp_set_adam(self._ensure(), m.ctypes.data, v.ctypes.data, step, m.size), "hb_mlp_set_adam")

    def train_step(self, x, y, lr: float, negative_weight: float = 1.0, high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD,
                   min_selected: int = 128, dropout: float = 0.0, dropout_seed: int = 0):
        """
        One reference training step (trainer.py:405-462) fused on the device: forward, high-loss selection,
        weighted BCE (mean over the selected rows), backward and -- when at least ``min_selected`` rows were
        selected -- Adam.  ``dropout`` > 0 applies the train-mode input dropout first (the trainer does; parity tests do not).
        Returns (probabilities cuda [B,1], stats cuda f32[4] = loss, n_selected, stepped, high_loss_rate).
        """
        import torch

        x = self.apply_dropout(x, dropout, dropout_seed)
        x = x.reshape(x.shape[0], -1).contiguous()
        assert x.is_cuda and x.dtype == torch.float32 and y.is_cuda and y.dtype == torch.int64
        b = x.shape[0]
        prob = torch.empty((b, 1), dtype=torch.float32, device=x.device)
        stats = torch.empty(4, dtype=torch.float32, device=x.device)
        lib = _native.load()
        with torch.cuda.device(x.device):
            nbytes = lib.hb_mlp_workspace_bytes(b, 1)
            ws = self._ws(nbytes)
            _native.check(lib.hb_mlp_train_step(self._ensure(), x.data_ptr(), y.data_ptr(), b, float(lr), float(negative_weight),
                                                float(high_loss_threshold), int(min_selected), prob.data_ptr(), stats.data_ptr(),
                                                ws.data_ptr(), ws.numel(), _native.stream_ptr(x.device)), "hb_mlp_train_step")
        return prob, stats

    # -- the same step split for data-parallel training (heybuddy_b200/dp.py drives these) ----------------------------
    def dp_select(self, x, y, high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD):
        """Forward + high-loss selection of this rank's shard -> (prob cuda [b,1], stats cuda f32[4]; stats[1] = rows selected)."""
        import torch

        x = x.reshape(x.shape[0], -1).contiguous()
        assert x.is_cuda and x.dtype == torch.float32 and y.is_cuda and y.dtype == torch.int64
        b = x.shape[0]
        prob = torch.empty((b, 1), dtype=torch.float32, device=x.device)
        stats = torch.empty(4, dtype=torch.float32, device=x.device)
        lib = _native.load()
        with torch.cuda.device(x.device):
            ws = self._ws(lib.hb_mlp_workspace_bytes(b, 1))
            _native.check(lib.hb_mlp_select(self._ensure(), x.data_ptr(), y.data_ptr(), b, float(high_loss_threshold), prob.data_ptr(),
                                            stats.data_ptr(), ws.data_ptr(), ws.numel(), _native.stream_ptr(x.device)), "hb_mlp_select")
        self._dp = (x, y, prob, stats, ws)     # the workspace holds the forward activations until dp_backward
        return prob, stats

    def dp_backward(self, n_selected_total, negative_weight: float = 1.0, high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD,
                    min_selected: int = 128):
        """Gradients of (weighted BCE of this shard's selected rows) / n_selected_total -> the model's gradient buffer."""
        import torch

        x, y, prob, stats, ws = self._dp
        assert n_selected_total.is_cuda and n_selected_total.dtype == torch.float32 and n_selected_total.numel() == 1
        lib = _native.load()
        with torch.cuda.device(x.device):
            _native.check(lib.hb_mlp_backward(self._ensure(), x.data_ptr(), y.data_ptr(), x.shape[0], float(negative_weight),
                                              float(high_loss_threshold), n_selected_total.data_ptr(), int(min_selected), prob.data_ptr(),
                                              stats.data_ptr(), ws.data_ptr(), ws.numel(), _native.stream_ptr(x.device)), "hb_mlp_backward")
        return stats

    def dp_grads(self, buf=None, to_model: bool = False):
        """Copy the packed gradients model -> ``buf`` (a cuda f32 [n_params] tensor, allocated when None) or back."""
        import torch

        lib = _native.load()
        n = int(lib.hb_mlp_num_params())
        if buf is None:
            buf = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(lib.hb_mlp_grads_copy(self._ensure(), buf.data_ptr(), n, int(to_model), _native.stream_ptr(self.device)),
                          "hb_mlp_grads_copy")
        return buf

    def dp_local_step(self, x, y, negative_weight: float = 1.0, high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD):
        """
        First half of the one-collective data-parallel step: forward, selection, unnormalised loss sum and gradients of this shard.
        Returns (prob [b,1], stats f32[4], exchange f32 [n_params + 2] = {gradients | loss sum | rows selected}); all-reduce (SUM)
        ``exchange`` over the ranks, then ``dp_apply``.
        """
        import torch

        x = x.reshape(x.shape[0], -1).contiguous()
        assert x.is_cuda and x.dtype == torch.float32 and y.is_cuda and y.dtype == torch.int64
        b = x.shape[0]
        lib = _native.load()
        prob = torch.empty((b, 1), dtype=torch.float32, device=x.device)
        stats = torch.empty(4, dtype=torch.float32, device=x.device)
        if getattr(self, "_exchange", None) is None or self._exchange.device != x.device:
            self._exchange = torch.empty(int(lib.hb_mlp_num_params()) + 2, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = self._ws(lib.hb_mlp_workspace_bytes(b, 1))
            _native.check(lib.hb_mlp_local_step(self._ensure(), x.data_ptr(), y.data_ptr(), b, float(negative_weight), float(high_loss_threshold),
                                                prob.data_ptr(), stats.data_ptr(), self._exchange.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _native.stream_ptr(x.device)), "hb_mlp_local_step")
        return prob, stats, self._exchange

    def dp_apply(self, exchange, lr: float, min_selected: int, stats):
        """Second half: divide by the global count, fill stats[0..2] with the global loss / count / stepped flag, Adam."""
        import torch

        lib = _native.load()
        with torch.cuda.device(self.device):
            _native.check(lib.hb_mlp_apply_exchange(self._ensure(), exchange.data_ptr(), float(lr), int(min_selected), stats.data_ptr(),
                                                    _native.stream_ptr(self.device)), "hb_mlp_apply_exchange")
        return stats

    def dp_adam(self, lr: float, stats):
        """Adam on the (all-reduced) gradient buffer; skipped on the device when stats[2] == 0."""
        import torch

        lib = _native.load()
        with torch.cuda.device(self.device):
            _native.check(lib.hb_mlp_adam(self._ensure(), float(lr), stats.data_ptr(), _native.stream_ptr(self.device)), "hb_mlp_adam")

    # -- inference mixin (wakeword.py:129-169) ---------------------------------------------------------------
    @property
    def speech_embeddings(self):
        from heybuddy_b200.embeddings import get_speech_embeddings

        if not hasattr(self, "_speech_embeddings"):
            self._speech_embeddings = get_speech_embeddings(device_id=self.device.index or 0)
        return self._speech_embeddings

    def predict(self, audio: Any, threshold: float = 0.5, embedding_spectrogram_batch_size: int = 32, embedding_batch_size: int = 32,
                return_scores: bool = False, min_frames: int = 23040) -> Union[Tuple[bool, ...], Tuple[float, ...]]:
        """Predicts on one or more audio clips (centre-pads to ``min_frames``, wakeword.py:141-156)."""
        import torch

        audio_tensor, _ = audio_to_bct_tensor(audio, sample_rate=16000)
        n, c, t = audio_tensor.shape
        if t < min_frames:
            pad = min_frames - t
            left = int(pad / 2)
            audio_tensor = torch.cat([torch.zeros(n, c, left, dtype=audio_tensor.dtype), audio_tensor,
                                      torch.zeros(n, c, pad - left, dtype=audio_tensor.dtype)], dim=-1)
        emb = self.speech_embeddings(audio_tensor, embedding_batch_size=embedding_batch_size,
                                     spectrogram_batch_size=embedding_spectrogram_batch_size)
        preds = self(emb).cpu().numpy()
        return tuple(preds.flatten()) if return_scores else tuple((preds > threshold).flatten())


class MultiWakeWordModel:
    """
    BASELINE config 5: N wake-word models evaluated on the same rolling ``[16, 96]`` embedding buffer
    (browser semantics, src/ts/src/hey-buddy.ts:350-413; the Python reference runs one thread per model and
    re-featurizes per model, util/model_util.py:62-93).  Featurize once, then ``hb_mlp_forward_multi``: ONE stacked
    ``[M*128, 1536]`` first-layer GEMM for all the models (norm_in folded into the weights) + the 96-wide remainder batched over
    the models -- a launch chain whose length does not depend on M.
    """

    def __init__(self, models: List[WakeWordMLPModel]):
        assert models
        self.models = models
        self._ws = None

    def __call__(self, x):
        """cuda f32 ``[B,16,96]`` -> cuda f32 ``[M, B]``."""
        import torch

        x = x.reshape(x.shape[0], -1).contiguous()
        b, m = x.shape[0], len(self.models)
        out = torch.empty((m, b), dtype=torch.float32, device=x.device)
        lib = _native.load()
        handles = (ctypes.c_void_p * m)(*[mod._ensure() for mod in self.models])
        with torch.cuda.device(x.device):
            nbytes = lib.hb_mlp_multi_workspace_bytes(m, b)
            if self._ws is None or self._ws.numel() < nbytes:
                self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=x.device)
            _native.check(lib.hb_mlp_forward_multi(handles, m, x.data_ptr(), out.data_ptr(), b, self._ws.data_ptr(), self._ws.numel(),
                                                   _native.stream_ptr(x.device)), "hb_mlp_forward_multi")
        return out
