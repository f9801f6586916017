"""
Ring-1 model callable: the B200 replacement of ``PretrainedONNXModel``
(reference util/onnx_util.py:12-96).

Same surface -- ``__init__(device_id=None, load=False)``, ``from_file``, ``load``,
``unload``, ``loaded``, ``device``, ``__call__(**named_inputs, retry=True) -> list[np.ndarray]``,
one automatic retry after ``unload()`` -- but ``__call__`` runs hand-written sm_100a kernels
through the C-ABI instead of an onnxruntime session.  ``device_id=None`` means "the current
CUDA device" (the reference's CPU provider has no equivalent here: there is no CPU path).
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from heybuddy_b200 import _native

__all__ = ["PretrainedNativeModel", "PretrainedONNXModel"]


class PretrainedNativeModel:
    input_name: str = "input"

    def __init__(self, device_id: Optional[int] = None, load: bool = False) -> None:
        self.loaded = False
        self.device_id = device_id
        self._pretrained_model_path: Optional[str] = None
        if load:
            self.load()

    @property
    def device(self):
        return _native.require_cuda(self.device_id)

    @property
    def pretrained_model_path(self) -> Optional[str]:
        return self._pretrained_model_path

    @classmethod
    def from_file(cls, pretrained_model_path: str, device_id: Optional[int] = None, load: bool = False):
        instance = cls(device_id=device_id, load=False)
        instance._pretrained_model_path = pretrained_model_path
        if load:
            instance.load()
        return instance

    # -- subclass hooks ------------------------------------------------------------------
    def _load(self) -> None:
        raise NotImplementedError

    def _unload(self) -> None:
        pass

    def _run(self, **inputs: np.ndarray) -> List[np.ndarray]:
        raise NotImplementedError

    # -- reference surface -----------------------------------------------------------------
    def load(self) -> None:
        if self.loaded:
            return
        _native.load()
        self._load()
        self.loaded = True

    def unload(self) -> None:
        self.loaded = False
        self._unload()

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        if not self.loaded:
            self.load()
        retry = kwargs.pop("retry", True)
        try:
            return self._run(**kwargs)
        except _native.NativeError:
            if retry:
                self.unload()
                return self(*args, retry=False, **kwargs)
            raise


# Drop-in alias so reference code importing the old name keeps working.
PretrainedONNXModel = PretrainedNativeModel
