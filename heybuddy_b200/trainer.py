"""
``WakeWordTrainer`` -- the reference's training loop (``heybuddy/trainer.py:314-608, 764-1007``) on top of
the fused device step ``WakeWordMLPModel.train_step``.

Kept: cosine LR with linear warm-up and hold (trainer.py:127-156), high-loss selection, weighted BCE with a
(dynamic) negative weight, "no update below 128 selected rows" with the reference's accumulation counter
(:451-458), validation-driven negative-weight adjustment (:531-536), the three stages with lr x0.5,
steps x2, batch x0.5 (:918-926), checkpoints as ``torch.save(state_dict)`` (:186-198).
Dropped on purpose: the per-step ``gc.collect() + empty_cache() + synchronize()`` (:592-594), wandb,
matplotlib and torchmetrics reporting (host-side, out of scope).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np

from heybuddy_b200.constants import *  # noqa: F401,F403
from heybuddy_b200.wakeword import WakeWordMLPModel

__all__ = ["WakeWordTrainer", "get_learning_rate"]


def get_learning_rate(step: int, warmup_steps: int = 0, hold_steps: int = 0, total_steps: int = 0,
                      start_learning_rate: float = 0.0, target_learning_rate: float = DEFAULT_LEARNING_RATE) -> float:
    """Cosine decay with warm-up and hold (trainer.py:127-156)."""
    lr = 0.5 * target_learning_rate * (1 + np.cos(np.pi * (step - warmup_steps - hold_steps) / float(total_steps - warmup_steps - hold_steps)))
    warm = target_learning_rate * (step / warmup_steps) if warmup_steps > 0 else 0.0
    if hold_steps > 0:
        lr = np.where(step > warmup_steps + hold_steps, lr, target_learning_rate)
    return float(np.where(step < warmup_steps, warm, lr))


class WakeWordTrainer:
    def __init__(self, model: Optional[WakeWordMLPModel] = None, learning_rate: float = DEFAULT_LEARNING_RATE,
                 checkpoint_dir: Optional[str] = None, device_id: Optional[int] = None, distributed: Optional[bool] = None) -> None:
        self.model = model or WakeWordMLPModel(device_id=device_id)
        if distributed is None:     # data-parallel when launched under torchrun with more than one rank
            import torch.distributed as dist

            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.distributed = bool(distributed)
        self.learning_rate = learning_rate
        self.checkpoint_dir = checkpoint_dir
        self.history: Dict[str, List[float]] = {"loss": [], "learning_rate": [], "high_loss_rate": [], "negative_weight": [], "stepped": []}

    @property
    def device(self):
        return self.model.device

    def save_checkpoint(self, name: str) -> Optional[str]:
        if not self.checkpoint_dir:
            return None
        import torch

        os.makedirs(self.checkpoint_dir, exist_ok=True)
        path = os.path.join(self.checkpoint_dir, f"{name}.pt")
        torch.save(self.model.state_dict(), path)
        return path

    def evaluate(self, dataset: Iterable[Tuple[Any, Any]], max_batches: Optional[int] = None,
                 activation_threshold: float = DEFAULT_ACTIVATION_THRESHOLD) -> Dict[str, float]:
        """Recall / false-positive rate over an iterator of (x, y) (trainer.py:483-560; 1.44 s per sample for FP/hour)."""
        import torch

        tp = fn = fp = neg = 0
        for i, (x, y) in enumerate(dataset):
            if max_batches is not None and i >= max_batches:
                break
            p = self.model(x.to(self.device, dtype=torch.float32)).squeeze(1)
            y = y.to(self.device)
            pred = p > activation_threshold
            tp += int((pred & (y == 1)).sum())
            fn += int((~pred & (y == 1)).sum())
            fp += int((pred & (y == 0)).sum())
            neg += int((y == 0).sum())
        hours = neg * 1.44 / 3600.0
        return {"recall": tp / max(tp + fn, 1), "false_positive_rate": fp / max(neg, 1),
                "false_positives_per_hour": fp / hours if hours > 0 else 0.0}

    def train_epoch(self, training: Iterable[Tuple[Any, Any]], validation: Optional[Iterable[Tuple[Any, Any]]] = None,
                    num_steps: int = DEFAULT_STEPS, warmup_steps: Optional[int] = None, hold_steps: Optional[int] = None,
                    learning_rate: Optional[float] = None, negative_weight: float = DEFAULT_NEGATIVE_WEIGHT,
                    high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD, validation_steps: int = DEFAULT_VALIDATION_STEPS,
                    target_false_positive_rate: float = DEFAULT_TARGET_FALSE_POSITIVE_RATE,
                    dynamic_negative_weight: bool = DEFAULT_DYNAMIC_NEGATIVE_WEIGHT,
                    negative_weight_adjust_ratio: float = DEFAULT_NEGATIVE_WEIGHT_ADJUST_RATIO,
                    validation_batches: Optional[int] = 8, sync_every: int = 1) -> float:
        """One stage.  Returns the negative weight at the end of the stage (carried to the next one)."""
        import torch

        lr0 = self.learning_rate if learning_rate is None else learning_rate
        warmup_steps = int(num_steps / 5.0) if warmup_steps is None else warmup_steps
        hold_steps = int(num_steps / 3.0) if hold_steps is None else hold_steps
        accumulated = 0
        for step, (x, y) in enumerate(training):
            if step >= num_steps:
                break
            lr = get_learning_rate(step, warmup_steps, hold_steps, num_steps, target_learning_rate=lr0)
            x = x.to(self.device, dtype=torch.float32, non_blocking=True)
            y = y.to(self.device, dtype=torch.int64, non_blocking=True)
            # reference: no backward/step until >= 128 selected rows have accumulated over consecutive steps (:441-458)
            if self.distributed:
                # every rank feeds its shard of the global batch; selection count, loss and gradients are reduced over the ranks
                from heybuddy_b200.dp import distributed_train_step

                _, stats = distributed_train_step(self.model, x, y, lr, negative_weight, high_loss_threshold,
                                                  min_selected=max(1, 128 - accumulated))
            else:
                _, stats = self.model.train_step(x, y, lr, negative_weight, high_loss_threshold, min_selected=max(1, 128 - accumulated))
            loss, n_sel, stepped, rate = stats.tolist()
            accumulated = 0 if stepped else accumulated + int(n_sel)
            for k, v in (("loss", loss), ("learning_rate", lr), ("high_loss_rate", rate), ("negative_weight", negative_weight), ("stepped", stepped)):
                self.history[k].append(float(v))
            if validation is not None and validation_steps and step > 0 and step % validation_steps == 0:
                metrics = self.evaluate(validation, max_batches=validation_batches)
                if dynamic_negative_weight:
                    if metrics["false_positives_per_hour"] > target_false_positive_rate:
                        negative_weight *= negative_weight_adjust_ratio
                    else:
                        negative_weight = max(negative_weight / negative_weight_adjust_ratio, 1e-3)
        return negative_weight

    def __call__(self, training: Any, validation: Any = None, num_steps: int = DEFAULT_STEPS, num_stages: int = DEFAULT_STAGES,
                 learning_rate: Optional[float] = None, negative_weight: float = DEFAULT_NEGATIVE_WEIGHT, **kwargs: Any) -> None:
        """Three-stage schedule (trainer.py:918-926): after each stage lr x0.5, steps x2, batch size x0.5."""
        lr = self.learning_rate if learning_rate is None else learning_rate
        steps = num_steps
        for stage in range(num_stages):
            negative_weight = self.train_epoch(training, validation, num_steps=steps, learning_rate=lr,
                                               negative_weight=negative_weight, **kwargs)
            self.save_checkpoint(f"stage_{stage}")
            lr *= DEFAULT_LEARNING_RATE_ADJUST_RATIO
            steps = int(steps * DEFAULT_STEP_ADJUST_RATIO)
            if hasattr(training, "multiply_batch_size"):
                training.multiply_batch_size(DEFAULT_BATCH_SIZE_ADJUST_RATIO)
        self.save_checkpoint("final")
