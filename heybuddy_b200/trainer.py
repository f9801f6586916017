"""
``WakeWordTrainer`` -- the reference's training loop (``heybuddy/trainer.py:314-608, 764-1007``) on top of
the fused device step ``WakeWordMLPModel.train_step``.

Kept (each line cites what it mirrors):
  * cosine LR with linear warm-up and hold (trainer.py:127-156);
  * high-loss selection and weighted BCE, mean over the selected rows (:405-446, inside the fused step);
  * "no update below 128 selected rows": a step that leaves the running count of selected rows under 128 does not update, the
    counter ``accumulation_steps`` grows, and the step that finally fires has its loss -- hence its gradients -- divided by that
    counter (:441-458; gradients are NOT accumulated across the skipped steps in the reference either, only the divisor is);
  * the negative weight: a float, or a per-step schedule list with "use the last value when too short" (:427-433); with
    ``dynamic_negative_weight`` the validation false-positive rate doubles it or halves it with a floor of 1.0 (:531-536),
    without it the stage runs ``np.linspace(1, max_negative_weight, num_steps)`` (:849-854);
  * false positives counted with ``>=`` (``y - p <= -threshold``, :290-299);
  * train-mode ``nn.Dropout(0.1)`` on the classifier input (wakeword.py:197,338) -- drawn on the device from the draw table's
    Philox generator (``hb_mlp_dropout``); ``dropout=0`` reproduces the parity configuration;
  * the stages: lr x0.5, steps x2 (never below ``validation_steps``), batch x0.5, the dynamic negative weight carried over
    (:918-926);
  * checkpoints ``<name>.pt`` = ``torch.save(model.state_dict())`` and ``<name>_optimizer.pt`` = a ``torch.optim.Adam``
    state dict (:186-198) -- interchangeable with the reference's -- and ``resume(name)`` picking the newest model / optimizer
    pair written within 2 s of each other (:54-118).
Dropped on purpose: the per-step ``gc.collect() + empty_cache() + synchronize()`` (:592-594), wandb, matplotlib and torchmetrics
reporting (host-side, out of scope; ``history`` keeps the per-step series).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

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
                 checkpoint_dir: Optional[str] = None, device_id: Optional[int] = None, distributed: Optional[bool] = None,
                 dropout: Optional[float] = None, seed: int = 0) -> None:
        self.model = model or WakeWordMLPModel(device_id=device_id)
        if distributed is None:     # data-parallel when launched under torchrun with more than one rank
            import torch.distributed as dist

            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.distributed = bool(distributed)
        self.learning_rate = learning_rate
        self.checkpoint_dir = checkpoint_dir
        self.dropout = self.model.dropout if dropout is None else float(dropout)   # train-mode input dropout (wakeword.py:197)
        self.seed = seed
        self.history: Dict[str, List[float]] = {"loss": [], "learning_rate": [], "high_loss_rate": [], "negative_weight": [], "stepped": []}

    @property
    def device(self):
        return self.model.device

    # -- checkpoints (trainer.py:54-118, 186-198) ------------------------------------------------------------------------------
    def save_checkpoint(self, name: str, optimizer: bool = True) -> Optional[str]:
        if not self.checkpoint_dir:
            return None
        import torch

        os.makedirs(self.checkpoint_dir, exist_ok=True)
        path = os.path.join(self.checkpoint_dir, f"{name}.pt")
        torch.save(self.model.state_dict(), path)
        if optimizer:
            torch.save(self.model.optimizer_state_dict(self.learning_rate), os.path.join(self.checkpoint_dir, f"{name}_optimizer.pt"))
        return path

    def resume(self, name: str) -> None:
        """Loads the newest ``<name>*.pt`` / ``<name>*_optimizer.pt`` pair written within 2 s of each other."""
        import torch

        files = os.listdir(self.checkpoint_dir) if self.checkpoint_dir and os.path.isdir(self.checkpoint_dir) else []
        stamp = lambda f: os.path.getmtime(os.path.join(self.checkpoint_dir, f))
        models = sorted(((f, stamp(f)) for f in files if f.startswith(name) and f.endswith(".pt") and not f.endswith("_optimizer.pt")),
                        key=lambda x: x[1], reverse=True)
        optims = sorted(((f, stamp(f)) for f in files if f.startswith(name) and f.endswith("_optimizer.pt")), key=lambda x: x[1], reverse=True)
        for mf, mt in models:
            for of, ot in optims:
                if abs(mt - ot) < 2:
                    self.model.load_state_dict(torch.load(os.path.join(self.checkpoint_dir, mf), weights_only=True, map_location="cpu"))
                    self.model.load_optimizer_state_dict(torch.load(os.path.join(self.checkpoint_dir, of), weights_only=True, map_location="cpu"))
                    return
        raise FileNotFoundError(f"Checkpoint {name} not found.")

    # -- metrics ---------------------------------------------------------------------------------------------------------------
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
            pred = p >= activation_threshold          # num_false_positives: y - p <= -threshold (trainer.py:290-299)
            tp += int((pred & (y == 1)).sum())
            fn += int((~pred & (y == 1)).sum())
            fp += int((pred & (y == 0)).sum())
            neg += int((y == 0).sum())
        hours = neg * 1.44 / 3600.0
        return {"recall": tp / max(tp + fn, 1), "false_positive_rate": fp / max(neg, 1),
                "false_positives_per_hour": fp / hours if hours > 0 else 0.0}

    # -- one stage ---------------------------------------------------------------------------------------------------------------
    def train_epoch(self, training: Iterable[Tuple[Any, Any]], validation: Optional[Iterable[Tuple[Any, Any]]] = None,
                    num_steps: int = DEFAULT_STEPS, warmup_steps: Optional[int] = None, hold_steps: Optional[int] = None,
                    learning_rate: Optional[float] = None, negative_weight: Union[float, Sequence[float]] = DEFAULT_NEGATIVE_WEIGHT,
                    high_loss_threshold: float = DEFAULT_HIGH_LOSS_THRESHOLD, validation_steps: int = DEFAULT_VALIDATION_STEPS,
                    target_false_positive_rate: float = DEFAULT_TARGET_FALSE_POSITIVE_RATE,
                    dynamic_negative_weight: bool = DEFAULT_DYNAMIC_NEGATIVE_WEIGHT,
                    negative_weight_adjust_ratio: Optional[float] = DEFAULT_NEGATIVE_WEIGHT_ADJUST_RATIO,
                    validation_batches: Optional[int] = 8, sync_every: int = 1,
                    negative_weight_schedule: Optional[Union[float, Sequence[float]]] = None) -> float:
        """
        One stage (trainer.py:314-608).  ``negative_weight`` (alias ``negative_weight_schedule``, the reference's name): a float or a
        per-step list.  Returns the negative weight at the end of the stage (carried to the next one).
        """
        import torch

        schedule = negative_weight if negative_weight_schedule is None else negative_weight_schedule
        is_list = not isinstance(schedule, (int, float))
        if is_list:
            schedule = [float(v) for v in schedule]
            assert not (dynamic_negative_weight and negative_weight_adjust_ratio is not None and validation is not None), \
                "Negative weight schedule must be a float when using dynamic negative weight adjustment."   # trainer.py:532
        lr0 = self.learning_rate if learning_rate is None else learning_rate
        warmup_steps = num_steps // 5 if warmup_steps is None else warmup_steps
        hold_steps = num_steps // 3 if hold_steps is None else hold_steps
        accumulated_samples, accumulation_steps = 0, 1
        weight_now = float(schedule[0]) if is_list and schedule else (float(schedule) if not is_list else DEFAULT_NEGATIVE_WEIGHT)
        for step, (x, y) in enumerate(training):
            if step >= num_steps:
                break
            lr = get_learning_rate(step, warmup_steps, hold_steps, num_steps, target_learning_rate=lr0)
            if is_list:
                weight_now = schedule[step] if step < len(schedule) else schedule[-1]      # "use the last value" (trainer.py:429-433)
            else:
                weight_now = float(schedule)
            x = x.to(self.device, dtype=torch.float32, non_blocking=True)
            y = y.to(self.device, dtype=torch.int64, non_blocking=True)
            # reference: no backward / step until >= 128 selected rows have accumulated over consecutive steps; the loss of the
            # step that fires is divided by the accumulation counter (:441-458)
            self.model.set_loss_scale(1.0 / accumulation_steps)
            min_selected = max(1, 128 - accumulated_samples)
            if self.distributed:
                # every rank feeds its shard of the global batch; selection count, loss and gradients are reduced over the ranks
                from heybuddy_b200.dp import distributed_train_step

                _, stats = distributed_train_step(self.model, self.model.apply_dropout(x, self.dropout, self.seed), y, lr, weight_now,
                                                  high_loss_threshold, min_selected=min_selected)
            else:
                _, stats = self.model.train_step(x, y, lr, weight_now, high_loss_threshold, min_selected=min_selected,
                                                 dropout=self.dropout, dropout_seed=self.seed)
            loss, n_sel, stepped, rate = stats.tolist()
            if stepped:
                accumulated_samples, accumulation_steps = 0, 1
            elif n_sel > 0:                         # a step with no selected rows at all changes nothing (trainer.py:443)
                accumulated_samples += int(n_sel)
                accumulation_steps += 1
            for k, v in (("loss", loss), ("learning_rate", lr), ("high_loss_rate", rate), ("negative_weight", weight_now), ("stepped", stepped)):
                self.history[k].append(float(v))
            if validation is not None and validation_steps and step > 0 and step % validation_steps == 0:
                metrics = self.evaluate(validation, max_batches=validation_batches)
                if dynamic_negative_weight and negative_weight_adjust_ratio is not None and not is_list:
                    if metrics["false_positives_per_hour"] > target_false_positive_rate:
                        schedule = float(schedule) * negative_weight_adjust_ratio
                    else:
                        schedule = max(1.0, float(schedule) / negative_weight_adjust_ratio)        # trainer.py:536
        self.model.set_loss_scale(1.0)
        return float(schedule) if not is_list else float(weight_now)

    # -- the schedule (trainer.py:764-1007) ------------------------------------------------------------------------------------------
    def __call__(self, training: Any, validation: Any = None, num_steps: int = DEFAULT_STEPS, num_stages: int = DEFAULT_STAGES,
                 learning_rate: Optional[float] = None, max_negative_weight: float = DEFAULT_NEGATIVE_WEIGHT,
                 validation_steps: int = DEFAULT_VALIDATION_STEPS, dynamic_negative_weight: bool = DEFAULT_DYNAMIC_NEGATIVE_WEIGHT,
                 negative_weight_adjust_ratio: float = DEFAULT_NEGATIVE_WEIGHT_ADJUST_RATIO,
                 batch_size_adjust_ratio: float = DEFAULT_BATCH_SIZE_ADJUST_RATIO,
                 learning_rate_adjust_ratio: float = DEFAULT_LEARNING_RATE_ADJUST_RATIO, step_adjust_ratio: float = DEFAULT_STEP_ADJUST_RATIO,
                 name: str = "heybuddy", negative_weight: Optional[float] = None, **kwargs: Any) -> None:
        """Three-stage schedule (trainer.py:846-926): after each stage lr x0.5, steps x2, batch size x0.5."""
        lr = self.learning_rate if learning_rate is None else learning_rate
        if negative_weight is not None:           # round-1 spelling of max_negative_weight
            max_negative_weight = negative_weight
        steps = num_steps
        for stage in range(num_stages):
            if dynamic_negative_weight:
                weights: Union[float, List[float]] = float(max_negative_weight)
                ratio: Optional[float] = negative_weight_adjust_ratio
            else:
                weights = np.linspace(1, max_negative_weight, steps).tolist()        # trainer.py:853
                ratio = None
            self.learning_rate = lr
            last = self.train_epoch(training, validation, num_steps=steps, learning_rate=lr, negative_weight=weights,
                                    dynamic_negative_weight=dynamic_negative_weight, negative_weight_adjust_ratio=ratio,
                                    validation_steps=validation_steps, **kwargs)
            self.save_checkpoint(f"{name}_{stage}")
            lr *= learning_rate_adjust_ratio
            steps = max(validation_steps, int(steps * step_adjust_ratio))
            if validation is not None and dynamic_negative_weight:
                max_negative_weight = last                                               # trainer.py:921-922
            if hasattr(training, "multiply_batch_size"):
                training.multiply_batch_size(batch_size_adjust_ratio)
        self.save_checkpoint(f"{name}_final")
