"""
``WakeWordTrainingDatasetIterator`` -- threaded batcher over precalculated datasets producing
``(x f32[sum n,16,96], y i64[sum n])`` with positives first (label 1) then negatives (label 0)
(reference ``heybuddy/dataset/training.py:29-277``).

Same surface (``positive=[(dataset, n)]``, ``negative=[...]``, ``num_batch_threads``, ``max_queued_batches``,
``max_samples``, ``start``; ``start/stop/iterate/summary/metadata/multiply_batch_size``) and the factory
classmethods ``default / testing / validation / all`` on top of ``TrainingFeaturesGenerator``.  The reference
starts its batcher threads before ``positive``/``negative`` are assigned (training.py:170-178, a thread dies
and is resurrected); here fields are assigned first.  The hosted negative sets need network downloads: pass
``negative_datasets`` (iterators or names of local ``.npy`` files) instead.
"""
from __future__ import annotations

from queue import Empty, Queue
from threading import Event, Thread
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from heybuddy_b200.constants import *  # noqa: F401,F403
from heybuddy_b200.dataset.precalculated import PrecalculatedDatasetIterator
from heybuddy_b200.util import logger

__all__ = ["TrainingDatasetIterator", "WakeWordTrainingDatasetIterator"]


class TrainingDatasetIterator:
    def __init__(self, max_samples: Optional[int] = None, num_batch_threads: int = 2, max_queued_batches: int = 100,
                 start: bool = True, **kwargs: Any) -> None:
        self.total_yielded_samples = 0
        self.max_samples = max_samples
        self.num_batch_threads = num_batch_threads
        self.threads: List[Tuple[Thread, Event]] = []
        self.queue: "Queue[Tuple[Any, Any]]" = Queue(max_queued_batches)
        self.started = False
        if start:
            self.start()

    def metadata(self) -> Dict[str, Any]:
        return {"max_samples": self.max_samples, "num_batch_threads": self.num_batch_threads}

    def _spawn(self) -> Tuple[Thread, Event]:
        ev = Event()
        th = Thread(target=self._generate_batches, args=(ev,), daemon=True)
        th.start()
        return th, ev

    def start(self) -> None:
        if self.started:
            return
        self.started = True
        logger.info(f"Starting batch generation with {self.num_batch_threads} threads")
        self.threads = [self._spawn() for _ in range(self.num_batch_threads)]

    def check_restart(self) -> None:
        if not self.started:
            self.start()
            return
        for i, (thread, event) in enumerate(self.threads):
            if not thread.is_alive():
                logger.warning(f"Batch generation thread {i} has stopped, restarting")
                self.threads[i] = self._spawn()

    def stop(self) -> None:
        for _, ev in self.threads:
            ev.set()
        for th, _ in self.threads:
            th.join()
        self.threads.clear()
        with self.queue.mutex:
            self.queue.queue.clear()
        self.started = False

    def iterate(self) -> Iterator[Tuple[Any, Any]]:
        yielded = 0
        while True:
            try:
                yield self.queue.get(timeout=1)
                yielded += 1
                self.total_yielded_samples += 1
                if self.max_samples is not None and yielded >= self.max_samples:
                    break
                if self.total_yielded_samples % 10 == 0:
                    self.check_restart()
            except Empty:
                self.check_restart()

    def __iter__(self) -> Iterator[Tuple[Any, Any]]:
        return self.iterate()

    def _generate_batches(self, stop_event: Event) -> None:
        raise NotImplementedError("Subclasses must implement this method")


class WakeWordTrainingDatasetIterator(TrainingDatasetIterator):
    def __init__(self, max_samples: Optional[int] = None, num_batch_threads: int = 2, max_queued_batches: int = 100,
                 start: bool = True, positive: Sequence[Tuple[PrecalculatedDatasetIterator, int]] = (),
                 negative: Sequence[Tuple[PrecalculatedDatasetIterator, int]] = (), pin_memory: bool = False) -> None:
        assert positive or negative, "At least one positive or negative dataset is required"
        self.positive = list(positive)   # assigned BEFORE the threads start
        self.negative = list(negative)
        self.pin_memory = pin_memory
        super().__init__(max_samples=max_samples, num_batch_threads=num_batch_threads,
                         max_queued_batches=max_queued_batches, start=start)

    def metadata(self) -> Dict[str, Any]:
        def rows(sets):
            return [{"length": len(ds), "batch_size": n, "metadata": ds.metadata() if isinstance(ds, PrecalculatedDatasetIterator) else None}
                    for ds, n in sets]
        return {**super().metadata(), "positive": rows(self.positive), "negative": rows(self.negative)}

    def summary(self) -> str:
        lines = [f"Total batches yielded: {self.total_yielded_samples}"]
        for label, sets in (("Positive", self.positive), ("Negative", self.negative)):
            for i, (ds, n) in enumerate(sets):
                lines.append(f"{label} dataset {i+1}: {ds.total_taken} samples taken out of {len(ds)} unique samples "
                             f"({n} per batch, {ds.total_taken / len(ds):.2%} seen)")
        return "\n".join(lines)

    def multiply_batch_size(self, ratio: float) -> None:
        restart = self.started
        if self.started:
            self.stop()
        self.positive = [(ds, max(1, int(n * ratio))) for ds, n in self.positive]
        self.negative = [(ds, max(1, int(n * ratio))) for ds, n in self.negative]
        if restart:
            self.start()

    def half_batch_size(self) -> None:
        self.multiply_batch_size(0.5)

    def double_batch_size(self) -> None:
        self.multiply_batch_size(2)

    def make_batch(self) -> Tuple[Any, Any]:
        """One batch: positives (label 1) then negatives (label 0), training.py:254-262."""
        import torch

        samples = [ds.take(n) for ds, n in self.positive] + [ds.take(n) for ds, n in self.negative]
        labels = [np.ones(n) for _, n in self.positive] + [np.zeros(n) for _, n in self.negative]
        x = torch.from_numpy(np.concatenate(samples))
        y = torch.from_numpy(np.concatenate(labels).astype(np.int64))
        if self.pin_memory and torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        return x, y

    def _generate_batches(self, stop_event: Event) -> None:
        while not stop_event.is_set():
            x, y = self.make_batch()
            while self.queue.full():
                if stop_event.is_set():
                    return
                stop_event.wait(0.1)
            self.queue.put((x, y))

    # -- factories (training.py:279-905) --------------------------------------------------------------------
    @staticmethod
    def _negatives(negative_datasets, per_batch: int) -> List[Tuple[PrecalculatedDatasetIterator, int]]:
        out = []
        for ds in negative_datasets or []:
            out.append((ds if isinstance(ds, PrecalculatedDatasetIterator) else PrecalculatedDatasetIterator(ds), per_batch))
        return out

    @classmethod
    def default(cls, wake_phrase: str, num_positive_samples: int = DEFAULT_POSITIVE_SAMPLES,
                num_adversarial_samples: int = DEFAULT_ADVERSARIAL_SAMPLES, positive_per_batch: int = DEFAULT_POSITIVE_BATCH_SIZE,
                negative_per_batch: int = DEFAULT_NEGATIVE_BATCH_SIZE, adversarial_per_batch: int = DEFAULT_ADVERSARIAL_BATCH_SIZE,
                use_cache: bool = True, num_batch_threads: int = DEFAULT_BATCH_THREADS, start: bool = True,
                negative_datasets: Optional[Sequence[Union[str, PrecalculatedDatasetIterator]]] = None, testing: bool = False,
                num_negative_samples: Optional[int] = None, **feature_kwargs: Any) -> "WakeWordTrainingDatasetIterator":
        from heybuddy_b200.dataset.features import TrainingFeaturesGenerator

        if num_negative_samples is not None:  # stale spelling used by the reference's own test
            num_adversarial_samples = num_negative_samples
        positive, adversarial = TrainingFeaturesGenerator.get_training_features(
            wake_phrase, num_positive_samples=num_positive_samples, num_adversarial_samples=num_adversarial_samples,
            testing=testing, use_cache=use_cache, **feature_kwargs)
        return cls(positive=[(positive, positive_per_batch)],
                   negative=[(adversarial, adversarial_per_batch)] + cls._negatives(negative_datasets, negative_per_batch),
                   num_batch_threads=num_batch_threads, start=start)

    @classmethod
    def testing(cls, wake_phrase: str, num_positive_samples: int = DEFAULT_TESTING_POSITIVE_SAMPLES,
                num_adversarial_samples: int = DEFAULT_TESTING_ADVERSARIAL_SAMPLES, **kwargs: Any) -> "WakeWordTrainingDatasetIterator":
        return cls.default(wake_phrase, num_positive_samples=num_positive_samples, num_adversarial_samples=num_adversarial_samples,
                           testing=True, **kwargs)

    @classmethod
    def validation(cls, wake_phrase: str, num_positive_samples: int = DEFAULT_VALIDATION_SAMPLES,
                   positive_per_batch: int = DEFAULT_VALIDATION_POSITIVE_BATCH_SIZE,
                   negative_per_batch: int = DEFAULT_VALIDATION_NEGATIVE_BATCH_SIZE, use_cache: bool = True,
                   num_batch_threads: int = DEFAULT_BATCH_THREADS, start: bool = True,
                   negative_datasets: Optional[Sequence[Union[str, PrecalculatedDatasetIterator]]] = None,
                   **feature_kwargs: Any) -> "WakeWordTrainingDatasetIterator":
        from heybuddy_b200.dataset.features import TrainingFeaturesGenerator

        positive = TrainingFeaturesGenerator.get_validation_features(wake_phrase, num_positive_samples=num_positive_samples,
                                                                     use_cache=use_cache, **feature_kwargs)
        return cls(positive=[(positive, positive_per_batch)], negative=cls._negatives(negative_datasets, negative_per_batch),
                   num_batch_threads=num_batch_threads, start=start)

    @classmethod
    def all(cls, wake_phrase: str, **kwargs: Any):
        """(training, testing, validation), none started (training.py:705-905)."""
        kwargs = {**kwargs, "start": False}
        val_kwargs = {k: v for k, v in kwargs.items() if k not in ("num_adversarial_samples", "adversarial_per_batch")}
        return cls.default(wake_phrase, **kwargs), cls.testing(wake_phrase, **kwargs), cls.validation(wake_phrase, **val_kwargs)
