"""
CPU oracle for the featurization hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic the reference performs on the
path (augmentation -> log-mel -> speech embeddings -> classifier).  It exists to
check the CUDA kernels; it is NOT a fallback.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``heybuddy_b200/`` imports it, and the product
path raises if the CUDA extension is missing.

Parity status:

* windowing / ordering / .npy / batch iterator  -- PINNED: ``oracle.pipeline`` is
  checked against the reference's own unmodified ``SpeechEmbeddings`` driven
  with injected callables (fixtures ``tests/golden/pipeline_order.npz``,
  generator ``tests/golden/make_golden.py``).
* classifier forward                              -- PINNED: checked against the
  reference ``WakeWordMLPModel`` loaded with in-repo trained weights
  (``tests/golden/classifier_*.npz``).
* background-noise SNR mix                        -- PINNED to the third-party
  dependency ``torchaudio.functional.add_noise`` (torchaudio 2.11.0 present
  here; reference floor torchaudio>=2.3, environment.yml:28); fixture
  ``tests/golden/add_noise.npz``.
* mel values, embedding CNN interior, gain / coloured noise / reverb
  -- **PARITY UNPINNED**: the algorithms live in artefacts and packages that are
  absent from /root/reference and from this image (mel-spectrogram.onnx,
  speech-embedding.onnx, torch_audiomentations>=0.11, speechbrain>=1.0).  The
  oracle restates their published behaviour and anchors
  on the reference's call sites and pinned shapes.
* K9 (``oracle.k9``): SevenBandParametricEQ, TanhDistortion, BandStopFilter
  -- **PARITY UNPINNED** (audiomentations, torch_audiomentations, julius absent;
  scipy ``sosfilt`` / numpy ``percentile`` / torch do the arithmetic the libraries
  delegate).  PitchShift: the four stages are the library's own calls
  (``torch.stft``, torchaudio ``TimeStretch`` / ``Resample``, ``torch.istft``: present,
  PINNED); ``torch_pitch_shift``'s glue around them is restated (UNPINNED).
"""
