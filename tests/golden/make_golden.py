"""
Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box, and nothing in tests/, smoke() or bench.py reads it at run time):

    python tests/golden/make_golden.py

What it does (SURVEY.md 8c recipe):
  * copies ``/root/reference/src/python/heybuddy`` to a scratch dir (the package
    writes into its own tree at import, so it is never imported in place);
  * installs four stub modules (av, soundfile, numpy.compat, piper_phonemize) so the
    package imports in this image;
  * drives the reference's own, unmodified classes and writes small .npz files:

    pipeline_order.npz   reference ``SpeechEmbeddings.__call__`` with the oracle's mel /
                         embedding callables injected behind ``.spectrogram`` /
                         ``.embeddings`` (ring 1) -> pins windowing and output order.
    classifier_hey_buddy.npz  reference ``WakeWordMLPModel`` (eval) loaded with the in-repo trained
                         weights ``src/ts/models/hey-buddy.onnx`` -> probabilities, and one
                         training-step loss + gradients computed with the trainer's own
                         selection / weighting lines (trainer.py:405-462).
    classifier_zero_answers.npz  p(zeros[1,16,96]) for all 7 in-repo weight sets.
    add_noise.npz        ``torchaudio.functional.add_noise`` (the reference's dependency,
                         call site augmented.py:272-276) on seeded inputs.
    precalculated.npz    reference ``PrecalculatedDatasetIterator.from_array/take`` and
                         ``WakeWordTrainingDatasetIterator._generate_batches`` layout.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    scratch = tempfile.mkdtemp(prefix="heybuddy_ref_")
    shutil.copytree(os.path.join(REF, "src", "python", "heybuddy"), os.path.join(scratch, "heybuddy"))
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    for m in ("av", "soundfile", "numpy.compat", "piper_phonemize"):
        sys.modules[m] = types.ModuleType(m)
    sys.modules["piper_phonemize"].phonemize_espeak = lambda *a, **k: []
    np.compat = sys.modules["numpy.compat"]
    sys.path.insert(0, scratch)
    return scratch


def main() -> None:
    import torch
    import torchaudio

    from heybuddy_b200 import spec
    from oracle import classifier as ocls
    from oracle import embed as oembed
    from oracle import mel as omel

    scratch = import_reference()
    from heybuddy.embeddings import SpeechEmbeddings  # reference
    from heybuddy.wakeword import WakeWordMLPModel  # reference
    from heybuddy.dataset.precalculated import PrecalculatedDatasetIterator  # reference
    from heybuddy.dataset.training import WakeWordTrainingDatasetIterator  # reference

    weights = spec.init_embedding_weights()

    # ---- 1. pipeline order -------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(1001))
    clips = (0.1 * rng.standard_normal((3, spec.CLIP_SAMPLES))).clip(-1, 1).astype(np.float32)
    mel_calls, emb_calls = [], []

    def mel_fn(audio):
        mel_calls.append(tuple(audio.shape))
        m = omel.mel_spectrogram(audio)
        return m if m.shape[0] > 1 else m[0]  # the reference squeezes (spectrogram.py:32)

    def emb_fn(windows):
        emb_calls.append(tuple(windows.shape))
        return oembed.speech_embedding_model(windows, weights)

    s = SpeechEmbeddings()
    s.spectrogram = mel_fn
    s.embeddings = emb_fn
    emb, mels = s([torch.from_numpy(c.copy()) for c in clips], return_spectrograms=True)
    one_emb, one_mel = s(torch.from_numpy(clips[0, :17280].copy()), return_spectrograms=True)
    np.savez_compressed(
        os.path.join(HERE, "pipeline_order.npz"),
        seed=1001, embeddings=emb.astype(np.float32), spectrogram_shape=np.array(mels.shape),
        spectrogram_head=mels[:, :8].astype(np.float32), spectrogram_tail=mels[:, -8:].astype(np.float32),
        one_embeddings=one_emb.astype(np.float32), one_spectrogram_shape=np.array(one_mel.shape),
        mel_calls=np.array(mel_calls), emb_call_sizes=np.array([c[0] for c in emb_calls]),
    )
    print("pipeline_order", emb.shape, mels.shape, one_emb.shape, one_mel.shape, mel_calls[:5], emb_calls[:3])

    # ---- 2. classifier -------------------------------------------------------------------
    zero_answers = {}
    for fn in sorted(os.listdir(os.path.join(REF, "src", "ts", "models"))):
        if not fn.endswith(".onnx"):
            continue
        init = ocls.read_onnx_initializers(os.path.join(REF, "src", "ts", "models", fn))
        model = WakeWordMLPModel()
        model.load_state_dict({k: torch.from_numpy(v) for k, v in init.items()}, strict=True)
        model.eval()
        with torch.no_grad():
            zero_answers[fn[:-5]] = float(model(torch.zeros(1, 16, 96))[0, 0])
        if fn == "hey-buddy.onnx":
            rng = np.random.Generator(np.random.PCG64(4001))
            x = rng.standard_normal((64, 16, 96)).astype(np.float32)
            x[:16] += 0.5 * rng.standard_normal((1, 1, 96)).astype(np.float32)
            y = np.zeros(64, dtype=np.int64)
            y[:16] = 1
            with torch.no_grad():
                prob = model(torch.from_numpy(x)).numpy()
            # one training step's loss + grads with the trainer's own lines (trainer.py:405-446)
            thr, neg_w = 1e-4, 0.7
            model.zero_grad()
            xt, yt = torch.from_numpy(x), torch.from_numpy(y)
            y_pred = model(xt)
            negative_high_loss = y_pred[(yt == 0) & (y_pred.squeeze() >= thr)]
            positive_high_loss = y_pred[(yt == 1) & (y_pred.squeeze() < 1 - thr)]
            ysel = torch.cat([yt[(yt == 0) & (y_pred.squeeze() >= thr)], yt[(yt == 1) & (y_pred.squeeze() < 1 - thr)]]).to(dtype=torch.float32)
            y_pred = torch.cat([negative_high_loss, positive_high_loss])
            weight = torch.ones(ysel.shape[0]) * neg_w
            weight[ysel == 1] = 1.0
            loss = torch.nn.functional.binary_cross_entropy(y_pred, ysel.unsqueeze(1), weight.unsqueeze(1))
            loss.backward()
            grads = {f"grad::{k}": v.grad.numpy().copy() for k, v in model.named_parameters()
                     if not k.startswith("mlp_in.") and not k.startswith("norm_in.")}
            grads["gradnorm::mlp_in.hidden.weight"] = np.array(model.mlp_in.hidden.weight.grad.norm().item())
            grads["gradnorm::mlp_in.gate.weight"] = np.array(model.mlp_in.gate.weight.grad.norm().item())
            grads["gradnorm::norm_in.weight"] = np.array(model.norm_in.weight.grad.norm().item())
            grads["grad::mlp_in.hidden.bias"] = model.mlp_in.hidden.bias.grad.numpy().copy()
            np.savez_compressed(
                os.path.join(HERE, "classifier_hey_buddy.npz"),
                x_seed=4001, prob=prob.astype(np.float32), loss=np.array(float(loss)), n_selected=np.array(int(y_pred.shape[0])),
                negative_weight=np.array(neg_w), threshold=np.array(thr),
                **{f"param::{k}": v for k, v in init.items()}, **grads,
            )
            print("classifier", prob[:4, 0], float(loss), int(y_pred.shape[0]))
    np.savez(os.path.join(HERE, "classifier_zero_answers.npz"), **{k.replace("-", "_"): np.array(v) for k, v in zero_answers.items()})
    print("zero answers", zero_answers)

    # ---- 3. add_noise ----------------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(2002))
    wav = rng.standard_normal((4, 4096)).astype(np.float32) * 0.1
    noi = rng.standard_normal((4, 4096)).astype(np.float32) * np.array([[0.01], [0.3], [1.0], [5.0]], dtype=np.float32)
    snr = np.array([-10.0, 0.0, 7.5, 15.0], dtype=np.float32)
    out = torchaudio.functional.add_noise(torch.from_numpy(wav), torch.from_numpy(noi), torch.from_numpy(snr)).numpy()
    np.savez_compressed(os.path.join(HERE, "add_noise.npz"), seed=2002, snr=snr, out=out)
    print("add_noise", out.shape)

    # ---- 4. .npy store + batch layout --------------------------------------------------------
    d = tempfile.mkdtemp(prefix="heybuddy_npy_")
    import heybuddy.dataset.precalculated as refpre
    refpre.LOCAL_DIR = d
    arr = np.arange(7 * 16 * 96, dtype=np.float32).reshape(7, 16, 96)
    # from_array re-opens from the *default* directory (quirk, precalculated.py:482-491): save there.
    default_dir = PrecalculatedDatasetIterator.__init__.__defaults__[0]
    os.makedirs(default_dir, exist_ok=True)
    it = PrecalculatedDatasetIterator.from_array(arr, "golden_tmp", directory=default_dir, ordered=True)
    with open(os.path.join(default_dir, "golden_tmp.npy"), "rb") as fh:
        header = fh.read(128)
    took = [it.take(3).copy(), it.take(3).copy(), it.take(3).copy()]  # third wraps around
    neg = PrecalculatedDatasetIterator("golden_tmp", directory=default_dir, ordered=True)
    tr = WakeWordTrainingDatasetIterator(positive=[(it, 2)], negative=[(neg, 3)], num_batch_threads=1, start=False)
    tr.start()
    x, y = next(iter(tr))
    tr.stop()
    np.savez_compressed(
        os.path.join(HERE, "precalculated.npz"), header=np.frombuffer(header, dtype=np.uint8),
        take0=took[0][:, 0, 0], take1=took[1][:, 0, 0], take2=took[2][:, 0, 0],
        batch_x_shape=np.array(x.shape), batch_x_dtype=str(x.dtype), batch_y=y.numpy(), batch_y_dtype=str(y.dtype),
        batch_x_first=x.numpy()[:, 0, 0],
    )
    print("precalculated", header[:10], [t[:, 0, 0] for t in took], x.shape, y)
    shutil.rmtree(scratch, ignore_errors=True)
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
