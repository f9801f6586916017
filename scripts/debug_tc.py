"""GPU debug: per-layer parity of the tcgen05 embedding path against the oracle (f16 operand emulation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from heybuddy_b200 import spec
from heybuddy_b200.embeddings import SpeechEmbeddingModel
from oracle import mel as omel, embed as oembed

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 141
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.Generator(np.random.PCG64(2))
t = 512 + 160 * (frames - 1)
audio = (0.1 * rng.standard_normal((B, t))).clip(-1, 1).astype(np.float32) * np.float32(spec.AUDIO_SCALE)
m = omel.mel_spectrogram(audio)
w = spec.init_embedding_weights()
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
md = torch.from_numpy(m).cuda()

x = torch.from_numpy(m).double()[:, None]
def rnd(v): return v.to(torch.float16).to(torch.float64)
for li, (name, kh, kw, cin, cout, pad, act, pool) in enumerate(spec.EMBEDDING_LAYERS[:16]):
    wt = torch.from_numpy(w[f"{name}.weight"]).double().permute(3, 2, 0, 1)
    xin, win = (x, wt) if li == 0 else (rnd(x), rnd(wt))
    x = F.conv2d(xin, win, torch.from_numpy(w[f"{name}.bias"]).double(), padding=(0, kw // 2) if pad == "same" else 0)
    if act: x = F.leaky_relu(x, spec.LEAKY_SLOPE)
    if pool: x = F.max_pool2d(x, pool, pool)
    want = x.permute(0, 2, 3, 1).numpy()
    try:
        got = model.activation_device(md, li).cpu().numpy()
    except Exception as exc:
        print(f"layer {li:2d} {name:10s} ERROR {exc}")
        break
    err = np.abs(got - want)
    bad = np.argwhere(err > 1e-2 * np.abs(want).max())
    print(f"layer {li:2d} {name:10s} shape {got.shape} max|err|/max|ref| {err.max()/np.abs(want).max():.3e} finite {np.isfinite(got).all()} nbad {len(bad)}" + (f" first bad {bad[0]} got {got[tuple(bad[0])]:.4f} want {want[tuple(bad[0])]:.4f}" if len(bad) else ""))
    if len(bad) and "-v" in sys.argv:
        rows = sorted(set(int(b[1]) for b in bad))
        print("   bad rows:", rows[:40], " bad f:", sorted(set(int(b[2]) for b in bad))[:40], " bad c:", sorted(set(int(b[3]) for b in bad))[:40])
offs = spec.embedding_frame_offsets(spec.CLIP_SAMPLES) if frames >= 136 else [0]
got = model.run_clips_device(md, offs).cpu().numpy()
want = np.stack([oembed.speech_embedding_model(m[:, o:o + 76, :, None], w, dtype=torch.float64) for o in offs], axis=1)
err = got - want
print("final", got.shape, "max", np.abs(err).max() / np.abs(want).max(), "l2", np.linalg.norm(err) / np.linalg.norm(want))
