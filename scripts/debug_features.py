import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import DrawTable
from heybuddy_b200.dataset.features import SyntheticSpeechSource, TrainingFeaturesGenerator
from heybuddy_b200.pipeline import RaggedClips
from oracle import augment as oaug, embed as oembed, mel as omel, pipeline as opipe
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_features_gpu import _banks
rng = np.random.default_rng(3)
noise, rirs = _banks(rng)
gen = TrainingFeaturesGenerator(device_id=0, use_autoconfigure=False, augment_batch_size=8, augment_background_dataset=noise,
                                augment_impulse_dataset=rirs, precision="fp32", seed=2004, source=SyntheticSpeechSource(9))
got = gen(40)
pipe, aug = gen._pipeline(True)
clips = SyntheticSpeechSource(9)(40)
table = DrawTable.build([c.shape[0] for c in clips], aug.cfg, 2004, aug.noise_bank.clip_lengths, len(aug.rir_bank))
stream = aug.noise_bank.stream.cpu().numpy()
audio, i0 = [], 0
for d, ncur, ridx in zip(table.batches, table.noise_clip_cursor, table.rir_index):
    b = len(d.pad_before)
    fixed = np.stack([oaug.to_target_length(c, int(p)) for c, p in zip(clips[i0:i0 + b], d.pad_before)])
    off = aug.noise_bank.offset_of_clip(ncur) if d.background_apply else 0
    audio.append(oaug.augment_batch(fixed, colored_base=d.colored_base if d.colored_apply else None, colored_snr_db=d.colored_snr_db,
        gain_db=d.gain_db if d.gain_apply else None,
        noise=stream[off:off + b * spec.CLIP_SAMPLES].reshape(b, -1) if d.background_apply else None, noise_snr_db=d.noise_snr_db,
        rir=aug.rir_bank.kernels_host[ridx] if d.reverb_apply else None))
    print("batch", d.index, "colored", d.colored_apply, "gain", round(d.gain_db,2), "bg", d.background_apply, ncur, "rev", d.reverb_apply, ridx)
    i0 += b
audio = np.concatenate(audio)
# device audio through the same pipeline
chunk = pipe.upload(RaggedClips.from_list(clips), table)
emb_d, audio_d = pipe.run_device(chunk, keep_audio=True)
audio_d = audio_d.cpu().numpy()
for i in range(0, 40, 4):
    print(i, "audio err", np.abs(audio_d[i] - audio[i]).max() / np.abs(audio[i]).max(), "amp", np.abs(audio[i]).max())
weights = spec.init_embedding_weights()
want = opipe.speech_embeddings([a for a in audio], omel.mel_spectrogram, lambda w: oembed.speech_embedding_model(w, weights, dtype=torch.float64))
e = np.abs(got - want).reshape(40, -1).max(axis=1) / np.abs(want).max()
print("per clip emb err", np.round(e, 4))
e2 = np.abs(emb_d.cpu().numpy() - want).reshape(40, -1).max(axis=1) / np.abs(want).max()
print("per clip emb err (upload path)", np.round(e2, 4))
m_d = pipe.speech.spectrogram.run_device(torch.from_numpy(audio).cuda(), scale=spec.AUDIO_SCALE).cpu().numpy()
m_o = omel.mel_spectrogram(audio * np.float32(spec.AUDIO_SCALE))
print("mel err per clip", np.round(np.abs(m_d - m_o).reshape(40, -1).max(axis=1), 5))
print("mel min per clip", np.round(m_o.reshape(40,-1).min(axis=1), 2))
