/*
 * heybuddy_b200 -- C-ABI of the B200-native featurization hot path.
 *
 * The reference (therealadityashankar/hey-buddy) has no FFI: its boundary is the
 * Python "ring 1" model callable (src/python/heybuddy/util/onnx_util.py:83-96,
 * InferenceSession.run) and the stock PyTorch / torchaudio / speechbrain ops the
 * hot path calls.  Each entry point below replaces one of those call sites and is
 * what a ctypes stub on the reference side binds.
 *
 * Conventions
 *   - every function returns 0 (HB_OK) or a negative HB_ERR_* code; the message is
 *     available from hb_last_error() (thread-local);
 *   - pointers named *_dev are device pointers on the current CUDA device, pointers
 *     named *_host are host pointers; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous on that stream and re-entrant across streams;
 *   - no torch types, no global state except immutable constant tables and the
 *     explicit model handles.
 */
#ifndef HEYBUDDY_B200_H
#define HEYBUDDY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_INVALID (-1)
#define HB_ERR_CUDA (-2)
#define HB_ERR_UNSUPPORTED (-3)

#define HB_ABI_VERSION 1

/* Precision mode of the embedding conv stack. */
#define HB_EMBED_FP32 0 /* CUDA-core fp32 direct convolution (parity mode)          */
#define HB_EMBED_F16 1  /* tcgen05 kind::f16, fp16 operands, fp32 TMEM accumulators */

typedef struct hb_embed_model hb_embed_model; /* device-resident conv weights */
typedef struct hb_mlp_model hb_mlp_model;     /* device-resident classifier   */

/* ---- library ---------------------------------------------------------------------- */
int hb_abi_version(void);
const char* hb_last_error(void);
/* The tensor-core kernels bound their barrier waits and flag a timeout instead of hanging; this synchronises the device and reports
 * any such flag (HB_OK otherwise).  For tests and smoke runs. */
int hb_check_kernels(void);
/* Number of kernels this library has launched in this process (all streams); bench.py reports the
 * per-step delta as "gpu_launches". */
int64_t hb_launch_count(void);
/* Fills the constant tables (Hann window, FFT twiddles, mel filterbank) on the current
 * device.  hann_host: f32[512] (periodic Hann(400) zero-padded), melfb_host: f32[257*32]
 * row-major [bin][mel].  Must be called once per device before hb_mel_f32. */
int hb_init_tables(const float* hann_host, const float* melfb_host);

/* ---- K6: log-mel spectrogram ----------------------------------------------------------
 * Replaces MelSpectrogramModel.__call__ -> ORT session.run(mel-spectrogram.onnx)
 * (src/python/heybuddy/spectrogram.py:23-32) including the `/10 + 2` post-scale.
 *   audio_dev  f32 [B][audio_row_stride], first T samples of each row are used
 *   scale      multiplied into every sample on load (32767 for the Python path,
 *              embeddings.py:182; 1 if the caller already scaled)
 *   mel_dev    f32 [B][F][32], F = hb_mel_frames(T) = 1 + (T-512)/160
 */
int hb_mel_frames(int T);
int hb_mel_f32(const float* audio_dev, int64_t audio_row_stride, float scale, float* mel_dev,
               int B, int T, void* stream);

/* ---- K7: speech-embedding conv stack ----------------------------------------------------
 * Replaces SpeechEmbeddingModel.__call__ -> ORT session.run(speech-embedding.onnx)
 * (src/python/heybuddy/embeddings.py:32-42).
 *
 * hb_embed_create: weights_host is the 20 conv layers packed in table order, each as
 * kernel f32[kh][kw][cin][cout] followed by bias f32[cout] (274,440 floats).
 */
int hb_embed_create(hb_embed_model** out, const float* weights_host, int64_t n_floats);
int hb_embed_destroy(hb_embed_model* m);
int64_t hb_embed_num_params(void);

/* Ring-1 shape: n windows f32 [n][76][32] -> f32 [n][96]. */
int64_t hb_embed_windows_workspace_bytes(int n, int mode);
int hb_embed_windows(const hb_embed_model* m, int mode, const float* windows_dev, float* out_dev,
                     int n, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Whole-clip shape (fully convolutional): mel f32 [B][F][32] ->
 * f32 [B][n_slots][96] where slot s is the embedding of frames
 * [slot_offsets[s], slot_offsets[s]+76).  Offsets must be multiples of 4 and
 * offset + 76 <= F.  Evaluated once per clip instead of once per window. */
int64_t hb_embed_clips_workspace_bytes(int B, int F, int mode);
int hb_embed_clips(const hb_embed_model* m, int mode, const float* mel_dev, int B, int F,
                   const int32_t* slot_offsets_host, int n_slots, float* out_dev,
                   void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Parity hook: runs conv layers 0..layer (0..19) fully convolutionally over mel
 * f32 [B][F][32] and writes the activation after that layer (after its pool; pool phase 0)
 * as f32 NHWC [B][T_l][F_l][C_l] into out_dev (capacity in floats).  Returns the number of
 * floats written, or <0.  Same workspace size as hb_embed_clips. */
int64_t hb_embed_activation(const hb_embed_model* m, int mode, const float* mel_dev, int B, int F, int layer,
                            float* out_dev, int64_t out_capacity, void* workspace_dev, int64_t workspace_bytes,
                            void* stream);

/* ---- K1-K4: fused augmentation ------------------------------------------------------------
 * Replaces, per clip of an augmentation batch (execute_augment_batch,
 * src/python/heybuddy/dataset/augmented.py:363-392): torch_audiomentations AddColoredNoise
 * (:107-115) and Gain (:116-120) in mode="per_batch", torchaudio.functional.add_noise
 * (:272-276) and speechbrain reverberate (:388-392) -- in that order, one kernel, the clip
 * resident in shared memory.  Batch semantics (one coloured pattern / gain / RIR per batch,
 * consecutive noise rows) live in the per-clip descriptors the host fills from the draw table.
 */
typedef struct hb_clip_aug {
    int64_t noise_offset;  /* first sample of this clip's noise row in noise_bank_dev; < 0: no background */
    int32_t colored_index; /* row of colored_bases_dev ([n][16000] unit-RMS patterns);   < 0: none        */
    int32_t rir_index;     /* row of rir_spec_bank_dev ([n][T/2+1] complex spectra);     < 0: no reverb   */
    float colored_snr_db;
    float gain;            /* linear factor 10^(dB/20); 1 when Gain was not drawn                          */
    float noise_snr_db;
    int32_t reserved;
} hb_clip_aug;

/* Spectra of rotated RIRs for the exact-length circular convolution (once per RIR bank):
 * kernels_dev f32 [n][T] (rotated so the peak is at lag 0, zero padded; augment host code)
 * -> spec_dev f32 [n][T/2+1][2].  T must be even with T/2 = 2^a 3^b 5^c and T <= 23040. */
int hb_rir_spectrum(const float* kernels_dev, float* spec_dev, int n, int T, void* stream);
/* Coloured-noise patterns of k augmentation batches, generated on the device (torch_audiomentations AddColoredNoise
 * `_gen_noise`, call site augmented.py:107-115): for batch ids batch_ids_dev[i] the 16000-sample N(0,1) pattern comes from
 * stream 3 of the draw table's Philox4x32-10 counters (heybuddy_b200/dataset/draws.py: same seed, same counters, so the
 * host restatement and the device agree), is shaped by 1 / linspace(1, sqrt(8000), 8001)^f_decay between an rfft / irfft
 * pair and scaled to unit RMS -> out_dev f32 [k][16000].  Nothing but two k-entry arrays crosses PCIe. */
int hb_colored_bases(uint64_t seed, const int64_t* batch_ids_dev, const float* f_decay_dev, int k, float* out_dev, void* stream);
int hb_augment_clips_f32(const float* clips_dev,          /* f32 [n][T] length-fixed clips            */
                         const float* noise_bank_dev,     /* f32 noise stream or NULL                  */
                         const float* colored_bases_dev,  /* f32 [..][16000] or NULL                   */
                         const float* rir_spec_bank_dev,  /* from hb_rir_spectrum or NULL              */
                         const hb_clip_aug* params_dev,   /* [n] device array                          */
                         float* out_dev, int n, int T, void* stream);

/* ---- K9: the reference's per-clip numpy augmentations (audiomentations.Compose, augmented.py:79-90, applied at :325-328) ----
 * In place on the f32 [n][T] length-fixed clips, only on the clips listed in clip_index_dev (the ones whose coin came up).
 * hb_k9_eq_f32:   SevenBandParametricEQ = seven cascaded biquads, causal, zero initial state (scipy.signal.sosfilt); sos_dev f64
 *                 [k][7][5] = (b0, b1, b2, a1, a2) / a0 per section (computed on the host from the draws).  Each biquad runs in the
 *                 trapezoidal state-variable form (same transfer function, float32, within 4e-6 of float64 sosfilt).
 * hb_k9_tanh_f32: TanhDistortion = tanh(x * 0.5 / (percentile(|x|, 100 - 99 amount) + 1e-6)), RMS-matched to the input. */
int hb_k9_eq_f32(float* clips_dev, const int32_t* clip_index_dev, const double* sos_dev, int k, int T, void* stream);
int hb_k9_tanh_f32(float* clips_dev, const int32_t* clip_index_dev, const float* amount_dev, int k, int T, void* stream);

/* ---- K9: the two batch transforms ahead of AddColoredNoise / Gain (torch_audiomentations.Compose, augmented.py:93-106, applied at
 * :369-372; mode per_batch).  Both work in place on the clips listed in clip_index_dev.
 * hb_k9_bandstop_f32: BandStopFilter = x - julius.bandpass_filter(x): one FIR (difference of two Hann-windowed-sinc low-passes of
 *                 2 h + 1 taps, replicate padding) per batch, run as a partitioned overlap-save convolution on the clip's
 *                 exact-length FFT.  meta_dev i32 [k][3] = (first spectrum row, partitions, h) per clip; spec_dev = hb_rir_spectrum
 *                 of the partition rows (T / 2 taps each, zero-padded to T); scratch_dev f32 [k][T].
 * hb_k9_pitch_f32: PitchShift = torch_pitch_shift.pitch_shift: torch.stft(250, 7, rectangular) -> torchaudio phase vocoder ->
 *                 torch.istft -> torchaudio sinc resampling -> crop / zero-pad.  A plan holds one ratio's host-computed tables:
 *                 the vocoder's time steps (idx0 = floor(t), idx1 = floor(t + 1), alpha = t mod 1; float32, as the library
 *                 computes them), the phase advance per bin, and the resampling kernel f32 [up][2 width + orig]. */
int hb_k9_bandstop_f32(float* clips_dev, const int32_t* clip_index_dev, const int32_t* meta_dev, const float* spec_dev,
                       float* scratch_dev, int k, int T, void* stream);
typedef struct hb_pitch_plan hb_pitch_plan;
int hb_pitch_plan_create(hb_pitch_plan** out, int T, int n_fft, int hop, int frames_out, const int32_t* idx0_host,
                         const int32_t* idx1_host, const float* alpha_host, const float* phase_advance_host, int orig, int up,
                         int width, const float* kernel_host);
int hb_pitch_plan_destroy(hb_pitch_plan* plan);
int64_t hb_k9_pitch_workspace_bytes(const hb_pitch_plan* plan, int k);
int hb_k9_pitch_f32(const hb_pitch_plan* plan, float* clips_dev, const int32_t* clip_index_dev, int k, void* workspace_dev,
                    int64_t workspace_bytes, void* stream);

/* a1: int16 ragged clips -> length-fixed f32 [n][T] (to_target_length, augmented.py:200-232):
 * /32768, front-truncate when longer than T, otherwise zero-pad with pad_before[b] zeros on
 * the left.  samples_dev: concatenated int16 samples; offsets_dev i64[n+1]; pad_before_dev i32[n]. */
int hb_fix_length_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                      float* out_dev, int n, int T, void* stream);

/* a1 + K1-K4 in one kernel (T = 23040 only): the clip goes from the ragged int16 samples through the length fix into the
 * augmentation kernel's shared memory; the f32 [n][T] intermediate of hb_fix_length_i16 never reaches HBM.  Same arithmetic
 * as hb_fix_length_i16 followed by hb_augment_clips_f32 (bit-identical).  Other T: HB_ERR_UNSUPPORTED (use the two calls). */
int hb_augment_clips_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                         const float* noise_bank_dev, const float* colored_bases_dev, const float* rir_spec_bank_dev,
                         const hb_clip_aug* params_dev, float* out_dev, int n, int T, void* stream);

/* Production mode of the front end: a1 + K1-K4 + K6 in ONE kernel (T = 23040 only).  The augmented clip
 * never reaches HBM -- it goes from the augmentation's shared memory straight into the 141 log-mel frames (audio x scale,
 * MelSpectrogramModel's x/10 + 2 folded in) -> mel_dev f32 [n][141][32].  Bit-identical to hb_augment_clips_i16 followed by
 * hb_mel_f32(scale) (the parity-mode pair, which crosses the f32 [n][T] intermediate the reference crosses through the host:
 * augmented.py:416-421 -> embeddings.py:178-190).  Needs hb_init_tables.  Other T: HB_ERR_UNSUPPORTED. */
int hb_augment_mel_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                       const float* noise_bank_dev, const float* colored_bases_dev, const float* rir_spec_bank_dev,
                       const hb_clip_aug* params_dev, float scale, float* mel_dev, int n, int T, void* stream);

/* The whole featurization path of one chunk in one call (the fused entry SURVEY.md 8b ring 3 proposes): ragged int16 clips
 * -> length fix + augmentation + log-mel (hb_augment_mel_i16, audio x 32767) -> embeddings f32 [n][n_slots][96].  Equivalent to
 * hb_augment_clips_i16, hb_mel_f32(scale = 32767) and hb_embed_clips run back to back on `stream` (bit-identical); T = 23040. */
int64_t hb_featurize_workspace_bytes(int n, int T, int mode);
int hb_featurize_i16(const hb_embed_model* m, int mode, const int16_t* samples_dev, const int64_t* offsets_dev,
                     const int32_t* pad_before_dev, const float* noise_bank_dev, const float* colored_bases_dev,
                     const float* rir_spec_bank_dev, const hb_clip_aug* params_dev, const int32_t* slot_offsets_host,
                     int n_slots, float* out_dev, int n, int T, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ---- K8: wake-word classifier ---------------------------------------------------------------
 * Replaces WakeWordMLPModel.forward (src/python/heybuddy/wakeword.py:334-348) and the
 * loss/backward/Adam of WakeWordTrainer.train_epoch (trainer.py:405-462).
 * params are packed in state_dict order (heybuddy_b200/spec.py classifier_param_shapes).
 */
int64_t hb_mlp_num_params(void);
int hb_mlp_create(hb_mlp_model** out, const float* params_host, int64_t n_floats);
int hb_mlp_destroy(hb_mlp_model* m);
int hb_mlp_get_params(const hb_mlp_model* m, float* params_host, int64_t n_floats);
int hb_mlp_set_params(hb_mlp_model* m, const float* params_host, int64_t n_floats);
int64_t hb_mlp_workspace_bytes(int B, int training);
/* x_dev f32 [B][1536] -> prob_dev f32 [B]. */
int hb_mlp_forward(const hb_mlp_model* m, const float* x_dev, float* prob_dev, int B,
                   void* workspace_dev, int64_t workspace_bytes, void* stream);
/* One training step: forward, high-loss selection, weighted BCE (mean over selected
 * rows), backward, and -- unless fewer than min_selected rows were selected -- Adam.
 * stats_dev f32[4] = {loss, n_selected, stepped(0/1), high_loss_rate}. */
int hb_mlp_train_step(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B,
                      float lr, float negative_weight, float high_loss_threshold, int min_selected,
                      float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
/* The same step in data-parallel form (the reference trains on one device).  Every rank holds a replica
 * and a shard of the batch:
 *   hb_mlp_select    forward + selection; stats_dev[1] = rows selected in this shard
 *   (all-reduce SUM of stats_dev[1] over the ranks -> n_selected_total_dev)
 *   hb_mlp_backward  BCE / n_total and its gradients for this shard -> the model's gradient buffer; stats_dev[0] = this
 *                    shard's part of the mean loss, stats_dev[2] = 1 when n_total >= min_selected
 *   hb_mlp_grads_copy(to_model=0) -> all-reduce SUM -> hb_mlp_grads_copy(to_model=1)
 *   hb_mlp_adam      the Adam update (skipped when stats_dev[2] == 0), identical on every rank
 * With one rank and n_selected_total_dev = stats_dev + 1 this is exactly hb_mlp_train_step. */
int hb_mlp_select(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float high_loss_threshold,
                  float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
int hb_mlp_backward(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight,
                    float high_loss_threshold, const float* n_selected_total_dev, int min_selected,
                    const float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
int hb_mlp_grads_copy(hb_mlp_model* m, float* buf_dev, int64_t n_floats, int to_model, void* stream);
int hb_mlp_adam(hb_mlp_model* m, float lr, const float* stats_dev, void* stream);
/* The same with ONE collective per step: loss and gradients are linear in 1 / n_total, so a rank computes them unnormalised and the
 * division follows the exchange.
 *   hb_mlp_local_step      forward, selection, unnormalised loss sum and gradients of this shard ->
 *                          exchange_dev f32 [hb_mlp_num_params() + 2] = {gradients | loss sum | rows selected};
 *                          stats_dev[1] / [3] = this shard's selected rows / high-loss rate
 *   (all-reduce SUM of exchange_dev over the ranks)
 *   hb_mlp_apply_exchange  gradients (-> the model's gradient buffer) and loss divided by the global count, stats_dev[0..2] = {mean loss,
 *                          rows selected on all ranks, stepped}, Adam (skipped when the count is below min_selected)
 * Replaces trainer.py:405-462 on N devices; not available in the HB_MLP_STAGED / HB_MLP_FMA parity modes. */
int hb_mlp_local_step(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight, float high_loss_threshold,
                      float* prob_dev, float* stats_dev, float* exchange_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
int hb_mlp_apply_exchange(hb_mlp_model* m, const float* exchange_dev, float lr, int min_selected, float* stats_dev, void* stream);
/* The reference divides the loss of the step that finally fires by its accumulation counter (1 + the number of preceding steps
 * that selected fewer than 128 rows and were skipped, trainer.py:441-458): scale = 1 / accumulation_steps, applied to the loss
 * and its gradients of the following hb_mlp_train_step / hb_mlp_backward calls (default 1). */
int hb_mlp_set_loss_scale(hb_mlp_model* m, float scale);
/* Adam state = torch.optim.Adam's exp_avg / exp_avg_sq (packed like params) and step: `<name>_optimizer.pt` checkpoints and
 * Trainer.resume (trainer.py:54-118, 186-198). */
int hb_mlp_get_adam(const hb_mlp_model* m, float* exp_avg_host, float* exp_avg_sq_host, int* step_host, int64_t n_floats);
int hb_mlp_set_adam(hb_mlp_model* m, const float* exp_avg_host, const float* exp_avg_sq_host, int step, int64_t n_floats);
/* nn.Dropout(p) on the classifier input, train mode only (wakeword.py:197,338): y = x * keep / (1 - p) with keep drawn per
 * element from Philox4x32-10 (key = seed, counter = (element / 4, 7, call)); n a multiple of 4, 16-byte aligned buffers. */
int hb_mlp_dropout(const float* x_dev, float* y_dev, int64_t n, float p, uint64_t seed, uint64_t call, void* stream);
/* Gradients of the last hb_mlp_train_step / hb_mlp_backward (packed like params), for parity tests. */
int hb_mlp_get_grads(const hb_mlp_model* m, float* grads_host, int64_t n_floats);

/* config 5 (src/ts/src/hey-buddy.ts:350-413: every wake-word model runs on the SAME [16, 96] buffer): M models evaluated on
 * the same inputs in one launch chain whose length does not depend on M -- the shared (x - mean) * rstd, norm_in's affine folded
 * into ONE stacked [M*128, 1536] first-layer GEMM (hidden + gate of every model), then the 96-wide remainder batched over the
 * models.  x_dev [B][1536] -> prob_dev [M][B].  The models' live parameters are read on every call. */
int64_t hb_mlp_multi_workspace_bytes(int M, int B);
/* y = x W^T + b on tcgen05 with fp32 accuracy (three TF32 passes, csrc/gemm_tf32.cu): the product behind torch.nn.Linear in the gated
 * MLPs (wakeword.py:334-348).  x f32 [M][K], w f32 [N][K], bias f32 [N] or NULL, y f32 [M][N]; row strides lda / ldb / ldc in floats.
 * K a multiple of 32; N and the strides multiples of 4; 16-byte aligned pointers.  hb_mlp_forward / _train_step / _forward_multi use it
 * for their x W^T products (HB_MLP_FMA=1 keeps them on the fp32 FMA kernels). */
int hb_linear_tf32x3(const float* x_dev, int lda, const float* w_dev, int ldb, const float* bias_dev, float* y_dev, int ldc, int M, int N,
                     int K, void* stream);
int hb_mlp_forward_multi(hb_mlp_model* const* models, int M, const float* x_dev, float* prob_dev, int B,
                         void* workspace_dev, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEYBUDDY_B200_H */
