/*
 * encodec_b200.h -- C ABI of the B200-native EnCodec codec forward pass.
 *
 * The reference (ellen660/encodec, a fork of facebookresearch/encodec 0.1.2a3) is 100% Python: it has
 * no FFI / plugin interface for this path. The drop-in boundary is therefore its
 * Python nn.Module surface, mirrored by the `encodec_b200` package, which binds the entry points below
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add). Each entry point
 * cites the reference code whose arithmetic it replaces (paths relative to /root/reference/encodec).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `ecb_last_error()` then returns a
 *     thread-local message. Nothing here calls exit()/abort().
 *   - all tensor pointers are DEVICE pointers on the current CUDA device, fp32 unless noted,
 *     contiguous, in the REFERENCE's boundary layouts ([B, C, T] channels-first; codes int64).
 *     Channels-last layouts are internal to the library.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - `workspace` is caller-owned scratch of at least the size the matching *_workspace_bytes query
 *     returned; it may be reused across calls on the same stream.
 */
#ifndef ENCODEC_B200_H
#define ENCODEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECB_MAX_RATIOS 8

/* Hyper-parameters of SEANetEncoder/SEANetDecoder (modules/seanet.py:92-96,176-182) and
 * ResidualVectorQuantizer (quantization/vq.py:46-56) that reach the kernels. */
typedef struct ecb_spec {
  int32_t channels;             /* audio channels, 1 or 2 */
  int32_t causal;               /* 1: causal padding/trim rules, 0: asymmetric (conv.py:211-219,252-262) */
  int32_t group_norm;           /* norm after every conv (conv.py:38-52): 0 weight_norm (folded at load, Identity norm),
                                 * 1 time_group_norm = GroupNorm(1, C), 2 layer_norm = ConvLayerNorm (norm.py:16-30) */
  int32_t n_filters;            /* 32 */
  int32_t dimension;            /* latent dimension, 128 (EnCodec) or 256 (the fork's 10 Hz models) */
  int32_t n_ratios;             /* number of up/down-sampling stages */
  int32_t ratios[ECB_MAX_RATIOS]; /* decoder order, e.g. {8,5,4,2} or the fork's {6,5,5,2,1}; the encoder uses them reversed.
                                   * n_filters * 2^n_ratios (the LSTM width) must be 512 or 1024 */
  int32_t kernel_size;          /* 7 */
  int32_t last_kernel_size;     /* 7 */
  int32_t residual_kernel_size; /* 3 */
  int32_t compress;             /* 2 */
  int32_t lstm_layers;          /* 2 (0 disables the SLSTM) */
  int32_t bins;                 /* codebook size, 1024 */
  int32_t n_q;                  /* number of RVQ layers held (n_q_max) */
} ecb_spec;

typedef struct ecb_codec ecb_codec; /* opaque: prepared (folded, repacked) weights on one device */

const char* ecb_last_error(void);
int ecb_version(void);

/* ---- lifetime ------------------------------------------------------------------------------- */
int ecb_codec_create(const ecb_spec* spec, ecb_codec** out);
void ecb_codec_destroy(ecb_codec* codec);

/* ---- weights: replaces nn.Module.load_state_dict + the per-forward weight_norm hook ----------------
 * `key` is a reference state_dict key, e.g. "encoder.model.3.conv.conv.weight_g",
 * "decoder.model.6.convtr.convtr.weight_v", "encoder.model.13.lstm.weight_hh_l0",
 * "decoder.model.15.conv.conv.weight", "encoder.model.0.conv.norm.weight",
 * "quantizer.vq.layers.7._codebook.embed". `data` is a device pointer holding `numel` floats in the
 * reference's shape. Unknown keys that belong to training state (cluster_size, embed_avg, inited)
 * are accepted and ignored; any other unknown key is an error. */
int ecb_codec_load_tensor(ecb_codec* codec, const char* key, const float* data, int64_t numel, void* stream);
/* Folds weight-norm (conv.py:28-29: w = g * v / ||v||, norm over all dims but 0), repacks every
 * conv / convtr / LSTM weight into the kernels' layouts and precomputes ||E||^2 per codebook entry.
 * Fails if a tensor the spec requires was never loaded. */
int ecb_codec_finalize(ecb_codec* codec, void* stream);

/* Decoder operand scheme of the tensor-core convolutions. 1: one TF32 pass with operands rounded to TF32 by their
 * producers (nothing downstream of the decoder is discrete). 0: split-operand TF32 (3 products, fp32-accurate), like the
 * encoder. -1 (the state of a fresh codec): automatic -- weight-norm models (24 kHz) run one TF32 pass: decoded audio within
 * ~1e-4 max-abs / 2.5e-5 RMS of the reference on the golden cases and on real speech at three loudness levels (north_star
 * bar 1e-3 / 1e-4), decoder convs ~1.6x faster (ECB_DEC_SPLIT=3 in the environment turns this off); GroupNorm models
 * (48 kHz, whose output is scaled back to the segment's loudness) and LayerNorm models keep split operands, because TF32's
 * ~3e-4 relative error measured 1.3e-4 .. 1.6e-4 RMS there. With an explicit 1 a GroupNorm decoder runs its >= 128-channel
 * convs in TF32 (+10 % throughput, outside the RMS bar). The encoder and the quantiser are always fp32-accurate. */
int ecb_codec_set_decoder_precision(ecb_codec* codec, int32_t tf32_single_pass);

/* ---- SEANetEncoder.forward (modules/seanet.py:145-146; SConv1d conv.py:202-221; SLSTM lstm.py:22-28)
 * x: `n_items` items; item i starts at x + (i / n_seg) * x_batch_stride + (i % n_seg) * x_seg_stride,
 * channel c at + c * x_chan_stride, `length` samples each (this is how the 48 kHz model's overlapping
 * segments, model.py:168-170, are read in place). If `scale_out` != NULL the per-item loudness scale of
 * EncodecModel._encode_frame (model.py:180-185) is computed, written to scale_out[n_items] and the input
 * is divided by it on the fly. emb_out: [n_items, dimension, T_f] channels-first, T_f = ceil(length/hop).
 * If emb_frames_out != NULL the same values are also written frames-major [n_items * T_f, dimension]. */
size_t ecb_encoder_workspace_bytes(const ecb_codec* codec, int64_t n_items, int64_t length);
int ecb_encoder_forward(ecb_codec* codec, const float* x, int64_t n_items, int64_t n_seg, int64_t length,
                        int64_t x_batch_stride, int64_t x_seg_stride, int64_t x_chan_stride,
                        float* scale_out, float* emb_out, float* emb_frames_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- SEANetDecoder.forward (modules/seanet.py:242-253; SConvTranspose1d conv.py:241-263) -----------
 * Exactly one of z ([n_items, dimension, T_f] channels-first) / z_frames ([n_items*T_f, dimension]) is
 * non-NULL. out: [n_items, channels, T_f * hop]. If `scale` != NULL, out[i] *= scale[i]
 * (EncodecModel._decode_frame, model.py:244-245). */
size_t ecb_decoder_workspace_bytes(const ecb_codec* codec, int64_t n_items, int64_t n_frames);
int ecb_decoder_forward(ecb_codec* codec, const float* z, const float* z_frames, int64_t n_items,
                        int64_t n_frames, const float* scale, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- ResidualVectorQuantization.forward / encode (quantization/core_vq.py:385-432) with
 *      EuclideanCodebook.quantize / dequantize (core_vq.py:178-202) fused per layer ------------------
 * Stand-alone form: frames [n, dim] (frames-major), codebooks [n_q, bins, dim], e2 [n_q, bins] = ||E||^2
 * (from ecb_rvq_prepare). codes [n_q, n] int64; quantized [n, dim] or NULL; quantized_stack
 * [n_q, n, dim] or NULL (vq.py:80-89 intermediate_results). dim must be 128. */
int ecb_rvq_prepare(const float* codebooks, int64_t n_q, int64_t bins, int64_t dim, float* e2, void* stream);
int ecb_rvq_encode_frames(const float* frames, int64_t n, int64_t dim, const float* codebooks, const float* e2,
                          int64_t n_q, int64_t bins, int64_t* codes, float* quantized, float* quantized_stack,
                          void* stream);
/* ResidualVectorQuantization.decode (core_vq.py:434-445): quantized[n, dim] = sum_i E_i[codes[i, n]],
 * summed in layer order. */
int ecb_rvq_decode_frames(const int64_t* codes, int64_t n, int64_t dim, const float* codebooks, int64_t n_q,
                          int64_t bins, float* quantized, void* stream);
/* Same, on a codec's own codebooks (loaded through ecb_codec_load_tensor), in the reference's layouts:
 * x [B, dim, T] -> codes [n_q, B, T] int64, quantized [B, dim, T]; optional frames-major copy of
 * quantized for the decoder; optional stack [n_q, B, dim, T]. x_frames (frames-major) may be given
 * instead of x. */
size_t ecb_codec_rvq_workspace_bytes(const ecb_codec* codec, int64_t batch, int64_t n_frames);
int ecb_codec_rvq_forward(ecb_codec* codec, const float* x, const float* x_frames, int64_t batch,
                          int64_t n_frames, int64_t n_q, int64_t* codes, float* quantized,
                          float* quantized_frames, float* quantized_stack,
                          void* workspace, size_t workspace_bytes, void* stream);
int ecb_codec_rvq_decode(ecb_codec* codec, const int64_t* codes, int64_t batch, int64_t n_frames, int64_t n_q,
                         float* quantized, float* quantized_frames, void* stream);

/* ---- utils._linear_overlap_add (utils.py:17-56) ------------------------------------------------------
 * frames: [batch, n_seg, channels, seg_len]; segment s holds seg_lens[s] <= seg_len valid samples (the
 * rest of its row is ignored) and starts at output offset s * stride. seg_lens is a DEVICE int32 array.
 * The triangle weight is built for seg_len (the first frame's length), a shorter frame uses its head,
 * exactly as utils.py:45-54. out: [batch, channels, total],
 * total = stride * (n_seg - 1) + seg_lens[n_seg - 1]. Requires stride * 2 >= seg_len. */
int ecb_overlap_add(const float* frames, const int32_t* seg_lens, int64_t batch, int64_t channels, int64_t n_seg,
                    int64_t seg_len, int64_t stride, float* out, int64_t total, void* stream);

/* ---- .ecdc code stream without entropy coding: binary.BitPacker / BitUnpacker (binary.py:55-122) as used by
 * compress_to_file / decompress_from_file with use_lm=False (compress.py:66-89,130-155) ---------------------------
 * One frame: values are pushed time-major (for t: for k: codes[k][t]), `bits` each, least-significant bit first; the last
 * partial byte is zero-padded. codes element (k, t) is at codes[k * k_stride + t * t_stride] (int64, DEVICE). out / in are
 * DEVICE byte buffers of ecb_packed_bytes(...) bytes. Byte-exact with the reference's BitPacker. 1 <= bits <= 24. */
int64_t ecb_packed_bytes(int64_t n_codebooks, int64_t n_frames, int32_t bits);
int ecb_pack_codes(const int64_t* codes, int64_t k_stride, int64_t t_stride, int64_t n_codebooks, int64_t n_frames,
                   int32_t bits, uint8_t* out, void* stream);
int ecb_unpack_codes(const uint8_t* in, int64_t n_bytes, int64_t n_codebooks, int64_t n_frames, int32_t bits, int64_t* codes,
                     int64_t k_stride, int64_t t_stride, void* stream);

/* ---- layout helpers ([B, C, T] <-> [B, T, C]) -------------------------------------------------------- */
int ecb_transpose_bct_to_btc(const float* in, float* out, int64_t batch, int64_t chans, int64_t len, void* stream);
int ecb_transpose_btc_to_bct(const float* in, float* out, int64_t batch, int64_t len, int64_t chans, void* stream);

/* Diagnostic (tests only): the next encoder/decoder forward copies the channels-last activation produced by
 * `stage` (encoder: 0 first conv, 1+2i residual block i (after ELU), 2+2i down-sampling conv i, 50 LSTM (after
 * skip + ELU); decoder: 100 first conv, 101 LSTM, 102+2i transposed conv i, 103+2i residual block i) into
 * buf (at most `capacity` floats). buf == NULL disables the tap. */
void ecb_debug_tap(float* buf, int64_t capacity, int32_t stage);
/* Diagnostic: while buf != NULL the persistent tensor-core LSTM kernel (lstm_tc.cu) records clock64 stamps of CTA 0 for
 * steps 20..27 into buf ([3 roles][8 steps][16 events] int64, device memory); tools/lstm_trace.py prints them. */
void ecb_debug_lstm_trace(long long* buf);

/* The fp32-accurate convs run on fp16 PAIR operands (a = a1 + 2^-11 a2 with fp16 a1, a2: the accuracy of the split-TF32
 * scheme at half the tensor-core time; csrc/tc_conv.cu, SPLIT = 2). fp16 ends at +-65504: a model whose activations exceed
 * that would be clipped. Returns the number of operand tiles in which the conversion saturated since the last reset on the
 * current device (0 for every configuration of SURVEY section 8; a non-zero count means: run with ECB_F16_PAIR=0, the
 * split-TF32 scheme, which has the fp32 range). Synchronises the device. Replaces nothing in the reference (its CPU path is
 * plain fp32, /root/reference/encodec/modules/conv.py:116). */
int64_t ecb_f16_saturation_count(int32_t reset);

/* Diagnostic: the encoder's SLSTM alone on x [B][T][H] -> out [B][T][H] (tools/lstm_bench.py). */
size_t ecb_debug_lstm_workspace_bytes(const ecb_codec* codec, int64_t batch, int64_t T);
int ecb_debug_lstm(ecb_codec* codec, const float* x, float* out, int64_t batch, int64_t T, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Diagnostic (tests only): one launch of the tensor-core implicit-GEMM convolution (csrc/tc_conv.cu) on caller
 * buffers. a0 points at (item 0, sample a0_first, channel 0) of a channels-last activation from which a0_rows
 * samples per item are addressable (reads outside are zero); output row m reads samples
 * m*stride - pad_left ... + taps - 1 of source 0, then row m of the optional 1-tap source a1. w is [taps*C0 + C1][N]
 * (N contiguous). out_raw / out_elu point at (item 0, row 0) of [M][N] outputs with `halo` reflected rows written
 * before and after each item. split = 3: fp32-accurate split-operand TF32, 1: single TF32 pass. Synchronises. */
int ecb_debug_tc_conv(const float* a0, int64_t a0_item_stride, int32_t C0, int64_t a0_first, int64_t a0_rows,
                      int32_t taps, int32_t stride, int32_t pad_left, const float* a1, int64_t a1_item_stride,
                      int32_t C1, int64_t a1_rows, const float* w, const float* bias, int32_t N, int64_t M,
                      int32_t n_items, float* out_raw, float* out_elu, int64_t out_item_stride, int32_t halo,
                      int32_t round_out, int32_t split, void* stream);

/* ---- entropy-coded .ecdc stream: LM + arithmetic coder --------------------------------
 * compress_to_file / decompress_from_file with use_lm=True (compress.py:63-87,125-152): LMModel (model.py:45-83) over
 * StreamingTransformerEncoder (modules/transformer.py:62-119), build_stable_quantized_cdf and ArithmeticCoder /
 * ArithmeticDecoder (quantization/ac.py:18-260). Compression evaluates every step of every frame in one batched pass
 * and returns the coder's two cdf values per symbol; decompression keeps the whole step loop (LM step -> cdfs -> decoder
 * pull -> codes of the next step's input) on the device. Both produce bit-identical probabilities by construction. */
typedef struct ecb_lm_spec {
  int32_t n_q;            /* codebooks the model was built for (LMModel n_q, model.py:55) */
  int32_t card;           /* codebook cardinality (2..2048) */
  int32_t dim;            /* transformer dimension (even, 32..256) */
  int32_t n_layers;       /* StreamingTransformerEncoder num_layers */
  int32_t n_heads;        /* num_heads; dim / n_heads <= 32 */
  int32_t hidden;         /* int(dim * hidden_scale) (32..1024) */
  int32_t past_context;   /* rows of streaming state kept per layer (transformer.py:117), 1..1400 */
  float max_period;       /* create_sin_embedding max_period (transformer.py:16) */
} ecb_lm_spec;
typedef struct ecb_lm ecb_lm; /* opaque: LM weights on one device */

int ecb_lm_create(const ecb_lm_spec* spec, ecb_lm** out);
void ecb_lm_destroy(ecb_lm* lm);
/* key: a key of the reference's LMModel.state_dict() ("transformer.norm_in.weight", "transformer.layers.0.self_attn.
 * in_proj_weight", "emb.3.weight", "linears.3.bias", ...); data: DEVICE float32, copied. Replaces load_state_dict. */
int ecb_lm_load_tensor(ecb_lm* lm, const char* key, const float* data, int64_t numel, void* stream);
/* All tensors present? pos_divisor: optional HOST float32 [dim / 2] = max_period ** (j / (dim / 2 - 1)) as the caller's
 * framework rounds it (transformer.py:23 is a float32 pow); null: computed in double precision and rounded once. */
int ecb_lm_finalize(ecb_lm* lm, const float* pos_divisor, void* stream);
/* Streaming state of n_items independent streams of <= capacity steps (the reference's `states`, transformer.py:99-118,
 * kept as projected keys / values): float32 [n_layers][n_items][capacity + 1][2 * dim], caller-owned, any content. */
size_t ecb_lm_cache_bytes(ecb_lm* lm, int64_t n_items, int64_t capacity);
size_t ecb_lm_workspace_bytes(ecb_lm* lm, int64_t n_rows, int64_t n_codebooks);
/* LMModel.forward (model.py:65-83) for steps t0 .. t0 + n_t - 1 of n_items streams whose steps < t0 are in `cache`.
 * tokens (DEVICE int64): element (item, k, t) at tokens[item * item_stride + k * k_stride + t * t_stride];
 *   tokens_are_codes = 0: the LM's input indices of the n_t steps (1 + previous code, 0 = none), t counted from t0;
 *   tokens_are_codes = 1: the codes of the whole frame, t absolute (step t is fed 1 + code[t - 1], 0 at t = 0: compress.py:69,78).
 * Outputs (DEVICE, each may be null): probas float32 [n_items][n_t][K][card] (the reference returns it permuted to
 * [B, card, K, T]); cdf int32, same shape = build_stable_quantized_cdf(probas[..], 24, check=False) (ac.py:18-53);
 * sym_ranges int32 [n_items][n_t][K][2] (needs tokens_are_codes = 1) = (cdf[s - 1] or 0, cdf[s]) of the code s at
 * (item, k, t): the two numbers ArithmeticCoder.push reads (ac.py:143-144). workspace: ecb_lm_workspace_bytes(n_items * n_t, K). */
int ecb_lm_forward(ecb_lm* lm, const int64_t* tokens, int64_t item_stride, int64_t k_stride, int64_t t_stride,
                   int32_t tokens_are_codes, int64_t n_items, int64_t n_codebooks, int64_t t0, int64_t n_t, float* cache,
                   int64_t capacity, float* probas, int32_t* cdf, int32_t* sym_ranges, void* workspace,
                   size_t workspace_bytes, void* stream);
/* The decoding loop of decompress_from_file for ONE frame (compress.py:125-152), enqueued as a whole: for every step the LM,
 * the quantised cdfs, ArithmeticDecoder.pull for the K codebooks, codes[k][t]. data: DEVICE bytes of the stream, this
 * frame's coder starts at byte first_byte; codes: DEVICE int64 [K][n_steps]; cache: capacity >= n_steps, 1 item;
 * result: DEVICE int64 [2] = (status: 0 ok, 1 "The stream ended sooner than expected", 2 "Binary search failed",
 * 3 range overflow, 4 bad cdf; bytes of data consumed up to the end of this frame, i.e. where the next frame starts).
 * workspace: ecb_lm_workspace_bytes(1, K). */
int ecb_lm_decode_frame(ecb_lm* lm, const uint8_t* data, int64_t n_bytes, int64_t first_byte, int64_t n_codebooks,
                        int64_t n_steps, int64_t* codes, float* cache, int64_t capacity, int64_t* result, void* workspace,
                        size_t workspace_bytes, void* stream);
/* ArithmeticDecoder.pull on the DEVICE alone -- the warp-parallel decoder of ecb_lm_decode_frame's loop against given cdfs:
 * symbol i is decoded against cdfs[i * card .. (i + 1) * card) (DEVICE int32). symbols: DEVICE int64 [n]; result: DEVICE
 * int64 [8] (result[0] = status as above, result[1] = bytes consumed, the rest is scratch). Bit-exact with the reference. */
int ecb_ac_decode_device(const uint8_t* data, int64_t n_bytes, const int32_t* cdfs, int64_t n, int32_t card,
                         int32_t total_range_bits, int64_t* symbols, int64_t* result, void* stream);
/* build_stable_quantized_cdf (ac.py:18-53, check=False) alone, in the float32 arithmetic of the reference's CPU tensors:
 * pdf DEVICE float32 [n_rows][card] -> cdf DEVICE int32 [n_rows][card]. Bit-exact. */
int ecb_quantized_cdf(const float* pdf, int64_t n_rows, int32_t card, int32_t total_range_bits, int32_t* cdf, void* stream);
/* ArithmeticCoder (ac.py:56-166) on the HOST (no GPU involved): n symbols given as their cdf ranges [low, high_exclusive)
 * (sym_ranges: HOST int32 [n][2], e.g. copied back from ecb_lm_forward), pushed in order, then flush(). out: HOST buffer of
 * `capacity` bytes (4 * n + 16 always suffices); *out_len = the bytes the reference writes. Bit-exact. */
int ecb_ac_encode(const int32_t* sym_ranges, int64_t n, int32_t total_range_bits, uint8_t* out, int64_t capacity,
                  int64_t* out_len);
/* ArithmeticDecoder (ac.py:169-260) on the HOST -- the same function the device loop of ecb_lm_decode_frame runs: symbol i
 * is decoded against cdfs[i * card .. (i + 1) * card) (HOST int32). symbols: HOST int32 [n]; *bytes_consumed (may be null):
 * bytes read from data. Errors carry the reference's messages. */
int ecb_ac_decode(const uint8_t* data, int64_t n_bytes, const int32_t* cdfs, int64_t n, int32_t card,
                  int32_t total_range_bits, int32_t* symbols, int64_t* bytes_consumed);

/* Per-kernel-class timing for bench.py's roofline leg. Between ecb_profile_begin() and ecb_profile_end()
 * every launch of this library is bracketed by CUDA events on its stream; ecb_profile_end() synchronises
 * them and returns one entry per kernel class that ran: launch count, summed device milliseconds, and the
 * summed ALGORITHMIC flops / bytes of those launches (2*M*N*K of the dense contraction; activations +
 * weights read once, outputs written once). Returns the number of entries written. */
typedef struct ecb_prof_entry {
  char name[32];
  int64_t launches;
  double ms;
  double flops;
  double bytes;
} ecb_prof_entry;
void ecb_profile_begin(void);
int ecb_profile_end(ecb_prof_entry* out, int capacity);

/* Number of kernel launches issued by this library since process start (bench.py's gpu_launches). */
int64_t ecb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ENCODEC_B200_H */
