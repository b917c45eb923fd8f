/* snacb -- B200-native SNAC 24 kHz decode for Orpheus/Canopy audio-token streams.  C ABI.
 *
 * This is the drop-in boundary for ONE path of Demon-Sheriff/tts-inference: token window ->
 * SNAC codes -> VQ decode -> conv decoder -> int16 PCM.  The reference has no FFI of its own
 * (it is pure Python calling the pip package `snac`); every entry point below names the
 * reference code it replaces.  Plain pointers and sizes only, no torch / C++ types.
 *
 * Conventions
 *   - all functions return 0 on success, a negative snacb_status otherwise; they never abort
 *     and never throw; snacb_last_error(h) gives the text of the last failure on a handle;
 *   - pointers are DEVICE pointers unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - a handle is bound to one GPU and is not re-entrant: one caller at a time (the reference
 *     serialises callers with a global asyncio.Lock, vllm_inference/modal_audio_stream.py:83).
 */
#ifndef SNACB_H_
#define SNACB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNACB_VERSION 100

typedef struct snacb_handle_s* snacb_handle;
typedef struct snacb_batcher_s* snacb_batcher;
typedef struct snacb_ingest_s* snacb_ingest;

typedef enum {
    SNACB_OK = 0,
    SNACB_ERR_ARG = -1,      /* bad argument (null pointer, negative size, ...)            */
    SNACB_ERR_CUDA = -2,     /* a CUDA call failed; see snacb_last_error                    */
    SNACB_ERR_NO_GPU = -3,   /* no CUDA device / device is not sm_100                       */
    SNACB_ERR_STATE = -4,    /* call not valid in the handle's current state                */
    SNACB_ERR_NOMEM = -5
} snacb_status;

/* decode flags */
#define SNACB_RAW_IDS        0x1   /* tokens are raw LLM ids; subtract 128266 on device
                                      (modal_audio_stream.py:103,366).  Otherwise they are
                                      already `id - 128266`, as convert_to_audio receives them */
#define SNACB_EXTRACT_SLICE  0x2   /* keep samples [2048:4096] when more than 4096 were decoded
                                      (modal_audio_stream.py:94-95,195-198)                    */
#define SNACB_FP32           0x4   /* fp32 CUDA-core arithmetic end to end (config 1).  Default: tensor-core
                                      contractions (tcgen05, kind::f16) with fp16 operands, fp32 accumulate  */
#define SNACB_KEEP_TAPS      0x8   /* debug: keep every stage's output for snacb_debug_tap     */
#define SNACB_STREAM_FP32    0x10  /* tensor-core path: keep the residual stream in fp32 between kernels  */
#define SNACB_BF16           0x20  /* tensor-core path: every contraction as a bf16 tcgen05.mma on exactly split operands
                                      (A = hi + lo, W = hi + lo: three MMAs), activations stored in fp16 -- the bf16x3 path,
                                      47.7 dB; one kernel per layer, ~3x slower than the default (DESIGN.md "precision")   */
#define SNACB_UNFUSED        0x40  /* tensor-core path: one kernel per layer instead of the fused
                                      NoiseBlock + ResidualUnit chain (A/B checks, per-stage taps)          */

/* Folded fp32 weights of the decode half of snac_24khz, host pointers.  Weight-norm is already
 * folded (w = g * v / ||v||, norm over dim 0: per OUTPUT channel for Conv1d, per INPUT channel for
 * ConvTranspose1d); the reference recomputes that on every forward (snac WNConv1d), we fold once at
 * load (init_snac, modal_audio_stream.py:106-129).  Layouts are the PyTorch ones. */
typedef struct {
    const float* alpha1;   /* [C]        ResidualUnit.block.0 Snake alpha            */
    const float* dw_w;     /* [C][1][7]  ResidualUnit.block.1 depthwise conv weight  */
    const float* dw_b;     /* [C]                                                     */
    const float* alpha2;   /* [C]        ResidualUnit.block.2 Snake alpha            */
    const float* pw_w;     /* [C][C][1]  ResidualUnit.block.3 1x1 conv weight         */
    const float* pw_b;     /* [C]                                                     */
} snacb_resunit_weights;

typedef struct {
    const float* alpha;    /* [Cin]           DecoderBlock.block.0 Snake alpha        */
    const float* convt_w;  /* [Cin][Cout][2s] DecoderBlock.block.1 ConvTranspose1d    */
    const float* convt_b;  /* [Cout]                                                  */
    const float* noise_w;  /* [Cout][Cout][1] DecoderBlock.block.2 NoiseBlock.linear (no bias) */
    snacb_resunit_weights res[3];   /* dilations 1, 3, 9                              */
} snacb_block_weights;

typedef struct {
    const float* codebook[3];    /* [4096][8]     quantizer.quantizers.i.codebook.weight   */
    const float* out_proj_w[3];  /* [768][8][1]   quantizer.quantizers.i.out_proj (folded) */
    const float* out_proj_b[3];  /* [768]                                                   */
    const float* stem_dw_w;      /* [768][1][7]   decoder.model.0                           */
    const float* stem_dw_b;      /* [768]                                                   */
    const float* stem_pw_w;      /* [1024][768][1] decoder.model.1                          */
    const float* stem_pw_b;      /* [1024]                                                  */
    snacb_block_weights block[4];/* decoder.model.2..5: 1024->512 s8, ->256 s8, ->128 s4, ->64 s2 */
    const float* tail_alpha;     /* [64]          decoder.model.6                           */
    const float* tail_w;         /* [1][64][7]    decoder.model.7                           */
    const float* tail_b;         /* [1]                                                     */
} snacb_weights;

int snacb_version(void);

/* Replaces init_snac() (modal_audio_stream.py:106-129) / load_models (tensorrt_tts/inference.py:152-165):
 * uploads and packs the weights on GPU `device` (fp32 copies, bf16 K-major copies for the tensor-core
 * path, ConvTranspose weights re-indexed per output phase).  Fails with SNACB_ERR_NO_GPU when there
 * is no sm_100 device -- there is no CPU fallback. */
int snacb_create(snacb_handle* out, const snacb_weights* w, int device);
void snacb_destroy(snacb_handle h);
const char* snacb_last_error(snacb_handle h);   /* h may be NULL: error of the last failed create */

/* Token -> code redistribution only (modal_audio_stream.py:156-188; redistribute_codes,
 * tensorrt_tts/inference.py:54-93).  tok is [B][ntok] int32, the first 7*(ntok/7) entries of each row
 * are used; c0 [B][F], c1 [B][2F], c2 [B][4F] int32 with F = ntok/7.  Bit-exact. */
int snacb_unpack(snacb_handle h, const int32_t* tok, int B, int ntok, int flags,
                 int32_t* c0, int32_t* c1, int32_t* c2, void* stream);

/* The whole helper, batched: convert_to_audio (modal_audio_stream.py:132-202) / decode_snac
 * (tensorrt_tts/inference.py:96-112) for B independent token rows of `frames` frames each.
 *   tok        [B][tok_stride] int32 (tok_stride >= 7*frames)
 *   noise      NULL -> NoiseBlock noise from the built-in counter RNG keyed by (seed, stream, block, t);
 *              else 4 device pointers, float [B][frames*4*{8,64,256,512}] (one scalar per (stream, t))
 *   pcm        int16 [B][n] with n = 2048 if SNACB_EXTRACT_SLICE and frames*2048 > 4096, else frames*2048
 *   wave       optional float [B][n], the tanh output before quantisation (parity checks)
 * Asynchronous on `stream`; no host synchronisation inside. */
int snacb_decode(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                 const float* const* noise, uint64_t seed, int16_t* pcm, float* wave, void* stream);

/* snacb_decode with the built-in NoiseBlock noise of row i keyed by stream_keys[i] (device int32 [B], >= 0) instead of
 * by the row's position in the batch: a stream then draws the same noise for the same time step whichever rows share
 * its launch and however many frames are decoded -- what a policy that re-decodes growing prefixes of many streams
 * needs (tts_inference_b200/policy.py).  NULL keys = snacb_decode. */
int snacb_decode_keyed(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                       const float* const* noise, uint64_t seed, const int32_t* stream_keys, int16_t* pcm, float* wave,
                       void* stream);

/* snacb_decode_keyed that writes only samples [sample_lo, sample_hi) of every row (pcm / wave are [B][sample_hi -
 * sample_lo]; SNACB_EXTRACT_SLICE is ignored) and, on the tensor-core path, computes only their receptive field in
 * blocks 1-3 (the dead-sample trimming of the sliced call, for an arbitrary range).  Bit-identical to slicing the full
 * decode.  A policy that re-decodes a growing prefix and emits only the new stable samples pays for those samples, not
 * for the whole prefix (block 0 and the stem, ~10 % of a decode, are still computed in full). */
int snacb_decode_range(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                       const float* const* noise, uint64_t seed, const int32_t* stream_keys, int sample_lo, int sample_hi,
                       int16_t* pcm, float* wave, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stateful streaming session: incremental decode of growing streams with per-stream,
 * per-stage state kept in HBM, for the streaming policies the reference documents (re-decode of the whole prefix every
 * N frames with lookahead, tensorrt_tts/PIPELINE_REPORT.md:475-511; sliding 28/7 windows,
 * vllm_inference/modal_audio_stream.py:86-95, 352-396).  A step that adds k frames to a stream computes, in every
 * stage, only the rows that became FINAL with them (their whole receptive field lies inside the known tokens) and emits
 * the samples that became final: 4k latent steps of work, no recompute of the prefix, and -- the NoiseBlock noise being
 * keyed by (seed, block, stream key, t) -- the concatenated output is BIT-IDENTICAL to one snacb_decode_keyed of the
 * finished stream.  Non-final samples lag the newest token by the decoder's receptive field (2.5 frames = 5050 samples), not by a
 * fixed 5-frame lookahead.
 *   n_slots, window_frames  every slot keeps a SLIDING WINDOW of its stream's activations: window_frames frames (rounded
 *                         up to a multiple of 32; ~1.45 MB per frame per slot, snacb_session_bytes).  When a step does
 *                         not fit, the window keeps its last 8 frames (more than the deepest stage's lag plus every
 *                         stage's halo), moves them to the front and goes on: streams are UNBOUNDED in length, memory is
 *                         per window.  A step may add at most window_frames - 16 frames to a non-empty window.
 *                         flags: SNACB_RAW_IDS, SNACB_BF16.
 *   snacb_session_step    appends new_frames frames (new_tok [n][tok_stride] int32, device) to slots
 *                         [slot0, slot0 + n), which must all be at the same position, and writes each slot's newly
 *                         final samples to pcm [n][pcm_stride] (device int16); *n_emitted = samples per slot (the same
 *                         for all n; snacb_session_next_emit tells it in advance).  final != 0 ends the streams: everything
 *                         up to 2048 * frames is emitted (the last receptive field sees the true end's zero padding and is
 *                         decoded by one stateless ranged decode of the window's tokens) and the slots need
 *                         snacb_session_reset before reuse.  stream_keys [n] (device) as in snacb_decode_keyed; NULL =
 *                         the slot index.  Asynchronous on `stream`.
 * --------------------------------------------------------------------------------------------- */
typedef struct snacb_session_s* snacb_session;
int snacb_session_create(snacb_handle h, int n_slots, int window_frames, int flags, snacb_session* out);
void snacb_session_destroy(snacb_session s);
int64_t snacb_session_bytes(snacb_session s);
int snacb_session_max_frames(snacb_session s);               /* the window, in frames (after rounding) */
int snacb_session_reset(snacb_session s, int slot0, int n);
int64_t snacb_session_frames(snacb_session s, int slot);     /* frames ingested so far (whole stream) */
int64_t snacb_session_emitted(snacb_session s, int slot);    /* samples emitted so far (whole stream) */
int snacb_session_next_emit(snacb_session s, int slot, int new_frames, int final);
int snacb_session_step(snacb_session s, int slot0, int n, const int32_t* new_tok, int tok_stride, int new_frames, int final,
                       uint64_t seed, const int32_t* stream_keys, int16_t* pcm, int pcm_stride, int* n_emitted,
                       void* stream);
/* The same step for an ARBITRARY set of slots (slots_host [n], any order, need not be contiguous) that may be at DIFFERENT
 * positions of their streams: past a stream's first three frames every stage's frontier is affine in the frame count, so
 * all streams that advance by the same number of frames share one launch sequence, each with its own row offset, buffer
 * slot and window origin (per-stream maps read by every kernel).  This is the call for asynchronous streams: one step
 * serves every stream that has `new_frames` new frames, wherever it is.  Streams holding fewer than 3 frames can only be
 * grouped with streams at exactly the same position.  Not for end of stream (use snacb_session_step with final != 0).
 * new_tok [n][tok_stride] (device), stream_keys [n] (device) or NULL = slot indices, pcm [n][pcm_stride]. */
int snacb_session_step_multi(snacb_session s, int n, const int32_t* slots_host, const int32_t* new_tok, int tok_stride,
                             int new_frames, uint64_t seed, const int32_t* stream_keys, int16_t* pcm, int pcm_stride,
                             int* n_emitted, void* stream);
/* Host-side frontier table of a session (no GPU needed): the number of FINAL rows of every stage once `frames` frames are
 * known, snac_24khz strides 8/8/4/2; bit bi of chain_mask = block bi runs the fused chain kernel.  out (>= 22 ints):
 * stem, then per block {ConvTranspose out, NoiseBlock out, ResidualUnit 0/1/2 out} (the last three 0 for a chain block
 * except res2 = the block output), then the emitted samples.  Returns the number of ints written. */
int snacb_debug_session_frontier(int frames, int chain_mask, int32_t* out, int cap);

/* Same with HOST buffers (pinned or pageable): copies the tokens in, decodes, copies the PCM out and
 * synchronises -- the boundary the reference's helper has (torch.tensor(..., device=) in,
 * .cpu().numpy().tobytes() out; modal_audio_stream.py:176-202). */
int snacb_decode_host(snacb_handle h, const int32_t* tok_host, int B, int tok_stride, int frames, int flags,
                      uint64_t seed, int16_t* pcm_host);

/* The same boundary, pipelined for a serving loop: _submit queues copy-in + decode + copy-out and returns; _wait
 * blocks until the OLDEST outstanding submit has its PCM in `pcm_host`.  At most two submits may be outstanding.
 * Calling submit(step i+1) and then wait() (for step i) every step overlaps the device->host copy of step i (copy stream) with the
 * decode of step i+1; each step still pays its own host<->device copies.  Use pinned host buffers (pageable ones make
 * the copies synchronous), and a distinct `pcm_host` for the two steps in flight. */
int snacb_decode_host_submit(snacb_handle h, const int32_t* tok_host, int B, int tok_stride, int frames, int flags,
                             uint64_t seed, int16_t* pcm_host);
int snacb_decode_host_wait(snacb_handle h);

/* Number of PCM samples per stream that snacb_decode writes for (frames, flags). */
int snacb_samples_out(int frames, int flags);

/* Workspace policy: streams are processed in groups sized so that one activation buffer stays under
 * `bytes` (default 1 GiB = 1024 four-frame windows per group; three such buffers are allocated on demand).
 * Larger groups amortise the 24 launches of the pipeline; measured on B200, L2 residency of a small group
 * does not pay for its launch overhead.  0 keeps the current value. */
int snacb_set_group_bytes(snacb_handle h, size_t bytes);

/* Counters since creation: kernels launched by this library, streams decoded. */
int snacb_stats(snacb_handle h, uint64_t* kernel_launches, uint64_t* streams_decoded);

/* Which formulation each DecoderBlock's fused chain kernel runs with fp16 operands, decided from the checkpoint when the
 * handle is created: modes[0..3] = 0 per-layer kernels (block 0, C = 512), 1 general variant (Snake evaluated as
 * x + (alpha + 1e-9)^-1 sin^2(alpha x) in fp32), 2 alpha-folded variant (alpha multiplied into neighbouring weights; only
 * when every Snake alpha of the block has magnitude in [2^-8, 2^6] and the folded parameters fit fp16).  A checkpoint
 * with alpha = 0, tiny, or huge values therefore still decodes correctly (alpha = 0 gives snake(x) = x as in the
 * reference), on the slower general variant. */
int snacb_chain_modes(snacb_handle h, int32_t* modes);

/* Per-launch CUDA-event timing of the pipeline stages (measurement aid; adds two event records per
 * launch, so do not time a headline number with it on).  snacb_profile(h,1) starts a fresh recording,
 * snacb_profile_report synchronises and writes one line per stage: "<name> <launches> <total_ms>". */
int snacb_profile(snacb_handle h, int enable);
int snacb_profile_report(snacb_handle h, char* buf, size_t cap);

/* Debug taps (SNACB_KEEP_TAPS): stage outputs of the LAST group decoded, converted to fp32,
 * channel-last [rows][cols].  names: "stem_dw","stem","b{0..3}.convt","b{i}.noise","b{i}.res{0..2}". */
int snacb_debug_tap_count(snacb_handle h);
int snacb_debug_tap_info(snacb_handle h, int idx, char* name, int name_cap, int64_t* rows, int64_t* cols);
int snacb_debug_tap_copy(snacb_handle h, int idx, float* dst_host, size_t dst_elems);

/* Host-side schedule of the fused NoiseBlock + ResidualUnit chain kernel (C = 64, 128 or 256 channels): for each
 * of the 3 ResidualUnits (dilation 1, 3, 9) and each of up to 16 warps, up to 4 spans {first_row, octets, chunk};
 * a span covers rows first_row + k*dilation, k < 8*octets, of one 64-channel chunk.  out: int16[3][16][4][3].
 * Returns (tile height in rows incl. the 40-row halo either side) | (warps per CTA << 16), or SNACB_ERR_ARG.
 * No GPU needed. */
int snacb_debug_chain_spans(int C, int16_t* out, int cap);
/* The schedule of any tile type (kernels_chain.cu): own_end = first tile row past the rows the tile owns (0: a full
 * tile; smaller for the SHORT last tile of a row range, which skips everything past its right halo); carry_top = 1: the
 * tile is not the first of its strip, owns its rows from row 0 on and takes the three class rows above row 0 from the
 * previous tile (top spans have field 4 = 1 + steps above row 0).  out: int16[3][16][4][4] {first_row, octets, chunk,
 * top}.  Returns as above; 0 when such a tile does not fit the table (the decoder then falls back to full / halo tiles). */
int snacb_debug_chain_spans_ex(int C, int own_end, int carry_top, int16_t* out, int cap);
/* Strip plan of the chain kernel for t_n rows per stream, S streams, `slots` CTA slots: out4 = {tiles per strip, strips
 * per stream, tiles of a stream's last strip, rows owned by its last tile (0 = full)}.  No GPU needed. */
int snacb_debug_chain_plan(int C, int t_n, int S, int slots, int32_t* out4);

/* The same for the warp-specialised, block-pipelined chain kernel (kernels_chain_ws.cu; C = 64 or 128, enabled with
 * SNACB_CHAIN_WS=1): spans live inside ONE 128-row block and count QUADS (4 steps): {first_row (block-relative),
 * quads, chunk} for each of the 8 prologue warps.  Returns (tile height in rows) | (prologue warps << 16). */
int snacb_debug_chain_ws_spans(int C, int16_t* out, int cap);
/* 1 when the library was built with SNACB_EXPERIMENTS=1 (python -m tts_inference_b200.build): k_chain_ws is then compiled
 * and SNACB_CHAIN_WS=1 selects it.  The default build leaves measured-and-dropped kernel variants out of the product. */
int snacb_experiments_built(void);

/* ---------------------------------------------------------------------------------------------
 * Streamer: the multi-stream front of the stateful session, what snacb_batcher_* is for window decodes.  Producers push
 * token ids per stream (THREAD-SAFE, also while a tick runs); tick() -- single caller -- hands every stream's new whole
 * frames to one snacb_session_step_multi per distinct frame count (streams at different positions share the launches),
 * flushes the streams that ended, and returns the newly final samples: chunk i belongs to stream ids[i], its PCM is at
 * pcm_host + offsets[i], lengths[i] samples; a stream's chunks over successive ticks concatenate to the batch decode of
 * its tokens, bit for bit.  A stream holds one of max_streams session slots from its first push until the tick that
 * flushes it; a push for a new stream when all slots are taken returns SNACB_ERR_STATE.  min_frames: new whole frames a
 * stream needs before a tick serves it (1 = every frame, the reference's sliding cadence).  Replaces the per-stream
 * Python buffering of stream_audio (modal_audio_stream.py:340-409) for any number of concurrent streams.
 * --------------------------------------------------------------------------------------------- */
typedef struct snacb_streamer_s* snacb_streamer;
int snacb_streamer_create(snacb_streamer* out, snacb_handle h, int max_streams, int window_frames, int flags, int min_frames);
void snacb_streamer_destroy(snacb_streamer s);
int snacb_streamer_push(snacb_streamer s, uint64_t stream_id, const int32_t* tokens_host, int n);
int snacb_streamer_end(snacb_streamer s, uint64_t stream_id);
int snacb_streamer_active(snacb_streamer s);            /* streams holding a slot */
/* Returns the number of chunks written (>= 0) or a negative status. */
int snacb_streamer_tick(snacb_streamer s, uint64_t seed, int max_chunks, uint64_t* ids, int64_t* offsets, int32_t* lengths,
                        int16_t* pcm_host, size_t pcm_capacity);

/* ---------------------------------------------------------------------------------------------
 * SNAC ENCODE path: audio -> codes, the other half of the codec.  The reference never
 * calls it at inference (it only decodes); upstream it is snac.SNAC.encode = preprocess (right-pad to a multiple of
 * 2048 samples) -> Encoder (conv k7 1 -> 48; 4 EncoderBlocks: 3 ResidualUnits d = 1/3/9, Snake, strided conv k = 2s,
 * s = 2/4/8/8, width doubling to 768; depthwise conv k7) -> ResidualVectorQuantize (3 levels, strides 4/2/1: avg-pool,
 * in_proj 768 -> 8, nearest L2-normalised code of 4096, residual -= out_proj(code)).  fp32 CUDA-core kernels
 * (csrc/encoder.cu): the codes are an argmax, so the latent has to track the fp32 reference.  Oracle:
 * oracle/snac_enc_ref.py.  Weight-norm folded, PyTorch layouts, host pointers.
 * --------------------------------------------------------------------------------------------- */
typedef struct snacb_encoder_s* snacb_encoder;
typedef struct {
    const float* alpha;    /* [C]            EncoderBlock.block.3 Snake alpha (C = block input width)   */
    const float* conv_w;   /* [2C][C][2s]    EncoderBlock.block.4 strided conv                         */
    const float* conv_b;   /* [2C]                                                                     */
    snacb_resunit_weights res[3];   /* EncoderBlock.block.0..2, width C, dilations 1, 3, 9            */
} snacb_encblock_weights;
typedef struct {
    const float* conv0_w;        /* [48][1][7]    encoder.block.0                                      */
    const float* conv0_b;        /* [48]                                                               */
    snacb_encblock_weights block[4];   /* encoder.block.1..4: 48 -> 96 s2, -> 192 s4, -> 384 s8, -> 768 s8 */
    const float* final_w;        /* [768][1][7]   encoder.block.5 (depthwise)                          */
    const float* final_b;        /* [768]                                                              */
    const float* in_proj_w[3];   /* [8][768][1]   quantizer.quantizers.i.in_proj                       */
    const float* in_proj_b[3];   /* [8]                                                                */
    const float* codebook[3];    /* [4096][8]                                                          */
    const float* out_proj_w[3];  /* [768][8][1]                                                        */
    const float* out_proj_b[3];  /* [768]                                                              */
} snacb_encoder_weights;
int snacb_encoder_create(snacb_encoder* out, const snacb_encoder_weights* w, int device);
void snacb_encoder_destroy(snacb_encoder e);
const char* snacb_encoder_last_error(snacb_encoder e);   /* e may be NULL: error of the last failed create */
uint64_t snacb_encoder_launches(snacb_encoder e);
/* Frames (= codes of level 0) that n_samples of audio encode to: ceil(n / 2048). */
int snacb_encode_frames(int n_samples);
/* audio float [B][audio_stride] (device), the first n_samples of each row are used, right-padded with zeros to
 * F = snacb_encode_frames(n_samples) frames.  c0 [B][F], c1 [B][2F], c2 [B][4F] int32.  Optional outputs (tests):
 * latent [B][4F][768] = the encoder output z (channel-last), best_dist [B][F + 2F + 4F] = winning distance per code,
 * level by level.  Asynchronous on `stream`. */
int snacb_encode(snacb_encoder e, const float* audio, int B, int n_samples, int audio_stride, int32_t* c0, int32_t* c1,
                 int32_t* c2, float* latent, float* best_dist, void* stream);
/* codes -> the 7 token ids per frame the LLM vocabulary uses (inverse of snacb_unpack; modal_audio_stream.py:156-188):
 * tok [B][7 * frames] = code + 4096 * position (+ 128266 with SNACB_RAW_IDS). */
int snacb_pack_tokens(const int32_t* c0, const int32_t* c1, const int32_t* c2, int B, int frames, int flags, int32_t* tok,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batcher: the multi-stream replacement of stream_audio's per-stream buffer policy
 * (modal_audio_stream.py:352-396), which decodes one stream at a time under a global lock.
 * Many producers push token ids; flush() packs every ready window of every stream into ONE decode.
 *   policy 0 (chunk)  : the shipped rule -- pop 28 codes, decode, emit all 8192 samples; at end of
 *                       stream decode the remaining whole frames.
 *   policy 1 (sliding): the rule the constants describe (modal_audio_stream.py:86-95) -- once 28 codes
 *                       are buffered, every 7 new codes decode the last 28 and emit samples [2048:4096].
 * --------------------------------------------------------------------------------------------- */
/* h may be NULL: the batcher then only queues (push / end / take), flushing returns SNACB_ERR_STATE. */
int snacb_batcher_create(snacb_batcher* out, snacb_handle h, int policy, int flags, int max_windows);
void snacb_batcher_destroy(snacb_batcher b);
/* Append n raw token ids (flags & SNACB_RAW_IDS) or codes to a stream.  THREAD-SAFE: any number of producer threads may
 * push concurrently, also while a flush runs (streams are sharded over independently locked tables). */
int snacb_batcher_push(snacb_batcher b, uint64_t stream_id, const int32_t* tokens_host, int n);
/* Mark a stream finished: policy 0 queues its remaining whole frames.  The id stays known as ended -- a later push to it
 * returns SNACB_ERR_STATE instead of silently starting a new stream -- until snacb_batcher_forget(id) releases it. */
int snacb_batcher_end(snacb_batcher b, uint64_t stream_id);
int snacb_batcher_forget(snacb_batcher b, uint64_t stream_id);
/* Decode every ready window in one launch sequence per chunk length, STRAIGHT into pcm_host (use pinned memory: the
 * copy-out is then an asynchronous DMA).  Returns the number of chunks produced (>= 0) or a negative status.  Chunk i:
 * stream ids[i], PCM at pcm_host + offsets[i], lengths[i] samples; chunks of one stream appear in time order.  If the
 * decode fails the windows go back to the head of their queues: nothing is lost, the call may be retried.
 * The flush calls are SINGLE-CALLER (one flusher thread); producers keep pushing meanwhile.
 *   snacb_batcher_flush         blocking: submit + wait
 *   snacb_batcher_flush_submit  returns once copy-in + decode + copy-out are queued on the GPU (at most two outstanding);
 *                               ids / offsets / lengths are final on return, the PCM is not
 *   snacb_batcher_flush_wait    blocks until the OLDEST outstanding submit's PCM is in its pcm_host
 * A serving loop that calls submit(tick i + 1) and then wait() (tick i) overlaps the copy-out of one tick with the
 * decode of the next; give the two ticks in flight distinct pcm_host buffers. */
int snacb_batcher_flush(snacb_batcher b, uint64_t seed, int max_chunks, uint64_t* ids, int64_t* offsets,
                        int32_t* lengths, int16_t* pcm_host, size_t pcm_capacity);
int snacb_batcher_flush_submit(snacb_batcher b, uint64_t seed, int max_chunks, uint64_t* ids, int64_t* offsets,
                               int32_t* lengths, int16_t* pcm_host, size_t pcm_capacity);
int snacb_batcher_flush_wait(snacb_batcher b);
int snacb_batcher_pending(snacb_batcher b);     /* windows ready to decode right now */
/* Pop up to max_windows ready windows WITHOUT decoding them (a caller that routes windows itself, e.g. into
 * snacb_decode_keyed): ids[i], frames[i] (4, or 1..3 for an end-of-stream remainder), tok[i][28] zero padded. */
int snacb_batcher_take(snacb_batcher b, int max_windows, uint64_t* ids, int32_t* frames, int32_t* tok);

/* ---------------------------------------------------------------------------------------------
 * Device-side token ingest: the LLM loop's sampled token ids go from a device
 * tensor straight into ready decode windows -- no Python ints, no per-token host work.  Per stream it restates
 *   generate_audio_tokens  (modal_audio_stream.py:313-333; same rule tensorrt_tts/inference.py:231-241):
 *                          skip everything up to and including the first TOKEN_SOS (128257), stop at the first
 *                          TOKEN_EOS (128258), pass every other id on;
 *   stream_audio           (modal_audio_stream.py:352-396): whenever 28 ids are buffered pop them as one window;
 *                          when the stream ends emit the remaining whole frames (1..3) and drop the rest.
 * Stream slots are indices 0..max_streams-1; a slot is reusable after snacb_ingest_reset.  Ids stay RAW (decode the
 * windows with SNACB_RAW_IDS).  Integer work, bit-exact against oracle/ingest_ref.py and the golden vectors
 * produced by the reference's own two functions (tests/golden/make_golden_ingest.py).
 * --------------------------------------------------------------------------------------------- */
int snacb_ingest_create(snacb_ingest* out, int device, int max_streams);
void snacb_ingest_destroy(snacb_ingest g);
/* Forget streams [first, first + n): state back to "waiting for TOKEN_SOS", buffer emptied. */
int snacb_ingest_reset(snacb_ingest g, int first, int n, void* stream);
/* Rows of win_tok / win_stream that one step over S streams x n_tok tokens can fill. */
int snacb_ingest_window_capacity(int S, int n_tok);
/* One step of the LLM loop for streams 0..S-1.
 *   tok          [S][n_tok] int32 token ids sampled this step (device)
 *   n_valid      optional [S]: how many of a stream's n_tok ids are real (NULL: all)
 *   finish       optional [S] bytes: the stream's generator ended without TOKEN_EOS (max_tokens) -> flush it
 *   win_tok      [win_cap][28], win_stream [win_cap]: full windows in (stream, time) order
 *   tail_tok     [S][21] zero padded, tail_stream [S], tail_frames [S]: end-of-stream remainders of 1..3 frames
 *   counts       [2]: number of full windows, number of tails written by this step
 * win_cap must be >= snacb_ingest_window_capacity(S, n_tok).  Asynchronous on `stream`. */
int snacb_ingest_step(snacb_ingest g, const int32_t* tok, int S, int n_tok, const int32_t* n_valid,
                      const uint8_t* finish, int32_t* win_tok, int32_t* win_stream, int win_cap, int32_t* tail_tok,
                      int32_t* tail_stream, int32_t* tail_frames, int32_t* counts, void* stream);
/* Debug / tests: copy per-stream state (0 waiting for SOS, 1 in speech, 2 ended) and buffered-id count to the host. */
int snacb_ingest_state(snacb_ingest g, int32_t* state_host, int32_t* count_host, int n);

/* ---------------------------------------------------------------------------------------------
 * Egress formats, device to device, byte-exact against the Python stdlib calls
 * the reference makes.
 *   snacb_pcm_to_base64  n_chunks PCM chunks of `samples` int16 each -> n_chunks strings of
 *                        snacb_base64_len(2*samples) ASCII bytes (RFC 4648, '=' padded, no terminator):
 *                        base64.b64encode(audio_chunk) of the /ws/audio endpoint (modal_audio_stream.py:483-487)
 *   snacb_pcm_to_wav     n PCM strings -> n records of 44 + 2*samples bytes: the RIFF/WAVE file that
 *                        wave.open(..., "wb") writes for 1 channel, 2 bytes per sample, `sample_rate`
 *                        (modal_audio_stream.py:561-566, 650-657)
 * --------------------------------------------------------------------------------------------- */
long long snacb_base64_len(long long bytes);
int snacb_pcm_to_base64(const int16_t* pcm, long long n_chunks, long long samples, uint8_t* out, void* stream);
int snacb_pcm_to_wav(const int16_t* pcm, long long n, long long samples, int sample_rate, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNACB_H_ */
