/*
 * atmvfi.h - C ABI of libatmvfi_b200.so: the B200 (sm_100a) kernels behind the ATM-VFI model forward.
 *
 * The reference (Gancheekim/ATM-VFI) has no FFI of its own: its "operator boundary" is the set of
 * torch / torch.nn.functional calls made by network/network_base.py, network/attention.py and
 * network/flow_warp.py.  Every entry point below replaces one family of those
 * calls and cites it.  Conventions:
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - no hidden allocation: the caller owns every buffer;
 *   - return 0 on success, non-zero on error; atmvfi_last_error() returns the message (thread-local);
 *   - feature maps are channels-last ("NHWC": [B][H][W][pitch], pitch >= C floats, pitch % 4 == 0),
 *     3-channel images, flows and masks are planar ("NCHW") like the reference's public tensors;
 *   - ROW WINDOWS (spatial row slabs): every operator that walks a grid takes `y0, y1` and produces only
 *     rows [y0, y1) of every image of its OUTPUT grid (y1 == 0: all rows).  Inputs are always described by their FULL
 *     extent: padding, align_corners scales, window geometry and warp coordinates stay global, the kernel simply reads the
 *     halo rows it needs from the full-size input.  For window-major tensors a "row" is a row of windows.
 *   - fp32 storage everywhere; `precision` selects the multiply datapath of the GEMM-shaped ops:
 *     ATMVFI_FP32 = CUDA-core FFMA, ATMVFI_TF32 = tcgen05.mma kind::tf32 (fp32 accumulate in TMEM),
 *     ATMVFI_TF32X3 = "3xTF32": fp32-tolerance products on the same tensor cores, each a*b issued as
 *     a_hi*b_lo + a_lo*b_hi + a_hi*b_hi (x_hi = tf32(x), x_lo = tf32(x - x_hi)); weights packed as hi|lo chunk pairs
 *     (pack.pack_tc_x3), activations split on the fly in shared memory; feature maps are stored un-rounded.
 *     ATMVFI_F16 = fp16 STORAGE of the channels-last feature maps + tcgen05.mma kind::f16 (fp32 accumulate): same 10-bit
 *     mantissa as TF32 at half the bytes and twice the MMA rate.  The mode is thread-local (atmvfi_set_activation_f16): while
 *     it is on, every `float*` that names a channels-last feature map in layernorm / window_gather_ln / dwconv3x3_gelu /
 *     flow_warp_nhwc (src, out) / conv3x3_first (out) / pack5_planar (out) / window_attention_tc (out) points to __half
 *     elements and its pitch counts halves.  Planar images, flows, masks, q|k|v and the 5-channel motion heads stay fp32.
 *
 * There is no CPU fallback behind any of these symbols.
 */
#ifndef ATMVFI_H_
#define ATMVFI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATMVFI_ABI_VERSION 3
#define ATMVFI_MAX_SRC 4

enum { ATMVFI_FP32 = 0, ATMVFI_TF32 = 1, ATMVFI_TF32X3 = 2, ATMVFI_F16 = 3 };

/* output row mapping of atmvfi_gemm_conv */
enum {
  ATMVFI_OUT_PIXEL = 0,      /* output pixel (b,y,x) -> row of `out` (plain conv / linear)                      */
  ATMVFI_OUT_SHUFFLE2 = 1,   /* ConvTranspose2d(k=2,s=2): column block q=(dy*2+dx) of pixel (y,x) -> (2y+dy,2x+dx) */
  ATMVFI_OUT_WINDOW_REV = 2, /* rows are window-major tokens; undo partition / roll / centre pad (attention.py:17-25,65-71,323-331) */
  ATMVFI_OUT_QKV_HEADS = 3   /* fused q|k|v linear (Cout = 3C) written in the HEAD-MAJOR layout the attention kernels can fetch with
                                TMA: with R = B*Hout*Wout rows, hd = C/heads,
                                  Q[h][r][d] at ((0*heads + h)*R + r)*hd + d,   K[h][r][d] at ((heads + h)*R + r)*hd + d,
                                  V^T[h][d][r] at 2*C*R + (h*hd + d)*R + r      (floats from `out`; out_pitch is ignored) */
};

/* window geometry shared by the transformer kernels (attention.py:28-62, 275-305) */
typedef struct {
  int32_t B2;        /* images on the batch axis (2 x pairs: frame-0 images first, network_base.py:451) */
  int32_t H, W;      /* token grid                                                                    */
  int32_t ws;        /* window side                                                                   */
  int32_t shift;     /* cyclic shift (0 or ws/2)                                                      */
  int32_t Hp, Wp;    /* grid after centre padding to a multiple of ws                                 */
  int32_t pad_top, pad_left;
} atmvfi_window_geom;

typedef struct {
  const float* ptr;  /* NHWC source                                                 */
  int32_t C;         /* channels taken from this source                             */
  int32_t pitch;     /* floats between consecutive pixels                           */
} atmvfi_src;

/*
 * Implicit-GEMM convolution / linear layer with a fused epilogue.  Replaces
 *   nn.Conv2d 3x3 / 1x1 (+PReLU)            network_base.py:20-25, 38-54, 155-159, 192-196, 203-260
 *   nn.ConvTranspose2d k2 s2 (+PReLU)       network_base.py:27-32, 202-221, 243-255
 *   nn.Linear (+bias, +residual)            attention.py:93-96, 138-141, 349-351
 *   torch.cat on the channel axis of inputs network_base.py:81, 384, 410, 418-428, 506  (the sources are read in place)
 * A[m][k]: m = output pixel, k = (tap, source, channel);  W: packed by the host (see atmvfi_pack.py).
 * out = act( A*W + bias (+ residual) ), optionally also out2 = prelu(out, slope2).
 */
typedef struct {
  int32_t nsrc;
  atmvfi_src src[ATMVFI_MAX_SRC];
  int32_t B, Hin, Win;           /* all sources share the spatial shape        */
  int32_t ksize;                 /* 1 or 3 (deconv uses ksize=1 + SHUFFLE2)     */
  int32_t stride, dil;           /* pad = dil*(ksize-1)/2                       */
  int32_t Hout, Wout;            /* conv output grid (before SHUFFLE2)          */
  int32_t Cout;                  /* true output channels (per shuffle block)    */
  const float* weight;           /* packed weights, layout depends on precision */
  int32_t ldw;                   /* FP32: floats per k-row (= padded N)         */
  const float* bias;             /* [Cout] or NULL                              */
  const float* prelu;            /* [Cout] slopes or NULL                       */
  const float* residual;         /* added after bias, rows like `out`, or NULL  */
  int32_t res_pitch;
  float* out;
  int32_t out_pitch;
  float* out2;                   /* optional second output = prelu(out, prelu2) */
  const float* prelu2;
  int32_t out2_pitch;
  int32_t out_mode;              /* ATMVFI_OUT_*                                */
  atmvfi_window_geom win;        /* used by ATMVFI_OUT_WINDOW_REV               */
  int32_t precision;             /* ATMVFI_FP32 / ATMVFI_TF32 / ATMVFI_TF32X3   */
  const void* tma_host;          /* TF32 / TF32X3: host pointer to the plan made by atmvfi_gemm_conv_plan, else NULL */
  int32_t row_begin, row_end;    /* row window on the GEMM grid [B][Hout][Wout] (window-major sources: rows of windows,
                                    i.e. Hout = B2-images x window rows); row_end == 0: all rows */
  int32_t qkv_heads;             /* ATMVFI_OUT_QKV_HEADS: number of attention heads */
  /* ATMVFI_F16 only: sources, residual and out2 are fp16 maps (pitches in halves, pitch % 8 == 0); `out` is fp16 unless out_f32;
   * head32 (optional) receives an fp32 copy of output channels [head32_c0, Cout) - the flows / occlusion logit of a motion head */
  int32_t out_f32;
  float* head32;
  int32_t head32_pitch, head32_c0;
  /* Tensor-core paths, Cout % 4 != 0: the caller owns the pad lanes [Cout, round_up(Cout, 4)) of every output row (out, out2) and
   * passes bias / prelu / prelu2 arrays padded to round_up(Cout, 4) floats; the kernel may then store whole 4-channel vectors (the
   * pad lanes receive zeros).  Lets layers with 101 / 197 / 389 channels use the vectorised epilogue. */
  int32_t pad_stores;
  /* bias / prelu / prelu2 point to arrays padded (zeros / ones) to a multiple of this many floats (0: exactly Cout).  >= 32 lets the
   * tensor-core kernel take its TMA-store epilogue, which reads parameters in whole 32-channel chunks. */
  int32_t param_pad;
} atmvfi_gemm_conv_desc;

const char* atmvfi_last_error(void);
int atmvfi_abi_version(void);
/* Thread-local mode for subsequent launches from this thread: when on, kernels that produce channels-last feature
 * maps round their outputs to the nearest TF32 value (the tcgen05 tf32 datapath truncates operands otherwise).
 * The host runtime turns it on for precision ATMVFI_TF32 and off for ATMVFI_FP32 / ATMVFI_TF32X3. */
void atmvfi_set_output_rounding(int on);
/* Thread-local: channels-last feature maps are fp16 (see ATMVFI_F16 above).  The host runtime sets it per launch sequence. */
void atmvfi_set_activation_f16(int on);
/* fp32 rows [rows][C] -> fp16 rows, zero-filling channels [C, zero_fill_to): small fp32 side products (per-token motion, motion
 * heads) that are also a source of an fp16 GEMM. */
int atmvfi_cast_f32_to_f16(const float* in, int in_pitch, void* out, int out_pitch, int64_t rows, int C, int zero_fill_to, void* stream);
/* Fills name[] (<=255 chars) with the device name and returns the SM count, or -1 without a usable sm_100 device. */
int atmvfi_device_info(int device, char* name, int* cc_major, int* cc_minor);

int atmvfi_gemm_conv(const atmvfi_gemm_conv_desc* d, void* stream);
/* TF32 path: build the TMA tensor maps once per (layer, shape).  plan_host must hold atmvfi_gemm_conv_plan_bytes() bytes. */
int atmvfi_gemm_conv_plan_bytes(void);
int atmvfi_gemm_conv_plan(const atmvfi_gemm_conv_desc* d, void* plan_host);

/* LayerNorm over channels of `rows` tokens (nn.LayerNorm, network_base.py:55,84; attention.py:235,333). */
int atmvfi_layernorm(const float* in, int in_pitch, float* out, int out_pitch, int64_t rows, int C,
                     const float* gamma, const float* beta, float eps, void* stream);   /* row windows: offset the pointers */

/*
 * pad_if_needed + torch.roll + window_partition + norm1 in one pass (attention.py:273-316):
 * reads tokens [B2][H][W][C] and writes window-major rows [B2*nW*ws*ws][C]; centre-pad tokens are
 * LayerNorm(0) = beta, exactly as in the reference where padding precedes norm1.
 */
int atmvfi_window_gather_ln(const float* tok, int tok_pitch, float* win, int win_pitch, int C,
                            const atmvfi_window_geom* g, const float* gamma, const float* beta, float eps,
                            int wy0, int wy1 /* window rows [wy0, wy1) of every image; wy1 == 0: all */, void* stream);

/*
 * Window attention with the attention-to-motion reduction (attention.py:187-213, 370-390).
 * qkv: window-major rows [rows][3C] (q | k | v, head-major inside each).  cross != 0: queries of a window
 * attend to the keys/values of the same window of the OTHER frame (attention.py:318).  The additive
 * -100 masks of the centre padding and of the cyclic shift are evaluated from `g` on the fly.
 * out: window-major [rows][C] = softmax(q k^T * hd^-0.5 + mask) v, heads concatenated.
 * motion (cross only, may be NULL): NHWC [B2/2][H][W][motion_pitch]; this block's 4 channels
 * (frame0 x,y, frame1 x,y) start at motion_off; values = head-mix MLP( sum_j attn_ij * relative_coord_ij ).
 * relative_coord: [2][N][N] buffer from the state-dict (attention.py:150-157).
 * scratch: rows*heads*2 floats of workspace for the per-head motion (needed only with motion).
 */
int atmvfi_window_attention(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                            const atmvfi_window_geom* g, int cross, const float* relative_coord,
                            const float* mix_w0, const float* mix_b0, const float* mix_w2, const float* mix_b2,
                            float* motion, int motion_pitch, int motion_off, float* scratch,
                            int wy0, int wy1 /* window rows of every image */,
                            int qkv_layout /* 0: rows [rows][3C] with qkv_pitch; 1: head-major (ATMVFI_OUT_QKV_HEADS) */, void* stream);

/* Same contract on the tensor cores (tcgen05 kind::tf32, accumulators in TMEM; Q, K, V^T staged in shared memory as
 * TF32).  rc_closed_form != 0 asserts that relative_coord holds the reference's own buffer contents (key position -
 * query position, attention.py:150-165), which the kernel then evaluates arithmetically instead of loading it.
 * Shapes outside the kernel's envelope (N > 256 tokens per window, head dim > 96) run on the CUDA-core kernel. */
int atmvfi_window_attention_tc(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                               const atmvfi_window_geom* g, int cross, const float* relative_coord, int rc_closed_form,
                               const float* mix_w0, const float* mix_b0, const float* mix_w2, const float* mix_b2,
                               float* motion, int motion_pitch, int motion_off, float* scratch,
                               int wy0, int wy1, int qkv_layout, void* stream);

/* Debug aid: with ATMVFI_ATTN_PROF=1 in the environment, thread 0 of every CTA of the tcgen05 attention kernel accumulates the
 * clock cycles of its six phases (staging, QK^T, softmax max, softmax exp + P, PV, output); this reads and clears the totals.
 * Returns non-zero when profiling is off. */
int atmvfi_attn_prof_read(unsigned long long* out6_host);

/* Mlp middle: depth-wise 3x3 (pad 1) + bias + exact-erf GELU on NHWC tokens (attention.py:74-85,118-119). */
int atmvfi_dwconv3x3_gelu(const float* in, float* out, int B, int H, int W, int C, int pitch,
                          const float* w9c /* [9][C] */, const float* bias, int y0, int y1, void* stream);

/* Fused Mlp tail (attention.py:74-85, 118-123 and the block residual at attention.py:333, 494):
 *     out = residual + fc2( GELU( DWConv3x3(hidden) + b_dw ) ) + b_fc2
 * on the tensor cores, without ever storing the activated hidden map: it is produced tile by tile in shared memory as the A operand
 * of a tcgen05 GEMM (csrc/mlp_tail_tc.cu).  Same arithmetic as atmvfi_dwconv3x3_gelu followed by atmvfi_gemm_conv with a residual.
 * hidden [B][H][W][hid_pitch] (Ch channels), residual / out [B][H][W][pitch] (C channels): fp32 maps for ATMVFI_TF32, fp16 maps for
 * ATMVFI_F16.  w10: [10][Ch] fp32 = the nine depth-wise taps (row ky*3+kx) followed by the depth-wise bias.  w_fc2: the packed
 * tensor-core operand of fc2, [w_rows][Ch] K-major (fp32 values pre-rounded to TF32, or fp16), w_rows >= C.  bias_fc2: padded with
 * zeros to a multiple of 32 floats past C + 352.  Requires Ch % 32 == 0 (fp16: % 64), C % 32 == 0, 16-byte aligned operands.
 * [y0, y1): row window of the output (it reads hidden rows y0 - 1 .. y1). */
int atmvfi_mlp_tail(const void* hidden, int hid_pitch, int B, int H, int W, int Ch, const float* w10, const void* w_fc2, int w_rows,
                    const float* bias_fc2, const void* residual, int res_pitch, void* out, int out_pitch, int C, int precision,
                    int y0, int y1, void* stream);
/* Debug aid (ATMVFI_MT_PROF=1): per-role cycle counters of CTA 0 of atmvfi_mlp_tail (csrc/mlp_tail_tc.cu); reads and clears them. */
int atmvfi_mlp_tail_prof_read(unsigned long long* out16_host);

/* First encoder layer (feat_extracts.0.0, network_base.py:103): Conv2d(3 -> Cout, k3, p1) + PReLU read straight from
 * the planar frame, written channels-last.  wk: [27][ldw] with row = (ky*3+kx)*3 + c (the FP32 packing of atmvfi_gemm_conv). */
int atmvfi_conv3x3_first(const float* img, const float* wk, int ldw, const float* bias, const float* prelu, float* out,
                         int out_pitch, int B, int H, int W, int Cout, int y0, int y1, void* stream);

/* Five planar [B,3,H,W] images -> channels [0,15) of an NHWC buffer (channel 15 zeroed): the image part of
 * torch.cat([feat, im0, I_t_0, im1, I_t_1, I_t], 1) at network_base.py:418, in one coalesced pass. */
int atmvfi_pack5_planar(const float* s0, const float* s1, const float* s2, const float* s3, const float* s4, float* out,
                        int out_pitch, int B, int H, int W, int y0, int y1, void* stream);

/* flow_warp.flow_warp (flow_warp.py:50-60) on planar tensors: out[b,c] = bilinear(img[b,c], grid+flow[b]). */
int atmvfi_flow_warp_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W, int y0, int y1,
                          void* stream);

/* Same sampling on an NHWC map; the flow is read from channels [flow_off, flow_off+2) of an NHWC head. */
int atmvfi_flow_warp_nhwc(const float* src, int src_pitch, const float* head, int head_pitch, int flow_off,
                          float* out, int out_pitch, int B, int C, int H, int W, int y0, int y1, void* stream);

/*
 * Fused pair warp + occlusion blend (network_base.py:385-387, 464-466, 496-498, 514-525):
 * head: NHWC 5 channels (flow0.xy, flow1.xy, occlusion logit) starting at head_off.
 * w0 = warp(im0, flow0), w1 = warp(im1, flow1), it = sigmoid(l)*w0 + (1-sigmoid(l))*w1  (all planar [B,3,H,W]).
 * Optional planar exports (NULL to skip): flow0/flow1 [B,2,H,W], occ1/occ2 [B,1,H,W].
 */
int atmvfi_warp_blend(const float* im0, const float* im1, const float* head, int head_pitch, int head_off,
                      float* w0, float* w1, float* it, float* flow0, float* flow1, float* occ1, float* occ2,
                      int B, int H, int W, int y0, int y1, void* stream);

/* F.interpolate(bilinear, align_corners=True) on `planes` planar images; values multiplied by `scale`
 * (network_base.py:11-18 uses x2 with scale 2; :445-446 uses x0.5 with scale 1). */
int atmvfi_resize_bilinear_ac(const float* in, float* out, int planes, int Hin, int Win, int Hout, int Wout,
                              float scale, int y0, int y1, void* stream);

/* planar [B][C][H][W] -> channels [chan_off, chan_off+C) of an NHWC buffer (replaces torch.cat with images). */
int atmvfi_nchw_to_nhwc(const float* in, float* out, int out_pitch, int chan_off, int B, int C, int H, int W,
                        int zero_fill_to /* also zero channels [chan_off+C, zero_fill_to) */, int y0, int y1, void* stream);

/* channels [0, C) of an NHWC map (pointer already offset to the first channel) -> planar [B][C][H][W]. */
int atmvfi_nhwc_to_nchw(const float* in, int in_pitch, float* out, int B, int C, int H, int W, void* stream);

/* Multi-scale global-motion ensemble (network_base.py:548-615).
 * l1_mean: out[s] = mean_i |a[s][i] - b[s][i]| over n elements per sample (nn.L1Loss + torch.mean(dim=[1,2,3]), :560-561);
 * deterministic two-stage reduction; scratch holds atmvfi_l1_mean_scratch_floats(samples) floats.
 * select_min3: out[s] = the candidate c_k[s] (n floats per sample) with the smallest loss, first minimum wins like the
 * reference's if / elif / else chain (:596-611) - evaluated on the device, no host round trip. */
int atmvfi_l1_mean_scratch_floats(int samples);
int atmvfi_l1_mean(const float* a, const float* b, float* out, float* scratch, int samples, int64_t n, void* stream);
int atmvfi_select_min3(const float* l0, const float* l1, const float* l2, const float* c0, const float* c1, const float* c2,
                       float* out, int samples, int64_t n, void* stream);

/* Stream-ordered device-to-device copy (cudaMemcpyAsync); used by the video-stream plan to move a frame's encoder features
 * from the "frame 1" half of the batch axis to the "frame 0" half instead of recomputing them (demo_2x.py:129-168 encodes every
 * interior frame twice). */
int atmvfi_copy(void* dst, const void* src, size_t bytes, void* stream);

/* I_t += 2*sigmoid(res)-1 ; clamp (network_base.py:429, 532-533).  res: NHWC 3 channels. */
int atmvfi_residual_finish(const float* res, int res_pitch, const float* it, float* it_sum, float* it_clamped,
                           int B, int H, int W, int y0, int y1, void* stream);

/* One level of the global-motion pyramid warp (network_base.py:480-485) for both frames in one launch: the x2 align_corners
 * up-sampling of the coarser level's flows with doubled values (upsample_flow, network_base.py:11-18; `upsample` = 1: flow0 / flow1 are
 * [B,2,H/2,W/2] and the up-sampled flows are also written to flow*_out when non-null) fused into the backward warps of im0 / im1
 * ([B,3,H,W] planar) that consume them (flow_warp.py:50-60).  Source tiles are staged in shared memory; results are bit-identical to
 * atmvfi_resize_bilinear_ac + atmvfi_flow_warp_nchw. */
int atmvfi_pyramid_warp(const float* im0, const float* im1, const float* flow0, const float* flow1, int upsample, float* out0,
                        float* out1, float* flow0_out, float* flow1_out, int B, int H, int W, int y0, int y1, void* stream);

/* demo_2x.inference_2frame host arithmetic on the device (demo_2x.py:64-75, 79-85):
 * uint8 HWC (optionally BGR) -> fp32 planar RGB / 255, replicate-padded by (left, top) to Hp x Wp, and back
 * with round-half-even (np.round) and clipping to [0,255]. */
int atmvfi_u8_to_planar(const uint8_t* in, float* out, int H, int W, int Hp, int Wp, int top, int left,
                        int bgr, void* stream);
int atmvfi_planar_to_u8(const float* in, uint8_t* out, int H, int W, int Hp, int Wp, int top, int left,
                        int bgr, void* stream);
/* atmvfi_u8_to_planar restricted to rows [y0, y1) of the PADDED frame: only source rows clamp(y - top, 0, H-1) of `in` are read, so a
 * rank of the row-slab mode uploads just its share of the uint8 frame (the planar rows then travel to the peers over NVLink). */
int atmvfi_u8_to_planar_rows(const uint8_t* in, float* out, int H, int W, int Hp, int Wp, int top, int left,
                             int bgr, int y0, int y1, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Spatial row-slab mode over NVLink (BASELINE.json configs[3]: one 4096x2160 pair on 2/4/8 GPUs).
 * The reference has no multi-GPU code; this replaces nothing in it.  One process per GPU.  Every rank allocates one
 * arena with the SAME layout, exports it through CUDA IPC and maps the peers' arenas, so the peer copy of a buffer is
 * peer_base + (ptr - my_base).  Rows that a consumer needs from a neighbour are PUSHED into the consumer's copy of the
 * buffer by the producer (16-byte stores through the peer mapping), followed by a system-scope release of a flag in the
 * consumer's memory.  Flags hold the step number ("epoch", a device-resident counter), so a captured CUDA graph
 * replays without argument patching.
 * ------------------------------------------------------------------------------------------------------------------ */
#define ATMVFI_IPC_HANDLE_BYTES 64
#define ATMVFI_P2P_MAX_PIECES 16
#define ATMVFI_P2P_MAX_PEERS 8

int atmvfi_arena_alloc(size_t bytes, void** ptr);              /* cudaMalloc (IPC-exportable, unlike a caching allocator block) */
int atmvfi_arena_free(void* ptr);
int atmvfi_ipc_export(void* ptr, unsigned char* handle64);     /* cudaIpcGetMemHandle */
int atmvfi_ipc_open(const unsigned char* handle64, void** peer_ptr);   /* cudaIpcOpenMemHandle, lazy peer access */
int atmvfi_ipc_close(void* peer_ptr);

/* `nchunks` chunks of `chunk_bytes`, `chunk_stride` bytes apart on both sides (rows [y0,y1) of every image / plane). */
typedef struct {
  const void* src;       /* local                               */
  void* dst;             /* peer-mapped address (or local)      */
  uint64_t chunk_bytes;  /* multiple of 4 (16-byte vectors when everything is 16-byte aligned) */
  uint64_t chunk_stride;
  uint32_t nchunks;
  uint32_t reserved;
} atmvfi_p2p_piece;

/* Which GPU holds which rows of a buffer: rows [row_lo[i], row_lo[i+1]) are read at (local address + byte_delta[i]) in the
 * peer-mapped address space; byte_delta 0 = the local copy. */
typedef struct {
  int32_t nseg;
  int32_t row_lo[ATMVFI_P2P_MAX_PEERS * 2 + 1];
  int64_t byte_delta[ATMVFI_P2P_MAX_PEERS * 2];
} atmvfi_row_owners;

/* atmvfi_flow_warp_nhwc whose SOURCE rows are read in place from the GPUs that own them (backward warps have a
 * data-dependent reach, so the 1/8-resolution feature maps are not gathered: each sample is one NVLink load). */
int atmvfi_flow_warp_nhwc_p2p(const float* src, int src_pitch, const float* head, int head_pitch, int flow_off,
                              float* out, int out_pitch, int B, int C, int H, int W, int y0, int y1,
                              const atmvfi_row_owners* owners, void* stream);

/* atmvfi_warp_blend whose two SOURCE images (same row layout) are read in place from the GPUs that own their rows. */
int atmvfi_warp_blend_p2p(const float* im0, const float* im1, const float* head, int head_pitch, int head_off,
                          float* w0, float* w1, float* it, float* flow0, float* flow1, float* occ1, float* occ2,
                          int B, int H, int W, int y0, int y1, const atmvfi_row_owners* owners, void* stream);

/* One exchange site: copy every piece, then store *epoch (release, system scope) into each signal flag (peer memory),
 * then wait until each wait flag (local memory, raised by a peer's call of this function) has reached *epoch.
 * `counter` is a zero-initialised word private to the site.  A wait that lasts longer than the time-out (below) sets the
 * sticky *error_word and returns; while *error_word is non-zero every exchange / step-begin launch of this rank returns
 * at once without pushing or signalling (the step is poisoned; the host must look at the word - SlabSession.check). */
int atmvfi_p2p_exchange(const atmvfi_p2p_piece* pieces, int npieces, uint32_t* const* signal_flags, int nsignal,
                        const uint32_t* const* wait_flags, int nwait, const uint32_t* epoch, uint32_t* counter,
                        uint32_t* error_word, void* stream);
/* Start of a step: ++*epoch, publish it to every peer and wait until every peer has published the same value (they
 * have finished the previous step, so their halo rows may be overwritten). */
int atmvfi_p2p_step_begin(uint32_t* epoch, uint32_t* const* signal_flags, int nsignal, const uint32_t* const* wait_flags,
                          int nwait, uint32_t* error_word, void* stream);

/* Time-out of every flag wait, in milliseconds (default 4000, or ATMVFI_P2P_TIMEOUT_MS); applies to launches issued - and
 * CUDA graphs captured - after the call. */
int atmvfi_p2p_set_timeout_ms(int ms);

#ifdef __cplusplus
}
#endif
#endif /* ATMVFI_H_ */
