/*
 * fmi_b200.h — C ABI of the B200-native (sm_100a) generator hot path of
 * syncdoth/face_mask_inpaint.
 *
 * Every entry point takes raw DEVICE pointers, plain sizes and a CUDA stream
 * (cudaStream_t passed as void*), launches asynchronously on that stream, never
 * synchronises the host, never allocates unless stated, and returns 0 on success
 * or a negative FMI_E* code (message via fmi_last_error(), thread-local).
 * No torch types cross this boundary.
 *
 * The reference interfaces each entry point replaces are cited as
 * file:line relative to the reference repository root.
 */
#ifndef FMI_B200_H_
#define FMI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of activation buffers */
#define FMI_F32 0
#define FMI_BF16 1
#define FMI_F16 2

/* tensor-core operand precision of the contraction kernels */
#define FMI_MMA_TF32 0 /* fp32 I/O contract: max rel err <= 1e-3 */
#define FMI_MMA_BF16 1 /* bf16 contract:     max rel err <= 2e-2 */

/* error codes */
#define FMI_OK 0
#define FMI_EINVAL (-1)   /* bad argument / unsupported configuration (never silent garbage) */
#define FMI_ECUDA (-2)    /* CUDA runtime / driver error                                  */
#define FMI_EARCH (-3)    /* device is not sm_100                                          */
#define FMI_ENOMEM (-4)   /* caller-provided workspace too small                           */

int fmi_version(void);
const char* fmi_last_error(void);
/* 0 when the current CUDA device can run these kernels (compute capability 10.x). */
int fmi_device_check(void);
/* Number of kernels this library has launched since it was loaded (all entry points). */
long long fmi_kernel_launch_count(void);
/* Optional CUDA-event timing of the dominant kernels on their own launch stream:
 *   kind 0 = attention main kernel, kind 1 = modulated-conv implicit GEMM, kind 2 = the robust attention kernel when
 *   it runs as the per-image fallback behind the fast one (normally an immediate exit).
 * fmi_profile_collect synchronises on the recorded events, returns their summed duration (ms) and
 * count, and clears the record. Off by default (no events are created). */
int fmi_profile_enable(int on);
int fmi_profile_collect(int kind, double* total_ms, int* launches);
/* Kinds 0 .. fmi_profile_kinds()-1 (names: fmi_profile_kind_name). fmi_profile_dump returns the per-launch records of one
 * kind in launch order — duration in ms, and the ALGORITHMIC FLOPs and HBM bytes (inputs once + outputs once) the launcher
 * stated for that launch — at most `cap` of them, and clears the record. Launches inside a CUDA-graph capture are not timed. */
int fmi_profile_kinds(void);
const char* fmi_profile_kind_name(int kind);
int fmi_profile_dump(int kind, double* ms_out, double* flops_out, double* bytes_out, int cap, int* n);

/* ---------------------------------------------------------------------------------------------
 * a5  fused bias + activation.
 * Replaces pybind `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 *   (modules/psp/stylegan2/op/fused_bias_act.cpp:11-21, kernel fused_bias_act_kernel.cu:18-49).
 *   y[i] = act'(x[i] + b[(i / step_b) % size_b]) * scale,  act*10+grad in {10,11,12,30,31,32};
 *   b == NULL / ref == NULL mean "absent" (the reference's numel()==0 convention, :62-63).
 * x, ref, y have element type `dtype`; b has element type `bias_dtype` (fp32 parameters with
 * bf16 activations are allowed — new capability, the reference requires equal types).
 * ------------------------------------------------------------------------------------------- */
int fmi_fused_bias_act(const void* x, const void* b, const void* ref, void* y, int act, int grad,
                       float alpha, float scale, int64_t size_x, int64_t step_b, int64_t size_b,
                       int dtype, int bias_dtype, void* stream);

/* Fused backward of fused_leaky_relu: FusedLeakyReLUFunctionBackward.forward
 *   (modules/psp/stylegan2/op/fused_act.py:18-38):
 *   grad_in = grad_out * (out > 0 ? 1 : alpha) * scale ; grad_bias[c] = sum_{n,inner} grad_in.
 * grad_bias is fp32 [size_b] and is ACCUMULATED into (caller zeroes it); may be NULL.
 * Layout [outer, size_b, step_b] contiguous. */
int fmi_bias_act_bwd(const void* grad_out, const void* out, void* grad_in, float* grad_bias,
                     float alpha, float scale, int64_t size_x, int64_t step_b, int64_t size_b,
                     int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  upfirdn2d.
 * Replaces pybind `upfirdn2d.upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y,
 *   down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)` (modules/psp/stylegan2/op/upfirdn2d.cpp:12-23,
 *   upfirdn2d_kernel.cu:140-272). Output [major,out_h,out_w,minor] with
 *   out = (in*up + pad0 + pad1 - k + down) / down  (upfirdn2d_kernel.cu:167-168).
 * Taps are applied flipped (true convolution, upfirdn2d_kernel.cu:77). kernel is fp32 on device.
 * EVERY (up, down, pad, kernel<=32x32) configuration is computed (the reference silently
 * returns uninitialised memory outside its six template modes).
 * ------------------------------------------------------------------------------------------- */
int fmi_upfirdn2d(const void* x, const float* kernel, void* y, int64_t major, int in_h, int in_w,
                  int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                  int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype, void* stream);
/* output extent helper (same formula), returns <0 when the extent would be empty */
int fmi_upfirdn2d_out_size(int in, int up, int down, int pad0, int pad1, int k);

/* ---------------------------------------------------------------------------------------------
 * a7  masked source/reference compositing.
 * scale_img: F.interpolate(mask, size, mode='bilinear', align_corners=True) (modules/model.py:10-12)
 * blend:     out = (1-m)*src + m*ref  (modules/model.py:99; psp_encoders.py:135-138)
 * mask is fp32 [N,1,Hm,Wm] at FULL resolution; it is sampled in-register at (H,W).
 * src/ref/out are [N,C,H,W] of `dtype`. Products are rounded separately (no FMA contraction) so
 * the fp32 result is bit-identical to the reference's separate mul/add kernels given equal m.
 * ------------------------------------------------------------------------------------------- */
int fmi_scale_mask(const float* mask, float* out, int N, int Hm, int Wm, int H, int W, void* stream);
int fmi_composite(const void* src, const void* ref, const float* mask, void* out, int N, int C,
                  int H, int W, int Hm, int Wm, int dtype, void* stream);
/* backward: g_src = (1-m)*g, g_ref = m*g (either may be NULL) */
int fmi_composite_bwd(const void* grad_out, const float* mask, void* grad_src, void* grad_ref, int N,
                      int C, int H, int W, int Hm, int Wm, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a1/a2  reference-guided attention (ExampleGuidedAttention.forward,
 *   modules/example_guided_att.py:21-41) and PICNet Auto_Attn.forward
 *   (modules/pluralistic_model/base_function.py:420-448), as ONE flash-style kernel family:
 *     q  = Wq x (+ bq)                      1x1 conv C -> d      (example_guided_att.py:27, base_function.py:429)
 *     P  = softmax_j(q_i . q_j)             no 1/sqrt(d) scale   (example_guided_att.py:30, base_function.py:432-433)
 *     O_g = v_g P^T   for value groups g    (example_guided_att.py:18, base_function.py:436,443)
 *     out_g[c,i] = a_g*w_i*O_g[c,i] + r_i*v_g[c,i]
 *        plain  group: w_i = 1,      r_i = b_g           (Auto_Attn: a=gamma, b=1; EGA src: a=1, b=0)
 *        masked group: w_i = 1-m_i,  r_i = m_i           (EGA ref: a=1; Auto_Attn pre: a=alpha)
 *   The S x S map never leaves the SM (TMEM/registers).
 *
 * fmi_attn_workspace_bytes: bytes of scratch the caller must provide (operand staging in the
 *   tensor-core layout: q^T [N,S,dpad] and the concatenated values [N,Cv,S] in bf16 / tf32).
 * fmi_attn_fwd arguments:
 *   x       [N,C,S]  query source features (fp32 or bf16 = dtype)
 *   wq      [d,C] fp32, bq [d] fp32 or NULL
 *   v0,v1   value groups [N,C0,S], [N,C1,S] (v1 may be NULL, C1 = 0); v0 may alias x
 *   mask    [N,S] fp32 (already at feature resolution) or NULL when no group is masked
 *   a0,a1   device pointers to fp32 scalars (NULL = 1.0); b0,b1 host floats
 *   masked0/masked1  0/1
 *   out0,out1  [N,C0,S], [N,C1,S] with batch strides out0_bs/out1_bs ELEMENTS (so both can live in
 *              one [N,C0+C1,S] concatenated tensor); element type = dtype
 *   lse     [N,S] fp32 row log-sum-exp (natural log) or NULL — saved for backward
 *   o_save  [N,C0+C1,S] (dtype) normalised attention output O before the blend, or NULL — saved for backward
 *   mma     FMI_MMA_TF32 | FMI_MMA_BF16
 * ------------------------------------------------------------------------------------------- */
int64_t fmi_attn_workspace_bytes(int N, int C, int d, int C0, int C1, int S, int mma);
int fmi_attn_fwd(const void* x, const float* wq, const float* bq, const void* v0, const void* v1,
                 const float* mask, const float* a0, float b0, int masked0, const float* a1, float b1,
                 int masked1, void* out0, int64_t out0_bs, void* out1, int64_t out1_bs, float* lse,
                 void* o_save, int N, int C, int d, int C0, int C1, int S, int dtype, int mma, void* workspace,
                 int64_t workspace_bytes, void* stream);

/* Backward of fmi_attn_fwd (training: train_reference_fill.py / train_psp.py go through it by autograd).
 *   Inputs as in fmi_attn_fwd plus: o_saved and lse from the forward, dout0/dout1 = gradients of the two
 *   output groups (dtype, batch strides in elements). Outputs (fp32, caller-allocated):
 *     dq  [N,S,dpad]  gradient of the projected query q^T (dpad = d rounded up to 64; columns >= d are zero)
 *     dv0 [N,C0,S], dv1 [N,C1,S]  gradients of the value groups incl. the blend's direct path (NULL = skip)
 *     da0, da1  device scalars, ACCUMULATED (gradients of gamma / alpha); NULL = skip
 *   The 1x1-conv gradients that follow from dq (dWq = dq x^T, dx += Wq^T dq, dbq = sum dq) are plain GEMMs
 *   left to the caller. Round-1 implementation: per image the S x S maps are materialised in the operand
 *   type inside `workspace` (3 S^2 elements) and every contraction is a tcgen05 GEMM. */
int64_t fmi_attn_bwd_workspace_bytes(int N, int C, int d, int C0, int C1, int S, int mma);
int fmi_attn_bwd(const void* x, const float* wq, const float* bq, const void* v0, const void* v1,
                 const float* mask, const float* a0, float b0, int masked0, const float* a1, float b1,
                 int masked1, const void* o_saved, const float* lse, const void* dout0, int64_t dout0_bs,
                 const void* dout1, int64_t dout1_bs, float* dq, float* dv0, float* dv1, float* da0,
                 float* da1, int N, int C, int d, int C0, int C1, int S, int dtype, int mma, void* workspace,
                 int64_t workspace_bytes, void* stream);

/* Opt-in materialisation of the S x S map that Auto_Attn returns (base_function.py:448):
 *   attn[n,i,j] = exp(q_i.q_j - lse_i). q^T is taken from the workspace of the matching
 *   fmi_attn_fwd call. attn is fp32 [N,S,S]. */
int fmi_attn_materialize(const void* workspace, const float* lse, float* attn, int N, int d, int S,
                         int mma, void* stream);

/* 1x1 convolution  y[n,o,s] = sum_c W[o,c] x[n,c,s] + b[o]  in fp32 SIMT (query conv,
 *   example_guided_att.py:9,27; out_conv, example_guided_att.py:13,38-39). */
int fmi_conv1x1(const void* x, const float* w, const float* b, void* y, int N, int Cin, int Cout,
                int S, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3/a6  modulated convolution (ModulatedConv2d.forward, modules/psp/stylegan2/model.py:241-279) with
 *   the StyledConv / NoiseInjection / FusedLeakyReLU / ToRGB glue (:289-294, :340-346, :360-369).
 *   Activations are kept in NHWC in the tensor-core operand type between layers (bf16 for
 *   FMI_MMA_BF16, tf32-rounded fp32 for FMI_MMA_TF32); per-sample modulated+demodulated weights are
 *   staged tap-major/K-major once per layer; the convolution is an implicit GEMM on tcgen05 with the
 *   im2col gather done by 4-D TMA boxes. See DESIGN.md.
 * ------------------------------------------------------------------------------------------- */
/* NCHW (dtype) -> NHWC (operand type of `mma`) and back: the module boundary. */
int fmi_nchw_to_nhwc(const void* x, void* y, int B, int C, int H, int W, int dtype, int mma, void* stream);
int fmi_nhwc_to_nchw(const void* x, void* y, int B, int C, int H, int W, int mma, int dtype, void* stream);

/* s[b,i] = latent[b,:] . mod_weight[i,:] / sqrt(K) + mod_bias[i]
 *   (EqualLinear `modulation`, model.py:159-167 with lr_mul = 1; called at :244). latent rows are
 *   latent_row_stride floats apart (so latent[:, layer] of a [B, n_latent, K] tensor needs no copy). */
int fmi_style_modulation(const float* latent, int64_t latent_row_stride, const float* mod_weight,
                         const float* mod_bias, float* s, int B, int K, int I, void* stream);

/* wp[b][tap][o][i] = scale*W[o,i,tap]*s[b,i]*demod[b,o],  demod = rsqrt(sum (scale*W*s)^2 + 1e-8)
 *   (model.py:245-249); weight is the parameter [O,I,k,k] fp32, k in {1,3}. */
int64_t fmi_modconv_weight_bytes(int B, int I, int O, int ksize, int mma);
int fmi_modconv_weight_prep(const float* weight, const float* s, void* wp, int B, int I, int O, int ksize,
                            int demodulate, int mma, void* stream);

/* One StyledConv (3x3) on NHWC operands: y = epi(modconv(x)). x [B,H,W,I], wp from
 *   fmi_modconv_weight_prep, y [B,OH,OW,O] with OH = H (plain) or 2H (upsample: conv_transpose2d
 *   stride 2 -> (2H+1)^2 intermediate in `workspace` -> 4x4 blur pad (1,1), model.py:255-263).
 *   act = 1: y = sqrt2*lrelu_0.2(conv + noise_w*noise[b|0,p] + act_bias[o])  (model.py:341-344, :294);
 *   act = 0: y = conv (ModulatedConv2d alone). noise is fp32 [B or 1, OH*OW] or NULL; noise_w a device
 *   scalar; blur_k the module's 4x4 `blur.kernel` buffer (already x4), required when upsample. */
int64_t fmi_styled_conv_workspace_bytes(int B, int O, int H, int W, int upsample, int mma);
int fmi_styled_conv_nhwc(const void* x, const void* wp, void* y, const float* noise, int noise_batched,
                         const float* noise_w, const float* act_bias, const float* blur_k, int B, int I,
                         int O, int H, int W, int upsample, int act, int mma, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* A plain StyledConv with the following ToRGB fused into its epilogue (StyledConv.forward :340-346 followed by
 *   ToRGB.forward :360-369 on its output, Generator.forward :537-539) for O <= 256: the 1x1 modulated conv to 3
 *   channels is accumulated from the activations the epilogue already holds, so the layer output is not re-read.
 *   rgb_w [B,3,O] = W[o,c]*s[b,c]/sqrt(O) from fmi_torgb_weights; rgb_skip [B,3,H/2,W/2] or NULL; rgb [B,3,H,W] fp32.
 *   y may be NULL: only the RGB image is produced (the last layer of an inference forward — nothing else reads its activations). */
int fmi_torgb_weights(const float* weight, const float* s, float* rgb_w, int B, int C, void* stream);
int fmi_styled_conv_torgb_nhwc(const void* x, const void* wp, void* y, const float* noise, int noise_batched,
                               const float* noise_w, const float* act_bias, const float* rgb_w,
                               const float* rgb_bias, const float* rgb_skip, const float* rgb_kernel, float* rgb,
                               int B, int I, int O, int H, int W, int mma, void* stream);

/* ToRGB (model.py:360-369): rgb[b,o,p] = sum_i (W[o,i]*s[b,i]/sqrt(I)) * x[b,p,i] + bias[o]
 *   + upfirdn2d(skip, blur_k, up=2, pad=(2,1)) (Upsample, model.py:30-49) when skip != NULL.
 *   x NHWC operand type; weight [3,I], s [B,I], bias [3], skip [B,3,H/2,W/2], rgb [B,3,H,W] fp32. */
int fmi_torgb_nhwc(const void* x, const float* weight, const float* s, const float* bias,
                   const float* skip, const float* blur_k, float* rgb, int B, int I, int H, int W, int mma,
                   void* stream);

/* Backward of fmi_styled_conv_nhwc (training: train_psp.py with train_decoder goes through it by autograd; the
 *   reference differentiates model.py:241-279 + :289-294 + fused_act.py:18-47 with ATen).
 *   Inputs: x, y (the forward output, needed when act = 1), dy [B,OH,OW,O] in the operand type, the fp32 parameter
 *   `weight` [O,I,3,3], the modulation s [B,I], noise as in the forward.
 *   Outputs: dx [B,H,W,I] operand type (NULL to skip), dweight [O,I,3,3] fp32 (through modulation AND demodulation),
 *   ds [B,I] fp32, and with act = 1: dnoise_w (1 float) and dbias [O] fp32. Outputs are overwritten, not accumulated.
 *   H and W must be powers of two >= 4; I, O multiples of 32, I <= 512. */
int64_t fmi_styled_conv_bwd_workspace_bytes(int B, int I, int O, int H, int W, int upsample, int act, int mma);
int fmi_styled_conv_bwd_nhwc(const void* x, const void* y, const void* dy, const float* weight, const float* s,
                             const float* noise, int noise_batched, const float* blur_k, void* dx, float* dweight,
                             float* ds, float* dnoise_w, float* dbias, int B, int I, int O, int H, int W, int upsample,
                             int act, int demodulate, int mma, void* workspace, int64_t workspace_bytes, void* stream);

/* Backward of the 1x1 modulated conv of ToRGB (model.py:360-369): dx[b,p,c] = sum_o rgb_w[b,o,c] * drgb[b,o,p]
 *   (operand type, NHWC), d_rgbw[b,o,c] = sum_p drgb[b,o,p] * x[b,p,c] (fp32, gradient w.r.t. the modulated weights
 *   rgb_w = fmi_torgb_weights), dbias[o] = sum drgb. The skip branch is differentiated by fmi_upfirdn2d. */
int fmi_torgb_bwd_nhwc(const void* x, const float* drgb, const float* rgb_w, void* dx, float* d_rgbw, float* dbias,
                       int B, int I, int H, int W, int mma, void* stream);

/* ---- f1: PICNet decoder conv blocks, inference ------------------------------------------------------
 * Replaces the cuDNN calls behind ResBlockDecoder.forward / Output.forward (modules/pluralistic_model/base_function.py:
 * 308-398) as assembled by ResGenerator.forward (network.py:247-268): nn.Conv2d(3,1,1), nn.ConvTranspose2d(3,2,1,1),
 * nn.InstanceNorm2d(affine=True), LeakyReLU, ReflectionPad2d(1), Tanh. Activations are NHWC in the tensor-core operand
 * type (mma = FMI_MMA_TF32: tf32-rounded fp32, FMI_MMA_BF16: bf16); "pixel stride" = elements between consecutive
 * pixels, so a tensor may be a channel slice of a wider buffer. */

/* wp[t][o][i_off + i] = weight[o,i,t] (Conv2d layout [O,I,3,3], transposed = 0) or weight[i,o,t] (ConvTranspose2d
 *   layout [I,O,3,3], transposed = 1) in the operand type; wp is [9][O_rows][I_row] (ksize = 3; [1][O_rows][I_row] for the
 *   1x1 convs of the ResBlock shortcuts, ksize = 1), rows / columns not written stay as
 *   the caller initialised them (zero). Writing two weights at different i_off concatenates them along the input
 *   channels. merged = 1 (transposed only, O_rows = 4*O): the layout of fmi_conv3x3_nhwc mode 3, wp [4][4*O][I_row] with
 *   slab = input shift 2*dy + dx and row = (2*py + px)*O + o. weight is the effective fp32 weight (SpectralNorm already applied: w_bar / sigma, external_function.py:55-57). */
int fmi_conv_weight_prep(const float* weight, void* wp, int O, int I, int transposed, int O_rows, int I_row, int i_off,
                         int merged, int ksize, int mma, void* stream);

/* The same for a SpectralNorm-wrapped conv (external_function.py:16-72): one power iteration on (w_bar, u, v) — u [Hh] and
 *   v [Wd] are updated in place as SpectralNorm._update_u_v does (Hh = w_bar.shape[0], Wd = numel / Hh) — and wp receives
 *   w_bar / sigma, sigma = u . (W v). scratch: (Wd + Hh) floats. */
int fmi_conv_weight_prep_sn(const float* w_bar, float* u, float* v, float* scratch, void* wp, int O, int I, int transposed,
                            int O_rows, int I_row, int i_off, int merged, int ksize, int mma, void* stream);

/* n calls of fmi_conv_weight_prep_sn as three launches (the power iterations of a network's convolutions are independent of
 *   its activations: they run once at the start of the forward). descs: n descriptors in DEVICE memory, 96 bytes each:
 *   { const float* w_bar; float* u; float* v; float* v_part ( 4*Wd floats ); float* u_raw ( Hh floats ); void* wp;
 *     int O, I, transposed, O_rows, I_row, i_off, merged, T ( ksize^2 ), Hh, Wd, 0, 0; }
 *   max_wd, max_hh, max_elems: the largest Wd, Hh and O*I*T among them (max_wd <= 12288). */
int fmi_conv_weight_prep_sn_batch(const void* descs, int n, int max_wd, int max_hh, int max_elems, int mma, void* stream);

/* NCHW (dtype) -> NHWC operand type into a channel slice: y[b, p, c] at y + (b*H*W + p) * y_pixel_stride + c. */
int fmi_nchw_to_nhwc_slice(const void* x, void* y, int B, int C, int H, int W, int64_t y_pixel_stride, int dtype,
                           int round_y, int mma, void* stream);

/* InstanceNorm2d statistics (biased variance over H*W per sample and channel, F.instance_norm) folded with the affine
 *   parameters: scale_shift[b][c] = (gamma[c]*rstd, beta[c] - mean*gamma[c]*rstd). sums: B*C*2 doubles of scratch. */
int fmi_instnorm_stats_nhwc(const void* x, int64_t x_pixel_stride, const float* gamma, const float* beta,
                            float* scale_shift, double* sums, int B, int C, int HW, float eps, int mma, void* stream);

/* y = leaky_relu(x * scale + shift, slope) per (sample, channel); scale_shift = NULL: activation only. */
int fmi_norm_act_nhwc(const void* x, int64_t x_pixel_stride, void* y, int64_t y_pixel_stride, const float* scale_shift,
                      int B, int C, int HW, float slope, int mma, void* stream);

/* nn.AvgPool2d(2, 2) on NHWC operand-type tensors (base_function.py:238-239, 290-298): y [B,H/2,W/2,C]. */
int fmi_avgpool2_nhwc(const void* x, int64_t x_pixel_stride, void* y, int64_t y_pixel_stride, int B, int C, int H, int W,
                      int round_y, int mma, void* stream);

/* Error-compensated TF32 operands ("3xTF32") for the fp32 contract of the PICNet conv blocks (the reference's strict-fp32
 * cuDNN run, torch.backends.cudnn.allow_tf32 = False, of base_function.py:207-398). kind::tf32 reads the upper 19 bits of an
 * fp32 operand; with x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly),
 *   x.w = hi_x.hi_w + hi_x.lo_w + lo_x.hi_w + O(2^-20 |x||w|), every product accumulated in fp32,
 * and the three terms are ONE fmi_conv3x3_nhwc call over 3*I input channels: activations [hi | hi | lo] (order 0), weights
 * [hi | lo | hi] (order 1) along the channel axis. x: `rows` rows (pixels, or the tap x output-channel rows of wp) of C fp32
 * values x_stride elements apart; y [rows][3*C]. C a multiple of 4.
 * fmi_set_tf32_exact(1) makes the FMI_MMA_TF32 producers that otherwise round to tf32 (fmi_conv_weight_prep*,
 * fmi_norm_act_nhwc) keep exact fp32 values for the split; it returns the previous setting. Process-wide, read at launch. */
int fmi_tf32_split3(const float* x, int64_t x_stride, float* y, int64_t rows, int C, int order, void* stream);
int fmi_set_tf32_exact(int on);
int fmi_get_tf32_exact(void);

/* ReflectionPad2d(1): fills the one-pixel border of y [B,H+2,W+2,C] from its interior. */
int fmi_reflect_border_nhwc(void* y, int B, int C, int H, int W, int mma, void* stream);

/* 3x3 convolution with batch-shared weights as a tcgen05 implicit GEMM.
 *   mode 0: Conv2d(3, stride 1, padding 1)                              x [B,H,W,*]      -> y [B,H,W,*]
 *   mode 1: Conv2d(3, stride 1, padding 0) on a pre-padded input        x [B,H+2,W+2,*]  -> y [B,H,W,*]
 *   mode 2: ConvTranspose2d(3, stride 2, padding 1, output_padding 1)   x [B,H,W,*]      -> y [B,2H,2W,*]
 *   mode 3: mode 2 as one GEMM over the 4 output-parity classes (O <= 64, wp from fmi_conv_weight_prep(merged = 1))
 *   mode 4: Conv2d(1, stride 1, padding 0), wp [1][O][I]
 *   act + 10 (act 2 only): y += acc + bias, the residual sum of a ResBlock onto the shortcut already stored in y
 *   wp [9][O][I] from fmi_conv_weight_prep (O a multiple of 32: pad with zero rows); bias [O] fp32 or NULL;
 *   act 2: y = acc + bias;  act 1: leaky_relu(acc + bias, slope);  act 3: tanh(acc + bias).
 *   round_y: 1 = y rounded to the operand type (tf32 / bf16); 0 (TF32 mode only) = exact fp32 values, for tensors that
 *   feed InstanceNorm statistics (a later MMA truncates them instead). Same flag on fmi_nchw_to_nhwc_slice.
 *   y (may be NULL if y_nchw is given): NHWC operand type, pixel stride y_pixel_stride, written to the interior of a
 *   buffer padded by y_pad (0 or 1) pixels per side. y_nchw (or NULL): fp32 [B, nchw_C, OH, OW], the first nchw_C channels. */
int fmi_conv3x3_nhwc(const void* x, int64_t x_pixel_stride, const void* wp, const float* bias, void* y,
                     int64_t y_pixel_stride, int y_pad, float* y_nchw, int nchw_C, int B, int I, int O, int H, int W,
                     int mode, int act, float slope, int round_y, int mma, void* stream);

/* Output block (base_function.py:369-398 after the leaky-ReLU and ReflectionPad2d(1) that fmi_conv3x3_nhwc(act = 1,
 *   y_pad = 1) + fmi_reflect_border_nhwc produce): img = tanh(Conv2d(C -> O <= 3, 3x3, padding 0)(xpad) + bias), fp32
 *   accumulation on the CUDA cores (HBM-bound: 128*C bytes per 54*C FLOPs). xpad [B,H+2,W+2,C] NHWC operand type, C in
 *   {16,32,64}; weight [O,C,3,3], bias [O] fp32. img [B,O,H,W] fp32 and / or pooled [B,O,H/4,W/4] (4x4 means =
 *   AdaptiveAvgPool2d of modules/model.py:111 for a 1024^2 -> 256^2 image); either may be NULL. scratch: 27*C + 4 floats. */
int fmi_output_conv_tanh(const void* xpad, const float* weight, const float* bias, float* img, float* pooled,
                         float* scratch, int B, int C, int O, int H, int W, int mma, void* stream);

/* ---- f1 in training: backward of the batch-shared convolutions of the PICNet conv blocks -------------------------------------
 * The reference differentiates SpectralNorm(nn.Conv2d) (base_function.py:207-305, network.py:73-172,296-365) with ATen / cuDNN
 * (cudnn_convolution_backward). Here the data gradient is the forward implicit GEMM on the flipped, transposed weights
 * (fmi_conv_nhwc with wp from fmi_conv_weight_prep(transposed = 1) of weight.flip(2, 3)) and the weight gradient is
 * fmi_conv_wgrad_nhwc: dwp[t][o][i] += sum_{b,p} dy[b,p,o] * x[b,p + off_t,i] — a tcgen05 GEMM whose contraction index is the
 * pixel (both operands MN-major NHWC boxes), split-K over pixel ranges and images, fp32 red.add into the caller's ZEROED dwp
 * [ksize^2][O][I]. x [B,H,W,I], dy [B,H,W,O]: dense NHWC in the operand type; ksize 1 or 3 (stride 1, padding ksize/2);
 * H, W powers of two >= 4; I, O multiples of 32.
 * transposed = 1: ConvTranspose2d(3, stride 2, padding 1, output_padding 1) (base_function.py:330-336): dy = the 4 pixel-parity
 * planes [4][B][H][W][O] (fmi_space_to_planes_nhwc) of the [B,2H,2W,O] output gradient; the data gradient of that layer is the
 * stride-2 convolution fmi_conv_nhwc(planes = 1) of the same planes. */
int fmi_conv_wgrad_nhwc(const void* x, const void* dy, float* dwp, int B, int I, int O, int H, int W, int ksize, int transposed,
                        int mma, void* stream);

/* Backward of y = leaky_relu(InstanceNorm2d(affine)(x)) as computed by fmi_instnorm_stats_nhwc + fmi_norm_act_nhwc (the norm +
 * activation pairs of ResBlockDecoder, base_function.py:338-344; the reference differentiates F.instance_norm / leaky_relu with
 * ATen): two passes over (dy, x), fp32 dense NHWC [B, HW, C]. scale_shift: the forward's [B][C][2]; mean_rstd [B][C][2] fp32;
 * dx [B, HW, C]; sums [B][C][2] doubles = (sum_p dz, sum_p dz * xhat) — dbeta[c] = sum_b sums[b][c][0], dgamma[c] = sum_b
 * sums[b][c][1]. C a multiple of 4, <= 1024. */
int fmi_instnorm_act_bwd_nhwc(const float* dy, const float* x, const float* scale_shift, const float* mean_rstd, float* dx,
                              double* sums, int B, int C, int HW, float slope, void* stream);

/* ---- f3: loss-side S x S and Gram products ----------------------------------------------------------------
 * C[b] = A[b] * B[b]^T, fp32 out, on the tcgen05 GEMM (csrc/gemm.cuh): replaces torch.bmm in contextual_loss (cosine similarities
 * of VGG features, external_function.py:249-252) and GramMatrix (:180-185), which the reference runs as fp32 SIMT GEMMs (matmul TF32
 * is off by default). A [batch, M, K] / B [batch, N, K]: K contiguous, row pitches lda / ldb and batch strides in ELEMENTS of the
 * operand type (mma = FMI_MMA_TF32: fp32 read as tf32; pass fmi_tf32_split3 operands, K tripled, for fp32-class products);
 * C [batch, M, ldc] fp32, accumulate = 1 adds to it. N a multiple of 8, rows 16-byte aligned. */
int fmi_gemm_nt(const void* A, int64_t lda, int64_t a_bs, const void* B, int64_t ldb, int64_t b_bs, float* C, int64_t ldc,
                int64_t c_bs, int batch, int M, int N, int K, int accumulate, int mma, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f2  the pSp encoder: IR-SE50 trunk, FPN adds and map2style heads of GradualStyleEncoder
 * (modules/psp/encoders/psp_encoders.py:13-37,100-152; units encoders/helpers.py:56-119), inference, NHWC in the operand type.
 *
 * fmi_conv_nhwc: Conv2d(ksize 1 or 3, padding ksize/2) on the tcgen05 implicit GEMM.
 *   x      input, I channels per pixel, addressed by (pixel, row, image) strides in ELEMENTS — a strided view (every second
 *          pixel and row) makes a stride-2 1x1 conv. With planes = 1 x holds the 4 pixel-parity planes [4][B][H][W][*] of the
 *          real input (fmi_space_to_planes_nhwc) and the conv is the 3x3 STRIDE-2 conv of that input. H, W = OUTPUT extents.
 *   wp     [sets][ksize^2][O][I] in the operand type (I contiguous; tap index ky*ksize + kx), sets = 1 when w_group = 0 (batch-
 *          shared weights), else B / w_group: images [g*w_group, (g+1)*w_group) use set g (heads as batch entries).
 *   bias   fp32 [O], or [9][O] with bias_classes = 9: indexed by the border class 3*vy + vx of the output pixel (vy: 0 first row,
 *          2 last row, 1 otherwise; vx alike) — the shift of a BatchNorm that precedes a zero-padded 3x3 conv (helpers.py:107-109).
 *          bias_per_set = 1 (with w_group >= 1): one such bias per weight set, concatenated.
 *   act    2: y = acc + bias; 1: leaky_relu(slope); 4: PReLU with slope_c [O] per channel (helpers.py:110). add_y: y += result.
 *   y      NHWC, O channels per pixel out of y_pixel_stride, dense rows / images. round_y as in fmi_conv3x3_nhwc.
 */
int fmi_conv_nhwc(const void* x, int64_t x_pixel_stride, int64_t x_row_stride, int64_t x_img_stride, const void* wp,
                  const float* bias, int bias_classes, const float* slope_c, float slope, void* y, int64_t y_pixel_stride,
                  int B, int I, int O, int H, int W, int ksize, int planes, int w_group, int bias_per_set, int act, int add_y,
                  int round_y, int mma, void* stream);
/* x [B][H][W][heads*C] (pixel stride x_pixel_stride) -> y [4][heads*B][H/2][W/2][C]: plane 2*(row&1) + (col&1), batch entry
 * head*B + b (the input layout of fmi_conv_nhwc(planes = 1)). */
int fmi_space_to_planes_nhwc(const void* x, int64_t x_pixel_stride, void* y, int B, int C, int H, int W, int heads, int mma,
                             void* stream);
/* SEModule (helpers.py:56-74): mean[b][c] = mean_hw r[b]; gate = sigmoid(w2 relu(w1 mean)), w1 [R][C], w2 [C][R] fp32;
 * r dense [B][HW][C]; mean, gate: fp32 [B][C] results; scratch: fp32 [B][32][C] (slab partial sums — no atomics, so the
 * result is bit-reproducible). */
int fmi_se_gate_nhwc(const void* r, const float* w1, const float* w2, float* scratch, float* mean, float* gate, int B, int C,
                     int R, int HW, int mma, void* stream);
/* y = r * gate[b][c] + sc (helpers.py:116-119): r, y dense [B][H][W][C]; sc read through (pixel, row, image) element strides. */
int fmi_se_scale_add_nhwc(const void* r, const float* gate, const void* sc, int64_t sc_pixel_stride, int64_t sc_row_stride,
                          int64_t sc_img_stride, void* y, int B, int C, int H, int W, int mma, void* stream);
/* y = bilinear_align_corners(x [B][h][w][C] -> OH x OW) + add  (psp_encoders.py:83-98 `_upsample_add`). */
int fmi_upsample_add_nhwc(const void* x, const void* add, void* y, int B, int C, int h, int w, int OH, int OW, int mma,
                          void* stream);

/* SpectralNorm in training (modules/pluralistic_model/external_function.py:30-42): one power iteration on (w_bar [Hh][Wd], u, v)
 * — u, v updated in place — and w_out = w_bar / sigma in the parameter's own layout (3 launches instead of the reference's ~13 ATen
 * launches per convolution and forward). snap [Hh + Wd + 1] keeps (u, v, sigma) of this call for the backward; scratch (Wd + Hh) floats.
 * Backward (u, v constant as in the reference, sigma = u^T W v): grad_w_bar = (g - <g, w_out> u v^T) / sigma; dot: one float. */
int fmi_spectral_norm_fwd(const float* w_bar, float* u, float* v, float* scratch, float* w_out, float* snap, int Hh, int Wd,
                          void* stream);
int fmi_spectral_norm_bwd(const float* g, const float* w_out, const float* snap, float* dot, float* grad_w_bar, int Hh, int Wd,
                          void* stream);

/* k x k mean (k = 2 or 4) of fp32 planes [planes][H][W] -> [planes][H/k][W/k]: the exact case of the generators' final
 * AdaptiveAvgPool2d (modules/psp/psp.py:33,113-114 `face_pool` 1024^2 -> 256^2; modules/model.py:79,111). */
int fmi_avgpool_planes(const float* x, float* y, int64_t planes, int H, int W, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMI_B200_H_ */
