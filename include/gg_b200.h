/*
 * gg_b200.h — C-ABI of the B200-native Graphical-GAN training hot path.
 *
 * This is the drop-in boundary of the repository: a shared library
 * (graphical-gan_b200/lib/libgg_b200.so) exporting stateless `extern "C"` entry points
 * that take raw device pointers, plain sizes and a cudaStream_t (passed as void*).
 * No torch / Python types appear here.  Every function returns 0 on success and a
 * non-zero code on failure (GG_ERR_* below, or 1000 + cudaError_t); gg_last_error()
 * returns a human readable message for the calling thread.
 *
 * The reference (zhenxuan00/graphical-gan) has no FFI: its arithmetic lives in
 * TensorFlow 1.x kernels reached from `tflib`.  Each entry point below names the
 * reference call site (file:line under /root/reference) whose TensorFlow kernel it
 * replaces.  Host code (graphical-gan_b200/gg/cabi.py) binds these with ctypes; see
 * INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - all floating tensors are fp32, contiguous, row-major in the stated shape;
 *   - image tensors are NHWC ("channels-last"); the NCHW boundary of tflib is handled
 *     by gg_transpose_b2d on the host side of the ABI;
 *   - conv filters use the reference's own memory layouts:
 *       Conv2D  filter (k,k,Cin,Cout)            tflib/ops/conv2d.py:75-83
 *       Deconv2D filter (k,k,Cout,Cin)           tflib/ops/deconv2d.py:60-69
 *     (a Deconv2D forward is gg_conv2d_dgrad with Ci:=Cout_deconv, Co:=Cin_deconv);
 *   - `act` is a GG_ACT_* code applied after the optional bias, `alpha` its parameter.
 */
#ifndef GG_B200_H_
#define GG_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status ---------------------------------------------------------------------- */
#define GG_OK 0
#define GG_ERR_BAD_ARG 1
#define GG_ERR_UNSUPPORTED 2
#define GG_ERR_WORKSPACE 3
#define GG_ERR_DRIVER 4
#define GG_ERR_CUDA_BASE 1000

const char* gg_last_error(void);
int gg_version(void);
/* number of kernels launched by this library in this process (all threads) */
long long gg_launch_count(void);
void gg_reset_launch_count(void);
/* 0 = auto (tcgen05 implicit GEMM where the shape allows), 1 = force the direct fp32
 * kernels, 2 = force tcgen05 (error if the shape is unsupported). */
int gg_set_conv_backend(int mode);
int gg_get_conv_backend(void);
/* which backend the most recent conv / gemm call on this thread used: 0 direct, 1 tcgen05 */
int gg_last_backend(void);
/* upper bound on the CTAs one tensor-core conv / dense launch may occupy (split-K is sized to it).  148 = the whole
 * GPU (default, best for a kernel running alone); the multi-stream executor sets 74 so two launches can overlap. */
/* programmatic dependent launch for the stream-ordered (non-cooperative) kernels: 1 = launch every kernel with
 * cudaLaunchAttributeProgrammaticStreamSerialization; kernels wait with griddepcontrol.wait after their prologue */
int gg_set_pdl(int on);
int gg_set_tc_max_ctas(int n);
/* cap the operand-ring depth of the tensor-core kernels (0 = as deep as fits; 3 lets two CTAs share an SM) */
int gg_set_tc_stages(int n);
/* launch configuration of the calling thread's most recent tensor-core conv / dense launch (parity tests assert that the
 * production shapes run under the plan's settings): out8 = {mode 0/1/2, tiles (grid.x), K splits (grid.y), n_tile,
 * ring stages, cluster flag, dynamic shared memory bytes, m_tiles} */
int gg_last_tc_info(int* out8);

/* ---- activation codes ------------------------------------------------------------ */
#define GG_ACT_NONE 0
#define GG_ACT_RELU 1
#define GG_ACT_LEAKY 2   /* max(alpha*x, x): script-level LeakyReLU, gmgan_inference_cifar10.py:122-123 */
#define GG_ACT_TANH 3
#define GG_ACT_SIGMOID 4

/* ---- convolution (tf.nn.conv2d, tflib/ops/conv2d.py:106-112; autodiff → Conv2DBackprop*) */
/* y[b,ho,wo,co] = act( sum_{r,s,ci} x[b, ho*stride+r-pad_t, wo*stride+s-pad_l, ci] * w[r,s,ci,co] + bias[co] )
 * out-of-range x reads are zero (TF SAME: pad_t = pad_total/2, the remainder goes after). */
int gg_conv2d_fwd(const float* x, const float* w, const float* bias, float* y,
                  int B, int H, int W, int Ci, int Co, int k, int stride,
                  int pad_t, int pad_l, int Ho, int Wo, int act, float alpha,
                  void* workspace, size_t workspace_bytes, void* stream);

/* dx[b,h,w,ci] = act( sum_{r,s,co : (h+pad_t-r)%stride==0, ...} dy[b,(h+pad_t-r)/stride,(w+pad_l-s)/stride,co] * w[r,s,ci,co] + bias[ci] )
 * = input gradient of gg_conv2d_fwd; with bias/act it is the Deconv2D forward
 * (tf.nn.conv2d_transpose, tflib/ops/deconv2d.py:101-114). */
int gg_conv2d_dgrad(const float* dy, const float* w, const float* bias, float* dx,
                    int B, int H, int W, int Ci, int Co, int k, int stride,
                    int pad_t, int pad_l, int Ho, int Wo, int act, float alpha,
                    void* workspace, size_t workspace_bytes, void* stream);

/* dw[r,s,ci,co] = sum_{b,ho,wo} x[b, ho*stride+r-pad_t, wo*stride+s-pad_l, ci] * dy[b,ho,wo,co]
 * (Conv2DBackpropFilter; for Deconv2D the roles of x and dy swap on the host side).
 * workspace: split-K partial sums, at least gg_conv2d_wgrad_workspace(...) bytes. */
int gg_conv2d_wgrad(const float* x, const float* dy, float* dw,
                    int B, int H, int W, int Ci, int Co, int k, int stride,
                    int pad_t, int pad_l, int Ho, int Wo,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t gg_conv2d_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
/* dx = act'(y_fwd) * dgrad(dy, w): the input-gradient of the conv followed by the gradient of the activation whose OUTPUT y_fwd
 * (same [B,H,W,Ci] layout as dx) fed the conv in the forward pass — autodiff of `LeakyReLU(Conv2D(...))` chains
 * (gmgan_inference_cifar10.py:122-123,276-288) in one launch.  Tensor-core path only: returns GG_ERR_UNSUPPORTED when
 * gg_conv2d_tc_supported(1, ...) is 0. */
int gg_conv2d_dgrad_actgrad(const float* dy, const float* w, float* dx, const float* y_fwd, int act, float alpha,
                            int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Convolution (mode 0, a = x) / transposed convolution (mode 1, a = dy; gg_conv2d_dgrad's argument meaning) that ALSO writes the
 * batch-norm statistics of its output: stats[t][0][c] = sum, stats[t][1][c] = sum of squares of out[., c] over the output rows of
 * m-tile t, t < gg_conv2d_stats_tiles(...), c < (mode 0 ? Co : Ci).  `Batchnorm` always follows a Conv2D / Deconv2D / Linear
 * (tflib/ops/batchnorm.py:29-30 applied to the output of conv2d.py:106-120, e.g. gmgan_inference_cifar10.py:173-180): the moments
 * pass over the activation disappears — gg_bn_apply(x, stats, S = tiles, count = rows, ...) folds the partial rows and
 * normalises.  A Linear layer is the 1x1 geometry (B = rows, H = W = Ho = Wo = k = stride = 1).  Tensor-core path only:
 * gg_conv2d_stats_tiles returns 0 when the shape does not run there, and gg_conv2d_bnstats then fails with GG_ERR_UNSUPPORTED. */
int gg_conv2d_stats_tiles(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo);
int gg_conv2d_bnstats(int mode, const float* a, const float* w, const float* bias, float* out, float* stats,
                      int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                      int act, float alpha, void* workspace, size_t workspace_bytes, void* stream);
/* 1 when mode (0 fwd, 1 dgrad, 2 wgrad) of this geometry runs on the tcgen05 kernels under the current backend setting */
int gg_conv2d_tc_supported(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
/* workspace bytes for mode 0 fwd / 1 dgrad / 2 wgrad.  The workspace holds split-K partial tiles and, in its first
 * bytes, per-tile arrival tickets: it must be ZERO-INITIALISED ONCE by the caller; every call leaves the tickets zero.
 * A smaller (or NULL) workspace is legal: the call then runs unsplit or on the direct kernels. */
size_t gg_conv2d_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);

/* ---- dense (tf.matmul, tflib/ops/linear.py:132-146) -------------------------------- */
/* C[M,N] = act( op(A) * op(B) + bias[N] ),  op(A) is [M,K], op(B) is [K,N]; row-major storage,
 * ta/tb = 1 means the stored matrix is the transpose ([K,M] resp. [N,K]). */
int gg_gemm(const float* A, const float* Bm, const float* bias, float* C,
            int M, int N, int K, int ta, int tb, int act, float alpha,
            void* workspace, size_t workspace_bytes, void* stream);
size_t gg_gemm_workspace(int M, int N, int K);

/* ---- batch normalisation (tf.nn.fused_batch_norm, tflib/ops/batchnorm.py:30; moments+batch_normalization :77-84) */
/* x is [R,C] (NHWC flattened, or [B,C] for axes=[0]); statistics are per column over the R rows.
 * stats: partial[s][0][c] = sum x, partial[s][1][c] = sum x^2 for row slice s of S (S returned by gg_bn_slices). */
int gg_bn_slices(int R, int C);
int gg_bn_stats(const float* x, float* partial, int R, int C, void* stream);
/* y = act((x-mean)*rstd*gamma+beta); mean/rstd are computed from `partial` (S slices, `count` rows
 * in total — count may exceed R when the partials were all-reduced over data-parallel ranks). */
int gg_bn_apply(const float* x, const float* partial, int S, float count,
                const float* gamma, const float* beta, float eps,
                float* y, float* mean_out, float* rstd_out,
                int R, int C, int act, float alpha, void* stream);
/* backward: g = act'(y)*dy ; partial[s][0][c] = sum g, partial[s][1][c] = sum g*xhat */
int gg_bn_bwd_reduce(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                     const float* gamma, const float* beta,
                     float* partial, int R, int C, int act, float alpha, void* stream);
/* dx = gamma*rstd*(g - sum_g/count - xhat*sum_gxhat/count); dgamma = sum_gxhat; dbeta = sum_g (local sums) */
int gg_bn_bwd_apply(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* partial, int S, float count,
                    float* dx, float* dgamma, float* dbeta,
                    int R, int C, int act, float alpha, void* stream);
/* sum S partial slices [S,2,C] into out[2,C] (used before the SyncBN all-reduce) */
int gg_bn_fold_partials(const float* partial, int S, float* out, int C, void* stream);
/* Single-kernel forms of the same two operations (batch statistics computed inside; no partial buffers): one
 * thread-block cluster per group of channels, rows split over the cluster's CTAs, partial sums exchanged through
 * distributed shared memory.  Used whenever the statistics need no cross-rank exchange (single GPU, or SyncBN off).
 * gg_bn_fused_supported: 1 when (R, C) is handled (C % 4 == 0), else use the multi-kernel entry points above. */
int gg_bn_fused_supported(int R, int C);
int gg_bn_fwd_fused(const float* x, const float* gamma, const float* beta, float eps,
                    float* y, float* mean_out, float* rstd_out, int R, int C, int act, float alpha, void* stream);
int gg_bn_bwd_fused(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                    const float* gamma, float* dx, float* dgamma, float* dbeta,
                    int R, int C, int act, float alpha, void* stream);

/* ---- elementwise / layout glue (script-level tf.* ops, SURVEY.md §8(a) a6,a7,a13) -------- */
#define GG_U_COPY 0
#define GG_U_RELU 1
#define GG_U_LEAKY 2
#define GG_U_TANH 3
#define GG_U_SIGMOID 4
#define GG_U_EXP 5
#define GG_U_LOG 6
#define GG_U_SQRT 7
#define GG_U_SQUARE 8
#define GG_U_NEG 9
#define GG_U_ABS 10
#define GG_U_AFFINE 11   /* a*x + b */
#define GG_U_POW 12      /* x^a */
#define GG_U_RSQRT 13
#define GG_U_RECIP 14
#define GG_U_BCE 15      /* max(x,0) - x*a + log1p(exp(-|x|)), a = label  (tf.nn.sigmoid_cross_entropy_with_logits) */
#define GG_U_CLIP 16     /* min(max(x,a),b) */
#define GG_U_SIGN 17
#define GG_U_SOFTSIGN 18
#define GG_U_DIVC 19     /* x / a  (true division, e.g. the /255. of the input decode) */
#define GG_U_RDIVC 20    /* a / x */
int gg_unary(int op, const float* x, float* y, long long n, float a, float b, void* stream);

#define GG_B_ADD 0
#define GG_B_SUB 1
#define GG_B_MUL 2
#define GG_B_DIV 3
#define GG_B_MAX 4
#define GG_B_MIN 5
#define GG_B_RELU_GRAD 6     /* (a > 0) ? b : 0          a = activation output, b = upstream grad */
#define GG_B_LEAKY_GRAD 7    /* (a > 0) ? b : alpha*b */
#define GG_B_TANH_GRAD 8     /* (1 - a*a) * b */
#define GG_B_SIGMOID_GRAD 9  /* a*(1-a) * b */
#define GG_B_BCE_GRAD 10     /* (sigmoid(a) - alpha) * b */
#define GG_B_GE_MASK 11      /* (a >= b) ? 1 : 0 */
#define GG_B_GT_MASK 12      /* (a >  b) ? 1 : 0 */
#define GG_B_ABS_GRAD 13     /* sign(a) * b */
#define GG_B_POW 14          /* a^b */
/* broadcasting binary op on up to 4 dims: out[i0,i1,i2,i3] = f(a[sa . i], b[sb . i]); strides in elements, 0 = broadcast */
int gg_binary(int op, const float* a, const float* b, float* out,
              const int* dims4, const int* sa4, const int* sb4, float alpha, void* stream);

/* y[o,i] = reduce_r x[o,r,i]; op 0 sum, 1 mean, 2 max */
int gg_reduce(int op, const float* x, float* y, int outer, int red, int inner, void* stream);
/* same, with a scratch buffer (gg_reduce_workspace bytes) that lets tall column reductions (bias gradients) use all SMs */
int gg_reduce_ws(int op, const float* x, float* y, int outer, int red, int inner, void* workspace, size_t workspace_bytes,
                 void* stream);
size_t gg_reduce_workspace(int outer, int red, int inner);
/* row-wise softmax of x[R,C] and its backward dx = y*(dy - sum(dy*y)) */
int gg_softmax_fwd(const float* x, float* y, int R, int C, void* stream);
int gg_softmax_bwd(const float* y, const float* dy, float* dx, int R, int C, void* stream);
/* y[b,c,r] = x[b,r,c]  (NHWC<->NCHW is a batched 2-D transpose) */
int gg_transpose_b2d(const float* x, float* y, int Bt, int R, int C, void* stream);
/* the same with non-contiguous batches (batch b of x at x + b*x_batch_stride, of y at y + b*y_batch_stride, strides in floats) and
 * an optional fused activation gradient y = act'(mask) * transpose(x) (mask: forward activation in the OUTPUT layout, contiguous):
 * the NCHW-flatten of `tf.reshape(output, [-1, 4*4*4*DIM])` written straight into `tf.concat([output, z_output], 1)`
 * (gmgan_inference_cifar10.py:292-294) and its mirror in the backward pass, each without the copy launches around it. */
int gg_transpose_b2d_ex(const float* x, float* y, int Bt, int R, int C, long long x_batch_stride, long long y_batch_stride,
                        const float* mask, int mask_act, float mask_alpha, void* stream);
/* generic permutation of up to 4 dims: y = transpose(x, perm), dims = shape of x */
int gg_transpose4(const float* x, float* y, const int* dims4, const int* perm4, void* stream);
/* strided 2-D copy (concat / slice building block): dst[r*dst_ld + c] = src[r*src_ld + c]; accumulate!=0 adds */
int gg_copy2d(const float* src, long long src_ld, float* dst, long long dst_ld,
              long long rows, long long cols, int accumulate, void* stream);
int gg_fill(float* x, long long n, float v, void* stream);
int gg_one_hot(const int32_t* idx, float* out, int n, int depth, void* stream);
int gg_argmax(const float* x, int32_t* idx, int R, int C, void* stream);
/* out[m,:] = table[idx[m],:] (zeros for idx[m] outside [0,depth)) + addend[m,:] (addend may be NULL):
 * tf.matmul(tf.one_hot(idx, depth), table) + noise — the mixture-prior sample `mu_k + eps` of gmgan_inference_cifar10.py:138-141 —
 * as a row gather */
int gg_gather_rows(const int32_t* idx, const float* table, const float* addend, float* out, int M, int N, int depth, void* stream);
/* y = a * float(x) + b ; x int32 (tf.cast of the int32 image placeholder, gmgan_inference_cifar10.py:341-342) */
int gg_cast_i32_f32(const int32_t* x, float* y, long long n, float a, float b, void* stream);
int gg_cast_u8_f32(const uint8_t* x, float* y, long long n, float a, float b, void* stream);
int gg_cast_f32_i32(const float* x, int32_t* y, long long n, void* stream);
/* y[i] = (int32) x[i]: a uint8 image batch (the dtype tflib/cifar10.py:8-48 and svhn / celebA loaders hand to
 * session.run) widened on the device into the graph's int32 placeholder (gmgan_inference_cifar10.py:341) — one byte per
 * pixel crosses PCIe instead of four.  Both pointers 16-byte aligned.  Bit exact. */
int gg_widen_u8_i32(const uint8_t* x, int32_t* y, long long n, void* stream);
/* out = sum_i in[i] over `count` same-sized tensors whose device pointers are listed in ptrs (host array) */
int gg_add_n(const float* const* ptrs, int count, float* out, long long n, void* stream);

/* ---- fused element-wise programs ----------------------------------------------------------------------------------------
 * The script-level glue of the reference is long chains of tiny tf.* ops: the Gumbel-softmax noise
 * `-tf.log(-tf.log(U + eps) + eps)` and the soft assignment around it (gmgan_inference_cifar10.py:155-163), the input
 * decode `2*((tf.cast(x, tf.float32)/255.)-.5)` (:341-342), the mixture-prior distances, and their gradients.  Each op is a
 * 2-3 us launch on a KB-sized tensor; a connected group of them runs as ONE launch of a small register program evaluated
 * per output element: values live in registers 0..GG_EW_REGS-1, registers 0..n_in-1 are loaded from the inputs (broadcast by
 * per-dimension element strides, 0 = broadcast; int32 inputs are converted like tf.cast), then `instr` runs in order — the
 * SAME arithmetic as gg_unary / gg_binary (bit-identical results) — and up to GG_EW_MAX_OUT registers are stored.
 * reduce_op != 0: the iteration space is [rows, red] (red = dims[3]) and output 0 is reduced along `red` with the summation
 * order of gg_reduce's row kernel (one CTA per row) instead of being stored per element: out[0][row] = sum / mean / max. */
#define GG_EW_MAX_IN 12
#define GG_EW_MAX_OUT 4
#define GG_EW_MAX_INSTR 40
#define GG_EW_REGS 32
typedef struct {
  int kind;        /* 0: regs[dst] = unary(op, regs[src0], a, b)   1: regs[dst] = binary(op, regs[src0], regs[src1], alpha = a) */
  int op;          /* GG_U_* / GG_B_* */
  int dst, src0, src1;
  float a, b;
} gg_ew_instr;
typedef struct {
  int n_in, n_out, n_instr;
  int flat;                              /* 1: every input has the (contiguous) shape of the iteration space */
  int reduce_op;                         /* 0 none, 1 sum, 2 mean, 3 max (output 0 only) */
  int dims[4];
  const void* in[GG_EW_MAX_IN];
  int in_is_int[GG_EW_MAX_IN];
  int in_stride[GG_EW_MAX_IN][4];
  float* out[GG_EW_MAX_OUT];
  int out_reg[GG_EW_MAX_OUT];
  gg_ew_instr instr[GG_EW_MAX_INSTR];
} gg_ew_program;
int gg_ew_run(const gg_ew_program* prog, void* stream);
int gg_ew_program_bytes(void);   /* sizeof(gg_ew_program): lets a foreign-language binding verify its struct layout */

/* ---- fused loss reductions (tflib/objs/gan_inference.py:85-101; tflib/utils/distance.py:3-7; gan_inference_svhn.py:353-354) */
/* out[0] (+)= weight * mean_i BCE(x_i, label); accumulate!=0 adds to out[0] */
int gg_bce_mean(const float* x, int n, float label, float weight, float* out, int accumulate, void* stream);
/* dx_i = (sigmoid(x_i)-label) * weight/n * (gscale ? gscale[0] : 1) */
int gg_bce_mean_grad(const float* x, int n, float label, float weight, const float* gscale, float* dx, int accumulate, void* stream);
/* out[0] = weight * mean((x-y)^2) (p=2) or mean(|x-y|) (p=1) */
int gg_dist_mean(const float* x, const float* y, long long n, int p, float weight, float* out, int accumulate, void* stream);
/* slopes[r] = sqrt(sum_c g[r,c]^2); out[0] = weight * mean_r (slopes[r]-1)^2 */
int gg_gp_slope_penalty(const float* g, int R, int C, float weight, float* slopes, float* out, void* stream);

/* ---- Adam (tf.train.AdamOptimizer, constructed at tflib/objs/gan_inference.py:108-117) --- */
/* TensorFlow ApplyAdam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m=b1 m+(1-b1)g; v=b2 v+(1-b2)g^2; p -= lr_t*m/(sqrt(v)+eps).
 * table (device): per tensor {param*, grad*, m*, v*, n} as gg_adam_entry; chunks (device): {tensor, offset} pairs of
 * GG_ADAM_CHUNK elements. state (device, 2 doubles + 1 int64): {b1^t, b2^t, t}; gg_adam_multi first advances t
 * (the advance is done by a single-thread kernel so CUDA-graph replays step correctly). grad_scale multiplies g
 * (1/world_size after a summing all-reduce). */
typedef struct { float* p; const float* g; float* m; float* v; long long n; } gg_adam_entry;
typedef struct { int tensor; int pad; long long offset; } gg_adam_chunk;
#define GG_ADAM_CHUNK 4096
int gg_adam_multi(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks, void* state,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* RMSProp (tf.train.RMSPropOptimizer defaults decay .9 momentum 0 eps 1e-10; gan_inference.py:8-13): ms in m */
/* gg_adam_multi = gg_adam_tick (advance {b1^t, b2^t, t} once per step) + gg_adam_apply (the update of the listed tensors with
 * the CURRENT state).  A step may apply disjoint tensor sets at different times: parameters whose gradients are ready early are
 * updated while the last backward kernels still run (gg/executor.py _emit_optimizer). */
int gg_adam_tick(void* state, float beta1, float beta2, void* stream);
int gg_adam_apply(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks, const void* state,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
int gg_rmsprop_multi(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks,
                     float lr, float decay, float eps, float grad_scale, void* stream);
/* multi-tensor gather/scatter between a table of tensors and one flat bucket (gradient all-reduce staging) */
int gg_pack_grads(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks,
                  const long long* flat_offsets, float* flat, int to_flat, void* stream);

/* ---- random numbers (tf.random_normal / tf.random_uniform / Categorical.sample) ---------- */
/* Philox4x32-10 keyed by (seed, stream_id), counter = (*tick_counter, element index). tick is a device int64
 * advanced by gg_rng_tick once per session.run so graph replays draw fresh numbers. */
int gg_rng_tick(void* tick_counter, void* stream);
int gg_rng_normal(float* out, long long n, float mean, float stddev, uint64_t seed, uint32_t stream_id,
                  const void* tick_counter, void* stream);
int gg_rng_uniform(float* out, long long n, float lo, float hi, uint64_t seed, uint32_t stream_id,
                   const void* tick_counter, void* stream);
/* idx[i] ~ Categorical(probs[0..K)), inverse-CDF on a Philox uniform */
int gg_rng_categorical(int32_t* idx, int n, const float* probs, int K, uint64_t seed, uint32_t stream_id,
                       const void* tick_counter, void* stream);

/* ---- small-message all-reduce over NVLink peer memory (SyncBN statistics; new functionality) ------ */
/* Each rank allocates one exchange buffer (gg_comm_alloc -> device pointer + 64-byte CUDA IPC handle), the handles are
 * exchanged by the host (torch.distributed all_gather), every rank maps its peers' buffers (gg_comm_open).
 * gg_allreduce_small: dst[i] = sum over ranks of src[i], n <= max_floats, one kernel, CUDA-graph capturable;
 * epoch_counter is a device uint32 owned by the caller (zero-initialised), peer_bufs_host a host array of `world` device
 * pointers ordered by rank (entry `rank` = the local buffer).  All ranks must issue the same sequence of calls. */
size_t gg_comm_buffer_bytes(int max_floats);
int gg_comm_alloc(int max_floats, void** buf_out, void* ipc_handle_out_64);
int gg_comm_open(const void* ipc_handle_64, void** buf_out);
/* zero-filled exchange arena of `bytes` (cudaMalloc + CUDA IPC handle), for the SyncBN call sites below */
int gg_comm_alloc_bytes(size_t bytes, void** buf_out, void* ipc_handle_out_64);

/* ---- data-parallel batch norm in ONE launch (SyncBN; tflib/ops/batchnorm.py:30,77-84 evaluated on the GLOBAL batch) -------
 * Same arithmetic as gg_bn_fwd_fused / gg_bn_bwd_fused on this rank's R rows, with the per-channel sums totalled over the
 * `world` ranks inside the kernel: every channel-group cluster writes its sums into all peers' arenas over NVLink and
 * spins on the peers' epoch flags (csrc/gg_bn_fused.cu).  peer_arenas_host[r] = rank r's arena as mapped in THIS process
 * (gg_comm_alloc_bytes / gg_comm_open); site_offset = this call site's region (gg_bn_dp_site_bytes(C, world) bytes,
 * 128-byte aligned, the same offset on every rank, never shared between two call sites).  Every rank must launch the same
 * call sites the same number of times.  dgamma / dbeta are the LOCAL sums (the gradient all-reduce totals them). */
size_t gg_bn_dp_site_bytes(int C, int world);
/* CTAs of the one-launch batch-norm kernels for an [R, C] input (0: unsupported shape) */
int gg_bn_fused_grid(int R, int C);
int gg_bn_fwd_fused_dp(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean_out,
                       float* rstd_out, int R, int C, int act, float alpha, void* const* peer_arenas_host, int rank, int world,
                       long long site_offset, void* stream);
int gg_bn_bwd_fused_dp(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma,
                       float* dx, float* dgamma, float* dbeta, int R, int C, int act, float alpha,
                       void* const* peer_arenas_host, int rank, int world, long long site_offset, void* stream);
int gg_allreduce_small(const float* src, float* dst, int n, void* const* peer_bufs_host, int rank, int world,
                       int max_floats, void* epoch_counter, void* stream);

/* ---- hardware probes used by the tests (not part of the training path) ---------------- */
/* TMA strided-box probe: loads x[b, h0 + 2*i, w0 + 2*j, c0..c0+32) for i<hb, j<wb into out[hb][wb][32] with a
 * tiled tensor map using elementStrides=(1,2,2,1), zero-filling out-of-range coordinates. `out` receives the raw shared-memory image
 * (128B-swizzled when swizzle128 != 0). */
int gg_probe_tma_strided(const float* x, int B, int H, int W, int C, int b, int h0, int w0, int c0,
                         int hb, int wb, int swizzle128, float* out, void* stream);
/* plain tcgen05 tf32 GEMM probe: D[128,N] = A[128,K] * B, with A K-major (row-major [128,K]) or MN-major
 * (stored [K,128]) and B K-major (stored [N,K]) or MN-major (stored [K,N]); K multiple of 32, N multiple of 32 <= 256 */
int gg_probe_umma_tf32(const float* A, const float* Bm, float* D, int N, int K, int a_mn_major, int b_mn_major,
                       int tma_tf32_convert, void* stream);

/* development aid: when set to a device buffer of >= 256 int64, CTA (0,0) of every tensor-core conv launch writes a
 * %globaltimer timeline (slot 0 start, 1+i TMA issue of k-block i, 64+i operands landed, 128 accumulator ready,
 * 129 partial written, 131 epilogue done).  NULL disables it. */
int gg_debug_set_buffer(void* device_buffer_256_int64);
/* measurement aid: on != 0 replaces EVERY kernel launch of the library by an empty one-warp kernel on the same stream, so a captured
 * step keeps its node / dependency structure but does no work (tools/exp_null_step.py measures the graph's own launch + dependency
 * cost this way).  Results are garbage while it is on. */
int gg_set_null_launch(int on);
/* development aid for the one-launch small-channel filter gradient (gg_conv2d_wgrad with Ci <= 4): when set to a device buffer
 * of >= 8 * 160 int64, CTA b writes %globaltimer stamps at [8*b + s]: s = 0 entry, 1 tiles staged, 2 FMA loops done,
 * 3 cluster rendezvous, 4 cluster partial in L2, 5 ticket taken, 6 (last cluster) dw written.  NULL disables it. */
int gg_debug_set_small_buffer(void* device_buffer_int64);
/* its launch plan for a geometry: out8 = {served, rows per unit, units, threads, pixel groups, clusters, dynamic smem bytes,
 * clusters of 8 CTAs the device keeps resident at once (cudaOccupancyMaxActiveClusters)} */
int gg_debug_small_wgrad_info(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo, int* out8);

/* ---- tooling: timed event nodes inside a captured CUDA graph (tools/trace_step.py) ------ */
int gg_trace_event_create(void** event_out);
int gg_trace_event_record(void* event, void* stream);   /* cudaEventRecordExternal: an event-record node under capture */
int gg_trace_event_elapsed_us(void* start, void* end, float* us_out);

#ifdef __cplusplus
}
#endif
#endif /* GG_B200_H_ */
