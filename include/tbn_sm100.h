/*
 * tbn_sm100.h -- C ABI of libtbn_sm100.so: the B200 (sm_100a) kernels of the Temporal Binding
 * Network forward hot path.
 *
 * The reference (tridivb/attention_based_tbn) is pure PyTorch: it has no FFI of its own, every
 * entry point below replaces a run of ATen/cuDNN library calls made from the reference's
 * Python.  The reference site each one replaces is cited as file:line (relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (never throws); the message of the
 *     last error on the calling thread is returned by tbn_last_error().
 *   - pointers are DEVICE pointers unless the name ends in _host; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).  Nothing synchronises the device.
 *   - activations are NHWC ("channels last"): element (n,y,x,c) of a tensor with channel stride
 *     `cs` and channel offset `co` lives at base[((n*H + y)*W + x)*cs + co + c].  A conv / pool
 *     reads a channel slice and writes a channel slice, which is how the inception-branch concat
 *     (reference core/models/bn_inception_audio.py:485-493, 593-600, ...) and the mid-fusion
 *     concat (core/models/model.py:250) are done in place.
 *   - dtype: TBN_F32 (0) or TBN_BF16 (1) is the STORAGE type of activations; accumulation is
 *     always fp32.
 */
#ifndef TBN_SM100_H
#define TBN_SM100_H

#ifdef __cplusplus
extern "C" {
#endif

#define TBN_F32 0
#define TBN_BF16 1

#define TBN_SEG_RELU 1
#define TBN_SEG_BIAS 2

#define TBN_POOL_AVG 0 /* 3x3 s1 p1, count_include_pad (always /9) */
#define TBN_POOL_MAX 1 /* 3x3 s1 p1, -inf padding */

/* Library version (major*10000 + minor*100 + patch). */
int tbn_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char* tbn_last_error(void);
/* SM count and compute capability of the current device. */
int tbn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Tuning / diagnostic knobs (process-wide).  Keys: "tc_halo" (1 = 3x3 convs read one halo tile per
 * 64-channel slice and feed the nine taps from it, 0 = one TMA box per tap), "tc_occ2" (1 = two CTAs per SM for N tiles <= 128 columns),
 * "tc_max_bn" (widest N tile, 32..256 columns), "tc_teams3" (1 = a third epilogue team where it shortens the longest team's share),
 * "tc_bres" (resident weights for cin <= 64 halo-mode 3x3 convs: 0 off, 1 N tiles <= 96 columns, 2 = default: CTA pairs for conv2_3x3), "tc_pdl" (1 = programmatic dependent launch),
 * "tc_cta2" (CTA pairs / tcgen05 cta_group::2: 0 never, 1 for N tiles wider than 128 columns, 2 always, 3 = default: wide tap-mode 3x3 convs only), "tc_debug" (ablation bits for profiling experiments; results are
 * WRONG when non-zero). */
int tbn_set_tuning(const char* key, int value);

/* Diagnostic: when dev_buf != NULL the persistent conv kernel writes per-role clock64 stamps of its first
 * `ctas` CTAs into it (ctas*stride uint64; layout in conv_tc.cu).  NULL switches tracing off. */
int tbn_debug_trace(void* dev_buf, int* ctas, int* stride, int* n);
/* Diagnostic: clock64 stamps of the fused stem's phases (CTA 0, first 8 tiles x 16 stamps) into dev_buf; NULL = off. */
int tbn_debug_trace_stem(void* dev_buf);

/* Diagnostic: enqueue a one-thread kernel that writes the device's %globaltimer (ns) to *slot (uint64, device). */
int tbn_debug_stamp(void* slot, void* stream);

/* Input layout conversion, NCHW fp32 -> NHWC dtype with channels zero-padded to c_pad.
 * Replaces the implicit layout of `input[m].view(b*n, c, h, w)` feeding conv1
 * (core/models/model.py:211-213; inputs are NCHW float32, core/dataset/transform.py:495-497). */
int tbn_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int w, int c_pad, int dtype, void* stream);

/* conv2d(+bias) -> BatchNorm2d(eval) -> ReLU as ONE kernel on CUDA cores (fp32 FMA), implicit GEMM.
 * BatchNorm is folded on the host: w_folded = s*w, bias_folded = beta + s*(b - mean),
 * s = gamma/sqrt(var + 1e-5).  Exact-precision path (fp32 parity mode) and checker of the
 * tensor-core path.  wk: fp32 [k*k*cin][cout] (tap-major, then input channel).
 * Replaces every conv/bn/relu triple of BNInception.features
 * (core/models/bn_inception_audio.py:24-31, 58-401; dataflow 437-1003). */
int tbn_conv_simt(const void* x, const float* wk, const float* bias, void* y,
                  int n, int h, int w, int cin, int x_cs, int x_co,
                  int cout, int y_cs, int y_co, int k, int stride, int pad, int relu,
                  int dtype, void* stream);

/* Same contract on the 5th-generation tensor cores: bf16 operands, fp32 accumulation in TMEM,
 * TMA-fed implicit GEMM (tcgen05.mma), bias + ReLU fused in the epilogue, output written at its
 * channel offset.  stride must be 1 and pad == k/2 (k = 1 or 3).
 * wp: bf16 [cout][k*k][cin_pad64] (K-major, cin zero-padded to a multiple of 64).
 * out_dtype: TBN_BF16 or TBN_F32. */
int tbn_conv_tc(const void* x, const void* wp, const float* bias, void* y,
                int n, int h, int w, int cin, int x_cs, int x_co,
                int cout, int y_cs, int y_co, int k, int relu, int out_dtype,
                int bias_period, void* stream);

/* 3x3 / stride 2 / pad 1 variant (inception_3c, inception_4e; bn_inception_audio.py:130-157, 302-329): the
 * TMA boxes traverse the input with element stride 2, so no patch matrix is materialised.  h, w are the INPUT
 * size; the output is ((h-1)/2+1) x ((w-1)/2+1).  Same weight packing as tbn_conv_tc(k=3). */
int tbn_conv_tc_s2(const void* x, const void* wp, const float* bias, void* y,
                   int n, int h, int w, int cin, int x_cs, int x_co,
                   int cout, int y_cs, int y_co, int relu, int out_dtype, void* stream);

/* Horizontally fused variant: one GEMM whose output columns are split over up to 4 destinations
 * (segment s receives `cols` consecutive output channels).  Used to run the same-input 1x1 convs
 * of an inception block (1x1, 3x3_reduce, double_3x3_reduce; bn_inception_audio.py:58-85, ...) as
 * ONE kernel that reads the block input once and still writes each branch where it belongs.
 * segs_host is a HOST array. */
typedef struct tbn_out_seg {
    void* y;   /* device pointer of the destination tensor */
    int y_cs;  /* its channel stride */
    int y_co;  /* channel offset of this segment inside it */
    int cols;  /* output channels of this segment (multiple of 16) */
    int flags; /* TBN_SEG_BIAS | TBN_SEG_RELU: epilogue applied to this segment */
} tbn_out_seg;
int tbn_conv_tc_multi(const void* x, const void* wp, const float* bias, const tbn_out_seg* segs_host, int nseg,
                      int n, int h, int w, int cin, int x_cs, int x_co, int k, int out_dtype, void* stream);

/* Grouped launch: up to 16 independent convolutions (each described like a tbn_conv_tc_multi call) handed over together.
 * Problems that agree on the kernel variant run as ONE persistent grid over the concatenated tile lists: the same layer
 * of the RGB / flow / audio backbones (the modality loop of core/models/model.py:208-213), or the independent 3x3 convs
 * of one inception block (bn_inception_audio.py dataflow 437-1003).  A launch then fills whole waves of the 148 SMs (the
 * 7x7 / 8x13 maps of inception 5a / 5b alone have 1-2.6 waves) and pays the pipeline fill / drain once.  Results are
 * bit-identical to separate tbn_conv_tc / _s2 / _multi calls.  descs (and every segs array) are HOST arrays. */
typedef struct tbn_conv_desc {
    const void* x;            /* NHWC bf16 input */
    const void* wp;           /* packed weights (tbn_conv_tc layout) */
    const float* bias;        /* folded bias, cout entries */
    const tbn_out_seg* segs;  /* output segments (1..4) */
    int nseg;
    int n, h, w;              /* input geometry */
    int cin, x_cs, x_co;
    int k, stride;            /* 1 or 3; 1, or 2 with k = 3 */
    int out_dtype;            /* TBN_BF16 / TBN_F32 */
} tbn_conv_desc;
int tbn_conv_tc_group(const tbn_conv_desc* descs, int n, void* stream);

/* Whole network stem in one kernel (cin = 1, 3 or 10): conv1_7x7_s2 + BN + ReLU -> pool1 (MaxPool 3x3/2, ceil) ->
 * conv2_3x3_reduce (1x1, 64 -> 64) + BN + ReLU (bn_inception_audio.py:21-31, 35-39; dataflow 417-426), NCHW in,
 * NHWC bf16 out of size pool(conv(h)) x pool(conv(w)) x 64.  Neither the conv1 output nor the pool1 output is ever
 * written to memory.  Inside the kernel the tensor-core operands are FP16 (input, both weight sets, pooled conv1 output:
 * the first layers difference large smooth inputs, fp16's extra mantissa bits cost nothing on kind::f16 and remove
 * most of the stem's share of the bf16 error); the output is bf16.  x: fp32 (x_dtype = TBN_F32, the reference's input
 * contract) or bf16 (TBN_BF16: half the host-link bytes; equals the fp32 path on inputs pre-rounded to bf16).
 * w1p: conv1 weights as FP16 [64 cout][K rounded up to 64], column (c*7 + r)*8 + s + 1 = w[c][r][s] (r, s = filter
 * row / column; the column with s + 1 == 0 of every (c, r) chunk is zero: an operand chunk is 8 consecutive input
 * columns), columns cin*56 and cin*56 + 1 = the folded conv1 bias as fp16 hi and lo parts (the operand has ones there,
 * so the bias add rides on the MMA); w2p: conv2_3x3_reduce weights FP16 [64][64] (K-major, the tbn_conv_tc packing);
 * bias2: its folded bias. */
int tbn_stem_fused_tc(const void* x, int x_dtype, const void* w1p, const void* w2p, const float* bias2, void* y, int n,
                      int cin, int h, int w, int y_cs, int y_co, void* stream);

/* AvgPool2d(3,1,1,count_include_pad) applied to a PRE-activation, then + bias and ReLU:
 * avgpool and the 1x1 pool_proj conv commute (both linear, zero padding), so the projection runs
 * first inside the block's fused GEMM and the pooling touches Cout instead of Cin channels
 * (bn_inception_audio.py:86-93, 122-129, ...). */
int tbn_avgpool3s1_bias_relu(const void* x, const float* bias, void* y, int n, int h, int w, int c,
                             int x_cs, int x_co, int y_cs, int y_co, int dtype, void* stream);

/* Patch gather ("im2col") for the strided convs (7x7/2 stem, 3x3/2 of 3c and 4e): writes
 * A[m][kpad], m = (n,oy,ox), column = (tap, cin), zero padded.  x is NHWC (TBN_F32/TBN_BF16) or,
 * with x_nchw != 0, the raw NCHW fp32 network input (fuses tbn_nchw_to_nhwc into the stem).
 * A is bf16.  Followed by tbn_conv_tc(k=1) on A viewed as an (1,1,M,kpad) image. */
int tbn_im2col(const void* x, void* a, int n, int h, int w, int cin, int x_cs, int x_co,
               int k, int stride, int pad, int kpad, int x_dtype, int x_nchw, void* stream);

/* MaxPool2d(3, stride 2, ceil_mode=True) (core/models/bn_inception_audio.py:21-23, 32-34, 155-157,
 * 327-329); writes into a channel slice (pass-through branch of the 3c/4e concat). */
int tbn_maxpool3s2(const void* x, void* y, int n, int h, int w, int c, int x_cs, int x_co,
                   int y_cs, int y_co, int dtype, void* stream);

/* 3x3 stride-1 pad-1 pooling: AvgPool2d(count_include_pad=True) (bn_inception_audio.py:86, 122, ...)
 * or MaxPool2d (5b, :394-396). */
int tbn_pool3s1(const void* x, void* y, int n, int h, int w, int c, int x_cs, int x_co,
                int y_cs, int y_co, int mode, int dtype, void* stream);

/* Feature pooling of the overridden BNInception.logits (core/models/bn_inception.py:16-35):
 * freq_only == 0: mean over (H,W)  -> y[n*y_ld + c]            (fp32; y may be a column block of
 *                                                               the mid-fusion row, model.py:250)
 * freq_only != 0: mean over H only -> y[(n*W + x)*y_ld + c]    (fp32, position-major). */
int tbn_avgpool_feat(const void* x, float* y, int n, int h, int w, int c, int y_ld, int freq_only,
                     int dtype, void* stream);
/* The same, additionally (or only: y may be NULL) writing every pooled row as the bf16 [hi | hi | lo] split operand of a
 * head tensor-core GEMM -- three blocks of c columns at split + row * split_ld (see tbn_split_bf16x3) -- so that the
 * positional-encoding, q.k and Fusion GEMMs (model.py:65, 229-237, 337-362) read their input without a split launch. */
int tbn_avgpool_feat_split(const void* x, float* y, void* split, long long split_ld, int n, int h, int w, int c,
                           int y_ld, int freq_only, int dtype, void* stream);

/* y[m][j] = act( sum_k x[m*ldx + k] * wt[j*ldw + k] + bias[(m % bias_period)*n + j] ), fp32 on CUDA
 * cores.  bias may be NULL; bias_period >= 1.  Replaces nn.Linear / Conv1d(k=1) of the heads:
 * pe[1] (model.py:65), MHA in/out projections (attention.py:48-57), Fusion (model.py:337-362),
 * Classifier (model.py:365-386). */
int tbn_linear(const float* x, const float* wt, const float* bias, float* y,
               int m, int n, int k, int ldx, int ldw, int ldy, int bias_period, int relu, void* stream);

/* fp32 (m, k) -> bf16 (m, 3k): per group of `group` columns the three blocks [hi | hi | lo] (hi = bf16(x),
 * lo = bf16(x - hi)).  With weights packed [hi | lo | hi] a bf16 tcgen05 GEMM (tbn_conv_tc, k = 1, cin = 3k) computes
 * the fp32 product to ~2^-16 relative: this is how the head GEMMs (pe[1], MHA projections, Fusion, Classifier;
 * model.py:65, 337-386, attention.py:48-57) run on the tensor cores without giving up fp32-class accuracy. */
int tbn_split_bf16x3(const float* x, void* out, int m, int k, int ldx, int group, void* stream);
/* The same into rows of a wider split buffer (leading dimension out_ld >= 3k elements). */
int tbn_split_bf16x3_ld(const float* x, void* out, long long out_ld, int m, int k, int ldx, int group, void* stream);

/* Split-K GEMM for tall-K, tiny-M products (the head GEMMs have M = frames): partial[s] (m, n) fp32, s < ksplit, =
 * x[:, K_s] (m, k bf16, row-major) @ wp[:, K_s]^T (n, k bf16) on tcgen05, then
 * y = act(sum_s partial[s] + bias[(row % bias_period)]) with leading dimension ldy. */
int tbn_gemm_tc_splitk(const void* x, const void* wp, void* partial, int m, int k, int n, int ksplit, void* stream);
int tbn_splitk_reduce(const float* partial, const float* bias, float* y, int m, int n, int ldy, int ksplit,
                      int bias_period, int relu, void* stream);
/* Split-K GEMM whose left operand has a row stride ldx >= k (a column block of a wider split buffer). */
int tbn_gemm_tc_splitk_ld(const void* x, int ldx, const void* wp, void* partial, int m, int k, int n, int ksplit,
                          void* stream);
/* Reduction with the rest of the head chain folded in: optional mean over groups of nseg consecutive activated rows
 * (the segment consensus of model.py:194-199 moved in front of the affine classifier it commutes with), result as fp32
 * (y, may be NULL) and / or as the bf16 [hi | hi | lo] operand of the next GEMM (split, may be NULL; three blocks of n
 * columns at split + row * split_ld). */
int tbn_splitk_reduce_ex(const float* partial, const float* bias, float* y, int m, int n, int ldy, int ksplit,
                         int bias_period, int relu, int nseg, void* split, long long split_ld, void* stream);

/* Grouped variant: group g uses x + g*gx, wt + g*gw, bias + g*gb, y + g*gy (element strides); used
 * for the per-head projections of the restructured attention (W_k,h^T q_h and W_v,h zbar_h). */
int tbn_linear_grouped(const float* x, const float* wt, const float* bias, float* y,
                       int m, int n, int k, int ldx, int ldw, int ldy, int groups,
                       long long gx, long long gw, long long gb, long long gy, int relu, void* stream);

/* GroupNorm(groups, c) over position-major z (f, s, c), in place: statistics over the
 * (c/groups) x s elements of each (sample, group), biased variance (model.py:66). */
int tbn_groupnorm(float* z, const float* gamma, const float* beta, int f, int s, int c, int groups,
                  float eps, void* stream);

/* Train-mode BatchNorm2d + ReLU, in place on a channel slice of an NHWC fp32 activation (m = n*h*w rows): normalises with
 * the BATCH statistics and updates running_mean / running_var (momentum, unbiased variance) like torch's
 * nn.BatchNorm2d in training mode -- the training forward of the reference (core/tools/train.py:68-76 calls
 * model.train(); partial-BN only freezes the affine parameters' gradients, model.py:164-176).  workspace: 2*c doubles.
 * running_mean / running_var may both be NULL (statistics not tracked). */
int tbn_bn_train_relu(float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      void* workspace, float momentum, float eps, long long m, int c, int cs, int co, int relu,
                      void* stream);

/* Fused single-query multi-head attention core (attention.py:48-57 -> torch MultiheadAttention):
 *   scores[f,h,p] = qk[f,h,:] . z[f,p,:]          (qk = W_k,h^T q_h / sqrt(d), folded by the caller)
 *   a[f,h,:]      = softmax_p(scores[f,h,:])
 *   wts[f,p]      = mean_h a[f,h,p]                (returned attention weights, (F,1,S))
 *   entropy[f]    = H((wts+1e-6)/sum)              (core/models/model.py:310, 323-324)
 *   zbar[f,h,:]   = sum_p a[f,h,p] z[f,p,:]        (value projection is applied afterwards)
 * z tiles are staged in shared memory; reductions use warp shuffles. */
int tbn_mha_core(const float* qk, const float* z, float* zbar, float* wts, float* entropy,
                 int f, int s, int e, int heads, void* stream);

/* The same with pe[2] = GroupNorm(groups, e) (model.py:66) applied to z[f] in shared memory first (no separate pass over
 * z in global memory; bit-identical to tbn_groupnorm + tbn_mha_core), and optionally the weighted value sums written as
 * the bf16 [hi | hi | lo] split operand (f, heads * 3e) of the out-projection tensor-core GEMM (zbar3; zbar may then be
 * NULL). */
int tbn_mha_core_gn(const float* qk, const float* z, const float* gamma, const float* beta, int groups, float eps,
                    float* zbar, void* zbar3, float* wts, float* entropy, int f, int s, int e, int heads, void* stream);

/* Training variant: mask (f, heads, s) = 0 or 1 / (1 - p) multiplies the attention probabilities after the softmax, the
 * dropout torch's MultiheadAttention applies in training (attention.py:48-57 with attn_dropout; the returned weights
 * are then the head mean of the dropped probabilities).  mask may be NULL.  psum (f, heads), optional: the sum of the
 * kept probabilities of every (frame, head) -- the weight the value bias gets, o_h = sum_s p_s (W_v z_s + b_v). */
int tbn_mha_core_train(const float* qk, const float* z, const float* mask, float* zbar, float* wts, float* entropy,
                       float* psum, int f, int s, int e, int heads, void* stream);

/* ---- backward pass of the heads (training with frozen backbones: reference model.py:160-163 freeze_mode "all";
 * core/tools/train.py:68-94 loss.backward()).  What torch autograd derives for model.py:214-262 -- Classifier
 * (model.py:365-386), Fusion (337-362), nn.MultiheadAttention (attention.py:48-57), pe = Conv1d + GroupNorm
 * (model.py:62-67), the MLP attentions (attention.py:60-145) -- as explicit fp32 kernels. ------------------------- */

/* c (m x n) = alpha * op(a) (m x k) * op(b) (k x n) + beta * c, row-major with leading dimensions; ta: a is stored
 * (k x m), tb: b is stored (n x k).  dX = dY W is (ta 0, tb 0), dW = dY^T X is (ta 1, tb 0). */
int tbn_sgemm(int ta, int tb, int m, int n, int k, float alpha, const float* a, long long lda, const float* b,
              long long ldb, float beta, float* c, long long ldc, void* stream);
/* out[j] = beta * out[j] + sum_i x[i][j]: bias gradients (fixed summation order). */
int tbn_colsum(const float* x, long long ld, int m, int n, float beta, float* out, void* stream);
/* Segment consensus backward (model.py:194-199): dlogits[b * nseg + s] = dscores[b] / nseg. */
int tbn_consensus_bwd(const float* dscores, long long ld, float* dlogits, int b, int nseg, int c, void* stream);
/* ReLU (+ the inverted dropout that follows it, model.py:354-360) backward from the layer output h, in place on dh:
 * dh = h > 0 ? dh * scale : 0 with scale = 1 / (1 - p). */
int tbn_relu_dropout_bwd(float* dh, const float* h, float scale, long long n, void* stream);
/* o[f][h*d + j] += bv[h*d + j] * (psum[f][h] - 1): value-bias weight under attention dropout (see tbn_mha_core_train). */
int tbn_value_bias_fix(float* o, const float* bv, const float* psum, int f, int e, int heads, void* stream);
/* GroupNorm(groups, c) backward over (f, s) rows: x pre-norm input, dy output gradient -> dx; per-frame partial
 * dgamma / dbeta tables (f, c) for tbn_colsum. */
int tbn_groupnorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, float* dgamma_f, float* dbeta_f,
                      int f, int s, int c, int groups, float eps, void* stream);
/* Single-query multi-head attention backward: q (f, e), k, v (f*s, e), mask (f, heads, s) or NULL, d_o (f, e) gradient
 * of the concatenated head outputs, d_w (f, s) gradient of the head-averaged weights or NULL -> dq, dk, dv. */
int tbn_mha_bwd(const float* q, const float* k, const float* v, const float* mask, const float* d_o, const float* d_w,
                float* dq, float* dk, float* dv, int f, int s, int e, int heads, void* stream);
/* feature[f] = sum_s a[f][s] w[f][s] backward w.r.t. the weights: dw = dw_in (or 0) + a . dy. */
int tbn_weighted_sum_bwd(const float* a, const float* dy, long long ld_dy, const float* dw_in, float* dw, int f, int s,
                         int c, void* stream);
/* Row softmax backward: dx = y * (dy - sum(y * dy)). */
int tbn_softmax_bwd(const float* y, const float* dy, float* dx, int rows, int cols, long long ld, void* stream);
/* da[(f, s)][c] = wts[f][s] * dy[f][c]: the weighted-sum attention backward w.r.t. the attended positions. */
int tbn_weighted_sum_bwd_a(const float* wts, const float* dy, long long ld_dy, float* da, int f, int s, int c,
                           void* stream);

/* ---- backward pass of the backbones (BNInception.features under model.train(): conv + BatchNorm(batch statistics) +
 * ReLU triples of bn_inception_audio.py:58-401, the pools, the feature pooling of bn_inception.py:16-35).  fp32. ---- */

/* Training BatchNorm + ReLU, out of place: u (m, c) contiguous conv output is kept for the backward; y = relu(bn(u)) goes
 * into a channel slice; batch mean / rstd are saved; running statistics are updated (momentum, unbiased variance). */
int tbn_bn_train_fwd(const float* u, float* y, int y_cs, int y_co, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float* save_mean, float* save_rstd, void* workspace,
                     float momentum, float eps, long long m, int c, void* stream);
/* Its backward: g = dy * [y > 0]; du = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)); dgamma = sum g * xhat,
 * dbeta = sum g (either may be NULL).  workspace: 2c doubles. */
int tbn_bn_train_bwd(const float* u, const float* dy, int dy_cs, int dy_co, const float* y, int y_cs, int y_co,
                     const float* gamma, const float* save_mean, const float* save_rstd, float* du, float* dgamma,
                     float* dbeta, void* workspace, long long m, int c, void* stream);
/* fp32 im2col: cols[(n, oy, ox)][tap * cin + ci] (zero padding), the A operand of the weight-gradient GEMM. */
int tbn_im2col_f32(const float* x, float* cols, int n, int h, int w, int cin, int x_cs, int x_co, int k, int stride,
                   int pad, void* stream);
/* c = alpha * op(a) * op(b) with the k range split over ksplit CTAs (atomic accumulation; c is zeroed first): weight
 * gradients reduce over all output positions into a small (k*k*cin, cout) result. */
int tbn_sgemm_splitk(int ta, int tb, int m, int n, long long k, float alpha, const float* a, long long lda,
                     const float* b, long long ldb, float* c, long long ldc, int ksplit, void* stream);
/* y (n, H, W, c) = x (n, oh, ow, c) placed at even rows / columns, zeros elsewhere (stride-2 conv dgrad). */
int tbn_upsample_zero2(const float* x, float* y, int n, int oh, int ow, int H, int W, int c, void* stream);
/* MaxPool2d(3) backward: stride2 != 0 -> (stride 2, ceil_mode, no padding), else (stride 1, pad 1).  dx[argmax] += dy
 * (atomic; dx must be initialised), first maximum in scan order as torch. */
int tbn_maxpool3_bwd(const float* x, int x_cs, int x_co, const float* dy, int dy_cs, int dy_co, float* dx, int dx_cs,
                     int dx_co, int n, int h, int w, int c, int stride2, void* stream);
/* dst[r][d_co + j] += src[r][s_co + j] over m rows, c channels (gradient accumulation into a concat slice). */
int tbn_add_view(float* dst, int d_cs, int d_co, const float* src, int s_cs, int s_co, long long m, int c, void* stream);
/* dst (m, c) bf16 contiguous = the fp32 channel slice src[r][s_co + j]: operand copy for the tensor-core convs of the
 * mixed-precision training path (bf16 products, fp32 accumulation; everything stored stays fp32). */
int tbn_cast_bf16_view(const float* src, int s_cs, int s_co, void* dst, long long m, int c, void* stream);
/* Transposed bf16 im2col: dst[tap * cin + ci][p], p = (n, oy, ox) < n*oh*ow, rows of mpad elements (zero beyond the last
 * pixel) -- the reduction-major operands of the tensor-core weight-gradient GEMM (tbn_gemm_tc_splitk with the output
 * positions as its k axis).  k = 1, stride 1: transpose-and-cast of an fp32 slice (the output gradient). */
int tbn_im2col_t_bf16(const float* x, int x_cs, int x_co, void* dst, long long mpad, int n, int h, int w, int cin, int k,
                      int stride, int pad, void* stream);
/* Feature pooling backward (bn_inception.py:16-35): broadcast of dfeat / (h w), or per column / h with freq_only. */
int tbn_avgpool_feat_bwd(const float* dfeat, long long ld, float* dx, int n, int h, int w, int c, int freq_only,
                         void* stream);

/* 10-crop path of the reference forward (core/models/model.py:243-248): the audio feature rows [0, rows) of the fusion
 * matrix x (k*rows x cols fp32 at leading dimension ld, already offset to the audio columns) are tiled k times,
 * `feature.repeat(k, 1)`, when the visual modalities carry k crops per segment. */
int tbn_repeat_rows(float* x, int rows, int cols, int ld, int k, void* stream);

/* Fixed / unimodal / prototype attention: y[f*y_ld + c] = sum_p a[f,p,c] * wts[f,p]
 * (core/models/model.py:224-228; attention.py:88-90, 141-144). */
int tbn_attn_weighted_sum(const float* a, const float* wts, float* y, int f, int s, int c, int y_ld,
                          void* stream);

/* Row softmax (unimodal / prototype attention logits, attention.py:85, 138). */
int tbn_softmax_rows(const float* x, float* y, int rows, int cols, void* stream);

/* Segment consensus: y[b][c] = mean_n x[(b*nseg + n)*c_total + c]  (model.py:194-199). */
int tbn_consensus(const float* x, float* y, int b, int nseg, int c, void* stream);

/* Evaluation metrics of the score consumer (core/utils/metric.py:132-157, Metric._get_correct_score): per clip the
 * `maxk` best classes (descending score, ties to the lower class index), correct[j*b + i] = (j-th best class of
 * clip i == target[i]) as uint8, preds[j*b + i] = that class (may be NULL), rank[i] = position of target[i] among
 * those maxk classes or maxk if it is not one of them (int32, may be NULL), and conf_mat[target[i]*c + best[i]] += 1
 * (fp32 (c,c), may be NULL).  scores (b,c) fp32, target (b) int64.  One launch, no host round trip. */
int tbn_topk_confusion(const float* scores, const long long* target, int b, int c, int maxk, unsigned char* correct,
                       long long* preds, int* rank, float* conf_mat, void* stream);

/* Audio front-end (core/dataset/dataset.py:459-497, _get_spectrogram: librosa.stft(n_fft, hop_length, win_length,
 * window="hann", center=True, pad_mode="constant") -> log(|X|^2 + eps)).  Three steps:
 *   tbn_stft_frames:   samples (n, seg_len) fp32 (row stride in elements) -> a3 (n*frames, 3*kp) bf16, kp = win rounded
 *                      up to 64, frames = 1 + (seg_len + 2*(n_fft/2) - n_fft) / hop; row = the `win` samples under the
 *                      centred window of one frame (zeros outside the segment), split as [hi | hi | lo];
 *   tbn_conv_tc(k=1):  g (n*frames, 2*bins) fp32 = a3 x dft3^T, dft3 (2*bins, 3*kp) bf16 = [hi | lo | hi] of
 *                      window*cos (rows 0..bins-1) and -window*sin (rows bins..2*bins-1);
 *   tbn_stft_logpower: out (n, bins, frames) fp32 = log(Re^2 + Im^2 + eps)  (the (n,1,bins,frames) NCHW model input). */
int tbn_stft_frames(const float* samples, long long row_stride, int n, int seg_len, int n_fft, int win, int hop, void* a3,
                    void* stream);
int tbn_stft_logpower(const float* g, float* out, int n, int frames, int bins, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TBN_SM100_H */
