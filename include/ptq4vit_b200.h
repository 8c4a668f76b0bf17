/* ptq4vit_b200 -- C ABI of the B200-native PTQ4ViT scale-factor search.
 *
 * Drop-in boundary: the reference has no FFI; its operator
 * surface for this path is the Python classes in quant_layers/{linear,matmul}.py.
 * A maintainer binds these entry points with ctypes from those classes (see
 * INTEGRATION.md); each function names the reference method it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated, fp32, row-major,
 * owned by the caller; nothing is allocated or freed by the library; work is
 * enqueued on `stream` (a cudaStream_t passed as void*) and the call returns
 * without synchronising.  Return value: 0 = ok, non-zero = error (message via
 * p4v_last_error()).  No exceptions cross the boundary.
 */
#ifndef PTQ4VIT_B200_H
#define PTQ4VIT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define P4V_API __attribute__((visibility("default")))
#else
#define P4V_API
#endif

#define P4V_OPERAND_AUTO 0
#define P4V_OPERAND_INT8 1 /* tcgen05.mma kind::i8, s32 accumulators            */
#define P4V_OPERAND_BF16 2 /* integer-valued bf16, kind::f16, exact f32 accum.   */
#define P4V_KERNEL_TCGEN05 0
#define P4V_KERNEL_SIMT 1 /* plain-CUDA cross-check kernel (bring-up / odd shapes) */

/* One wrapped Linear: mirrors the constructor of PTQSLQuantLinear /
 * PTQSLBatchingQuantLinear / PostGeluPTQSLBatchingQuantLinear
 * (quant_layers/linear.py:98-122, :350-363, :562-574). */
typedef struct p4v_linear_desc {
  int32_t rows;      /* M = images * tokens (all leading dims of x flattened)            */
  int32_t tokens;    /* tokens per image: rows / images (the reference means over them)  */
  int32_t in_features, out_features;
  int32_t n_V, n_H, n_a;
  int32_t w_bit, a_bit;
  int32_t eq_n;
  int32_t search_round;
  double eq_alpha, eq_beta; /* python floats in the reference: keep double */
  int32_t post_gelu; /* 1: twin-uniform activation quantizer (linear.py:557-642)         */
  int32_t has_bias;
  int32_t operand;   /* P4V_OPERAND_*  */
  int32_t kernel;    /* P4V_KERNEL_*   */
  int32_t init_layerwise; /* 1: every block starts from the layer-wise min-max step size (linear.py:382-383, :393-394) */
} p4v_linear_desc;

/* bytes of device workspace p4v_linear_* needs for this layer */
P4V_API int p4v_linear_workspace_bytes(const p4v_linear_desc* d, size_t* bytes);

/* number of floats of the optional score log: one [eq_n x groups] table per search
 * step in the reference's call order (W steps: groups = n_V, X steps: groups = 1). */
P4V_API int p4v_linear_score_log_floats(const p4v_linear_desc* d, size_t* n);

/* Replaces PTQSLBatchingQuantLinear.calibration_step2() (linear.py:536-555):
 *   _initialize_intervals (:380-397 / :576-599), candidate tables (:544-545),
 *   search_round x { _search_best_w_interval (:455-495), _search_best_a_interval
 *   (:497-533 / :609-642) }.
 * in : x [rows,in], weight [out,in], bias [out] or NULL, raw_out [rows,out], raw_grad [rows,out]
 * out: w_interval [n_V*n_H] (reference shape n_V,1,n_H,1), a_interval [n_a] (n_a,1),
 *      score_log (NULL or p4v_linear_score_log_floats floats).                         */
P4V_API int p4v_linear_calibrate(const p4v_linear_desc* d, const float* x, const float* weight, const float* bias,
                         const float* raw_out, const float* raw_grad, void* workspace, size_t workspace_bytes,
                         float* w_interval, float* a_interval, float* score_log, void* stream);

/* Step-wise surface (same workspace must be passed; begin() must come first):
 *   p4v_linear_begin      ~ _initialize_intervals + candidate tables
 *   p4v_linear_search_w   ~ _search_best_w_interval  (column blocks [h_begin,h_end))
 *   p4v_linear_search_a   ~ _search_best_a_interval  (activation chunks [a_begin,a_end))
 *   p4v_linear_intervals  -> copies the current step sizes out                          */
P4V_API int p4v_linear_begin(const p4v_linear_desc* d, const float* x, const float* weight, const float* bias,
                     const float* raw_out, const float* raw_grad, void* workspace, size_t workspace_bytes, void* stream);
P4V_API int p4v_linear_search_w(const p4v_linear_desc* d, const float* bias, const float* raw_out, const float* raw_grad,
                        void* workspace, int32_t h_begin, int32_t h_end, float* score_log, void* stream);
P4V_API int p4v_linear_search_a(const p4v_linear_desc* d, const float* bias, const float* raw_out, const float* raw_grad,
                        void* workspace, int32_t a_begin, int32_t a_end, float* score_log, void* stream);
P4V_API int p4v_linear_intervals(const p4v_linear_desc* d, void* workspace, float* w_interval, float* a_interval, void* stream);

/* Replaces quant_forward (linear.py:62-67 with quant_weight_bias :152-162 and
 * quant_input :164-169 / :601-607): out = fq(x) fq(W)^T + bias on the tensor cores. */
P4V_API int p4v_linear_quant_forward_workspace_bytes(const p4v_linear_desc* d, size_t* bytes);
P4V_API int p4v_linear_quant_forward(const p4v_linear_desc* d, const float* x, const float* weight, const float* bias,
                             const float* w_interval, const float* a_interval, void* workspace, size_t workspace_bytes,
                             float* out, void* stream);

/* One wrapped MatMul (head-wise groups, n_V = n_H = 1 per operand, as forced by
 * PTQSLBatchingQuantMatMul._get_padding_parameters, matmul.py:411-417, and used by
 * configs/PTQ4ViT.py:36-48).  A [batch,heads,S1,S2] @ B [batch,heads,S2,S3]. */
typedef struct p4v_matmul_desc {
  int32_t batch, heads, S1, S2, S3;
  int32_t A_bit, B_bit;
  int32_t eq_n;
  int32_t search_round;
  double eq_alpha, eq_beta;
  int32_t sos;       /* 1: split-of-softmax twin-uniform A (matmul.py:578-644)  */
  int32_t operand;
  int32_t kernel;
  int32_t init_layerwise; /* 1: every head starts from the layer-wise min-max step size (matmul.py:430-432) */
} p4v_matmul_desc;

P4V_API int p4v_matmul_workspace_bytes(const p4v_matmul_desc* d, size_t* bytes);
P4V_API int p4v_matmul_score_log_floats(const p4v_matmul_desc* d, size_t* n);
/* Replaces PTQSLBatchingQuantMatMul.calibration_step2() (matmul.py:565-576) and
 * SoSPTQSLBatchingQuantMatMul.calibration_step2() (matmul.py:633-644).
 * out: A_interval [heads] (sos: A_interval[0] = split/(qmax-1)), B_interval [heads], split [1] (sos) */
P4V_API int p4v_matmul_calibrate(const p4v_matmul_desc* d, const float* A, const float* B, const float* raw_out,
                         const float* raw_grad, void* workspace, size_t workspace_bytes, float* A_interval,
                         float* B_interval, float* split, float* score_log, void* stream);
/* Replaces quant_forward (matmul.py:140-145; SoS quant_input_A :595-598). */
P4V_API int p4v_matmul_quant_forward_workspace_bytes(const p4v_matmul_desc* d, size_t* bytes);
P4V_API int p4v_matmul_quant_forward(const p4v_matmul_desc* d, const float* A, const float* B, const float* A_interval,
                             const float* B_interval, const float* split, void* workspace, size_t workspace_bytes,
                             float* out, void* stream);

/* The patch-embedding convolution: ChannelwiseBatchingQuantConv2d with a_bit >= 32 (quant_layers/conv.py:444-614, wired
 * by configs/PTQ4ViT.py:52-54): one weight step size per output channel, activations left in FP32.  The caller passes
 * the im2col matrix of the FP32 input (torch.nn.functional.unfold, [images, positions, K], K = in_channels*kh*kw in the
 * kernel's own order), the kernel as [out_channels, K] and raw_out / raw_grad as [images, out_channels, positions]. */
typedef struct p4v_conv_desc {
  int32_t images, out_channels, K, positions;
  int32_t w_bit;
  int32_t eq_n;
  double eq_alpha, eq_beta;
  int32_t has_bias;
  int32_t kernel;    /* P4V_KERNEL_TCGEN05 only */
} p4v_conv_desc;
P4V_API int p4v_conv_workspace_bytes(const p4v_conv_desc* d, size_t* bytes);
/* Replaces ChannelwiseBatchingQuantConv2d.calibration_step2() (conv.py:591-603): _initialize_intervals (:482-496) and
 * _search_best_w_interval (:526-557).  out: w_interval [out_channels] (reference shape oc,1,1,1), score_log NULL or
 * [eq_n][out_channels].  The search is the same in every round when the activations are not quantised: it runs once. */
P4V_API int p4v_conv_calibrate(const p4v_conv_desc* d, const float* cols, const float* weight, const float* bias,
                       const float* raw_out, const float* raw_grad, void* workspace, size_t workspace_bytes,
                       float* w_interval, float* score_log, void* stream);

/* Integer export of a calibrated module (utils/integer.py:8-129): src [rows, cols] fp32 -> dst one byte per element.
 * mode 0: int8 = clamp(rne(x / delta), -q, q-1) (quantize_int_weight :8-18, quantize_matmul_input :27-42, plain
 * activations :64-69); mode 1: the post-GELU twin uint8 layout (:51-62); mode 2: the split-of-softmax twin uint8 layout
 * (:78-87).  delta[((row / rows_per_block) % n_row_blocks) * n_col_blocks + col / cols_per_block] is the step size of an
 * element (rows_per_block = 0: one row block). */
P4V_API int p4v_export_quantized(const float* src, long long rows, long long cols, const float* delta, int rows_per_block,
                         int n_row_blocks, int cols_per_block, int n_col_blocks, int mode, int bit, float d_neg,
                         const float* split, void* dst, void* stream);

P4V_API const char* p4v_last_error(void);
P4V_API int p4v_version(void);
/* number of kernel launches issued by this library in this process so far (for bench.py's gpu_launches) */
P4V_API long long p4v_launch_count(void);
/* Optional live timing of the sweep kernel (bench.py's roofline): while enabled every sweep launch is
 * bracketed by CUDA events on its own stream.  p4v_profile_collect synchronises those events and returns
 * the summed device time (ms), the number of sweep launches and the tensor-core operations (2*MAC) they
 * executed, then clears the record. */
P4V_API int p4v_profile_enable(int on);
P4V_API int p4v_profile_collect(double* sweep_ms, long long* sweep_launches, double* executed_ops);
/* The same record split by launch kind: out[0..2] = device ms of the bf16 slab sweeps, the int8 slab sweeps and the Gram
 * GEMMs, out[3..5] = tensor-core operations they executed, out[6..8] = launches, out[9..11] = the longest single launch
 * (ms, operations, kind).  n must be >= 12. */
P4V_API int p4v_profile_collect_kinds(double* out, int n);
/* Device self-test of the quantiser's division shortcut: evaluates round(v / delta) for n pseudo-random (v, delta)
 * pairs (plus pairs placed on and next to rounding ties) both with IEEE division, as the reference does
 * (quant_layers/linear.py:99-103 `(x / interval).round_()`), and with the reciprocal-based sequence the operand
 * image kernels use; writes the number of disagreements (must be 0).  Diagnostic only: unlike every other entry
 * point it allocates 8 bytes of device memory for the duration of the call and synchronises the stream. */
P4V_API int p4v_selftest_rint_div(unsigned long long n, unsigned long long seed, unsigned long long* mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif
