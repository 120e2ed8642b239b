/*
 * aurora_b200.h — C ABI of libaurora_b200.so: the sm_100a kernels behind Aurora's forward pass.
 *
 * The reference (microsoft/aurora) is pure Python/PyTorch and has no FFI; its seam is the Python
 * module boundary (Aurora.forward, aurora/model/aurora.py:265).  Every entry point below replaces a
 * group of ATen/cuBLAS/cuDNN/SDPA calls made by one reference function, cited as file:line relative
 * to the reference checkout.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every pointer is a DEVICE pointer owned by the caller (PyTorch allocates everything); the
 *     library never allocates device memory and keeps no pointer after a call returns.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - Return value: 0 (AB_OK) or a negative AbStatus; ab_last_error() gives a thread-local message.
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16); "f32" is IEEE float.
 *   - Row-major everywhere; `ld*` are leading dimensions in ELEMENTS.
 */
#ifndef AURORA_B200_H_
#define AURORA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB_ABI_VERSION 1

typedef enum AbStatus {
  AB_OK = 0,
  AB_ERR_INVALID_ARGUMENT = -1,
  AB_ERR_CUDA = -2,
  AB_ERR_UNSUPPORTED = -3
} AbStatus;

/* ABI version of the loaded library (AB_ABI_VERSION it was built with). */
int ab_version(void);

/* Thread-local description of the last failing call ("" if none). */
const char* ab_last_error(void);

/* Number of kernels this library has launched since load (all threads); bench.py reports the delta
 * over the timed region as `gpu_launches`. */
unsigned long long ab_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense projections — replaces every nn.Linear on the path:
 *   swin3d.py:153 (qkv), :169 (proj), :61-64 (MLP fc1 + GELU, fc2), :554 (PatchMerging reduction),
 *   :609,:612 (PatchSplitting lin1/lin2), perceiver.py:141-152 (to_q/to_kv/to_out), :79-84 (MLP),
 *   patchembed.py:112 (conv3d with kernel==stride, i.e. a GEMM), decoder.py:214,250 (heads).
 *
 *   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] ) + residual[m, n]
 *
 * A: bf16 [M, K] (lda), W: bf16 [N, K] (ldw; nn.Linear layout), fp32 accumulation in TMEM
 * (tcgen05.mma kind::f16, TMA-staged 128B-swizzled operand tiles).  bias (f32 [N]) and residual
 * (f32 [M, ldr]) are optional (NULL).  act: AB_ACT_NONE or AB_ACT_GELU_ERF (exact erf GELU, as
 * nn.GELU()).  Either or both outputs may be requested: out_f32 [M, ld_f32], out_bf16 [M, ld_bf16].
 * Requirements: K % 8 == 0, lda % 8 == 0, ldw % 8 == 0, A/W 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
enum { AB_ACT_NONE = 0, AB_ACT_GELU_ERF = 1 };

typedef struct AbGemm {
  const void* a;       /* bf16 [M, K] */
  const void* w;       /* bf16 [N, K] */
  const float* bias;   /* f32 [N] or NULL */
  const float* residual; /* f32 [M, ldr] or NULL */
  float* out_f32;      /* f32 [M, ld_f32] or NULL */
  void* out_bf16;      /* bf16 [M, ld_bf16] or NULL */
  int32_t m, n, k;
  int32_t lda, ldw, ldr, ld_f32, ld_bf16;
  int32_t act;
} AbGemm;

int ab_gemm_bf16(const AbGemm* g, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* AURORA_B200_H_ */
