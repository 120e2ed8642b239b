// Perceiver cross-attention over pressure levels, per horizontal location (tiny sequences: 3 x 13 in the
// encoder, 13 x 3 in the decoder), plus the small fp32 linear used for location-independent vectors.
//
//   ab_perceiver_attention : softmax(q k^T / sqrt(dh)) v per (location, head)      perceiver.py:139-152
//        The queries are location-independent parameters (encoder.py:185-186, decoder.py:226), so
//        to_q(latents) is computed once ([Lq, D] f32) instead of once per location.
//   ab_linear_small_f32    : y = act(x W^T + b) for a handful of rows (time MLP, adaLN modulation
//        GEMVs, level / lead-time / absolute-time embeddings)   film.py:27-28, swin3d.py:805-809,912-914
#include "common.h"
#include "ptx.cuh"

namespace ab {

struct PercArgs {
  const float* q;             // [Lq, D] f32
  const uint16_t* kv;         // [Lk * nloc, 2D]  row = ck * nloc + loc ; columns [k | v]  (bf16 or fp16)
  uint16_t* out;              // [Lq * nloc, D]   row = cq * nloc + loc
  long long nloc;
  int lq, lk, heads, dh, dim;
  int ld_kv, ld_out;
  float scale;
};

// One thread per (location, query, head); head fastest so that a group of `heads` threads reads / writes
// one full row contiguously.  kDH = head dim (32 or 64).
template <int kDH, int kHalf>
__global__ void __launch_bounds__(256) perceiver_attention_kernel(const PercArgs a) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = a.nloc * a.lq * a.heads;
  if (gid >= total) return;
  const int h = static_cast<int>(gid % a.heads);
  long long t = gid / a.heads;
  const int iq = static_cast<int>(t % a.lq);
  const long long loc = t / a.lq;
  float q[kDH];
  const float4* qp = reinterpret_cast<const float4*>(a.q + static_cast<long long>(iq) * a.dim + h * kDH);
#pragma unroll
  for (int i = 0; i < kDH / 4; ++i) {
    const float4 v = __ldg(qp + i);
    q[4 * i] = v.x * a.scale;
    q[4 * i + 1] = v.y * a.scale;
    q[4 * i + 2] = v.z * a.scale;
    q[4 * i + 3] = v.w * a.scale;
  }
  float acc[kDH];
#pragma unroll
  for (int i = 0; i < kDH; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int c = 0; c < a.lk; ++c) {
    const uint16_t* krow = a.kv + (static_cast<long long>(c) * a.nloc + loc) * a.ld_kv + h * kDH;
    const uint4* kp = reinterpret_cast<const uint4*>(krow);
    const uint4* vp = reinterpret_cast<const uint4*>(krow + a.dim);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kDH / 8; ++i) {
      float f[8];
      unpack16x8<kHalf>(__ldg(kp + i), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += q[8 * i + j] * f[j];
    }
    const float mn = fmaxf(m, s);
    const float alpha = __expf(m - mn);
    const float p = __expf(s - mn);
    l = l * alpha + p;
    m = mn;
#pragma unroll
    for (int i = 0; i < kDH / 8; ++i) {
      float f[8];
      unpack16x8<kHalf>(__ldg(vp + i), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[8 * i + j] = acc[8 * i + j] * alpha + p * f[j];
    }
  }
  const float inv = 1.f / l;
  uint4* op = reinterpret_cast<uint4*>(a.out + (static_cast<long long>(iq) * a.nloc + loc) * a.ld_out + h * kDH);
#pragma unroll
  for (int i = 0; i < kDH / 8; ++i) {
    uint4 u;
    u.x = pack16x2<kHalf>(acc[8 * i] * inv, acc[8 * i + 1] * inv);
    u.y = pack16x2<kHalf>(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
    u.z = pack16x2<kHalf>(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
    u.w = pack16x2<kHalf>(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
    op[i] = u;
  }
}

// y[r, n] = act(sum_k x[r, k] W[n, k] + b[n]) ; one warp per output element, lanes stride over K.
// act_in: apply SiLU to x on load (nn.Sequential(SiLU, Linear)); act_out: SiLU on the result.
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ y,
                                                           int rows, int n, int k, int silu_in, int silu_out) {
  const long long warp_id = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_id >= static_cast<long long>(rows) * n) return;
  const int col = static_cast<int>(warp_id % n);
  const int r = static_cast<int>(warp_id / n);
  const float* xr = x + static_cast<long long>(r) * k;
  const float* wr = w + static_cast<long long>(col) * k;
  float acc = 0.f;
  for (int i = lane; i < k; i += 32) {
    float xv = __ldg(xr + i);
    if (silu_in) xv = xv / (1.f + __expf(-xv));
    acc += xv * __ldg(wr + i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (b) acc += __ldg(b + col);
    if (silu_out) acc = acc / (1.f + __expf(-acc));
    y[static_cast<long long>(r) * n + col] = acc;
  }
}

}  // namespace ab

extern "C" int ab_perceiver_attention(const float* q, const void* kv_bf16, void* out_bf16, int64_t nloc, int32_t lq,
                                      int32_t lk, int32_t num_heads, int32_t head_dim, int32_t ld_kv,
                                      int32_t ld_out, int32_t dtype, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(q && kv_bf16 && out_bf16, "ab_perceiver_attention: null argument");
  AB_CHECK_ARG(head_dim == 32 || head_dim == 64, "ab_perceiver_attention: head_dim must be 32 or 64 (got %d)",
               head_dim);
  AB_CHECK_ARG(nloc > 0 && lq > 0 && lk > 0 && num_heads > 0, "ab_perceiver_attention: bad sizes");
  const int dim = num_heads * head_dim;
  AB_CHECK_ARG(ld_kv >= 2 * dim && ld_kv % 8 == 0 && ld_out >= dim && ld_out % 8 == 0,
               "ab_perceiver_attention: bad leading dimensions");
  PercArgs a;
  a.q = q;
  AB_CHECK_ARG(dtype == AB_DT_BF16 || dtype == AB_DT_F16, "ab_perceiver_attention: bad dtype");
  a.kv = reinterpret_cast<const uint16_t*>(kv_bf16);
  a.out = reinterpret_cast<uint16_t*>(out_bf16);
  a.nloc = nloc;
  a.lq = lq;
  a.lk = lk;
  a.heads = num_heads;
  a.dh = head_dim;
  a.dim = dim;
  a.ld_kv = ld_kv;
  a.ld_out = ld_out;
  a.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  const long long total = nloc * lq * num_heads;
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(total, 256));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 32) {
    if (dtype == AB_DT_F16) perceiver_attention_kernel<32, 1><<<grid, 256, 0, s>>>(a);
    else perceiver_attention_kernel<32, 0><<<grid, 256, 0, s>>>(a);
  } else {
    if (dtype == AB_DT_F16) perceiver_attention_kernel<64, 1><<<grid, 256, 0, s>>>(a);
    else perceiver_attention_kernel<64, 0><<<grid, 256, 0, s>>>(a);
  }
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_perceiver_attention");
  return AB_OK;
}

extern "C" int ab_linear_small_f32(const float* x, const float* w, const float* bias, float* y, int32_t rows,
                                   int32_t n, int32_t k, int32_t silu_in, int32_t silu_out, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(x && w && y, "ab_linear_small_f32: null argument");
  AB_CHECK_ARG(rows > 0 && n > 0 && k > 0, "ab_linear_small_f32: bad sizes");
  const long long warps = static_cast<long long>(rows) * n;
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(warps * 32, 256));
  linear_small_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, bias, y, rows, n, k, silu_in,
                                                                               silu_out);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_linear_small_f32");
  return AB_OK;
}
