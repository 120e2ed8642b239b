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

// One warp per (location, query): lane l owns features [l*E, (l+1)*E) of the D-wide row (E = D/32), i.e. a
// 1/lanes_per_head slice of one head; partial dot products are reduced over the lanes of a head with
// xor-shuffles, the softmax over the (<= 16) keys is online.  Every K / V / output row is therefore read or
// written by one warp as one contiguous, fully coalesced line.
template <int E, int kHalf>
__global__ void __launch_bounds__(256) perceiver_attention_kernel(const PercArgs a) {
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = a.nloc * a.lq;
  if (wid >= total) return;
  const int iq = static_cast<int>(wid % a.lq);  // queries of one location are adjacent: K/V rows stay in L1/L2
  const long long loc = wid / a.lq;
  const int lanes_per_head = a.dh / E;
  float q[E];
  {
    const float* qp = a.q + static_cast<long long>(iq) * a.dim + lane * E;
#pragma unroll
    for (int i = 0; i < E; i += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(qp + i));
      q[i] = v.x * a.scale;
      q[i + 1] = v.y * a.scale;
      q[i + 2] = v.z * a.scale;
      q[i + 3] = v.w * a.scale;
    }
  }
  float acc[E];
#pragma unroll
  for (int i = 0; i < E; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int c = 0; c < a.lk; ++c) {
    const uint16_t* krow = a.kv + (static_cast<long long>(c) * a.nloc + loc) * a.ld_kv + lane * E;
    float kf[E], vf[E];
    if constexpr (E >= 8) {
#pragma unroll
      for (int i = 0; i < E / 8; ++i) {
        float t[8];
        unpack16x8<kHalf>(__ldg(reinterpret_cast<const uint4*>(krow) + i), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) kf[8 * i + j] = t[j];
        unpack16x8<kHalf>(__ldg(reinterpret_cast<const uint4*>(krow + a.dim) + i), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) vf[8 * i + j] = t[j];
      }
    } else {
      const uint2 ku = __ldg(reinterpret_cast<const uint2*>(krow));
      const uint2 vu = __ldg(reinterpret_cast<const uint2*>(krow + a.dim));
      float2 t;
      t = unpack16x2<kHalf>(ku.x); kf[0] = t.x; kf[1] = t.y;
      t = unpack16x2<kHalf>(ku.y); kf[2] = t.x; kf[3] = t.y;
      t = unpack16x2<kHalf>(vu.x); vf[0] = t.x; vf[1] = t.y;
      t = unpack16x2<kHalf>(vu.y); vf[2] = t.x; vf[3] = t.y;
    }
    float sdot = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) sdot = fmaf(q[i], kf[i], sdot);
    for (int o = 1; o < lanes_per_head; o <<= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
    const float mn = fmaxf(m, sdot);
    const float alpha = __expf(m - mn);
    const float p = __expf(sdot - mn);
    l = l * alpha + p;
    m = mn;
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = fmaf(p, vf[i], acc[i] * alpha);
  }
  const float inv = 1.f / l;
  uint16_t* orow = a.out + (static_cast<long long>(iq) * a.nloc + loc) * a.ld_out + lane * E;
  if constexpr (E >= 8) {
#pragma unroll
    for (int i = 0; i < E / 8; ++i) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = acc[8 * i + j] * inv;
      reinterpret_cast<uint4*>(orow)[i] = pack16x8<kHalf>(t);
    }
  } else {
    uint2 u;
    u.x = pack16x2<kHalf>(acc[0] * inv, acc[1] * inv);
    u.y = pack16x2<kHalf>(acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<uint2*>(orow) = u;
  }
}

// Few keys, many queries (decoder: 13 level queries x 3 latent keys): one warp per LOCATION keeps the K / V rows
// of all (<= 4) keys in registers and loops over the queries, so every K / V row is read from memory once
// instead of once per query.
template <int E, int kHalf>
__global__ void __launch_bounds__(256) perceiver_attention_fewkeys_kernel(const PercArgs a) {
  constexpr int kMaxK = 4;
  // dim / E lanes (one or two warps) share a location; `lane` is the position inside that group
  const int lanes_per_loc = a.dim / E;
  const long long gl = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long loc = gl / lanes_per_loc;
  const int lane = static_cast<int>(gl % lanes_per_loc);
  if (loc >= a.nloc) return;
  const int lanes_per_head = a.dh / E;  // <= 32 and aligned, so a head never straddles a warp
  uint32_t kp[kMaxK][E / 2], vp[kMaxK][E / 2];  // packed 16-bit pairs
#pragma unroll
  for (int c = 0; c < kMaxK; ++c) {
    if (c < a.lk) {
      const uint16_t* krow = a.kv + (static_cast<long long>(c) * a.nloc + loc) * a.ld_kv + lane * E;
#pragma unroll
      for (int i = 0; i < E / 8; ++i) {
        const uint4 ku = __ldg(reinterpret_cast<const uint4*>(krow) + i);
        const uint4 vu = __ldg(reinterpret_cast<const uint4*>(krow + a.dim) + i);
        kp[c][4 * i] = ku.x; kp[c][4 * i + 1] = ku.y; kp[c][4 * i + 2] = ku.z; kp[c][4 * i + 3] = ku.w;
        vp[c][4 * i] = vu.x; vp[c][4 * i + 1] = vu.y; vp[c][4 * i + 2] = vu.z; vp[c][4 * i + 3] = vu.w;
      }
    }
  }
  for (int iq = 0; iq < a.lq; ++iq) {
    const float* qp = a.q + static_cast<long long>(iq) * a.dim + lane * E;
    float q[E];
#pragma unroll
    for (int i = 0; i < E; i += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(qp + i));
      q[i] = v.x * a.scale;
      q[i + 1] = v.y * a.scale;
      q[i + 2] = v.z * a.scale;
      q[i + 3] = v.w * a.scale;
    }
    float sc[kMaxK];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxK; ++c) {
      sc[c] = -INFINITY;
      if (c < a.lk) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < E / 2; ++i) {
          const float2 kk = unpack16x2<kHalf>(kp[c][i]);
          d = fmaf(q[2 * i], kk.x, d);
          d = fmaf(q[2 * i + 1], kk.y, d);
        }
        for (int o = 1; o < lanes_per_head; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        sc[c] = d;
        mx = fmaxf(mx, d);
      }
    }
    float l = 0.f;
    float acc[E];
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxK; ++c) {
      if (c < a.lk) {
        const float p = __expf(sc[c] - mx);
        l += p;
#pragma unroll
        for (int i = 0; i < E / 2; ++i) {
          const float2 vv = unpack16x2<kHalf>(vp[c][i]);
          acc[2 * i] = fmaf(p, vv.x, acc[2 * i]);
          acc[2 * i + 1] = fmaf(p, vv.y, acc[2 * i + 1]);
        }
      }
    }
    const float inv = 1.f / l;
    uint16_t* orow = a.out + (static_cast<long long>(iq) * a.nloc + loc) * a.ld_out + lane * E;
#pragma unroll
    for (int i = 0; i < E / 8; ++i) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = acc[8 * i + j] * inv;
      reinterpret_cast<uint4*>(orow)[i] = pack16x8<kHalf>(t);
    }
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
  const int e = dim / 32;
  AB_CHECK_ARG(dim % 32 == 0 && (e == 4 || e == 8 || e == 16 || e == 32) && head_dim % e == 0 &&
                   ((head_dim / e) & (head_dim / e - 1)) == 0,
               "ab_perceiver_attention: unsupported width (dim=%d head_dim=%d): need dim/32 in {4,8,16,32} dividing "
               "head_dim by a power of two", dim, head_dim);
  if (lk <= 4 && lq > lk && dim % 256 == 0) {
    // decoder shape: keep the few K / V rows in registers; dim / E lanes per location
    const int e_fk = dim % 512 == 0 ? 16 : 8;
    const unsigned grid_fk = static_cast<unsigned>(ceil_div_ll(nloc * (dim / e_fk), 256));
    cudaStream_t sfk = reinterpret_cast<cudaStream_t>(stream);
    const bool hfk = dtype == AB_DT_F16;
#define AB_PERC_FK(EE)                                                                 \
  do {                                                                                 \
    if (hfk) perceiver_attention_fewkeys_kernel<EE, 1><<<grid_fk, 256, 0, sfk>>>(a);   \
    else perceiver_attention_fewkeys_kernel<EE, 0><<<grid_fk, 256, 0, sfk>>>(a);       \
  } while (0)
    if (e_fk == 8) AB_PERC_FK(8);
    else AB_PERC_FK(16);
#undef AB_PERC_FK
    AB_COUNT_LAUNCH(1);
    AB_CHECK_LAUNCH("ab_perceiver_attention");
    return AB_OK;
  }
  const long long total = nloc * lq;  // warps
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(total * 32, 256));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool hf = dtype == AB_DT_F16;
#define AB_PERC(EE)                                                          \
  do {                                                                       \
    if (hf) perceiver_attention_kernel<EE, 1><<<grid, 256, 0, s>>>(a);       \
    else perceiver_attention_kernel<EE, 0><<<grid, 256, 0, s>>>(a);          \
  } while (0)
  if (e == 4) AB_PERC(4);
  else if (e == 8) AB_PERC(8);
  else if (e == 16) AB_PERC(16);
  else AB_PERC(32);
#undef AB_PERC
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
