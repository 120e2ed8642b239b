// Row-wise memory-bound kernels of the Swin backbone / Perceiver blocks (one warp per token row, 16-byte
// vector accesses, fp32 statistics):
//
//   ab_ln_mod_residual : out = residual + LN(y) * scale + shift (+ add_rows)      film.py:48-49 with
//                        swin3d.py:507-508; perceiver.py:225-232; encoder.py:320,346-363
//   ab_patch_merge_ln  : 2x2 gather (+ zero pad to even) -> LayerNorm(4D)          swin3d.py:526-553
//   ab_patch_split_ln  : pixel-shuffle 2x2 -> crop -> LayerNorm(D/2)               swin3d.py:574-611
//
// All are HBM-bound: algorithmic bytes per element are listed in DESIGN.md.
#include "common.h"
#include "ptx.cuh"

namespace ab {

constexpr int kRowWarps = 8;  // rows (warps) per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 float8_to_bf16(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

struct LnArgs {
  const uint16_t* y;       // [rows, ld_y] bf16 or fp16
  const float* scale;      // [D] or null (=1)
  const float* shift;      // [D] or null (=0)
  const float* residual;   // [*, ld_res] or null
  const float* add_rows;   // [add_mod, D] or null
  float* out_f32;
  uint16_t* out_bf16;      // bf16 or fp16
  long long rows;
  int dim;
  int ld_y, ld_res, ld_f32, ld_bf16;
  long long res_div, res_mod;  // residual row = (row / res_div) % res_mod  (res_mod == 0: residual row = row)
  long long add_mod;           // add row = row % add_mod
  float eps;
};

// kChunks: 16-byte (8 x bf16) chunks cached per lane; D <= 256 * kChunks.
template <int kChunks, int kHalfIn, int kHalfOut>
__global__ void __launch_bounds__(kRowWarps * 32) ln_mod_residual_kernel(const LnArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int nchunk = a.dim >> 3;
  const uint4* yrow = reinterpret_cast<const uint4*>(a.y + row * a.ld_y);
  float v[kChunks][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunk) {
      uint4 u = __ldg(yrow + ch);
      unpack16x8<kHalfIn>(u, v[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[c][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / a.dim;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    if (lane + c * 32 < nchunk) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[c][i] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.dim + a.eps);
  long long rrow = row;
  if (a.res_mod > 0) rrow = (row / a.res_div) % a.res_mod;
  const long long arow = a.add_rows ? row % a.add_mod : 0;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunk) {
      const int col = ch * 8;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd;
      if (a.scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.scale + col));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.scale + col + 4));
        o[0] *= s0.x; o[1] *= s0.y; o[2] *= s0.z; o[3] *= s0.w;
        o[4] *= s1.x; o[5] *= s1.y; o[6] *= s1.z; o[7] *= s1.w;
      }
      if (a.shift) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.shift + col));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shift + col + 4));
        o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
        o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
      }
      if (a.residual) {
        const float* r = a.residual + rrow * a.ld_res + col;
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(r));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(r + 4));
        o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
        o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
      }
      if (a.add_rows) {
        const float* r = a.add_rows + arow * a.dim + col;
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(r));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(r + 4));
        o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
        o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
      }
      if (a.out_f32) {
        float* p = a.out_f32 + row * a.ld_f32 + col;
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (a.out_bf16) *reinterpret_cast<uint4*>(a.out_bf16 + row * a.ld_bf16 + col) = pack16x8<kHalfOut>(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PatchMerging3D front half: gather 2x2 -> LayerNorm(4D) (affine) -> bf16 A operand of `reduction`.
// ---------------------------------------------------------------------------------------------
struct MergeArgs {
  const float* x;  // [B, C, H, W, D]
  const float* gamma;
  const float* beta;
  __nv_bfloat16* out;  // [B*C*H2*W2, 4D]
  int batch, c, h, w, d, h2, w2;
  float eps;
};

__global__ void __launch_bounds__(kRowWarps * 32) patch_merge_ln_kernel(const MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  const long long rows = static_cast<long long>(a.batch) * a.c * a.h2 * a.w2;
  if (row >= rows) return;
  const int wo = static_cast<int>(row % a.w2);
  long long t = row / a.w2;
  const int ho = static_cast<int>(t % a.h2);
  t /= a.h2;  // t = b * C + c
  const int d4 = a.d >> 2;   // float4 per source token
  const int n4 = a.d;        // float4 per output row (4D / 4)
  // pass 1: mean; pass 2: variance; pass 3: write (re-reads hit L1/L2).
  auto load4 = [&](int i4) -> float4 {
    const int q = i4 / d4;  // quadrant = hh * 2 + ww  (feature order (h w D), swin3d.py:540)
    const int hs = ho * 2 + (q >> 1), wsrc = wo * 2 + (q & 1);
    if (hs >= a.h || wsrc >= a.w) return make_float4(0.f, 0.f, 0.f, 0.f);  // bottom/right zero pad
    const float* p = a.x + ((t * a.h + hs) * a.w + wsrc) * static_cast<long long>(a.d);
    return __ldg(reinterpret_cast<const float4*>(p) + (i4 - q * d4));
  };
  float sum = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = load4(i);
    sum += v.x + v.y + v.z + v.w;
  }
  const float dim = 4.f * a.d;
  const float mean = warp_sum(sum) / dim;
  float sq = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = load4(i);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float rstd = rsqrtf(warp_sum(sq) / dim + a.eps);
  __nv_bfloat16* o = a.out + row * (4ll * a.d);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = load4(i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.beta) + i);
    uint2 pk;
    pk.x = pack_bf16x2((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y);
    pk.y = pack_bf16x2((v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w);
    *reinterpret_cast<uint2*>(o + i * 4) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// PatchSplitting3D middle: pixel shuffle of lin1's output, crop, LayerNorm(D/2) -> bf16 A of lin2.
// ---------------------------------------------------------------------------------------------
struct SplitArgs {
  const __nv_bfloat16* y;  // [B*C*H*W, 2D]
  const float* gamma;
  const float* beta;
  __nv_bfloat16* out;  // [B*C*Ho*Wo, D/2]
  int batch, c, h, w, d_half, ho, wo;
  float eps;
};

__global__ void __launch_bounds__(kRowWarps * 32) patch_split_ln_kernel(const SplitArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  const long long rows = static_cast<long long>(a.batch) * a.c * a.ho * a.wo;
  if (row >= rows) return;
  const int wq = static_cast<int>(row % a.wo);
  long long t = row / a.wo;
  const int hq = static_cast<int>(t % a.ho);
  t /= a.ho;  // b * C + c
  // The merge padding is bottom/right only (pad in {0,1} -> front = 0), so the crop keeps the origin.
  const long long src = (t * a.h + (hq >> 1)) * a.w + (wq >> 1);
  const int quad = (hq & 1) * 2 + (wq & 1);
  const uint4* p = reinterpret_cast<const uint4*>(a.y + src * (4ll * a.d_half) + static_cast<long long>(quad) * a.d_half);
  const int nchunk = a.d_half >> 3;
  float sum = 0.f;
  for (int i = lane; i < nchunk; i += 32) {
    float f[8];
    bf16x8_to_float(__ldg(p + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += f[k];
  }
  const float mean = warp_sum(sum) / a.d_half;
  float sq = 0.f;
  for (int i = lane; i < nchunk; i += 32) {
    float f[8];
    bf16x8_to_float(__ldg(p + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sq += (f[k] - mean) * (f[k] - mean);
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.d_half + a.eps);
  uint4* o = reinterpret_cast<uint4*>(a.out + row * static_cast<long long>(a.d_half));
  for (int i = lane; i < nchunk; i += 32) {
    float f[8];
    bf16x8_to_float(__ldg(p + i), f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + 2 * i);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + 2 * i + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta) + 2 * i);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta) + 2 * i + 1);
    float r[8];
    r[0] = (f[0] - mean) * rstd * g0.x + b0.x;
    r[1] = (f[1] - mean) * rstd * g0.y + b0.y;
    r[2] = (f[2] - mean) * rstd * g0.z + b0.z;
    r[3] = (f[3] - mean) * rstd * g0.w + b0.w;
    r[4] = (f[4] - mean) * rstd * g1.x + b1.x;
    r[5] = (f[5] - mean) * rstd * g1.y + b1.y;
    r[6] = (f[6] - mean) * rstd * g1.z + b1.z;
    r[7] = (f[7] - mean) * rstd * g1.w + b1.w;
    o[i] = float8_to_bf16(r);
  }
}

}  // namespace ab

extern "C" int ab_ln_mod_residual(const AbLnModResidual* p, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(p && p->y, "ab_ln_mod_residual: null input");
  AB_CHECK_ARG(p->rows > 0 && p->dim > 0 && p->dim % 8 == 0 && p->dim <= 2048,
               "ab_ln_mod_residual: dim must be a multiple of 8 and <= 2048 (got %d)", p->dim);
  AB_CHECK_ARG((p->in_dtype == AB_DT_BF16 || p->in_dtype == AB_DT_F16) &&
                   (p->out_dtype == AB_DT_BF16 || p->out_dtype == AB_DT_F16),
               "ab_ln_mod_residual: dtypes must be AB_DT_BF16 or AB_DT_F16");
  AB_CHECK_ARG(p->ld_y % 8 == 0 && (!p->residual || p->ld_res % 4 == 0) && (!p->out_f32 || p->ld_f32 % 4 == 0) &&
                   (!p->out_bf16 || p->ld_bf16 % 8 == 0),
               "ab_ln_mod_residual: leading dimensions must keep 16-byte alignment");
  AB_CHECK_ARG(p->out_f32 || p->out_bf16, "ab_ln_mod_residual: no output requested");
  AB_CHECK_ARG(!p->add_rows || p->add_mod > 0, "ab_ln_mod_residual: add_rows needs add_mod > 0");
  AB_CHECK_ARG(p->res_mod == 0 || p->res_div > 0, "ab_ln_mod_residual: res_mod needs res_div > 0");
  LnArgs a;
  a.y = reinterpret_cast<const uint16_t*>(p->y);
  a.scale = p->scale;
  a.shift = p->shift;
  a.residual = p->residual;
  a.add_rows = p->add_rows;
  a.out_f32 = p->out_f32;
  a.out_bf16 = reinterpret_cast<uint16_t*>(p->out_bf16);
  a.rows = p->rows;
  a.dim = p->dim;
  a.ld_y = p->ld_y;
  a.ld_res = p->ld_res;
  a.ld_f32 = p->ld_f32;
  a.ld_bf16 = p->ld_bf16;
  a.res_div = p->res_div;
  a.res_mod = p->res_mod;
  a.add_mod = p->add_mod;
  a.eps = p->eps;
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(p->rows, kRowWarps));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int hi = p->in_dtype == AB_DT_F16, ho = p->out_dtype == AB_DT_F16;
#define AB_LN_LAUNCH(CH)                                                                              \
  do {                                                                                                \
    if (hi && ho) ln_mod_residual_kernel<CH, 1, 1><<<grid, kRowWarps * 32, 0, s>>>(a);                \
    else if (hi) ln_mod_residual_kernel<CH, 1, 0><<<grid, kRowWarps * 32, 0, s>>>(a);                 \
    else if (ho) ln_mod_residual_kernel<CH, 0, 1><<<grid, kRowWarps * 32, 0, s>>>(a);                 \
    else ln_mod_residual_kernel<CH, 0, 0><<<grid, kRowWarps * 32, 0, s>>>(a);                         \
  } while (0)
  if (p->dim <= 256) AB_LN_LAUNCH(1);
  else if (p->dim <= 512) AB_LN_LAUNCH(2);
  else if (p->dim <= 1024) AB_LN_LAUNCH(4);
  else AB_LN_LAUNCH(8);
#undef AB_LN_LAUNCH
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_ln_mod_residual");
  return AB_OK;
}

extern "C" int ab_patch_merge_ln(const float* x, const float* gamma, const float* beta, void* out_bf16,
                                 int32_t batch, int32_t c, int32_t h, int32_t w, int32_t d, float eps,
                                 void* stream) {
  using namespace ab;
  AB_CHECK_ARG(x && gamma && beta && out_bf16, "ab_patch_merge_ln: null argument");
  AB_CHECK_ARG(batch > 0 && c > 0 && h > 1 && w > 1 && d > 0 && d % 8 == 0,
               "ab_patch_merge_ln: need H, W > 1 and D %% 8 == 0 (h=%d w=%d d=%d)", h, w, d);
  MergeArgs a;
  a.x = x;
  a.gamma = gamma;
  a.beta = beta;
  a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  a.batch = batch;
  a.c = c;
  a.h = h;
  a.w = w;
  a.d = d;
  a.h2 = (h + 1) / 2;
  a.w2 = (w + 1) / 2;
  a.eps = eps;
  const long long rows = static_cast<long long>(batch) * c * a.h2 * a.w2;
  patch_merge_ln_kernel<<<static_cast<unsigned>(ceil_div_ll(rows, kRowWarps)), kRowWarps * 32, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(a);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_patch_merge_ln");
  return AB_OK;
}

extern "C" int ab_patch_split_ln(const void* y_bf16, const float* gamma, const float* beta, void* out_bf16,
                                 int32_t batch, int32_t c, int32_t h, int32_t w, int32_t d, int32_t crop_h,
                                 int32_t crop_w, float eps, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(y_bf16 && gamma && beta && out_bf16, "ab_patch_split_ln: null argument");
  AB_CHECK_ARG(batch > 0 && c > 0 && h > 0 && w > 0 && d % 16 == 0, "ab_patch_split_ln: D %% 16 != 0 (d=%d)", d);
  AB_CHECK_ARG(crop_h >= 0 && crop_h <= 1 && crop_w >= 0 && crop_w <= 1, "ab_patch_split_ln: crop must be 0 or 1");
  SplitArgs a;
  a.y = reinterpret_cast<const __nv_bfloat16*>(y_bf16);
  a.gamma = gamma;
  a.beta = beta;
  a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  a.batch = batch;
  a.c = c;
  a.h = h;
  a.w = w;
  a.d_half = d / 2;
  a.ho = 2 * h - crop_h;
  a.wo = 2 * w - crop_w;
  a.eps = eps;
  const long long rows = static_cast<long long>(batch) * c * a.ho * a.wo;
  patch_split_ln_kernel<<<static_cast<unsigned>(ceil_div_ll(rows, kRowWarps)), kRowWarps * 32, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(a);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_patch_split_ln");
  return AB_OK;
}
