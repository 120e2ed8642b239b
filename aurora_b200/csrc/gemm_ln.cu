// GEMM with the AdaptiveLayerNorm + residual of a Swin block fused into its epilogue:
//
//   out[m, :] = residual[m, :] + LN( A[m, :] · W^T + bias ) * scale + shift          (fp32 stream + 16-bit copy)
//
// i.e. `x = shortcut + norm1(attn.proj(.), c)` and `x = x + norm2(mlp.fc2(.), c)` of aurora/model/swin3d.py:507-508
// with film.py:48-49 (LN without affine, eps 1e-5; scale = scale_bias + scale(c), shift = shift(c), both precomputed
// per model time step).  Unfused, the projection writes y in 16 bit and a row kernel reads y + the fp32 stream back and
// writes the stream + its 16-bit copy: 14 B of HBM traffic per element; fused it is 10 B and one launch less.
//
// LayerNorm needs whole rows, so a cluster owns whole rows: N = D = 512 * P columns, P CTA PAIRS per cluster.
//   * every pair runs the cta_group::2 pipeline of gemm.cu on ITS 512 columns: M = 256 rows per cluster tile (128 per
//     CTA), two accumulators of 256 fp32 columns = all 512 TMEM columns of both SMs (no accumulator double buffering:
//     the epilogue is HBM-bound, the next tile's operands are prefetched under it);
//   * epilogue, one thread per row and accumulator half (warp w: TMEM lane quadrant w & 3, half (w - 4) >> 2):
//       pass A  sum(y)            -> mean      (partials meet in shared memory; P = 2: + one DSMEM exchange)
//       pass B  sum((y - mean)^2) -> rstd      (two-pass variance, as the row kernel it replaces)
//       pass C  residual + (y - mean) * rstd * scale + shift -> fp32 stream (may alias the residual) + 16-bit copy
//     three reads of the accumulator out of TMEM (16 TB/s aggregate) instead of one trip through HBM.
//   * P = 2 (D = 1024): CTA r exchanges its per-row partial sums with CTA r ^ 2 (same rows, other 512 columns) through
//     distributed shared memory: remote store + remote mbarrier arrive (release.cluster) / acquire.cluster wait.
#include "common.h"
#include "ptx.cuh"

namespace ab {
namespace gln {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kHalfN = 256;          // one accumulator / one MMA N
constexpr int kPairN = 2 * kHalfN;   // columns per CTA pair
constexpr int kLoadN = 128;          // W rows staged per CTA and accumulator half
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 128 + kNumEpiWarps * 32;
constexpr int kStageBytesA = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kSubB = kLoadN * kBlockK * 2;          // 16 KB
constexpr int kStageBytes = kStageBytesA + 2 * kSubB;  // 48 KB
constexpr int kStages = 4;
constexpr int kOffVec = kStages * kStageBytes;         // bias | scale | shift of this pair's 512 columns (f32)
constexpr int kVecBytes = 3 * kPairN * 4;
constexpr int kOffStat = kOffVec + kVecBytes;          // partA[2][128] partB[2][128] remA[128] remB[128] bias_sum[2]
constexpr int kStatBytes = (4 * 128 + 2 * 128 + 8) * 4;
constexpr int kOffBar = kOffStat + kStatBytes;
constexpr int kBarrierBytes = 256;
constexpr int kSmemBytes = kOffBar + kBarrierBytes + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory");

struct Args {
  const float* bias;
  const float* scale;
  const float* shift;
  const float* residual;
  float* out_f32;
  uint16_t* out_16;
  int m, n, k;
  int ldr, ld_f32, ld_16;
  int out_half;
  float eps;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// Remote (DSMEM) float store into CTA `cta` of this cluster at the same shared-memory offset as `p`.
__device__ __forceinline__ void st_remote_f32(float* p, uint32_t cta, float v) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.f32 [ra], %2;\n\t"
      "}\n" ::"r"(smem_u32(p)),
      "r"(cta), "f"(v)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (uint32_t spin = 1;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 0x3FFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
// tcgen05.commit arriving on the barrier at this offset in both CTAs of MY pair (cluster ranks pair_base, pair_base + 1)
__device__ __forceinline__ void umma_commit_pair_of(uint64_t* bar, uint32_t pair_base) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3u << pair_base))
      : "memory");
}

template <bool kHalfIn, int P>
__global__ void __cluster_dims__(2 * P, 1, 1) __launch_bounds__(kThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;  // per stage: two 16 KB sub-tiles (accumulator halves)
  float* s_bias = reinterpret_cast<float*>(smem + kOffVec);
  float* s_scale = s_bias + kPairN;
  float* s_shift = s_scale + kPairN;
  float* part_a = reinterpret_cast<float*>(smem + kOffStat);  // [2][128]
  float* part_b = part_a + 256;                               // [2][128]
  float* rem_a = part_b + 256;                                // [128]  partner CTA's row sums (P = 2)
  float* rem_b = rem_a + 128;
  float* bias_sum = rem_b + 128;                              // [2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;  // leader's
  uint64_t* stat_a_bar = tmem_empty_bar + 1;     // P = 2: the partner CTA's four half-0 warps arrive here
  uint64_t* stat_b_bar = stat_a_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_b_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1;            // which 512 columns
  const uint32_t pair_base = rank & ~1u;      // cluster rank of my pair's leader
  const bool leader = (rank & 1u) == 0;
  const int n_base = static_cast<int>(pair) * kPairN;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader: own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 2 * kNumEpiWarps);
    mbar_init(stat_a_bar, 4);
    mbar_init(stat_b_bar, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  // this pair's slice of the per-column vectors
  for (int i = threadIdx.x; i < kPairN; i += kThreads) {
    s_bias[i] = g.bias != nullptr ? __ldg(g.bias + n_base + i) : 0.f;
    s_scale[i] = g.scale != nullptr ? __ldg(g.scale + n_base + i) : 1.f;
    s_shift[i] = g.shift != nullptr ? __ldg(g.shift + n_base + i) : 0.f;
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 3) {  // sum of the bias over each accumulator half (pass A adds it once per row)
    for (int h = 0; h < 2; ++h) {
      float s = 0.f;
      for (int i = lane; i < kHalfN; i += 32) s += s_bias[h * kHalfN + i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) bias_sum[h] = s;
    }
  }
  __syncthreads();

  const int num_tiles = (g.m + 2 * kBlockM - 1) / (2 * kBlockM);
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;
  const int cluster_id = blockIdx.x / (2 * P);
  const int num_clusters = gridDim.x / (2 * P);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (every CTA): its 128 A rows, its 128 W rows of both accumulator halves =====
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = tile * (2 * kBlockM) + static_cast<int>(rank & 1u) * kBlockM;
        const int n0 = n_base + static_cast<int>(rank & 1u) * kLoadN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(&empty_bar[s], ((it / kStages) & 1u) ^ 1u);
          uint8_t* sb = smem_b + s * (2 * kSubB);
          tma_load_2d_pair(smem_a + s * kStageBytesA, &tmap_a, &full_bar[s], kb * kBlockK, m0);
          tma_load_2d_pair(sb, &tmap_w, &full_bar[s], kb * kBlockK, n0);
          tma_load_2d_pair(sb + kSubB, &tmap_w, &full_bar[s], kb * kBlockK, n0 + kHalfN);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * kStageBytes);
          else mbar_arrive_remote(&full_bar[s], pair_base);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (the leader CTA of every pair) =====
      constexpr uint32_t idesc = umma_idesc_f16kind_f32(2 * kBlockM, kHalfN, kHalfIn);
      uint32_t it = 0, tc = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tc) {
        mbar_wait(tmem_empty_bar, (tc & 1u) ^ 1u);  // both CTAs have drained the previous tile out of TMEM
        tc_fence_after_sync();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(&full_bar[s], (it / kStages) & 1u);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + s * kStageBytesA);
          const uint32_t b_addr = smem_u32(smem_b + s * (2 * kSubB));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16_ss_pair(tmem_base + h * kHalfN, umma_desc_k_sw128(a_addr + k * kUmmaK * 2),
                                umma_desc_k_sw128(b_addr + h * kSubB + k * kUmmaK * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair_of(&empty_bar[s], pair_base);
        }
        umma_commit_pair_of(tmem_full_bar, pair_base);
      }
      if (tc > 0) mbar_wait(tmem_empty_bar, (tc - 1) & 1u);  // last remote arrivals land before the CTA may exit
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = (row, accumulator half) =====
    const int q = warp & 3;            // TMEM lane quadrant
    const int ch = (warp - 4) >> 2;    // accumulator half: columns [ch * 256, +256) of this pair's 512
    const int r = q * 32 + lane;       // row inside this CTA's 128
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ch * kHalfN;
    const float inv_n = 1.f / static_cast<float>(g.n);
    const float* sb = s_bias + ch * kHalfN;
    const float* ssc = s_scale + ch * kHalfN;
    const float* ssh = s_shift + ch * kHalfN;
    uint32_t tc = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tc) {
      const int row = tile * (2 * kBlockM) + static_cast<int>(rank & 1u) * kBlockM + r;
      const bool live = row < g.m;
      mbar_wait(tmem_full_bar, tc & 1u);
      tc_fence_after_sync();
      // ---- pass A: mean ----
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 1
      for (int c = 0; c < kHalfN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          s0 += __uint_as_float(v[j]);
          s1 += __uint_as_float(v[j + 1]);
          s2 += __uint_as_float(v[j + 2]);
          s3 += __uint_as_float(v[j + 3]);
        }
      }
      part_a[ch * 128 + r] = (s0 + s1) + (s2 + s3) + bias_sum[ch];
      named_bar_sync(1 + q, 64);  // the two warps that share these 32 rows
      float total = part_a[r] + part_a[128 + r];
      if constexpr (P > 1) {
        if (ch == 0) {
          st_remote_f32(&rem_a[r], rank ^ 2u, total);
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_release(stat_a_bar, rank ^ 2u);
        }
        mbar_wait_acquire_cluster(stat_a_bar, tc & 1u);
        total += rem_a[r];
      }
      const float mean = total * inv_n;
      // ---- pass B: variance around the mean ----
      s0 = s1 = s2 = s3 = 0.f;
#pragma unroll 1
      for (int c = 0; c < kHalfN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + c + j);
          const float d0 = __uint_as_float(v[j]) + (b4.x - mean), d1 = __uint_as_float(v[j + 1]) + (b4.y - mean);
          const float d2 = __uint_as_float(v[j + 2]) + (b4.z - mean), d3 = __uint_as_float(v[j + 3]) + (b4.w - mean);
          s0 = fmaf(d0, d0, s0);
          s1 = fmaf(d1, d1, s1);
          s2 = fmaf(d2, d2, s2);
          s3 = fmaf(d3, d3, s3);
        }
      }
      part_b[ch * 128 + r] = (s0 + s1) + (s2 + s3);
      named_bar_sync(1 + q, 64);
      total = part_b[r] + part_b[128 + r];
      if constexpr (P > 1) {
        if (ch == 0) {
          st_remote_f32(&rem_b[r], rank ^ 2u, total);
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_release(stat_b_bar, rank ^ 2u);
        }
        mbar_wait_acquire_cluster(stat_b_bar, tc & 1u);
        total += rem_b[r];
      }
      const float rstd = rsqrtf(total * inv_n + g.eps);
      // ---- pass C: normalise, modulate, add the residual, store the fp32 stream and its 16-bit copy ----
      const size_t col0 = static_cast<size_t>(n_base + ch * kHalfN);
      const float* res_row = g.residual != nullptr ? g.residual + static_cast<size_t>(row) * g.ldr + col0 : nullptr;
      float* o32 = g.out_f32 != nullptr ? g.out_f32 + static_cast<size_t>(row) * g.ld_f32 + col0 : nullptr;
      uint16_t* o16 = g.out_16 != nullptr ? g.out_16 + static_cast<size_t>(row) * g.ld_16 + col0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < kHalfN; c += 32) {
        float4 rr[8];
        if (live && res_row != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rr[j] = *reinterpret_cast<const float4*>(res_row + c + 4 * j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) rr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + c + 4 * j);
          const float4 sc4 = *reinterpret_cast<const float4*>(ssc + c + 4 * j);
          const float4 sh4 = *reinterpret_cast<const float4*>(ssh + c + 4 * j);
          f[4 * j + 0] = rr[j].x + fmaf((__uint_as_float(v[4 * j + 0]) + (b4.x - mean)) * rstd, sc4.x, sh4.x);
          f[4 * j + 1] = rr[j].y + fmaf((__uint_as_float(v[4 * j + 1]) + (b4.y - mean)) * rstd, sc4.y, sh4.y);
          f[4 * j + 2] = rr[j].z + fmaf((__uint_as_float(v[4 * j + 2]) + (b4.z - mean)) * rstd, sc4.z, sh4.z);
          f[4 * j + 3] = rr[j].w + fmaf((__uint_as_float(v[4 * j + 3]) + (b4.w - mean)) * rstd, sc4.w, sh4.w);
        }
        if (live) {
          if (o32 != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(o32 + c + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
          if (o16 != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 p;
              if (g.out_half) {
                p.x = pack_f16x2(f[8 * j + 0], f[8 * j + 1]);
                p.y = pack_f16x2(f[8 * j + 2], f[8 * j + 3]);
                p.z = pack_f16x2(f[8 * j + 4], f[8 * j + 5]);
                p.w = pack_f16x2(f[8 * j + 6], f[8 * j + 7]);
              } else {
                p.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
                p.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                p.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                p.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              }
              *reinterpret_cast<uint4*>(o16 + c + 8 * j) = p;
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tmem_empty_bar, pair_base);  // the pair leader's barrier collects both CTAs
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <bool kHalfIn, int P>
static int launch(const AbGemmLn* p, const Args& a, cudaStream_t stream) {
  CUtensorMap ta, tw;
  int rc = make_tmap_16bit_2d(&ta, p->a, p->m, p->k, p->lda, kBlockM, kBlockK, kHalfIn);
  if (rc != AB_OK) return rc;
  rc = make_tmap_16bit_2d(&tw, p->w, p->n, p->k, p->ldw, kLoadN, kBlockK, kHalfIn);
  if (rc != AB_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_kernel<kHalfIn, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) {
      set_error("ab_gemm_ln_residual: cudaFuncSetAttribute(smem=%d) failed: %s", kSmemBytes, cudaGetErrorString(e));
      return AB_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long tiles = ceil_div_ll(p->m, 2 * kBlockM);
  const long long max_clusters = sm_count() / (2 * P);
  const int clusters = static_cast<int>(tiles < max_clusters ? tiles : max_clusters);
  gemm_ln_kernel<kHalfIn, P><<<2 * P * clusters, kThreads, kSmemBytes, stream>>>(ta, tw, a);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_gemm_ln_residual");
  return AB_OK;
}

}  // namespace gln
}  // namespace ab

extern "C" int ab_gemm_ln_supported(int32_t n) { return (n == 512 || n == 1024) ? 1 : 0; }

extern "C" int ab_gemm_ln_residual(const AbGemmLn* p, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(p != nullptr && p->a != nullptr && p->w != nullptr, "ab_gemm_ln_residual: null operand");
  AB_CHECK_ARG(p->m > 0 && p->k > 0, "ab_gemm_ln_residual: bad shape m=%d k=%d", p->m, p->k);
  if (!ab_gemm_ln_supported(p->n)) {
    set_error("ab_gemm_ln_residual: N = %d is not supported (a cluster must own whole rows: N = 512 or 1024)", p->n);
    return AB_ERR_UNSUPPORTED;
  }
  AB_CHECK_ARG(p->k % 8 == 0 && p->lda % 8 == 0 && p->ldw % 8 == 0 && p->lda >= p->k && p->ldw >= p->k,
               "ab_gemm_ln_residual: K/lda/ldw must be multiples of 8 and ld >= K (k=%d lda=%d ldw=%d)", p->k, p->lda, p->ldw);
  AB_CHECK_ARG(p->out_f32 != nullptr || p->out_16 != nullptr, "ab_gemm_ln_residual: no output requested");
  AB_CHECK_ARG((p->in_dtype == AB_DT_BF16 || p->in_dtype == AB_DT_F16) && (p->out_dtype == AB_DT_BF16 || p->out_dtype == AB_DT_F16),
               "ab_gemm_ln_residual: dtypes must be AB_DT_BF16 or AB_DT_F16");
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  AB_CHECK_ARG(al16(p->residual) && al16(p->out_f32) && al16(p->out_16) &&
                   (p->residual == nullptr || (p->ldr % 4 == 0 && p->ldr >= p->n)) &&
                   (p->out_f32 == nullptr || (p->ld_f32 % 4 == 0 && p->ld_f32 >= p->n)) &&
                   (p->out_16 == nullptr || (p->ld_16 % 8 == 0 && p->ld_16 >= p->n)),
               "ab_gemm_ln_residual: residual / outputs must be 16-byte aligned with 16-byte-multiple pitches >= N");
  gln::Args a;
  a.bias = p->bias, a.scale = p->scale, a.shift = p->shift, a.residual = p->residual;
  a.out_f32 = p->out_f32, a.out_16 = reinterpret_cast<uint16_t*>(p->out_16);
  a.m = p->m, a.n = p->n, a.k = p->k, a.ldr = p->ldr, a.ld_f32 = p->ld_f32, a.ld_16 = p->ld_16;
  a.out_half = p->out_dtype == AB_DT_F16;
  a.eps = p->eps;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool half_in = p->in_dtype == AB_DT_F16;
  if (p->n == 512) return half_in ? gln::launch<true, 1>(p, a, s) : gln::launch<false, 1>(p, a, s);
  return half_in ? gln::launch<true, 2>(p, a, s) : gln::launch<false, 2>(p, a, s);
}
