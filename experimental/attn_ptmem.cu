// EXPERIMENTAL — round-2 staging area, NOT part of libaurora_b200.so and not on any product path.
//
// Variants "x1" / "x1b" of the tcgen05 / TMEM window-attention kernel, written against the findings of
// profiles/r01p_ncu_final_captures.md (the softmax warps are the per-item critical path; 26 % of their time is the
// P -> shared-memory store + proxy fence, 16 % is waiting for gathered tiles, ~10 % is item re-decoding in the epilogue):
//
//   * P never goes through shared memory: each softmax thread writes its 72 packed bf16 pairs to TMEM with
//     tcgen05.st, and P.V is issued with the A operand in TMEM (`tcgen05.mma ... [d], [a_tmem], b_desc, ...`);
//     no STS, no generic->async proxy fence, 54 KB of shared memory freed;
//   * the freed shared memory holds a 4th q/k/v stage (more run-ahead for the TMA row gathers);
//   * S tile 1 (query rows 128..143) is re-issued after P V, so its P can live in its own S columns;
//   * (batch, head) of an item travel through the stage metadata: no integer divisions in the epilogue.
//
// Expected effect (to be measured): the softmax loop body loses its P-store phase (26 % of the samples) and the
// epilogue's divisions (~10 %); in exchange the tile-1 warp's chain now contains P V(n) and S1(n+1) (its S tile can
// only be re-issued after the P V that reads P1 out of the same columns), roughly +700 cycles per item on ITS chain.
// With ~5 400 cycles per item today that suggests ~4 200 (-20 %).  If the tile-1 chain turns out to be the pole, the
// next step is to keep P1 in shared memory (2 KB, SS-mode P V for tile 1 only) and issue S1(n+1) early again.
//
// Whole-grid, full 144-token windows only (no latitude slabs yet).  Built and compared with the shipped kernel by
// experimental/probe_attn_x1.py (needs a B200).  It has been compiled for sm_100a but NEVER RUN: round 1 had no
// GPU time left when it was written.
//
// The shipped translation unit is included for its device helpers (index arithmetic, item decoding, Meta layout,
// swizzle); its extern "C" entry points are renamed so that both kernels can live in one test library.
#define ab_window_geometry x1_shipped_window_geometry
#define ab_window_index_map_host x1_shipped_window_index_map_host
#define ab_window_index_map x1_shipped_window_index_map
#define ab_window_attention x1_shipped_window_attention
#include "../aurora_b200/csrc/window_attention.cu"

namespace ab {
namespace x1 {
using namespace tc;

// Two variants of the same kernel:
//   <4, false>  "x1" : P of BOTH tiles in TMEM; four q/k/v stages; S tile 1 re-issued after P V (its P overwrites it)
//   <3, true>   "x1b": P of tile 0 in TMEM, P of tile 1 (16 live rows, 6 KB) through shared memory as in the shipped
//                      kernel, so S tile 1 is issued early again; three stages (the 4th does not fit next to P1)
constexpr int kMaxStagesX = 4;
constexpr int kMetaBytesX = 6144;
constexpr int kP1AreaBytes = 3 * kP1BlockBytes;               // 6 KB
template <int kSt, bool kP1Smem>
struct Lay {
  static constexpr int kOffP1 = kSt * kStageBytes;            // only used when kP1Smem
  static constexpr int kOffMeta = kOffP1 + (kP1Smem ? kP1AreaBytes : 0);
  static constexpr int kSmem = kOffMeta + kMetaBytesX + 1024;
  static_assert(kSmem <= 227 * 1024, "shared memory");
};
// TMEM columns (all multiples of 16): S0 [0,144)  P0 [144,216)  O0 [224,288)  S1 [288,432) with P1 = [288,360)  O1 [432,496)
constexpr uint32_t kColS0 = 0, kColP0 = 144, kColO0 = 224, kColS1 = 288, kColP1 = kColS1, kColO1 = 432;

struct MetaX {
  int lsrc[kMaxStagesX][kTok];
  int src[kMaxStagesX][kTok];
  alignas(16) uint8_t grp[kMaxStagesX][kTok + 16];
  int masked[kMaxStagesX];
  int head[kMaxStagesX];
  int batch[kMaxStagesX];
  uint64_t full[kMaxStagesX], empty[kMaxStagesX], s0_full, s1_full, s0_free, p_full, o_full, o_free;
  uint32_t tmem_slot;
};
static_assert(sizeof(MetaX) <= kMetaBytesX, "meta area");

// D[tmem] (+)= A[tmem] * B[smem]^T : A = 128 lanes x 8 columns of packed 16-bit pairs (K = 16) per instruction.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 32 / 8 consecutive 32-bit columns (thread t writes lane t).
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int kStagesX, bool kP1Smem>
__global__ void __launch_bounds__(kThreads, 1)
window_attention_x1_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_halo,
                           const AttnArgs a) {
  // Same roles as the shipped kernel (warps 0-1 loaders, 2 MMA issuer, 3 softmax of rows 128..143, 4-7 softmax of
  // rows 0..127); differences are marked X1.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  MetaX* meta = reinterpret_cast<MetaX*>(smem + Lay<kStagesX, kP1Smem>::kOffMeta);
  const WinGeom& g = a.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ld = 3 * a.dim;
  const long long n_windows = a.slab ? static_cast<long long>(g.nwin[0]) * a.kh_count * g.nwin[2]
                                     : static_cast<long long>(a.batch) * g.nwindows;
  const long long n_items = n_windows * a.num_heads;
  const int cnt = static_cast<int>((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x);  // items of this CTA

  if (warp == 0 && lane == 0 && a.box_rows > 0) prefetch_tmap(&tmap_qkv);
  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kStagesX; ++i) {
      mbar_init(&meta->full[i], 1);
      mbar_init(&meta->empty[i], 1);
    }
    mbar_init(&meta->s0_full, 1);
    mbar_init(&meta->s1_full, 1);
    // X1: only the four tile-0 warps release S0 (S1 is re-issued after P V in pipe order); X1b: all five warps release
    // both S tiles, as in the shipped kernel
    mbar_init(&meta->s0_free, kP1Smem ? 5 : 4);
    mbar_init(&meta->p_full, 5);
    mbar_init(&meta->o_full, 1);
    mbar_init(&meta->o_free, 5);
    fence_mbar_init();
  }
  if (warp == 2) {
    __syncwarp();
    tmem_alloc<512>(&meta->tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = meta->tmem_slot;

  if (warp < 2) {
    // ===== loaders: warp w takes items n = w (mod 2); item n lives in ring stage n % kStagesX =====
    for (int n = warp; n < cnt; n += 2) {
      const int st = n % kStagesX;
      uint8_t* stage = smem + st * kStageBytes;
      const uint32_t sq = smem_u32(stage), sk = sq + kTileBytes, sv = sk + kTileBytes;
      mbar_wait(&meta->empty[st], ((((n / kStagesX) & 1)) ^ 1));
      const TcItem it = tc_decode(a, blockIdx.x + static_cast<long long>(n) * gridDim.x);
      int grp0 = 0;
      tc_source_token(g, it.k0, it.k1, it.k2, 0, &grp0);
      int differs = 0;
      for (int t = lane; t < kTok; t += 32) {
        int grp, lrow, srow;
        const int src = tc_source_token(g, it.k0, it.k1, it.k2, t, &grp);
        tc_translate(a, src, &lrow, &srow);
        differs |= (grp != grp0);
        meta->lsrc[st][t] = lrow;
        meta->src[st][t] = srow;
        meta->grp[st][t] = static_cast<uint8_t>(grp);
      }
      const int masked = (__any_sync(0xffffffffu, differs) && g.shifted) ? 1 : 0;
      if (lane == 0) {
        meta->masked[st] = masked;
        meta->head[st] = it.head;   // X1: the epilogue takes (batch, head) from here instead of re-decoding the item
        meta->batch[st] = it.b;
      }
      __syncwarp();
      const long long row_base = static_cast<long long>(it.b) * a.tokens_per_batch;
      if (a.box_rows > 0) {
        // TMA gather: every run of `box_rows` consecutive in-window tokens along W is either all padding or one
        // contiguous piece of the token stream (the host picked box_rows as the gcd of all run boundaries), so
        // it is ONE 2-D box {64 columns, box_rows rows} of the [tokens, 3D] qkv matrix per q / k / v.  The
        // 128B swizzle is applied by the TMA unit, completion is counted on the stage's mbarrier.
        const int r = a.box_rows;
        const int groups_per_row = g.ws[2] / r;
        const int units = g.ws[0] * g.ws[1] * groups_per_row;  // (ic, ih, group)
        uint32_t bytes = 0;
        for (int u = lane; u < units; u += 32) {
          const int t0 = (u / groups_per_row) * g.ws[2] + (u % groups_per_row) * r;  // first window token of the run
          const int src = meta->lsrc[st][t0];
          const uint32_t off = static_cast<uint32_t>(t0) * kRowBytes;
          if (src >= 0) {
            const bool from_halo = (src & kHaloFlag) != 0;
            const int grow = from_halo ? (src & ~kHaloFlag) : static_cast<int>(row_base + src);
            const CUtensorMap* tm = from_halo ? &tmap_halo : &tmap_qkv;
            const int col = it.head * kHeadDim;
            tma_load_2d(stage + off, tm, &meta->full[st], col, grow);
            tma_load_2d(stage + kTileBytes + off, tm, &meta->full[st], col + a.dim, grow);
            tma_load_2d(stage + 2 * kTileBytes + off, tm, &meta->full[st], col + 2 * a.dim, grow);
            bytes += 3u * r * kRowBytes;
          }
        }
        if (g.nwindows * kTok != g.res[0] * g.res[1] * g.res[2]) {
          // zero-padded tokens (x = 0): q | k | v are the projection bias; filled by the whole warp, 16 B per lane
          for (int idx = lane; idx < kTok * 8; idx += 32) {
            const int t = idx >> 3, chunk = idx & 7;
            if (meta->lsrc[st][t] < 0) {
              const uint4* pb = reinterpret_cast<const uint4*>(a.pad_qkv + it.head * kHeadDim + chunk * 8);
              const uint32_t o2 = swz(t, chunk);
              *reinterpret_cast<uint4*>(stage + o2) = __ldg(pb);
              *reinterpret_cast<uint4*>(stage + kTileBytes + o2) = __ldg(pb + a.dim / 8);
              *reinterpret_cast<uint4*>(stage + 2 * kTileBytes + o2) = __ldg(pb + 2 * a.dim / 8);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        fence_proxy_async_smem();  // bias fills (generic proxy) -> tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&meta->full[st], bytes);
        continue;
      }
      for (int idx = lane; idx < kTok * 8; idx += 32) {
        const int t = idx >> 3, chunk = idx & 7;
        const int src = meta->lsrc[st][t];
        const uint32_t off = swz(t, chunk);
        const __nv_bfloat16* p;
        if (src < 0) p = a.pad_qkv + it.head * kHeadDim + chunk * 8;  // zero-padded token: bias
        else if (src & kHaloFlag) p = a.halo_qkv + static_cast<long long>(src & ~kHaloFlag) * ld + it.head * kHeadDim + chunk * 8;
        else p = a.qkv + (row_base + src) * ld + it.head * kHeadDim + chunk * 8;
        cp_async_16(sq + off, p);
        cp_async_16(sk + off, p + a.dim);
        cp_async_16(sv + off, p + 2 * a.dim);
      }
      cp_async_wait_all();
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->full[st]);
    }
  } else if (warp == 2) {
    if (lane == 0 && cnt > 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_s = umma_idesc_f16kind_f32(128, kTok, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, kHeadDim);  // A (now in TMEM) K-major, B = V MN-major
      // X1: S tile 0 and S tile 1 are issued (and signalled) separately.  S0(n+1) goes out as soon as the tile-0
      // warps hold S0(n) in registers (tensor work under the softmax, as before); S1(n+1) is issued AFTER P V(n),
      // because P of tile 1 lives in the first 72 columns of S1 and tcgen05.mma executes in issue order.
      auto issue_s0 = [&](int n) {
        const uint32_t qa = smem_u32(smem + (n % kStagesX) * kStageBytes);
        const uint32_t ka = qa + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + kColS0, umma_desc_k_sw128(qa + k * 32), umma_desc_k_sw128(ka + k * 32), idesc_s, k != 0);
        umma_commit(&meta->s0_full);
      };
      auto issue_s1 = [&](int n) {
        const uint32_t qa = smem_u32(smem + (n % kStagesX) * kStageBytes);
        const uint32_t ka = qa + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + kColS1, umma_desc_k_sw128(qa + kTile1Row0 * kRowBytes + k * 32),
                       umma_desc_k_sw128(ka + k * 32), idesc_s, k != 0);
        umma_commit(&meta->s1_full);
      };
      mbar_wait(&meta->full[0], 0);
      tc_fence_after_sync();
      issue_s0(0);
      issue_s1(0);
      for (int n = 0; n < cnt; ++n) {
        if (n + 1 < cnt) {
          mbar_wait(&meta->full[(n + 1) % kStagesX], ((n + 1) / kStagesX) & 1);
          mbar_wait(&meta->s0_free, n & 1);  // the tile-0 softmax warps have pulled S0(n) out of TMEM
          tc_fence_after_sync();
          issue_s0(n + 1);
          if constexpr (kP1Smem) issue_s1(n + 1);  // X1b: P1 is in shared memory, S1 may be overwritten right away
        }
        mbar_wait(&meta->p_full, n & 1);     // P(n) is in TMEM (tcgen05.st + wait::st + fence on the writer side)
        if (n > 0) mbar_wait(&meta->o_free, (n - 1) & 1);
        tc_fence_after_sync();
        const uint32_t va = smem_u32(smem + (n % kStagesX) * kStageBytes) + 2 * kTileBytes;
        // X1b: tile-1 A operand from shared memory; it starts 96 rows before its 16 live rows (don't-care rows that
        // fall into the last stage's V tile: mapped memory, accumulator rows never read), as in the shipped kernel
        const uint32_t p1 = smem_u32(smem + Lay<kStagesX, kP1Smem>::kOffP1) - kP1Rewind;
#pragma unroll
        for (int j = 0; j < kTok / 16; ++j) {  // 9 k-steps of 16 keys = 8 TMEM columns of packed bf16 pairs each
          const uint64_t dv = umma_desc_k_sw128(va + j * 16 * kRowBytes);
          umma_bf16_ts(tmem_base + kColO0, tmem_base + kColP0 + j * 8, dv, idesc_o, j != 0);
          if constexpr (kP1Smem)
            umma_bf16_ss(tmem_base + kColO1, umma_desc_k_sw128(p1 + (j >> 2) * kP1BlockBytes + (j & 3) * 32), dv, idesc_o,
                         j != 0);
          else
            umma_bf16_ts(tmem_base + kColO1, tmem_base + kColP1 + j * 8, dv, idesc_o, j != 0);
        }
        umma_commit(&meta->o_full);
        umma_commit(&meta->empty[n % kStagesX]);  // q / k / v of this stage are consumed
        if constexpr (!kP1Smem) {
          if (n + 1 < cnt) issue_s1(n + 1);      // overwrites P1(n) only after P V(n) above has read it
        }
      }
    }
  } else {
    // ===== softmax + epilogue (warps 3..7): thread = query row =====
    const int tile = (warp == 3) ? 1 : 0;
    const int lrow = (warp & 3) * 32 + lane;           // row inside the tile (TMEM lane)
    const int row = tile * kTile1Row0 + lrow;          // window token
    const bool valid = row < kTok;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (tile ? kColS1 : kColS0);
    const uint32_t o_addr = tmem_base + lane_addr + (tile ? kColO1 : kColO0);
    // X1: P goes to TMEM (this thread's lane, 72 columns of packed bf16 pairs): tile 0 has its own columns, tile 1
    // reuses the first 72 columns of its S tile.
    const uint32_t p_addr = tmem_base + lane_addr + (tile ? kColP1 : kColP0);
    uint64_t* const my_s_full = tile ? &meta->s1_full : &meta->s0_full;
    constexpr float kC = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) in the exp2 domain

    int prev_src = -1;
    float prev_inv = 0.f;
    int prev_b = 0, prev_head = 0;
    auto epilogue = [&](int n_prev) {
      mbar_wait(&meta->o_full, n_prev & 1);
      tc_fence_after_sync();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(o_addr, o0);
      tmem_ld_32x32b_x32(o_addr + 32, o1);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->o_free);
      if (valid && prev_src >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(a.out + (static_cast<long long>(prev_b) * a.tokens_per_batch + prev_src) * a.dim +
                                              prev_head * kHeadDim);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o0[8 * c + 0]) * prev_inv, __uint_as_float(o0[8 * c + 1]) * prev_inv);
          u.y = pack_bf16x2(__uint_as_float(o0[8 * c + 2]) * prev_inv, __uint_as_float(o0[8 * c + 3]) * prev_inv);
          u.z = pack_bf16x2(__uint_as_float(o0[8 * c + 4]) * prev_inv, __uint_as_float(o0[8 * c + 5]) * prev_inv);
          u.w = pack_bf16x2(__uint_as_float(o0[8 * c + 6]) * prev_inv, __uint_as_float(o0[8 * c + 7]) * prev_inv);
          dst[c] = u;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o1[8 * c + 0]) * prev_inv, __uint_as_float(o1[8 * c + 1]) * prev_inv);
          u.y = pack_bf16x2(__uint_as_float(o1[8 * c + 2]) * prev_inv, __uint_as_float(o1[8 * c + 3]) * prev_inv);
          u.z = pack_bf16x2(__uint_as_float(o1[8 * c + 4]) * prev_inv, __uint_as_float(o1[8 * c + 5]) * prev_inv);
          u.w = pack_bf16x2(__uint_as_float(o1[8 * c + 6]) * prev_inv, __uint_as_float(o1[8 * c + 7]) * prev_inv);
          dst[4 + c] = u;
        }
      }
    };

    for (int n = 0; n < cnt; ++n) {
      const int st = n % kStagesX;
      mbar_wait(&meta->full[st], (n / kStagesX) & 1);  // index map / group ids of this item are in smem
      const int my_src = valid ? meta->src[st][row] : -1;
      const int my_grp = valid ? meta->grp[st][row] : 0;
      const int masked = meta->masked[st];
      const int cur_b = meta->batch[st], cur_head = meta->head[st];
      mbar_wait(my_s_full, n & 1);
      tc_fence_after_sync();
      float sv[kTok];
      {
        uint32_t t0[32], t1[32], t2[32], t3[32], t4[16];
        tmem_ld_32x32b_x32(s_addr, t0);
        tmem_ld_32x32b_x32(s_addr + 32, t1);
        tmem_ld_32x32b_x32(s_addr + 64, t2);
        tmem_ld_32x32b_x32(s_addr + 96, t3);
        tmem_ld_32x32b_x16(s_addr + 128, t4);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sv[j] = __uint_as_float(t0[j]);
          sv[32 + j] = __uint_as_float(t1[j]);
          sv[64 + j] = __uint_as_float(t2[j]);
          sv[96 + j] = __uint_as_float(t3[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sv[128 + j] = __uint_as_float(t4[j]);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0 && (tile == 0 || kP1Smem)) mbar_arrive(&meta->s0_free);  // S(n) is in registers: S(n+1) may overwrite it
      if (masked) {
        // 0 / -100 on the scaled logits == 0 / -800 on the raw q.k products (scale 1/8)
        const uint4* g16 = reinterpret_cast<const uint4*>(meta->grp[st]);
        const uint32_t mine = static_cast<uint32_t>(my_grp) * 0x01010101u;
#pragma unroll
        for (int w16 = 0; w16 < kTok / 16; ++w16) {
          const uint4 gv = g16[w16];  // broadcast load: sixteen keys' group ids
          const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const uint32_t ne = __vcmpne4(gw[w4], mine);  // 0xff per key of another group
            const int j = 16 * w16 + 4 * w4;
            if (ne & 0x000000ffu) sv[j] -= 800.f;
            if (ne & 0x0000ff00u) sv[j + 1] -= 800.f;
            if (ne & 0x00ff0000u) sv[j + 2] -= 800.f;
            if (ne & 0xff000000u) sv[j + 3] -= 800.f;
          }
        }
      }
      float mxa[4] = {sv[0], sv[1], sv[2], sv[3]};  // four independent chains: one warp per scheduler has no TLP
#pragma unroll
      for (int j = 4; j < kTok; ++j) mxa[j & 3] = fmaxf(mxa[j & 3], sv[j]);
      const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      const float nm = -mx * kC;
      float suma[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[kTok / 2];
#pragma unroll
      for (int j = 0; j < kTok / 2; ++j) {
        const float e0 = ex2_approx(fmaf(sv[2 * j], kC, nm));
        const float e1 = ex2_approx(fmaf(sv[2 * j + 1], kC, nm));
        suma[j & 3] += e0 + e1;
        pk[j] = pack_bf16x2(e0, e1);
      }
      const float sum = (suma[0] + suma[1]) + (suma[2] + suma[3]);
      // P(n) may only replace P(n-1) once P V(n-1) has retired; that is what o_full(n-1) says.  Doing the
      // previous item's epilogue here keeps the tensor pipe busy with P V(n-1) / S(n+1) under this softmax.
      if (n > 0) epilogue(n - 1);
      // X1: P(n) -> TMEM.  tcgen05.st is warp-collective: every lane stores (rows >= 144 hold don't-care values whose
      // accumulator rows are never read).  P0(n) may replace P0(n-1) because epilogue(n-1) above has waited for
      // o_full(n-1), i.e. P V(n-1) has retired; P1(n) goes over S1(n), which this thread has already pulled out.
      if (kP1Smem && tile) {
        // X1b, tile 1: 16 live rows to K-major swizzled shared memory (three 2 KB blocks of 64 keys), generic -> async
        // proxy fence, exactly the shipped path
        const uint32_t prow = smem_u32(smem + Lay<kStagesX, kP1Smem>::kOffP1) + (lrow - 96) * kRowBytes;
#pragma unroll
        for (int ch = 0; ch < kTok / 8; ++ch) {
          const uint32_t addr = prow + (ch >> 3) * kP1BlockBytes + (((ch & 7) ^ (lrow & 7)) << 4);
          if (valid)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * ch]), "r"(pk[4 * ch + 1]),
                         "r"(pk[4 * ch + 2]), "r"(pk[4 * ch + 3])
                         : "memory");
        }
        fence_proxy_async_smem();
      } else {
        uint32_t c0[32], c1[32], c2[8];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          c0[j] = pk[j];
          c1[j] = pk[32 + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) c2[j] = pk[64 + j];
        tmem_st_32x32b_x32(p_addr, c0);
        tmem_st_32x32b_x32(p_addr + 32, c1);
        tmem_st_32x32b_x8(p_addr + 64, c2);
        tmem_st_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->p_full);
      prev_src = my_src;
      prev_inv = 1.f / sum;
      prev_b = cur_b;
      prev_head = cur_head;
    }
    if (cnt > 0) epilogue(cnt - 1);
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}


}  // namespace x1
}  // namespace ab

// Same argument struct as ab_window_attention (include/aurora_b200.h); whole grid + full 144-token windows only.
extern "C" int ab_window_attention_x1(const AbWindowAttention* p, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(p != nullptr && p->qkv != nullptr && p->out != nullptr, "ab_window_attention_x1: null argument");
  AB_CHECK_ARG(p->head_dim == kHeadDim && p->num_heads > 0 && p->batch > 0, "ab_window_attention_x1: bad heads / batch");
  AB_CHECK_ARG(p->slab_h_rows == 0 && p->bias == nullptr, "ab_window_attention_x1: no slabs / dense bias in this variant");
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.g = make_win_geom(p->res, p->window, p->shift, p->warped);
  AB_CHECK_ARG(a.g.ntok == tc::kTok, "ab_window_attention_x1: full 144-token windows only");
  const bool has_pad = a.g.nwindows * a.g.ntok != p->res[0] * p->res[1] * p->res[2];
  AB_CHECK_ARG(!has_pad || p->pad_qkv != nullptr, "ab_window_attention_x1: pad_qkv is required for a zero-padded grid");
  a.kh_count = a.g.nwin[1];
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(p->qkv);
  a.pad_qkv = reinterpret_cast<const __nv_bfloat16*>(p->pad_qkv);
  a.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  a.batch = p->batch;
  a.num_heads = p->num_heads;
  a.dim = p->num_heads * kHeadDim;
  a.tokens_per_batch = static_cast<long long>(p->res[0]) * p->res[1] * p->res[2];
  // AB_X1_VARIANT=b selects the "x1b" kernel (P of tile 1 through shared memory, three stages)
  const char* var = getenv("AB_X1_VARIANT");
  const bool vb = var != nullptr && var[0] == 'b';
  auto kern = vb ? x1::window_attention_x1_kernel<3, true> : x1::window_attention_x1_kernel<4, false>;
  const int smem_bytes = vb ? x1::Lay<3, true>::kSmem : x1::Lay<4, false>::kSmem;
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      set_error("ab_window_attention_x1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return AB_ERR_CUDA;
    }
  }
  // largest run length R such that every run of R in-window tokens along W is all padding or contiguous (as shipped)
  auto gcd = [](int x, int y) { while (y) { int t = x % y; x = y; y = t; } return x; };
  int r = a.g.ws[2];
  for (int kw = 0; kw < a.g.nwin[2]; ++kw) {
    int prev_valid = -1, prev_src = 0;
    for (int i = 0; i < a.g.ws[2]; ++i) {
      const int q = kw * a.g.ws[2] + i - a.g.lo[2];
      const int valid = q >= 0 && q < a.g.res[2];
      const int src = valid ? (q + a.g.ss[2]) % a.g.res[2] : -1;
      if (i > 0 && (valid != prev_valid || (valid && src != prev_src + 1))) r = gcd(r, i);
      prev_valid = valid;
      prev_src = src;
    }
  }
  CUtensorMap tq, th;
  memset(&tq, 0, sizeof(tq));
  memset(&th, 0, sizeof(th));
  const long long rows = static_cast<long long>(p->batch) * a.tokens_per_batch;
  if (make_tmap_16bit_2d(&tq, p->qkv, rows, 3ll * a.dim, 3ll * a.dim, r, kHeadDim, false) == AB_OK) a.box_rows = r;
  const long long items = static_cast<long long>(p->batch) * a.g.nwindows * p->num_heads;
  const unsigned grid = static_cast<unsigned>(items < sm_count() ? items : sm_count());
  kern<<<grid, tc::kThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tq, th, a);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_window_attention_x1");
  return AB_OK;
}
