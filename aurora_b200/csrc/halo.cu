// Halo exchange of a latitude-sharded forecast through PEER MEMORY (NVLink / NVSwitch), no NCCL on the step path.
//
// The reference has no multi-GPU path at all (its only distributed artefact is a
// world-size-1 DDP smoke test, reference tests/test_model.py:96-110).  What forces an exchange once ONE forecast is
// split by latitude is the shifted-window attention of aurora/model/swin3d.py:470-503: windows straddle the band
// boundaries, so every rank needs `halo` rows of its neighbours' qkv projection.
//
// Every rank owns one device allocation that its neighbours map with CUDA IPC (ab_ipc_export / ab_ipc_open):
//
//     ctrl  : uint32[64]   [0] flag "rows above me have landed"  (written by the rank above, value = round number)
//                          [1] flag "rows below me have landed"  (written by the rank below)
//                          [2] round number of MY last push      (local)
//                          [3] CTA completion counter of the running push (local)
//     slots : bf16 [2 parities][2 sides][C][halo][X]   side 0 = rows above my band, side 1 = rows below
//
// ab_halo_push : every rank copies its FIRST rows into the rank above's side-1 slot and its LAST rows into the rank
//                below's side-0 slot with plain 16-byte stores over NVLink — only as many rows as that neighbour's
//                windows reach into this band (0..halo, computed by both sides from the global geometry) and only the
//                byte range of every token the neighbour reads (k | v of the qkv projection: a rank never needs its
//                neighbours' queries) — then the last CTA to finish publishes the new round number in both
//                neighbours' flags (st.release.sys after __threadfence_system).
// ab_halo_wait : one warp spins (ld.acquire.sys) until both of MY flags reached my own round number.
//
// Both are ordinary kernels on the caller's stream, so the whole sharded step — 48 exchanges — is captured in ONE
// CUDA graph.  Slots alternate by parity of the exchange index: a neighbour can be at most one exchange ahead of me
// (its push n+2 is ordered after its wait n+1, which needs my push n+1, which my stream issues after my attention n
// has read parity slot n & 1), so two slots are enough and no "buffer free" handshake is needed.
#include "common.h"

#include <string.h>

namespace ab {

constexpr int kHaloThreads = 256;
constexpr int kCtrlFlagAbove = 0, kCtrlFlagBelow = 1, kCtrlRound = 2, kCtrlDone = 3;

struct HaloPushArgs {
  const uint4* src;       // local band [C, rows, W, src_tok16] in 16-byte units
  uint4* dst_above;       // rank above: its side-1 slot [C, slot_rows, W, tok16]
  uint4* dst_below;       // rank below: its side-0 slot
  uint32_t* flag_above;   // rank above: its ctrl[kCtrlFlagBelow]
  uint32_t* flag_below;   // rank below: its ctrl[kCtrlFlagAbove]
  uint32_t* ctrl;         // local control words
  int c, rows, slot_rows, w;
  int rows_to_above, rows_to_below;
  int src_tok16, off16, tok16;  // 16-byte units per source token, offset of the copied range, units copied per token
  unsigned total_ctas;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid = (ctas_per_chunk, C, 2): blockIdx.z = 0 -> first rows to the rank above, 1 -> last rows to the rank below.
__global__ void __launch_bounds__(kHaloThreads) halo_push_kernel(const HaloPushArgs a) {
  const int side = blockIdx.z;
  const int c = blockIdx.y;
  const int n = side == 0 ? a.rows_to_above : a.rows_to_below;
  // side 0: my rows [0, n) are the rows just BELOW the rank above -> its side-1 slot rows [0, n)
  // side 1: my rows [rows - n, rows) are the rows just ABOVE the rank below -> its side-0 slot rows [slot_rows - n, ..)
  const long long src_row0 = side == 0 ? 0 : (a.rows - n);
  const long long dst_row0 = side == 0 ? 0 : (a.slot_rows - n);
  const uint4* src = a.src + (static_cast<long long>(c) * a.rows + src_row0) * a.w * a.src_tok16 + a.off16;
  uint4* dst = (side == 0 ? a.dst_above : a.dst_below) + (static_cast<long long>(c) * a.slot_rows + dst_row0) * a.w * a.tok16;
  const long long units = static_cast<long long>(n) * a.w * a.tok16;  // contiguous in dst, token-strided in src
  const long long stride = static_cast<long long>(gridDim.x) * kHaloThreads;
  auto src_of = [&](long long i) -> const uint4* {
    const long long tok = i / a.tok16;
    return src + tok * a.src_tok16 + (i - tok * a.tok16);
  };
  long long i = static_cast<long long>(blockIdx.x) * kHaloThreads + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread: the stores cross NVLink, keep the pipe full
  for (; i + 3 * stride < units; i += 4 * stride) {
    const uint4 v0 = __ldg(src_of(i)), v1 = __ldg(src_of(i + stride)), v2 = __ldg(src_of(i + 2 * stride)),
                v3 = __ldg(src_of(i + 3 * stride));
    dst[i] = v0;
    dst[i + stride] = v1;
    dst[i + 2 * stride] = v2;
    dst[i + 3 * stride] = v3;
  }
  for (; i < units; i += stride) dst[i] = __ldg(src_of(i));

  __threadfence_system();  // my stores are ordered before whatever this CTA's thread 0 publishes below
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&a.ctrl[kCtrlDone], 1u);
    if (done == a.total_ctas - 1) {
      __threadfence_system();
      const uint32_t round = a.ctrl[kCtrlRound] + 1u;
      a.ctrl[kCtrlRound] = round;
      a.ctrl[kCtrlDone] = 0u;
      st_release_sys(a.flag_above, round);
      st_release_sys(a.flag_below, round);
    }
  }
}

__global__ void halo_wait_kernel(uint32_t* ctrl) {
  if (threadIdx.x >= 2) return;
  const uint32_t want = ctrl[kCtrlRound];  // my own push of this round precedes me on the stream
  const uint32_t* flag = ctrl + (threadIdx.x == 0 ? kCtrlFlagAbove : kCtrlFlagBelow);
  unsigned long long t0 = 0;
  for (unsigned spin = 1;; ++spin) {
    // signed distance: round numbers wrap after 2^32 exchanges
    if (static_cast<int32_t>(ld_acquire_sys(flag) - want) >= 0) break;
    if ((spin & 0xFFFu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();  // 20 s: a neighbour died; fail the launch instead of hanging
    }
  }
}

// ---- CUDA IPC through the driver entry points (no link-time dependency on libcuda) -------------------------------
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetAddressRangeFn address_range_fn() {
  static GetAddressRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GetAddressRangeFn>(p);
  }
  return fn;
}

}  // namespace ab

extern "C" int ab_ipc_export(const void* dev_ptr, uint8_t* handle, uint64_t* offset) {
  using namespace ab;
  static_assert(sizeof(cudaIpcMemHandle_t) == AB_IPC_HANDLE_BYTES, "IPC handle size");
  AB_CHECK_ARG(dev_ptr != nullptr && handle != nullptr && offset != nullptr, "ab_ipc_export: null argument");
  GetAddressRangeFn range = address_range_fn();
  if (!range) {
    set_error("ab_ipc_export: cuMemGetAddressRange is unavailable");
    return AB_ERR_CUDA;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = range(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr));
  if (r != CUDA_SUCCESS) {
    set_error("ab_ipc_export: cuMemGetAddressRange failed (CUresult %d)", static_cast<int>(r));
    return AB_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) {
    set_error("ab_ipc_export: cudaIpcGetMemHandle failed: %s (the buffer must come from cudaMalloc, not from an "
              "expandable-segments / VMM allocator)", cudaGetErrorString(e));
    cudaGetLastError();
    return AB_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  *offset = static_cast<uint64_t>(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return AB_OK;
}

extern "C" int ab_ipc_open(const uint8_t* handle, void** base_out) {
  using namespace ab;
  AB_CHECK_ARG(handle != nullptr && base_out != nullptr, "ab_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("ab_ipc_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return AB_ERR_CUDA;
  }
  *base_out = p;
  return AB_OK;
}

extern "C" int ab_ipc_close(void* base) {
  using namespace ab;
  AB_CHECK_ARG(base != nullptr, "ab_ipc_close: null argument");
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) {
    set_error("ab_ipc_close: cudaIpcCloseMemHandle failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return AB_ERR_CUDA;
  }
  return AB_OK;
}

extern "C" int ab_halo_push(const AbHaloPush* p, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(p != nullptr && p->local != nullptr && p->above_slot != nullptr && p->below_slot != nullptr &&
                   p->above_flag != nullptr && p->below_flag != nullptr && p->ctrl != nullptr,
               "ab_halo_push: null argument");
  AB_CHECK_ARG(p->c > 0 && p->w > 0 && p->slot_rows > 0 && p->rows > 0 && p->rows_to_above >= 0 && p->rows_to_below >= 0 &&
                   p->rows_to_above <= p->slot_rows && p->rows_to_below <= p->slot_rows &&
                   p->rows_to_above <= p->rows && p->rows_to_below <= p->rows,
               "ab_halo_push: bad shape c=%d rows=%d w=%d slot_rows=%d to_above=%d to_below=%d", p->c, p->rows, p->w,
               p->slot_rows, p->rows_to_above, p->rows_to_below);
  AB_CHECK_ARG(p->src_tok_bytes > 0 && p->tok_bytes > 0 && p->tok_off_bytes >= 0 &&
                   p->tok_off_bytes + p->tok_bytes <= p->src_tok_bytes && p->src_tok_bytes % 16 == 0 &&
                   p->tok_off_bytes % 16 == 0 && p->tok_bytes % 16 == 0,
               "ab_halo_push: token byte range [%lld, +%lld) of %lld must be 16-byte granular",
               static_cast<long long>(p->tok_off_bytes), static_cast<long long>(p->tok_bytes),
               static_cast<long long>(p->src_tok_bytes));
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  AB_CHECK_ARG(al16(p->local) && al16(p->above_slot) && al16(p->below_slot), "ab_halo_push: 16-byte alignment");
  HaloPushArgs a;
  a.src = reinterpret_cast<const uint4*>(p->local);
  a.dst_above = reinterpret_cast<uint4*>(p->above_slot);
  a.dst_below = reinterpret_cast<uint4*>(p->below_slot);
  a.flag_above = p->above_flag;
  a.flag_below = p->below_flag;
  a.ctrl = p->ctrl;
  a.c = p->c;
  a.rows = p->rows;
  a.slot_rows = p->slot_rows;
  a.w = p->w;
  a.rows_to_above = p->rows_to_above;
  a.rows_to_below = p->rows_to_below;
  a.src_tok16 = static_cast<int>(p->src_tok_bytes / 16);
  a.off16 = static_cast<int>(p->tok_off_bytes / 16);
  a.tok16 = static_cast<int>(p->tok_bytes / 16);
  const int nmax = p->rows_to_above > p->rows_to_below ? p->rows_to_above : p->rows_to_below;
  const long long units = static_cast<long long>(nmax) * a.w * a.tok16;
  // enough CTAs to keep ~900 GB/s of NVLink stores in flight (2 per SM over the whole grid); even when nothing is
  // sent one CTA per (level, side) runs so that the round is published
  long long per = (units + 4ll * kHaloThreads - 1) / (4ll * kHaloThreads);
  const long long cap = (2ll * sm_count() + 2ll * p->c - 1) / (2ll * p->c);
  if (per > cap) per = cap;
  if (per < 1) per = 1;
  dim3 grid(static_cast<unsigned>(per), static_cast<unsigned>(p->c), 2);
  a.total_ctas = grid.x * grid.y * grid.z;
  halo_push_kernel<<<grid, kHaloThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_halo_push");
  return AB_OK;
}

extern "C" int ab_halo_wait(uint32_t* ctrl, void* stream) {
  using namespace ab;
  AB_CHECK_ARG(ctrl != nullptr, "ab_halo_wait: null argument");
  halo_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ctrl);
  AB_COUNT_LAUNCH(1);
  AB_CHECK_LAUNCH("ab_halo_wait");
  return AB_OK;
}
