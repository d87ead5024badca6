// bf16 tensor-core path of the network tower (kernel family K6): 3x3 convolution C=128 -> 128 as an
// implicit GEMM on the 5th-generation tensor cores.
//
//   D[m, n] = sum_{tap, ci} A_tap[m, ci] * W[tap][ci][n],   m = (position, cell) row, n = out channel
//   M tile = 128 rows, N = 128, K = 9 taps x 128 channels = 18 k-blocks of 64
//
// * tcgen05.mma (cta_group::1, kind::f16, M128 N128 K16, bf16 x bf16 -> fp32) issued by ONE thread;
//   the accumulator lives in TMEM (2 stages x 128 columns) and is read back with tcgen05.ld.
// * B (weights) is pre-swizzled on the host into 16-KB K-major SWIZZLE_128B tiles; one bulk-TMA copy
//   (cp.async.bulk, completes on an mbarrier) brings a tile into a pipeline stage.
// * A (activations, bf16 [rows][128] in HBM/L2) has no im2col copy in memory: four producer warps
//   gather the shifted rows of the tap (zeros outside the 6x7 board) straight into the swizzled
//   stage (the shift is not expressible in a UMMA descriptor: rows come in groups of 8).
// * Warp roles (320 threads): 0-3 A producers, 4-7 epilogue (TMEM lanes 32*(w%4)..), 8 MMA issuer,
//   9 weight loader + TMEM allocator.  Three pipelines: smem full/empty (4 stages), TMEM full/empty.
// * Epilogue fused: + bias (folded BN), + residual, ReLU, fp32 -> bf16, straight to HBM.
#pragma once
#include <cuda.h>  // CUtensorMap (the encoder itself is fetched at run time with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "nnet.cuh"

namespace azb {

constexpr int kTcTileM = 128;
constexpr int kTcBlockK = 64;
#ifndef AZB_TC_STAGES
#define AZB_TC_STAGES 4
#endif
#ifndef AZB_TC_CPASYNC
#define AZB_TC_CPASYNC "cp.async.cg.shared.global"
#endif
constexpr int kTcStages = AZB_TC_STAGES;
constexpr int kTcKBlocks = 18;              // 9 taps x 2 halves of 64 input channels
constexpr uint32_t kTcTileBytes = 128 * 64 * 2;  // one 128-row operand tile of a k-block: 128 rows x 128 bytes
#ifndef AZB_TC_SUB
#define AZB_TC_SUB 2
#endif
constexpr int kTcSub = AZB_TC_SUB;                        // 128-row A sub-tiles per CTA tile: they share every B tile
constexpr int kTcCtaRows = kTcTileM * kTcSub;
constexpr uint32_t kTcStageBytes = (kTcSub + 1) * kTcTileBytes;  // A sub-tiles + B
constexpr int kTcThreads = 320;
#ifndef AZB_TC_WCOPIES
#define AZB_TC_WCOPIES 8
#endif
constexpr int kTcWeightCopies = AZB_TC_WCOPIES;
constexpr int kTcCluster = 4;               // CTAs sharing each weight tile by multicast
constexpr uint32_t kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// The same copy delivered to the same shared-memory offset of every CTA in cta_mask; each destination
// CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                       uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// >> 4 in [0,14), LBO [16,30) (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B) >> 4 in
// [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = BF16 (bits 7, 10),
// both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kIdescBf16M128N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ConvTcArgs {
  const __nv_bfloat16* in;        // [rows][128]
  const __nv_bfloat16* residual;  // [rows][128] or nullptr
  __nv_bfloat16* out;             // [rows][128]
  const uint8_t* w_tiles;         // [18][16384] pre-swizzled B tiles of this layer (replica 0)
  size_t w_copy_stride;           // byte distance between the kTcWeightCopies replicas
  const float* bias;              // [128]
  const uint32_t* count;          // positions this round (device), or nullptr
  uint32_t max_batch;
  unsigned long long* dbg;        // diagnostic (AZB200_TC_DEBUG=1): per-role cycle counters of CTA pair 0, or nullptr
  // k_conv3x3_tc3 only.  mode 0 (forward): out = ReLU(acc + bias [+ residual]).  mode 1 (backward data, run with the
  // tap-mirrored / transposed weight tiles): out = (acc [+ residual]) * (mask > 0), no bias, no ReLU.
  int mode;
  const __nv_bfloat16* mask;      // [rows][128] forward activation whose sign gates the gradient, or nullptr
};

// CL = CTAs per cluster.  With CL > 1 the CTAs of a cluster walk their M tiles in lock-step and share
// every weight tile: CTA r fetches the r-th 1/CL of it and multicasts that slice to all of them, and a
// pipeline stage is released only when every CTA of the cluster has consumed it (multicast commit).
template <int CL>
__global__ void __launch_bounds__(kTcThreads, 1) k_conv3x3_tc(ConvTcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-B alignment
  // stage s: [A sub-tile 0][A sub-tile 1][B]
  auto stage_a = [&](int s, int sub) { return base + s * kTcStageBytes + sub * kTcTileBytes; };
  auto stage_b = [&](int s) { return base + s * kTcStageBytes + kTcSub * kTcTileBytes; };
  const uint32_t bars = base + kTcStages * kTcStageBytes;
  auto bar_full_a = [&](int s) { return bars + 8u * s; };
  auto bar_full_b = [&](int s) { return bars + 64u + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 128u + 8u * s; };
  auto bar_acc_full = [&](int a) { return bars + 192u + 8u * a; };
  auto bar_acc_empty = [&](int a) { return bars + 208u + 8u * a; };
  const uint32_t tmem_slot = bars + 224u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_pos = g.count ? min(*g.count, g.max_batch) : g.max_batch;
  const uint32_t rows = n_pos * kCells;
  const uint32_t n_tiles = (rows + kTcCtaRows - 1) / kTcCtaRows;
  // tile schedule: group = cluster index, rank = CTA within the cluster; every CTA of a cluster runs
  // the same number of iterations (tiles past the end are all-zero rows whose output is dropped)
  const uint32_t rank = blockIdx.x % CL, group = blockIdx.x / CL, n_groups = gridDim.x / CL;
  const uint32_t iters = (n_tiles + n_groups * CL - 1) / (n_groups * CL);
  const uint16_t cl_mask = static_cast<uint16_t>((1u << CL) - 1u);
  auto tile_of = [&](uint32_t i) { return (i * n_groups + group) * CL + rank; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(bar_full_a(s), 128);
      mbar_init(bar_full_b(s), 1);
      mbar_init(bar_empty(s), CL);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full(a), 1);
      mbar_init(bar_acc_empty(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {  // TMEM: 2 accumulator stages x 2 sub-tiles x 128 fp32 columns = all 512
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers exist before anyone multicasts into them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ===== A producers: 8 consecutive lanes fetch the 8 16-byte chunks of one 128-byte row-half, so
    // every warp-wide cp.async covers 4 whole 128-byte lines; thread t owns chunk t%8 of rows t/8 + 16i =====
    const int t = threadIdx.x;
    const uint32_t j = t & 7u;
    uint32_t it = 0;
    for (uint32_t i = 0; i < iters; ++i) {
      const uint32_t tile = tile_of(i);
      if (CL == 1 && tile >= n_tiles) break;
      uint32_t rc[8 * kTcSub];  // per owned row: (r << 8) | c, or 0xFFFF when the row is past the end
#pragma unroll
      for (int i = 0; i < 8 * kTcSub; ++i) {
        const uint32_t m = tile * kTcCtaRows + (t >> 3) + 16u * i;
        const uint32_t cell = m % kCells;
        rc[i] = m < rows ? ((cell / 7u) << 8) | (cell % 7u) : 0xFFFFu;
      }
      for (int kb = 0; kb < kTcKBlocks; ++kb, ++it) {
        const int s = it % kTcStages;
        mbar_wait(bar_empty(s), ((it / kTcStages) & 1u) ^ 1u);
        const int tap = kb >> 1, dy = tap / 3 - 1, dx = tap % 3 - 1;
        const uint32_t stage = stage_a(s, 0);
        const __nv_bfloat16* colbase = g.in + (kb & 1) * kTcBlockK + j * 8;
#pragma unroll
        for (int i = 0; i < 8 * kTcSub; ++i) {
          const uint32_t row = (t >> 3) + 16u * i;  // 0..255; sub-tile = row / 128
          const int rr = static_cast<int>(rc[i] >> 8) + dy, cc = static_cast<int>(rc[i] & 0xFFu) + dx;
          const bool ok = rc[i] != 0xFFFFu && rr >= 0 && rr < 6 && cc >= 0 && cc < 7;
          const size_t srow = static_cast<size_t>(tile * kTcCtaRows + row) + dy * 7 + dx;
          const __nv_bfloat16* src = ok ? colbase + srow * kNetC : g.in;
          // swizzle-128B: chunk j of row lands at chunk (j ^ row%8); src-size 0 zero-fills.
          // (row >> 3) * 1024 walks straight from sub-tile 0 into sub-tile 1 (16 KB each)
          const uint32_t dst = stage + (row >> 3) * 1024u + (row & 7u) * 128u + ((j ^ (row & 7u)) << 4);
          asm volatile(AZB_TC_CPASYNC " [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
        }
        // The barrier gets this thread's arrival when the copies above have landed: no polling, so the
        // producers run ahead as far as the free stages allow (the consumer issues the proxy fence).
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full_a(s)) : "memory");
      }
    }
  } else if (warp < 8) {
    // ===== epilogue: TMEM -> registers -> bias / residual / ReLU -> bf16 -> HBM =====
    const int q = warp - 4;
    for (uint32_t ti = 0; ti < iters; ++ti) {
      const uint32_t tile = tile_of(ti);
      if (CL == 1 && tile >= n_tiles) break;
      const uint32_t a = ti & 1u;
      mbar_wait(bar_acc_full(a), (ti >> 1) & 1u);
      tc_fence_after();
      for (int sub = 0; sub < kTcSub; ++sub) {
      const uint32_t m = tile * kTcCtaRows + sub * kTcTileM + q * 32 + lane;
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 256u + sub * 128u + ch * 32u, acc);
        if (m < rows) {
          const float* bias = g.bias + ch * 32;
          uint32_t packed[16];
          uint4 res[4];
          if (g.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(g.residual + static_cast<size_t>(m) * kNetC + ch * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = rp[j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x0 = __uint_as_float(acc[2 * j]) + bias[2 * j];
            float x1 = __uint_as_float(acc[2 * j + 1]) + bias[2 * j + 1];
            if (g.residual) {
              const uint32_t rw = reinterpret_cast<const uint32_t*>(res)[j];
              x0 += __uint_as_float(rw << 16);
              x1 += __uint_as_float(rw & 0xFFFF0000u);
            }
            x0 = fmaxf(x0, 0.0f);
            x1 = fmaxf(x1, 0.0f);
            const __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
            packed[j] = *reinterpret_cast<const uint32_t*>(&p);
          }
          uint4* op = reinterpret_cast<uint4*>(g.out + static_cast<size_t>(m) * kNetC + ch * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) op[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty(a));
    }
  } else if (warp == 8) {
    // ===== MMA issuer: one elected thread =====
    uint32_t it = 0;
    for (uint32_t ti = 0; ti < iters; ++ti) {
      if (CL == 1 && tile_of(ti) >= n_tiles) break;
      const uint32_t a = ti & 1u;
      mbar_wait(bar_acc_empty(a), ((ti >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < kTcKBlocks; ++kb, ++it) {
        const int s = it % kTcStages;
        const uint32_t ph = (it / kTcStages) & 1u;
        mbar_wait(bar_full_a(s), ph);
        mbar_wait(bar_full_b(s), ph);
        fence_proxy_async();  // cp.async (generic proxy) writes of A -> the tensor core's async-proxy reads
        tc_fence_after();
        if (lane == 0) {
          const uint64_t bd = umma_desc_sw128(stage_b(s));
#pragma unroll
          for (int sub = 0; sub < kTcSub; ++sub) {
            const uint64_t ad = umma_desc_sw128(stage_a(s, sub));
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
              umma_bf16(tmem_base + a * 256u + sub * 128u, ad + 2u * k, bd + 2u * k, kIdescBf16M128N128, (kb | k) ? 1u : 0u);
          }
          // frees the stage (in every CTA of the cluster) when these MMAs have read it
          if (CL > 1) umma_commit_multicast(bar_empty(s), cl_mask);
          else umma_commit(bar_empty(s));
          if (kb == kTcKBlocks - 1) umma_commit(bar_acc_full(a));
        }
        __syncwarp();
      }
    }
  } else {
    // ===== weight loader: bulk TMA of pre-swizzled 16-KB B tiles =====
    uint32_t it = 0;
    constexpr uint32_t kSlice = kTcTileBytes / CL;
    for (uint32_t i = 0; i < iters; ++i) {
      if (CL == 1 && tile_of(i) >= n_tiles) break;
      for (int kb = 0; kb < kTcKBlocks; ++kb, ++it) {
        const int s = it % kTcStages;
        mbar_wait(bar_empty(s), ((it / kTcStages) & 1u) ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_full_b(s), kTcTileBytes);  // the whole tile: CL slices arrive
          const uint8_t* src = g.w_tiles + (group % kTcWeightCopies) * g.w_copy_stride +
                               static_cast<size_t>(kb) * kTcTileBytes + rank * kSlice;
          const uint32_t dst = stage_b(s) + rank * kSlice;
          if (CL > 1) tma_bulk_g2s_multicast(dst, src, kSlice, bar_full_b(s), cl_mask);
          else tma_bulk_g2s(dst, src, kSlice, bar_full_b(s));
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody leaves while a peer may still multicast into it
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ================================================================================================
// CTA-pair version: tcgen05.mma.cta_group::2 (M256 N128 K16), the layer's weights RESIDENT in shared
// memory, and the A operand gathered by TMA in IM2COL mode.
//
// Why: the single-CTA kernel above is bound by what each SM pulls through its load/store unit and its
// shared-memory port per k-block (A gathered with 16-byte cp.async at <= ~32 B/clk/SM, a streamed 16-KB
// weight tile, 64 KB read by the SS-mode MMAs).  Here a CTA pair (the two SMs of a TPC, cluster of 2)
// computes a 256-row tile together:
//  * each CTA holds HALF of the layer's weights (64 of the 128 output channels, all 18 k-blocks =
//    144 KB), loaded ONCE per launch with bulk TMA; the pair's tensor cores exchange the halves, so per
//    MMA an SM reads 4 KB (A) + 2 KB (B) of shared memory instead of 8 KB, and no weight tile is ever
//    streamed again;
//  * the activations are a 4-D tensor [position][6][7][128] to the TMA unit; ONE
//    cp.async.bulk.tensor.4d.im2col per k-block and CTA delivers the 128 consecutive (position, cell)
//    rows of the CTA's sub-tile, shifted by the tap (dy, dx), zero-filled outside the 6x7 board, 64
//    channels wide, already in the SWIZZLE_128B layout the UMMA descriptor expects — no gather warps,
//    no address arithmetic, no generic-proxy writes (so no proxy fence), and both CTAs' copies
//    complete on the LEADER's mbarrier (.cta_group::2), so the pair needs no relay either;
//  * accumulators: 2 x 128 TMEM columns in each CTA's own TMEM (its 128 rows).
// Roles (both CTAs unless noted): warps 0-7 epilogue (accumulator-empty arrivals go to the leader);
// warp 8 lane 0 of the leader (cluster rank 0) issues every MMA of the pair and the multicast commits;
// warp 9: TMEM allocation (cta_group::2: the same warp in both CTAs), then lane 0 preloads the weights
// and is the TMA producer.
// ================================================================================================
constexpr uint32_t kT2WTile = 64 * 128;                    // one k-block of the CTA's 64 output channels
constexpr uint32_t kT2WBytes = kTcKBlocks * kT2WTile;      // 147456: a CTA's resident half of a layer's weights
constexpr int kT2PairRows = 2 * kTcTileM;
constexpr uint32_t kIdescBf16M256N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `cta` of the cluster
// one lane of a converged warp (the same lane every time it is called by the same warp)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0u;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
  return remote;
}
// arrive on the barrier at the same offset in CTA `cta`.  Default semantics (release at CTA scope), as
// CUTLASS's umma_arrive_2x1SM_sm0: a cluster-scope release was measured at ~900 cycles per arrival.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(map_to_cta(local_bar, cta)) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
struct ActLayout {
  uint32_t pos_rows, pitch;  // 42, 7 or 56, 8
  __host__ __device__ size_t row(uint32_t pos, int y, int x) const { return static_cast<size_t>(pos) * pos_rows + y * pitch + x; }
};
constexpr ActLayout kActDense{42u, 7u};
constexpr ActLayout kActPadded{56u, 8u};

// ================================================================================================
// CTA-pair version 3: the A operand is fetched ONCE per tile and reused by all nine taps.
//
// k_conv3x3_tc2 asks the TMA unit for a shifted copy of the tile per tap and channel half (18 x 16 KB per
// 128 rows), and the unit's im2col rate (~5 cycles per 128-byte pixel row) ends up feeding the tensor
// core at ~40 % of what it could consume.  Here the activations live in the PADDED layout (ActLayout:
// 56 rows per position, board rows of 8 with a zero column, a zero row after every position), in which
// the tap (dy, dx) of output row R is input row R + dy * 8 + dx, zero padding included.  So a CTA
// loads rows [R0 - 16, R0 + 144) of its 128-row tile once per channel half (2 x 20 KB, plain 2-D TMA,
// SWIZZLE_128B, out-of-range rows zero-filled) and the 18 k-blocks of the tile are 18 views of the same
// stage: the UMMA descriptor's start address moves by (16 + dy * 8 + dx) rows, with the descriptor's
// base-offset field giving the swizzle phase of a start that is not 1024-byte aligned.  Per tile the
// issuing thread waits once, issues 72 MMAs and commits twice; 25 % of the rows of a tile are padding
// rows (computed, never stored).  Weights resident as in tc2 (144 KB per CTA); 2 stages of 40 KB
// (the next tile is fetched while the current one is computed); accumulators 2 x 128 TMEM columns.
// ================================================================================================
constexpr int kT3HaloRows = 16;                                   // rows fetched before the tile (>= 9, multiple of 8)
constexpr int kT3StageRows = kTcTileM + 2 * kT3HaloRows;          // 160
constexpr uint32_t kT3HalfBytes = kT3StageRows * 128;             // 20480: one channel half of a stage
constexpr uint32_t kT3StageBytes = 2 * kT3HalfBytes;              // 40960
constexpr int kT3Stages = 2;
constexpr uint32_t kT3SmemBytes = kT2WBytes + kT3Stages * kT3StageBytes + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias*/ +
                                  320 /*weight barriers, one per k-block and CTA*/;

// UMMA descriptor of a 128-row K-major SWIZZLE_128B operand that starts at an arbitrary 128-byte row of a
// 1024-byte aligned buffer: the base-offset field [49, 52) carries (start address >> 7) & 7.
__device__ __forceinline__ uint64_t umma_desc_sw128_rows(uint32_t smem_addr) {
  // Measured on B200: the swizzle XOR is taken from the absolute shared-memory address bits, so a start that is
  // not 1024-byte aligned needs NO base offset (with base offset = (addr >> 7) & 7 the taps with dx != 0 came
  // out wrong; with 0 the result is bit-identical to the im2col kernel).
  return umma_desc_sw128(smem_addr);
}
// two 32-column accumulator chunks at once, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, uint32_t (&r0)[32], uint32_t (&r1)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
        "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15]), "=r"(r0[16]),
        "=r"(r0[17]), "=r"(r0[18]), "=r"(r0[19]), "=r"(r0[20]), "=r"(r0[21]), "=r"(r0[22]), "=r"(r0[23]), "=r"(r0[24]),
        "=r"(r0[25]), "=r"(r0[26]), "=r"(r0[27]), "=r"(r0[28]), "=r"(r0[29]), "=r"(r0[30]), "=r"(r0[31])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
        "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15]), "=r"(r1[16]),
        "=r"(r1[17]), "=r"(r1[18]), "=r"(r1[19]), "=r"(r1[20]), "=r"(r1[21]), "=r"(r1[22]), "=r"(r1[23]), "=r"(r1[24]),
        "=r"(r1[25]), "=r"(r1[26]), "=r"(r1[27]), "=r"(r1[28]), "=r"(r1[29]), "=r"(r1[30]), "=r"(r1[31])
      : "r"(taddr + 32u));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256).  The epilogue's rows are 256 bytes apart, one per lane, so a
// warp-wide access touches 32 lines whatever its width: half as many accesses = half as many L1 wavefronts.
__device__ __forceinline__ void ld_global_256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
__device__ __forceinline__ void tma_tile2d_pair(uint32_t dst, const CUtensorMap* tmap, uint32_t mbar_cluster, int c, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar_cluster), "r"(c), "r"(row)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
k_conv3x3_tc3(ConvTcArgs g, const __grid_constant__ CUtensorMap tmap_in) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto w_tile = [&](int kb) { return base + kb * kT2WTile; };
  auto stage_a = [&](int s, int half) { return base + kT2WBytes + s * kT3StageBytes + half * kT3HalfBytes; };
  const uint32_t bars = base + kT2WBytes + kT3Stages * kT3StageBytes;
  auto bar_full = [&](int s) { return bars + 8u * s; };              // leader: 1 arrival (its expect_tx) + 2 x 40 KB of TMA bytes
  auto bar_empty = [&](int s) { return bars + 16u + 8u * s; };       // 1 arrival: the pair's MMAs have read the stage
  auto bar_acc_full = [&](int a) { return bars + 32u + 8u * a; };    // 1 arrival: the tile's MMAs are done
  auto bar_acc_empty = [&](int a) { return bars + 48u + 8u * a; };   // leader: 16 arrivals (one per epilogue warp of both CTAs)
  // weights: one barrier per k-block (the first tile's MMAs start when the first 8 KB have landed, the other 136 KB
  // arrive under them) and, on the leader, one more per k-block that the peer CTA's relay thread arrives on
  auto bar_w = [&](int kb) { return bars + 768u + 8u * kb; };          // 1 arrival (expect_tx) + 8 KB of bulk-copy bytes
  auto bar_w_peer = [&](int kb) { return bars + 768u + 144u + 8u * kb; };  // leader: 1 arrival from the peer's relay
  const uint32_t tmem_slot = bars + 80u;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));  // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t n_pos = g.count ? min(*g.count, g.max_batch) : g.max_batch;
  const uint32_t rows = n_pos * kActPadded.pos_rows;  // padded rows
  const uint32_t n_tiles = (rows + kT2PairRows - 1) / kT2PairRows;
  const uint32_t iters = pair < n_tiles ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0u;
  auto row0_of = [&](uint32_t i) { return (pair + i * n_pairs) * kT2PairRows + rank * kTcTileM; };
  const bool dbg_on = g.dbg != nullptr && blockIdx.x < 2;
  unsigned long long dbg_t[6] = {0, 0, 0, 0, 0, 0};
  long long dbg_c = 0;
#define AZB_DBG_T0() do { if (dbg_on) dbg_c = clock64(); } while (0)
#define AZB_DBG_ADD(k) do { if (dbg_on) { const long long n_ = clock64(); dbg_t[k] += n_ - dbg_c; dbg_c = n_; } } while (0)
  const long long dbg_start = dbg_on ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT3Stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full(a), 1);
      mbar_init(bar_acc_empty(a), 16);
    }
    for (int kb = 0; kb < kTcKBlocks; ++kb) {
      mbar_init(bar_w(kb), 1);
      mbar_init(bar_w_peer(kb), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_in)) : "memory");
  }
  if (threadIdx.x < kNetC) s_bias[threadIdx.x] = g.bias[threadIdx.x];  // (parameters: not written by the previous layer)
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (dbg_on && threadIdx.x == 0) g.dbg[rank * 16 + 13] = clock64() - dbg_start;

  if (warp < 8) {
    // ===== epilogue, 8 warps: TMEM lane quarter warp % 4, output channels 64 * (warp / 4) .. + 63 =====
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;
    for (uint32_t ti = 0; ti < iters; ++ti) {
      const uint32_t a = ti & 1u;
      const uint32_t m = row0_of(ti) + q * 32 + lane;        // padded row
      const uint32_t rem = m % kActPadded.pos_rows;
      const bool real = m < rows && (rem & 7u) != 7u && rem < 48u;  // not the zero column, not the zero row
      // the residual (64 channels = 8 x 16 B of my row) is requested before the accumulator is waited for
      uint4 res[8], msk[8];
      if (real && g.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(g.residual + static_cast<size_t>(m) * kNetC + half * 64);
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) ld_global_256(rp + jj, res[jj], res[jj + 1]);
      }
      if (real && g.mask) {
        const uint4* mp = reinterpret_cast<const uint4*>(g.mask + static_cast<size_t>(m) * kNetC + half * 64);
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) ld_global_256(mp + jj, msk[jj], msk[jj + 1]);
      }
      AZB_DBG_T0();
      mbar_wait(bar_acc_full(a), (ti >> 1) & 1u);
      AZB_DBG_ADD(0);
      tc_fence_after();
      uint32_t acc0[32], acc1[32];
      tmem_ld32x2(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 128u + half * 64u, acc0, acc1);
      tc_fence_before();
      __syncwarp();
      // the accumulator is in registers: the next tile may overwrite it.  ONE arrival per warp: 512 per-thread arrivals
      // on the leader's barrier (half of them remote) serialised into ~5 k cycles per tile once the MMA warp got faster
      if (lane == 0) mbar_arrive_cluster(bar_acc_empty(a), 0u);
      if (real) {
        uint4* op = reinterpret_cast<uint4*>(g.out + static_cast<size_t>(m) * kNetC + half * 64);
        uint4 prev = make_uint4(0u, 0u, 0u, 0u);  // an even chunk waits for the odd one: one 32-byte store per two chunks
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {  // 8 output channels per 16-byte chunk
          const uint32_t* acc = c8 < 4 ? acc0 : acc1;
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + half * 64 + c8 * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + half * 64 + c8 * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const uint32_t rw[4] = {res[c8].x, res[c8].y, res[c8].z, res[c8].w};
          const uint32_t mw[4] = {msk[c8].x, msk[c8].y, msk[c8].z, msk[c8].w};
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0 = __uint_as_float(acc[(c8 & 3) * 8 + 2 * e]), x1 = __uint_as_float(acc[(c8 & 3) * 8 + 2 * e + 1]);
            if (g.mode == 0) {
              x0 += bb[2 * e];
              x1 += bb[2 * e + 1];
            }
            if (g.residual) {
              x0 += __uint_as_float(rw[e] << 16);
              x1 += __uint_as_float(rw[e] & 0xFFFF0000u);
            }
            if (g.mode == 0) {
              x0 = fmaxf(x0, 0.0f);
              x1 = fmaxf(x1, 0.0f);
            } else if (g.mask) {  // d ReLU: the gradient passes where the forward activation was positive
              if (!(__uint_as_float(mw[e] << 16) > 0.0f)) x0 = 0.0f;
              if (!(__uint_as_float(mw[e] & 0xFFFF0000u) > 0.0f)) x1 = 0.0f;
            }
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(x0, x1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&p2);
          }
          if (c8 & 1) st_global_256(op + c8 - 1, prev, make_uint4(pk[0], pk[1], pk[2], pk[3]));
          else prev = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      AZB_DBG_ADD(1);
    }
    if (dbg_on && threadIdx.x == 128) {
      g.dbg[rank * 16 + 2] = dbg_t[0];
      g.dbg[rank * 16 + 3] = dbg_t[1];
      g.dbg[rank * 16 + 10] = clock64() - dbg_start;
    }
  } else if (warp == 8) {
    if (iters > 0 && rank == 1) {
      if (lane == 0) {
        // relay: tell the leader's MMA warp, k-block by k-block, that this CTA's half of the weights has landed
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          mbar_wait(bar_w(kb), 0u);
          mbar_arrive_cluster(bar_w_peer(kb), 0u);
        }
      }
    } else if (iters > 0) {
      // ===== MMA issuer of the pair: per tile one wait, 18 k-blocks x 4 MMAs on shifted views of the stage, two commits.
      // The WHOLE warp runs the loop and one elected lane issues: every address below is warp-uniform and, with the
      // k-block loop unrolled, a compile-time offset from two descriptors per tile.  (The first version ran the loop in
      // one lane with the k-block loop rolled: ~35 dependent ALU instructions + a register-to-uniform move per k-block
      // sat right at the 320 cycles the four MMAs take, and any addition to the loop made the issuing thread the limit.)
      const bool me = elect_one_sync();
      const bool dbg_mma = dbg_on && me;
      for (uint32_t ti = 0; ti < iters; ++ti) {
        const uint32_t a = ti & 1u;
        const int s = ti % kT3Stages;
        if (dbg_mma) dbg_c = clock64();
        mbar_wait(bar_acc_empty(a), ((ti >> 1) & 1u) ^ 1u);
        if (dbg_mma) { const long long n_ = clock64(); dbg_t[1] += n_ - dbg_c; dbg_c = n_; }
        mbar_wait(bar_full(s), (ti / kT3Stages) & 1u);
        if (dbg_mma) { const long long n_ = clock64(); dbg_t[0] += n_ - dbg_c; dbg_c = n_; }
        tc_fence_after();
        if (dbg_mma && ti == 0) g.dbg[15] = clock64() - dbg_start;
        const uint64_t a0 = umma_desc_sw128(stage_a(s, 0)), b0 = umma_desc_sw128(w_tile(0));
        const uint32_t acc = tmem_base + a * 128u;
#pragma unroll
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          if (ti == 0) {  // first tile: k-block kb of the weights, both halves
            if (dbg_mma) dbg_c = clock64();
            mbar_wait(bar_w(kb), 0u);
            mbar_wait(bar_w_peer(kb), 0u);
            if (dbg_mma) { const long long n_ = clock64(); dbg_t[5] += n_ - dbg_c; dbg_c = n_; }
          }
          const int tap = kb >> 1, shift = kT3HaloRows + (tap / 3 - 1) * 8 + (tap % 3 - 1);  // rows
          const uint64_t ad = a0 + static_cast<uint64_t>(((kb & 1) * kT3HalfBytes + shift * 128) >> 4);
          const uint64_t bd = b0 + static_cast<uint64_t>((kb * kT2WTile) >> 4);
#pragma unroll
          for (int k = 0; k < kTcBlockK / 16; ++k)
            if (me) umma2_bf16(acc, ad + 2u * k, bd + 2u * k, kIdescBf16M256N128, (kb | k) ? 1u : 0u);
        }
        if (dbg_mma) { const long long n_ = clock64(); dbg_t[3] += n_ - dbg_c; dbg_c = n_; }
        if (me) {
          umma2_commit_multicast(bar_empty(s), 3u);     // the stage may be refilled (both CTAs)
          umma2_commit_multicast(bar_acc_full(a), 3u);  // the accumulator is complete (both CTAs' epilogues)
        }
        if (dbg_mma) { const long long n_ = clock64(); dbg_t[4] += n_ - dbg_c; dbg_c = n_; }
      }
      if (dbg_mma) for (int k = 0; k < 6; ++k) g.dbg[4 + k] = dbg_t[k];
    }
    __syncwarp();
  } else if (warp == 9) {
    if (lane == 0 && iters > 0) {
      // ===== weight preload (once per launch), then the A producer: two 2-D TMA copies per tile =====
      // An SM takes in ~64 bytes per clock, so the 144 KB of weights are ~2.3 k cycles of its inbound path: only the
      // first kWFirst k-blocks go ahead of the grid dependency (they cover the wait and the first MMAs), the first
      // activation tile follows as soon as the previous layer is complete, and the remaining weights stream in behind
      // it, ahead of the MMAs that need them (a k-block is consumed in ~200 cycles and fetched in ~128).
      constexpr int kWFirst = 6;
      auto fetch_w = [&](int kb) {
        mbar_arrive_expect_tx(bar_w(kb), kT2WTile);
        tma_bulk_g2s(w_tile(kb), g.w_tiles + static_cast<size_t>(kb) * kTcTileBytes + rank * kT2WTile, kT2WTile, bar_w(kb));
      };
      for (int kb = 0; kb < kWFirst; ++kb) fetch_w(kb);
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous layer's output is complete and visible
      if (dbg_on) g.dbg[rank * 16 + 14] = clock64() - dbg_start;
      for (uint32_t i = 0; i < iters; ++i) {
        const int s = i % kT3Stages;
        AZB_DBG_T0();
        mbar_wait(bar_empty(s), ((i / kT3Stages) & 1u) ^ 1u);
        AZB_DBG_ADD(0);
        if (rank == 0) mbar_arrive_expect_tx(bar_full(s), 2u * kT3StageBytes);  // both CTAs' copies land on this barrier
        const int r0 = static_cast<int>(row0_of(i)) - kT3HaloRows;  // negative for the first tile: zero-filled
        const uint32_t full = map_to_cta(bar_full(s), 0u);
        tma_tile2d_pair(stage_a(s, 0), &tmap_in, full, 0, r0);
        tma_tile2d_pair(stage_a(s, 1), &tmap_in, full, kTcBlockK, r0);
        if (i == 0)
          for (int kb = kWFirst; kb < kTcKBlocks; ++kb) fetch_w(kb);
        AZB_DBG_ADD(1);
      }
      if (dbg_on) { g.dbg[rank * 16 + 0] = dbg_t[0]; g.dbg[rank * 16 + 1] = dbg_t[1]; }
    }
    __syncwarp();
  }
  if (dbg_on && threadIdx.x == 0) { g.dbg[rank * 16 + 11] = clock64() - dbg_start; g.dbg[rank * 16 + 12] = iters; }
#undef AZB_DBG_T0
#undef AZB_DBG_ADD

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ================================================================================================
// The whole residual tower in ONE launch (forward pass of the leaf evaluator): k_conv3x3_tc3's tile pipeline inside a
// loop over the 2R layers, on tiles that never need another CTA pair's rows.
//
// Why: at the sizes a search round produces (~1 k positions = 3 tiles per CTA pair) a layer of k_conv3x3_tc3 spends
// ~13 k cycles in its MMAs and ~11 k around them (AZB200_TC_DEBUG / AZB200_TOWER_DEBUG timelines, profiles/r2_tower.md):
// the launch boundary (a CTA fills the SM's shared memory, so the next layer's CTA cannot be resident before this one has
// left), barrier / TMEM set-up, the wait for the whole previous grid, the first tile's fetch, and the last tile's epilogue
// with nothing behind it.  A first version of this kernel kept the 256-row tiles and put a grid-wide arrive / wait on a
// global counter between the layers: the boundary (stores -> fence -> atomic -> every other CTA -> acquire -> TMA) cost
// the same ~6 k cycles as the kernel boundary it replaced.
//
// So the dependency itself is removed.  In the padded layout a position is 56 rows that end with a zero row, and a 3x3
// tap never reaches across it: a tile that starts on a position boundary and holds whole positions depends on NOTHING
// outside itself, in any layer.  A tile is therefore 4 positions = 224 rows of a 256-row MMA (rows 224..255 are computed
// and dropped: 12.5 % more MMA work than k_conv3x3_tc3's dense tiling), a CTA pair owns its tiles through all 2R layers,
// and the only ordering left is inside the pair — "this CTA's rows of tile i of layer L are in global memory", from the
// epilogue warps (generic-proxy stores) to the pair's two producer lanes (TMA = async-proxy reads of tile i of layer L + 1):
//   epilogue warp   stores of the tile -> __syncwarp -> one arrive on a CTA-local mbarrier (ring of 4, 8 arrivals each);
//   publisher warp  (one per CTA) waits for the ring entry -> fence.acq_rel.cluster, cumulative over the eight warps'
//                   stores -> relaxed store of G + 1 into its flag in BOTH CTAs' shared memory.  The fence is
//                   MEMBAR.GPU + CCTL.IVALL, ~2 k cycles: paid by the epilogue warps it made them, not the tensor pipe,
//                   the tile period;
//   producer lane   one volatile 8-byte read of the two flags (cached minimum) -> fence.acq_rel.cluster ->
//                   fence.proxy.async.global (one proxy fence on the causality chain orders the generic store and the
//                   TMA read) -> the two TMA copies.
// With two or more tiles per pair the flags are up long before the pipeline gets there.  The next layer's first tile is
// requested while the current layer's last tile is still in the tensor pipe, and that tile hands the weights back k-block by
// k-block (one tcgen05.commit per k-block), so the next layer's 144 KB stream in under it.  No grid barrier, no cooperative
// launch, pairs drift apart freely; set-up happens once.
// Phases of the tile barriers continue across layers (global tile number G = layer * iters + tile); the weight barriers
// complete one phase per layer.  Results are bit-identical to the layer-by-layer kernels (same K order per output row).
// ================================================================================================
// A unit of work is a run of WHOLE positions that nothing outside reaches into: 4 positions in one tile (224 rows kept of the
// 256 computed, 12.5 % padding) or 9 positions in two consecutive tiles of the same pair (504 of 512, 1.6 %).  The two-tile
// unit halves the padding but has half as many units to deal out and couples its two tiles (each needs 16 rows of the other
// from the layer before); the kernel takes whichever gives the pair with the most work fewer tiles (tower_unit_pos).
constexpr uint32_t kTwTilePos = 4;  // the grid is sized for the one-tile unit (the larger tile count)
__host__ __device__ inline uint32_t tower_unit_pos(uint32_t n_pos, uint32_t n_pairs) {
  const uint32_t t4 = (n_pos + 3u) / 4u, u9 = (n_pos + 8u) / 9u;
  const uint32_t m4 = (t4 + n_pairs - 1u) / n_pairs, m9 = 2u * ((u9 + n_pairs - 1u) / n_pairs);
  return m9 < m4 || (m9 == m4 && m9 > 2u) ? 9u : 4u;
}
struct TowerModel {
  __nv_bfloat16* act[3];   // padded activation buffers; act[0] holds the stem's output, the result is in act[R odd ? 2 : 0]
  const uint8_t* w_tiles;  // layer l at + l * kTcKBlocks * kTcTileBytes
  const float* bias;       // layer l at + l * 128
  const uint32_t* count;   // positions this round (device), or nullptr
};
// One launch may carry TWO models of the same shape (the arena: the two players' leaf batches of a round).  The CTA pairs
// are split between them in proportion to their tile counts, so two half-empty passes become one full one.
struct TowerTcArgs {
  TowerModel m[2];
  int n_models;            // 1 or 2
  uint32_t max_batch;      // per model
  int n_layers;            // 2 x residual blocks
  unsigned long long* dbg; // [8 + 8 * n_layers] diagnostic timeline of CTA 0 (AZB200_TOWER_DEBUG=1), or nullptr
};
constexpr uint32_t kTwSmemBytes = kT3SmemBytes + 512 + 192;  // a second bias buffer (layers alternate), 18 + 4 more barriers
constexpr int kTwThreads = kTcThreads + 32;                  // + the publisher warp

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// the smaller of two consecutive flags (one volatile read of this CTA's own shared memory: a cluster-scope load per flag
// was measured at ~200 cycles)
__device__ __forceinline__ uint32_t min2_volatile_shared(uint32_t cta_addr) {
  uint32_t a, b;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(cta_addr) : "memory");
  return min(a, b);
}
__device__ __forceinline__ void st_relaxed_cluster(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTwThreads, 1)
k_tower_tc3(TowerTcArgs g, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
            const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
            const __grid_constant__ CUtensorMap map4, const __grid_constant__ CUtensorMap map5) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto w_tile = [&](int kb) { return base + kb * kT2WTile; };
  auto stage_a = [&](int s, int half) { return base + kT2WBytes + s * kT3StageBytes + half * kT3HalfBytes; };
  const uint32_t bars = base + kT2WBytes + kT3Stages * kT3StageBytes;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 16u + 8u * s; };
  auto bar_acc_full = [&](int a) { return bars + 32u + 8u * a; };
  auto bar_acc_empty = [&](int a) { return bars + 48u + 8u * a; };
  const uint32_t tmem_slot = bars + 80u;
  const uint32_t done_w = bars + 128u;  // [2] u32 (16 bytes apart): tiles below this global number are in global memory, CTA r's rows at [4 r]
  auto bar_st = [&](uint32_t G) { return bars + 1744u + 8u * (G & 3u); };  // 8 arrivals: the epilogue warps have issued tile G's stores
  auto bar_w = [&](int kb) { return bars + 768u + 8u * kb; };
  auto bar_w_peer = [&](int kb) { return bars + 768u + 144u + 8u * kb; };
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));    // bias of the even layers
  float* s_bias1 = reinterpret_cast<float*>(smem_raw + (bars + 1088u - smem_u32(smem_raw)));  // ... of the odd layers
  auto bar_w_free = [&](int kb) { return bars + 1600u + 8u * kb; };  // 1 arrival per layer: its last tile is done with k-block kb

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // which model this pair works for, and as which of how many pairs
  uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1, model = 0u;
  const uint32_t n_pos0 = g.m[0].count ? min(*g.m[0].count, g.max_batch) : g.max_batch;
  uint32_t n_pos = n_pos0;
  if (g.n_models == 2) {
    const uint32_t n_pos1 = g.m[1].count ? min(*g.m[1].count, g.max_batch) : g.max_batch;
    const uint32_t t0 = (n_pos0 + 3u) / 4u, t1 = (n_pos1 + 3u) / 4u;
    uint32_t p0 = n_pairs;  // pairs of model 0
    if (t1 > 0u) p0 = t0 == 0u ? 0u : min(max((n_pairs * t0 + (t0 + t1) / 2u) / (t0 + t1), 1u), n_pairs - 1u);
    if (pair >= p0) {
      model = 1u;
      pair -= p0;
      n_pairs -= p0;
      n_pos = n_pos1;
    } else {
      n_pairs = p0;
    }
  }
  const TowerModel& gm = g.m[model];
  const uint32_t rows = n_pos * kActPadded.pos_rows;
  const uint32_t unit_pos = tower_unit_pos(n_pos, n_pairs), tpu = unit_pos == 9u ? 2u : 1u;  // tiles per unit
  const uint32_t unit_rows = unit_pos * kActPadded.pos_rows;
  const uint32_t n_units = (n_pos + unit_pos - 1u) / unit_pos;
  const uint32_t iters = pair < n_units ? tpu * ((n_units - pair + n_pairs - 1u) / n_pairs) : 0u;  // tiles of this pair per layer
  // tile i of the pair: sub-tile i % tpu of its unit i / tpu
  auto row0_of = [&](uint32_t i) { return (pair + (i / tpu) * n_pairs) * unit_rows + (i % tpu) * kT2PairRows + rank * kTcTileM; };
  const bool dbg_on = g.dbg != nullptr && blockIdx.x == 0;
  const long long t_start = dbg_on ? clock64() : 0;
  // diagnostic timeline of CTA 0 (AZB200_TOWER_DEBUG=1): g.dbg[8 + layer * 8 + k], cycles since the CTA started
#define AZB_TW_MARK(k) do { if (dbg_on) g.dbg[8 + layer * 8 + (k)] = clock64() - t_start; } while (0)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT3Stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full(a), 1);
      mbar_init(bar_acc_empty(a), 16);
    }
    for (int kb = 0; kb < kTcKBlocks; ++kb) {
      mbar_init(bar_w(kb), 1);
      mbar_init(bar_w_peer(kb), 1);
      mbar_init(bar_w_free(kb), 1);
    }
    for (uint32_t r = 0; r < 4u; ++r) mbar_init(bar_st(r), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(model ? &map3 : &map0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(model ? &map4 : &map1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(model ? &map5 : &map2)) : "memory");
  }
  if (threadIdx.x < 16) asm volatile("st.shared.u32 [%0], %1;" ::"r"(done_w + 4u * threadIdx.x), "r"(0u) : "memory");
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // buffer rotation of the residual blocks: conv1 reads X writes Y; conv2 reads Y, adds X, writes Z; then X <-> Z
  auto in_of = [&](int layer) { const int b = layer >> 1; return (layer & 1) ? 1 : ((b & 1) ? 2 : 0); };
  auto out_of = [&](int layer) { const int b = layer >> 1; return (layer & 1) ? ((b & 1) ? 0 : 2) : 1; };
  auto res_of = [&](int layer) { const int b = layer >> 1; return (b & 1) ? 2 : 0; };  // conv2 only

  if (iters == 0) {
    // nothing to do for this pair
  } else if (warp < 8) {
    // ===== epilogue, 8 warps: TMEM lane quarter warp % 4, output channels 64 * (warp / 4) .. + 63 =====
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;
    const uint32_t row_in_tile = rank * kTcTileM + q * 32 + lane;
    for (int layer = 0; layer < g.n_layers; ++layer) {
      float* sb = (layer & 1) ? s_bias1 : s_bias;
      if (threadIdx.x < kNetC) sb[threadIdx.x] = gm.bias[layer * kNetC + threadIdx.x];
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the layer's bias is in place (and everyone has left layer - 1)
      const __nv_bfloat16* residual = (layer & 1) ? gm.act[res_of(layer)] : nullptr;
      __nv_bfloat16* out = gm.act[out_of(layer)];
      for (uint32_t ti = 0; ti < iters; ++ti) {
        const uint32_t G = static_cast<uint32_t>(layer) * iters + ti, a = G & 1u;
        const uint32_t m = row0_of(ti) + q * 32 + lane;        // padded row
        const uint32_t rem = m % kActPadded.pos_rows;
        const bool kept = (ti % tpu) * kT2PairRows + row_in_tile < unit_rows;  // the rows past the unit belong to another unit
        const bool real = kept && m < rows && (rem & 7u) != 7u && rem < 48u;  // not the zero column, not the zero row
        uint4 res[8];
        if (real && residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(residual + static_cast<size_t>(m) * kNetC + half * 64);
#pragma unroll
          for (int jj = 0; jj < 8; jj += 2) ld_global_256(rp + jj, res[jj], res[jj + 1]);
        }
        mbar_wait(bar_acc_full(a), (G >> 1) & 1u);
        if (ti + 1 == iters && threadIdx.x == 0) AZB_TW_MARK(5);  // last accumulator complete
        tc_fence_after();
        uint32_t acc0[32], acc1[32];
        tmem_ld32x2(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 128u + half * 64u, acc0, acc1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar_acc_empty(a), 0u);  // one arrival per warp: the accumulator is in registers
        if (real) {
          uint4* op = reinterpret_cast<uint4*>(out + static_cast<size_t>(m) * kNetC + half * 64);
          uint4 prev = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t* acc = c8 < 4 ? acc0 : acc1;
            const float4 b0 = *reinterpret_cast<const float4*>(sb + half * 64 + c8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sb + half * 64 + c8 * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const uint32_t rw[4] = {res[c8].x, res[c8].y, res[c8].z, res[c8].w};
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0 = __uint_as_float(acc[(c8 & 3) * 8 + 2 * e]) + bb[2 * e];
              float x1 = __uint_as_float(acc[(c8 & 3) * 8 + 2 * e + 1]) + bb[2 * e + 1];
              if (residual) {
                x0 += __uint_as_float(rw[e] << 16);
                x1 += __uint_as_float(rw[e] & 0xFFFF0000u);
              }
              const __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f));
              pk[e] = *reinterpret_cast<const uint32_t*>(&p2);
            }
            if (c8 & 1) st_global_256(op + c8 - 1, prev, make_uint4(pk[0], pk[1], pk[2], pk[3]));
            else prev = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        // My stores of this tile are issued: tell the publisher warp (a CTA-scope release; the GPU-scope fence that makes
        // them visible to the pair's TMA reads costs ~2 k cycles, MEMBAR.GPU + CCTL.IVALL, and is the publisher's to pay —
        // on the epilogue warps it made them, not the tensor pipe, the tile period).
        if (layer + 1 < g.n_layers) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_st(G));
        }
        if (ti + 1 == iters && threadIdx.x == 0) AZB_TW_MARK(6);  // last tile's stores issued
      }
    }
  } else if (warp == 8) {
    if (rank == 1) {
      if (lane == 0) {
        // relay: tell the leader's MMA warp, k-block by k-block and layer by layer, that this CTA's half of the weights is in
        for (int layer = 0; layer < g.n_layers; ++layer)
          for (int kb = 0; kb < kTcKBlocks; ++kb) {
            mbar_wait(bar_w(kb), layer & 1);
            mbar_arrive_cluster(bar_w_peer(kb), 0u);
          }
      }
    } else {
      // ===== MMA issuer of the pair (the whole warp runs the loop, one elected lane issues) =====
      const bool me = elect_one_sync();
      for (int layer = 0; layer < g.n_layers; ++layer) {
        for (uint32_t ti = 0; ti < iters; ++ti) {
          const uint32_t G = static_cast<uint32_t>(layer) * iters + ti, a = G & 1u;
          const int s = static_cast<int>(G & 1u);
          mbar_wait(bar_acc_empty(a), ((G >> 1) & 1u) ^ 1u);
          mbar_wait(bar_full(s), (G >> 1) & 1u);
          tc_fence_after();
          if (ti == 0 && me) AZB_TW_MARK(2);  // first activation tile in
          const uint64_t a0 = umma_desc_sw128(stage_a(s, 0)), b0 = umma_desc_sw128(w_tile(0));
          const uint32_t acc = tmem_base + a * 128u;
#pragma unroll
          for (int kb = 0; kb < kTcKBlocks; ++kb) {
            if (ti == 0) {  // first tile of the layer: k-block kb of its weights, both halves
              mbar_wait(bar_w(kb), layer & 1);
              mbar_wait(bar_w_peer(kb), layer & 1);
            }
            const int tap = kb >> 1, shift = kT3HaloRows + (tap / 3 - 1) * 8 + (tap % 3 - 1);  // rows
            const uint64_t ad = a0 + static_cast<uint64_t>(((kb & 1) * kT3HalfBytes + shift * 128) >> 4);
            const uint64_t bd = b0 + static_cast<uint64_t>((kb * kT2WTile) >> 4);
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k)
              if (me) umma2_bf16(acc, ad + 2u * k, bd + 2u * k, kIdescBf16M256N128, (kb | k) ? 1u : 0u);
            // the layer's last tile hands the weights back k-block by k-block: the next layer's stream in under it
            if (me && ti + 1 == iters && layer + 1 < g.n_layers) umma2_commit_multicast(bar_w_free(kb), 3u);
          }
          if (me) {
            umma2_commit_multicast(bar_empty(s), 3u);     // the stage may be refilled (both CTAs)
            umma2_commit_multicast(bar_acc_full(a), 3u);  // the accumulator is complete (both CTAs' epilogues)
            if (ti == 0) AZB_TW_MARK(3);          // first tile issued (its MMAs wait for the weights k-block by k-block)
            if (ti + 1 == iters) AZB_TW_MARK(4);  // last tile issued
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    if (lane == 0) {
      // ===== publisher: "this CTA's rows of every tile below G are in global memory", for the pair's two producers =====
      // The eight epilogue warps' stores happen-before their arrivals, the arrivals before this thread's wait, and the
      // cluster-scope fence is cumulative over all of it; the flag stores follow the fence.
      const uint32_t own_flag = map_to_cta(done_w + 4u * rank, rank), peer_flag = map_to_cta(done_w + 4u * rank, rank ^ 1u);
      for (int layer = 0; layer + 1 < g.n_layers; ++layer)
        for (uint32_t ti = 0; ti < iters; ++ti) {
          const uint32_t G = static_cast<uint32_t>(layer) * iters + ti;
          mbar_wait(bar_st(G), (G >> 2) & 1u);
          fence_acq_rel_cluster();
          st_relaxed_cluster(own_flag, G + 1u);
          st_relaxed_cluster(peer_flag, G + 1u);
          if (ti + 1 == iters) AZB_TW_MARK(7);  // (diagnostic: the layer's last tile published)
        }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (lane == 0) {
      // ===== weights of every layer and the activation tiles =====
      // Layer 0 as k_conv3x3_tc3 (six k-blocks, the grid dependency, the first tile, the other twelve).  From then on the
      // first tile of a layer depends only on the same tile of the layer before, so it is requested while the previous
      // layer's last tile is still in the tensor pipe, and the weights follow k-block by k-block as that tile lets go of
      // them (with one tile per pair the weights go first: the tile has to wait for its own epilogue anyway).
      constexpr int kWFirst = 6;
      uint32_t seen = 0u;  // every epilogue warp of the pair has published the tiles below this global number
      for (int layer = 0; layer < g.n_layers; ++layer) {
        const uint8_t* wl = gm.w_tiles + static_cast<size_t>(layer) * kTcKBlocks * kTcTileBytes;
        auto fetch_w = [&](int kb) {
          if (layer > 0) mbar_wait(bar_w_free(kb), (layer - 1) & 1);  // the previous layer's last MMA on this k-block has retired
          mbar_arrive_expect_tx(bar_w(kb), kT2WTile);
          tma_bulk_g2s(w_tile(kb), wl + static_cast<size_t>(kb) * kTcTileBytes + rank * kT2WTile, kT2WTile, bar_w(kb));
        };
        const int in_idx = in_of(layer);
        const CUtensorMap* tm = model ? (in_idx == 0 ? &map3 : (in_idx == 1 ? &map4 : &map5)) : (in_idx == 0 ? &map0 : (in_idx == 1 ? &map1 : &map2));
        auto fetch_tile = [&](uint32_t i) {
          const uint32_t G = static_cast<uint32_t>(layer) * iters + i;
          const int s = static_cast<int>(G & 1u);
          if (layer > 0) {  // what tile i reads of the previous layer is in global memory: both CTAs' publishers said so
            // (a one-tile unit: the same tile; a two-tile unit: both tiles of the unit — each holds 16 rows the other needs)
            const uint32_t need = G - iters + 1u + (tpu == 2u && (i & 1u) == 0u ? 1u : 0u);
            if (seen < need) {
              do {
                seen = min2_volatile_shared(done_w);
              } while (seen < need);
              fence_acq_rel_cluster();     // what the pair's epilogue warps stored before the flags is visible to this thread ...
              fence_proxy_async_global();  // ... and to the TMA reads it issues from here on
            }
          }
          if (i == 0) AZB_TW_MARK(1);  // first tile of the layer may be fetched
          mbar_wait(bar_empty(s), ((G >> 1) & 1u) ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(bar_full(s), 2u * kT3StageBytes);  // both CTAs' copies land on this barrier
          const int r0 = static_cast<int>(row0_of(i)) - kT3HaloRows;  // negative for the first tile: zero-filled
          const uint32_t full = map_to_cta(bar_full(s), 0u);
          tma_tile2d_pair(stage_a(s, 0), tm, full, 0, r0);
          tma_tile2d_pair(stage_a(s, 1), tm, full, kTcBlockK, r0);
        };
        if (layer == 0) {
          AZB_TW_MARK(0);
          for (int kb = 0; kb < kWFirst; ++kb) fetch_w(kb);
          asm volatile("griddepcontrol.wait;" ::: "memory");  // the stem's output is complete and visible
          fetch_tile(0u);
          for (int kb = kWFirst; kb < kTcKBlocks; ++kb) fetch_w(kb);
        } else if (iters <= tpu) {  // the first tile waits for the previous layer's LAST tile: the weights go first
          for (int kb = 0; kb < kTcKBlocks; ++kb) fetch_w(kb);
          AZB_TW_MARK(0);  // all weights requested
          fetch_tile(0u);
        } else {
          fetch_tile(0u);
          for (int kb = 0; kb < kTcKBlocks; ++kb) fetch_w(kb);
          AZB_TW_MARK(0);  // all weights requested
        }
        for (uint32_t i = 1; i < iters; ++i) fetch_tile(i);
      }
    }
    __syncwarp();
  }
  if (dbg_on && threadIdx.x == 0) { g.dbg[0] = clock64() - t_start; g.dbg[1] = iters; }
#undef AZB_TW_MARK

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ================================================================================================
// Weight gradient of the 3x3 convolution on tcgen05 (building block of the training step):
//   dW[tap][ci][co] = sum over padded rows R of X[R + 8*dy + dx][ci] * dZ[R][co]
// In the padded layout the zero rows contribute nothing, so this is one plain product per tap,
// D[M = ci][N = co] += A[ci][K = row] * B[co][K = row], with BOTH operands MN-major: a stage half is
// [row][64 channels] = 128-byte rows, i.e. the canonical MN-major SWIZZLE_128B atom (64 elements along M/N x 8
// rows along K, 1024 B), the next 8 rows 1024 B further (SBO), the other 64 channels in the other half of the stage
// (LBO).  Per 128-row chunk a CTA loads X (with halo, as the forward pass) and dZ once (4 x 20 KB, 2-D TMA) and
// issues 8 MMAs (K = 16 rows each) per tap: M128 N128 K16, cta_group::1, the tap's accumulator in TMEM.
// TMEM holds 4 such accumulators, so the 9 taps are split over 3 CTA groups (3 taps = 384 columns each); CTA b
// works on taps 3*(b % 3) .. +2 and on the row chunks b / 3, b / 3 + grid / 3, ...; at the end every CTA adds its
// partial sums to dW in HBM with fp32 atomics.
// ================================================================================================
constexpr uint32_t kWgStageBytes = 4 * kT3HalfBytes;                     // X half 0/1, dZ half 0/1: 80 KB
constexpr int kWgStages = 2;
constexpr uint32_t kWgSmemBytes = kWgStages * kWgStageBytes + 1024 + 256;
constexpr int kWgThreads = 192;                                          // warps 0-3 epilogue, 4 MMA, 5 TMA producer + TMEM alloc
// A and B MN-major (bits 15, 16), D = F32, A = B = BF16, N = 128, M = 128
constexpr uint32_t kIdescBf16M128N128MN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// MN-major SWIZZLE_128B descriptor: LBO = distance between the two 64-channel halves, SBO = 8 rows = 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void tma_tile2d(uint32_t dst, const CUtensorMap* tmap, uint32_t mbar, int c, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar), "r"(c), "r"(row)
               : "memory");
}
struct WgradArgs {
  float* dw;            // [9][128 ci][128 co] fp32, accumulated into (zero it first)
  uint32_t n_pos;
};
__global__ void __launch_bounds__(kWgThreads, 1)
k_conv3x3_wgrad(WgradArgs g, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dz) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto stage_x = [&](int s, int half) { return base + s * kWgStageBytes + half * kT3HalfBytes; };
  auto stage_dz = [&](int s, int half) { return base + s * kWgStageBytes + (2 + half) * kT3HalfBytes; };
  const uint32_t bars = base + kWgStages * kWgStageBytes;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 16u + 8u * s; };
  const uint32_t bar_done = bars + 32u;
  const uint32_t tmem_slot = bars + 40u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tg = blockIdx.x % 3u, slice = blockIdx.x / 3u, n_slices = gridDim.x / 3u;
  const uint32_t rows = g.n_pos * kActPadded.pos_rows;
  const uint32_t n_chunks = (rows + kTcTileM - 1) / kTcTileM;
  const uint32_t iters = slice < n_chunks && blockIdx.x < 3u * n_slices ? (n_chunks - slice + n_slices - 1) / n_slices : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 5) {
    if (lane == 0) {
      for (uint32_t i = 0; i < iters; ++i) {
        const int s = i % kWgStages;
        mbar_wait(bar_empty(s), ((i / kWgStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_full(s), kWgStageBytes);
        const int r0 = static_cast<int>((slice + i * n_slices) * kTcTileM) - kT3HaloRows;
        tma_tile2d(stage_x(s, 0), &tmap_x, bar_full(s), 0, r0);
        tma_tile2d(stage_x(s, 1), &tmap_x, bar_full(s), kTcBlockK, r0);
        tma_tile2d(stage_dz(s, 0), &tmap_dz, bar_full(s), 0, r0);
        tma_tile2d(stage_dz(s, 1), &tmap_dz, bar_full(s), kTcBlockK, r0);
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    if (lane == 0) {
      for (uint32_t i = 0; i < iters; ++i) {
        const int s = i % kWgStages;
        mbar_wait(bar_full(s), (i / kWgStages) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int t3 = 0; t3 < 3; ++t3) {
          const int tap = static_cast<int>(tg) * 3 + t3, shift = kT3HaloRows + (tap / 3 - 1) * 8 + (tap % 3 - 1);
#pragma unroll
          for (int j = 0; j < kTcTileM / 16; ++j) {  // K = 16 rows per MMA
            const uint64_t ad = umma_desc_sw128_mn(stage_x(s, 0) + (shift + 16 * j) * 128, kT3HalfBytes);
            const uint64_t bd = umma_desc_sw128_mn(stage_dz(s, 0) + (kT3HaloRows + 16 * j) * 128, kT3HalfBytes);
            umma_bf16(tmem_base + t3 * 128u, ad, bd, kIdescBf16M128N128MN, (i | static_cast<uint32_t>(j)) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty(s));
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane = ci (warp q: lanes 32q..), columns = co: add this CTA's partial sums to dW =====
    if (iters > 0) {
      mbar_wait(bar_done, 0u);
      tc_fence_after();
      const int q = warp;
      const uint32_t ci = q * 32 + lane;
      for (int t3 = 0; t3 < 3; ++t3) {
        float* dst = g.dw + (static_cast<size_t>(tg * 3 + t3) * kNetC + ci) * kNetC;
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t3 * 128u + ch * 32u, acc);
#pragma unroll
          for (int c = 0; c < 32; ++c) atomicAdd(dst + ch * 32 + c, __uint_as_float(acc[c]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// fp32 dense [pos][42][128] <-> bf16 padded [pos][56][128] (test hooks and, later, the training step's inputs)
__global__ void k_dense_f32_to_padded_bf16(const float* __restrict__ in, uint32_t n_pos, __nv_bfloat16* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n_pos) * kCells * kNetC) return;
  const uint32_t c = i % kNetC, m = static_cast<uint32_t>(i / kNetC), pos = m / kCells, cell = m % kCells;
  out[ActLayout{56u, 8u}.row(pos, cell / 7, cell % 7) * kNetC + c] = __float2bfloat16_rn(in[i]);
}
__global__ void k_padded_bf16_to_dense_f32(const __nv_bfloat16* __restrict__ in, uint32_t n_pos, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n_pos) * kCells * kNetC) return;
  const uint32_t c = i % kNetC, m = static_cast<uint32_t>(i / kNetC), pos = m / kCells, cell = m % kCells;
  out[i] = __bfloat162float(in[ActLayout{56u, 8u}.row(pos, cell / 7, cell % 7) * kNetC + c]);
}

// Stem: conv3x3(2 -> 128) + ReLU from the bitboard planes, bf16 out.  The two input planes are
// binary, so a window row (3 cells x 2 planes = 6 bits) selects one of 64 precomputed partial sums:
// out = ReLU(bias + T[0][bits of row y-1] + T[1][bits of row y] + T[2][bits of row y+1]).  Each
// persistent CTA copies the 3 x 64 x 128 fp32 table (96 KB, built from the 18 x 128 stem weights when
// the parameters are uploaded) into shared memory, then every thread produces 8 channels of one
// (position, cell) row per step: HBM-write bound (256 B per row).
constexpr uint32_t kStemSmemBytes = 3 * 64 * kNetC * 4;
// Host side of the table (azb_nnet::upload): entry [dyi][cur3 + 8*opp3][c] = its set taps added in (dx, plane) order.
inline void stem_table_build(const float* prm, const NetLayout& L, float* tab /* [3][64][128] */) {
  for (uint32_t e = 0; e < 3u * 64u * kNetC; ++e) {
    const uint32_t c = e % kNetC, bits = (e / kNetC) % 64u, dyi = e / (kNetC * 64u);
    float acc = 0.0f;
    for (uint32_t dxi = 0; dxi < 3u; ++dxi)
      for (uint32_t pl = 0; pl < 2u; ++pl)
        if ((bits >> (dxi + 3u * pl)) & 1u) acc += prm[L.stem_w + ((dyi * 3u + dxi) * 2u + pl) * kNetC + c];
    tab[e] = acc;
  }
}
__global__ void __launch_bounds__(1024)
k_stem_bf16(const float* __restrict__ prm, NetLayout L, const float* __restrict__ tab, const uint4* __restrict__ states,
            const uint32_t* __restrict__ count, uint32_t max_batch, __nv_bfloat16* __restrict__ out, ActLayout lay) {
  extern __shared__ float stem_tab[];  // [3][64][128], copied from the table built at upload
  const uint32_t n_pos = count ? min(*count, max_batch) : max_batch;
  const size_t total = static_cast<size_t>(n_pos) * kCells * (kNetC / 8);
  if (static_cast<size_t>(blockIdx.x) * blockDim.x >= total) return;  // (the grid is sized for max_batch) no work: no 96-KB copy
  for (uint32_t e = threadIdx.x; e < 3u * 64u * kNetC / 4u; e += blockDim.x)
    reinterpret_cast<float4*>(stem_tab)[e] = reinterpret_cast<const float4*>(tab)[e];
  __syncthreads();
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t cg = idx % (kNetC / 8);
    const uint32_t m = static_cast<uint32_t>(idx / (kNetC / 8)), pos = m / kCells, cell = m % kCells;
    const int r = cell / 7, c = cell % 7;
    const uint4 st = states[pos];
    const uint64_t cur = (static_cast<uint64_t>(st.y) << 32) | st.x, opp = (static_cast<uint64_t>(st.w) << 32) | st.z;
    const float4 b0 = *reinterpret_cast<const float4*>(prm + L.stem_b + cg * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(prm + L.stem_b + cg * 8 + 4);
    float acc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int dyi = 0; dyi < 3; ++dyi) {
      const int rr = r + dyi - 1;
      if (rr < 0 || rr >= 6) continue;  // a window row off the board contributes nothing
      // bits c-1, c, c+1 of the board row (zero beyond the edges) for both planes: index = cur3 + 8 * opp3
      const uint32_t cur3 = ((static_cast<uint32_t>(cur >> (rr * 7)) & 0x7Fu) << 1 >> c) & 7u;
      const uint32_t opp3 = ((static_cast<uint32_t>(opp >> (rr * 7)) & 0x7Fu) << 1 >> c) & 7u;
      const float* t = stem_tab + (static_cast<uint32_t>(dyi) * 64u + cur3 + 8u * opp3) * kNetC + cg * 8;
      const float4 t0 = *reinterpret_cast<const float4*>(t), t1 = *reinterpret_cast<const float4*>(t + 4);
      acc[0] += t0.x; acc[1] += t0.y; acc[2] += t0.z; acc[3] += t0.w; acc[4] += t1.x; acc[5] += t1.y; acc[6] += t1.z; acc[7] += t1.w;
    }
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(acc[2 * k], 0.0f), fmaxf(acc[2 * k + 1], 0.0f));
      pk[k] = *reinterpret_cast<const uint32_t*>(&p2);
    }
    *reinterpret_cast<uint4*>(out + lay.row(pos, r, c) * kNetC + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Heads from the bf16 tower output.  A CTA of 256 threads takes 6 positions per step: thread = one
// (position, cell) row computes the three 1x1-convolution outputs of its cell (policy planes 0/1 and
// the value plane) straight from its 256 bytes of bf16 activations, the 3 x 128 weights coming from
// the kernel-parameter constant bank; the small FC layers, softmax and tanh follow from shared memory.
// Every output is accumulated in exactly the order of heads_from_smem (the fp32 path's head), so both
// paths produce the same head arithmetic on the same tower values.
struct HeadConvW {
  float w[kNetC][3];  // [in channel][policy plane 0, policy plane 1, value plane]
  float b[3];
};
constexpr int kHeadPos = 6;  // positions per CTA step: 6 x 42 = 252 rows on 256 threads
__global__ void __launch_bounds__(256)
k_heads_bf16(const float* __restrict__ prm, NetLayout L, const __grid_constant__ HeadConvW hw,
             const __nv_bfloat16* __restrict__ act, const uint32_t* __restrict__ count, uint32_t max_batch,
             float* __restrict__ pi_out, float* __restrict__ v_out, ActLayout lay) {
  __shared__ float pol[kHeadPos][84];   // plane*42 + cell
  __shared__ float val[kHeadPos][42];
  __shared__ float h1[kHeadPos][64];
  __shared__ float logit[kHeadPos][8];
  asm volatile("griddepcontrol.wait;" ::: "memory");  // (launched with programmatic stream serialization)
  const uint32_t n_pos = count ? min(*count, max_batch) : max_batch;
  const int tid = threadIdx.x;
  for (uint32_t p0 = blockIdx.x * kHeadPos; p0 < n_pos; p0 += gridDim.x * kHeadPos) {
    __syncthreads();
    const int lp = tid / kCells, cell = tid % kCells;
    if (tid < kHeadPos * kCells && p0 + lp < n_pos) {
      const uint4* src = reinterpret_cast<const uint4*>(act + lay.row(p0 + lp, cell / 7, cell % 7) * kNetC);
      float s0 = hw.b[0], s1 = hw.b[1], s2 = hw.b[2];
#pragma unroll
      for (int v8 = 0; v8 < kNetC / 8; ++v8) {
        const uint4 x = src[v8];
        const uint32_t wd[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = __uint_as_float(wd[k] << 16), hi = __uint_as_float(wd[k] & 0xFFFF0000u);
          const int ci = v8 * 8 + 2 * k;
          s0 = fmaf(lo, hw.w[ci][0], s0); s1 = fmaf(lo, hw.w[ci][1], s1); s2 = fmaf(lo, hw.w[ci][2], s2);
          s0 = fmaf(hi, hw.w[ci + 1][0], s0); s1 = fmaf(hi, hw.w[ci + 1][1], s1); s2 = fmaf(hi, hw.w[ci + 1][2], s2);
        }
      }
      pol[lp][cell] = fmaxf(s0, 0.0f);
      pol[lp][42 + cell] = fmaxf(s1, 0.0f);
      val[lp][cell] = fmaxf(s2, 0.0f);
    }
    __syncthreads();
    // FC(84 -> 7) of the policy head: 6 x 7 outputs; FC(42 -> 64) of the value head: 6 x 64 outputs
    for (int o = tid; o < kHeadPos * (7 + 64); o += 256) {
      const int q = o / (7 + 64), j = o % (7 + 64);
      if (p0 + q >= n_pos) continue;
      if (j < 7) {
        float s = prm[L.pol_fc_b + j];
        for (int i = 0; i < 84; ++i) s = fmaf(pol[q][i], prm[L.pol_fc_w + i * 7 + j], s);
        logit[q][j] = s;
      } else {
        const int jj = j - 7;
        float s = prm[L.val_fc1_b + jj];
        for (int i = 0; i < 42; ++i) s = fmaf(val[q][i], prm[L.val_fc1_w + i * 64 + jj], s);
        h1[q][jj] = fmaxf(s, 0.0f);
      }
    }
    __syncthreads();
    if (tid < 2 * kHeadPos) {
      const int q = tid >> 1;
      if (p0 + q < n_pos) {
        if (tid & 1) {
          float s = prm[L.val_fc2_b];
          for (int j = 0; j < 64; ++j) s = fmaf(h1[q][j], prm[L.val_fc2_w + j], s);
          v_out[p0 + q] = tanhf(s);
        } else {
          float m = logit[q][0];
          for (int a = 1; a < 7; ++a) m = fmaxf(m, logit[q][a]);
          float e[7], sum = 0.0f;
          for (int a = 0; a < 7; ++a) { e[a] = expf(logit[q][a] - m); sum += e[a]; }
          float* po = pi_out + static_cast<size_t>(p0 + q) * 8u;
          for (int a = 0; a < 7; ++a) po[a] = e[a] / sum;
          po[7] = 0.0f;
        }
      }
    }
  }
}

}  // namespace azb
