// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = A[M,K] * W[N,K]^T   (fp16 in, fp32 accumulate)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2-D, 128B-swizzled tiles, STAGES-deep mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, kind::f16)
//   warps 2..9  : epilogue.  Warp w owns TMEM lanes [32*(w%4), +32) = 32 accumulator rows (thread = row) and every
//                 other 64-column block of the tile (two warps share a lane quarter).  Each 32-column chunk is pulled
//                 with tcgen05.ld 32x32b (double-buffered in registers) and handed to the fused epilogue functor, which
//                 stages its 128-byte-per-row output slab in warp-private shared memory (128B swizzle) and ships it
//                 with a TMA store (and, for the residual variants, brings the residual in with a TMA load).
//
// Two TMEM accumulator stages (2 x BLOCK_N fp32 columns) let the epilogue of tile i overlap the MMAs of tile
// i+1.  Tiles are statically strided over the persistent grid (tile = blockIdx.x + k * gridDim.x, N fastest so
// concurrently running CTAs share the same A rows in L2).  A is [M,K] row-major (K contiguous), W is the
// torch nn.Linear layout [N,K] row-major -- both "K-major" UMMA operands, no transposes anywhere.
// K tails / M tails / N tails rely on TMA out-of-bounds zero fill (loads) and clipping (stores).
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 fp16 = 128 B = one swizzle row
#ifndef B200_GEMM_EPI_WARPS
#define B200_GEMM_EPI_WARPS 8
#endif
// Epilogue warps per CTA: 4 (one per TMEM lane quarter) or 8 (two per quarter, each owning every other 64-column
// block).  Measured on B200: 8 warps do not shorten the layer GEMMs (they are bound by operand bytes in flight, i.e.
// by pipeline depth x L2 latency), while 4 warps leave shared memory for one more pipeline stage.
constexpr int GEMM_EPI_WARPS = B200_GEMM_EPI_WARPS;
constexpr int GEMM_EPI_PARTS = GEMM_EPI_WARPS / 4;
static_assert(GEMM_EPI_WARPS == 4 || GEMM_EPI_WARPS == 8, "4 or 8 epilogue warps");
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_BAR_BYTES = 1024;
constexpr int GEMM_BIAS_BYTES = 8192;   // per-column epilogue vector (bias) of the whole GEMM, staged once per CTA: N <= 2048

template <int BLOCK_N, class Epi>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = GEMM_EPI_WARPS * Epi::SMEM_PER_WARP;  // SMEM_PER_WARP is a multiple of 1024
  static constexpr int budget = 227 * 1024 - 1024 /*alignment slack*/ - EPI_BYTES - GEMM_BIAS_BYTES - GEMM_BAR_BYTES;
  static constexpr int STAGES = (budget / STAGE_BYTES) > 6 ? 6 : (budget / STAGE_BYTES);
  static constexpr int TOTAL = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + GEMM_BIAS_BYTES + GEMM_BAR_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(Epi::SMEM_PER_WARP % 1024 == 0, "epilogue slabs must keep 1024-byte alignment (128B swizzle)");
};

constexpr uint32_t tmem_cols_pow2(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

// Per-warp epilogue context handed to the functor (lives in registers for the whole persistent loop).
struct EpiCtx {
  long long* trace;     // debug: per-chunk clock64 stamps (nullptr in production)
  uint8_t* smem;        // warp-private slab, Epi::SMEM_PER_WARP bytes, 1024-byte aligned
  uint64_t* bars;       // 4 warp-private mbarriers
  const CUtensorMap* map_c;
  const float* bias_all; // CTA-shared copy of the functor's per-column vector for columns [0, N)
  int lane;
  int M, N;
  int col_base;         // first column of the tile in flight
  int col_end;          // first column after the tile in flight (col_base + BLOCK_N)
  uint32_t seq;         // running chunk / block counter (buffer rotation + mbarrier parity), functor-defined
  bool primed;          // functor-defined: a prefetch for the chunk about to be processed is already in flight
};

// One tile's epilogue for one warp: 64-column blocks c = 64*part, 64*part + 64*PARTS, ... of the BLOCK_N accumulator
// columns, each pulled with one tcgen05.ld and handed to the functor as two 32-column chunks; `release` (accumulator
// free) runs as soon as this warp's last TMEM read has landed.
template <int BLOCK_N, class Epi, int PARTS = GEMM_EPI_PARTS, class Release>
__device__ __forceinline__ void epilogue_tile(EpiCtx& ctx, const typename Epi::Params& ep, uint32_t taddr, int row0,
                                              int col_base, int part, uint64_t* acc_full_bar, uint32_t acc_phase,
                                              Release release) {
  const bool live = row0 < ctx.M;  // warp-uniform: this warp's 32 rows exist
  ctx.col_base = col_base;
  ctx.col_end = col_base + BLOCK_N;
  if (live) Epi::tile_begin(ctx, ep, row0, col_base);
  mbar_wait(acc_full_bar, acc_phase);
  tc_fence_after();
  auto done_reading = [&]() {
    tc_fence_before();
    __syncwarp();
    if (ctx.lane == 0) release();
  };
  int tr_i = 0;
  auto run = [&](uint32_t (&raw)[32], int c, int cn) {
    if (ctx.trace && ctx.lane == 0) ctx.trace[tr_i++] = clock64();
    if (live && col_base + c < ctx.N)
      Epi::chunk(ctx, ep, raw, row0, col_base + c, (cn < BLOCK_N && col_base + cn < ctx.N) ? col_base + cn : -1);
    if (ctx.trace && ctx.lane == 0) ctx.trace[tr_i++] = clock64();
  };
  // One tcgen05.ld per 64-column block of this warp (two 32-column chunks for the functor): the epilogue of tile i runs
  // while the MMAs of tile i+1 own the TMEM port, which arbitrates per instruction -- fewer, wider loads wait less.
  uint32_t raw[64];
#pragma unroll 1
  for (int c = part * 64; c < BLOCK_N; c += 64 * PARTS) {
    const bool two = c + 32 < BLOCK_N;                       // (BLOCK_N = 96: the last block is a single chunk)
    const int cnext = c + 64 * PARTS;
    if (two) tmem_ld_32x64(taddr + c, raw);
    else tmem_ld_32x32(taddr + c, *reinterpret_cast<uint32_t(*)[32]>(raw));
    tmem_ld_wait();
    if (cnext >= BLOCK_N) done_reading();
    run(*reinterpret_cast<uint32_t(*)[32]>(raw), c, two ? c + 32 : cnext);
    if (two) run(*reinterpret_cast<uint32_t(*)[32]>(raw + 32), c + 32, cnext);
  }
  if (part * 64 >= BLOCK_N) done_reading();
  if (live) Epi::tile_end(ctx, ep, row0, col_base, taddr);
}

// Epi interface (all static, called by every lane of an epilogue warp, warp-uniform arguments):
//   SMEM_PER_WARP                                   bytes of warp-private shared memory
//   tile_begin(ctx, p, row0, col_base)              before waiting for the accumulator (prefetch residuals here)
//   chunk(ctx, p, v, row0, col0, next_col0)         v[32] = accumulator row (row0+lane), columns [col0, col0+32);
//                                                   next_col0 = first column of this warp's next chunk in the tile, or -1
//   tile_end(ctx, p, row0, col_base)                after the last chunk
//   finish(ctx)                                     once, before the CTA exits (drain async stores)
template <int BLOCK_N, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, int M, int N, int K,
                 const __grid_constant__ typename Epi::Params ep) {
  using SM = GemmSmem<BLOCK_N, Epi>;
  constexpr int STAGES = SM::STAGES;
  constexpr uint32_t ACC_STRIDE = (BLOCK_N <= 128) ? 128 : 256;      // TMEM columns between the two accumulators
  constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * ACC_STRIDE);
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N constraint (M=128)");
  static_assert(BLOCK_N % 32 == 0, "epilogue works on 32-column chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* epi_smem = smem + STAGES * SM::STAGE_BYTES;
  float* bias_all = reinterpret_cast<float*>(epi_smem + SM::EPI_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + SM::EPI_BYTES + GEMM_BIAS_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]
  uint64_t* epi_bars = bars + 2 * STAGES + 4;   // [GEMM_EPI_WARPS][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bars + GEMM_EPI_WARPS * 4);
  static_assert((2 * 6 + 4 + GEMM_EPI_WARPS * 4) * 8 + 8 <= GEMM_BAR_BYTES, "barrier area too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  pdl_launch_dependents();
  Epi::preload(ep, bias_all, N, threadIdx.x, blockDim.x);   // visible to the epilogue warps after the barrier below
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], GEMM_EPI_WARPS);
    }
    for (int s = 0; s < GEMM_EPI_WARPS * 4; ++s) mbar_init(&epi_bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * SM::STAGE_BYTES;
          uint8_t* sb = sa + SM::A_BYTES;
          mbar_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
          tma_load_2d(sb, &map_b, &full_bar[stage], kb * GEMM_BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * SM::STAGE_BYTES);
          const uint32_t sb = sa + SM::A_BYTES;
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;  // which 64-column blocks of the tile this warp owns
    EpiCtx ctx;
    ctx.smem = epi_smem + (warp - 2) * Epi::SMEM_PER_WARP;
    ctx.bars = epi_bars + (warp - 2) * 4;
    ctx.map_c = &map_c;
    ctx.bias_all = bias_all;
    ctx.lane = lane;
    ctx.M = M;
    ctx.N = N;
    ctx.seq = 0;
    ctx.primed = false;
    ctx.trace = nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * GEMM_BLOCK_M + q * 32;
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_tile<BLOCK_N, Epi>(ctx, ep, taddr, row0, n_blk * BLOCK_N, part, &acc_full[as], aphase,
                                  [&]() { mbar_arrive(&acc_empty[as]); });
    }
    Epi::finish(ctx);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace b200
