// tcgen05 implicit-GEMM kernel for the temporal convolutions.  See conv_gemm.cuh for the math.
//
// CTA = 384 threads, persistent over output tiles (128 rows x BLOCK_N channels):
//   warp 0 lane 0 : TMA producer  (A tile 128x64 bf16 + W tile BLOCK_Nx64 bf16 per k-block)
//   warp 1 lane 0 : tcgen05.mma issuer (4 x K=16 MMAs per k-block, accumulator in TMEM)
//   warp 2        : TMEM allocator / deallocator
//   warp 3 lane 0 : auxiliary producer (RES variant): TMA-loads the 128x64 residual tile(s) / the Z
//                   tile of every 64-column store block into shared memory ahead of the epilogue
//   warps 4..11   : epilogue, two groups of four warps (each group covers the 128 TMEM lanes); the
//                   64-column store blocks of a tile alternate between the groups, so two epilogue
//                   warps share every SM sub-partition and hide each other's latencies:
//                   tcgen05.ld -> BN affine / ReLU / residual / batch sums in registers -> bf16 pack
//                   into the group's SWIZZLE_128B staging tile(s) -> one TMA store per warp of its
//                   32 x 64 slice (coalesced 128-byte rows, clipped at the tensor edge by the map)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// PAIR variant (BLOCK_N = 256): the grid is launched in clusters of two CTAs that compute a
// 256-row x 256-column tile together with one tcgen05.mma.cta_group::2 stream issued by the
// cluster's rank-0 CTA ("leader").  Each CTA TMA-loads its own 128-row A tile and HALF of the W tile
// (128 of the 256 output channels) -- 16 + 16 KiB per k-block instead of 16 + 32 -- both crediting
// the leader's `full` barrier; the leader's tcgen05.commit multicasts to both CTAs' `empty` /
// accumulator-full barriers; the accumulator rows of each CTA live in its own TMEM and go through
// its own epilogue, which releases the accumulator stage on the leader's barrier.  Per SM this cuts
// the operand bytes per MMA by a third and deepens the operand pipeline (6 / 4 stages instead of
// 4 / 3), which is what the operand-supply-bound 1x1 layers need.
#include "conv_gemm.cuh"

#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"

namespace vp3d {

// WRES ("W resident"): layers whose whole weight slab for one N block fits in 128 KiB of shared
// memory (the expand conv: K = 128) load it once per CTA and stream only the A tiles; the grid is a
// multiple of the number of N blocks so that a CTA keeps its N block for all of its tiles.
// OUT2: two output planes (hi, lo) -> each epilogue group needs two staging tiles; the extra 32 KiB
// come out of the operand pipeline.
// LEAN (inference layers: affine + ReLU [+ one-plane TMA residual] -> one 16-bit plane): a dedicated
// epilogue with every option resolved at compile time.  With a residual it works IN PLACE: the
// result overwrites the residual tile it has just consumed and the bulk store reads from there, so
// no staging tiles are needed and the operand pipeline gets their 32 KiB (5 instead of 4 stages
// on CTA pairs).
template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool PAIR = false, bool LEAN = false>
struct GemmCfg {
  static_assert(!(RES && WRES), "W-resident variant has no residual path");
  static_assert(!(LEAN && OUT2), "the lean epilogue writes one plane");
  static_assert(!PAIR || (BLOCK_N == 256 && !WRES), "CTA pairs run 256-wide, streamed-W tiles");
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = (PAIR ? BLOCK_N / 2 : BLOCK_N) * kBlockK * 2;
  static constexpr uint32_t kStageBytes = WRES ? kABytes : kABytes + kBBytes;
  static constexpr uint32_t kWResBytes = WRES ? 128u * 1024u : 0u;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
  static constexpr uint32_t kTileBytes = kBlockM * 64 * 2;  // one 128 x 64 bf16 tile (16 KiB)
  // one (hi[, lo]) set per epilogue group; none when the result is staged in the residual tile
  static constexpr int kStoreTiles = (LEAN && RES) ? 0 : (OUT2 ? 4 : 2);
  // auxiliary (residual / Z) landing tiles: 3 next to 128x256 tiles, 4 next to narrower ones, so
  // that two-tile store blocks (hi+lo residual, or residual + Z) still get two stages in flight
  // (with two output planes the four staging tiles already take 64 KiB: one auxiliary stage only,
  // the second epilogue group covers the exposed load latency, and the operand pipeline keeps its
  // depth)
  // (CTA pairs stream half the W bytes per stage: a fourth landing tile fits next to 4 stages)
  // (in place: a slot stays busy until its store has drained -- four slots, two per group)
  static constexpr int kResSlots = RES ? ((BLOCK_N == 256 && !PAIR && !LEAN) ? 3 : 4) : 0;
  static constexpr uint32_t kFixedBytes = kWResBytes + (kStoreTiles + kResSlots) * kTileBytes;
  // per-channel affine (scale, shift) of the current N block: 2 x BLOCK_N floats
  static constexpr uint32_t kAffineBytes = 2 * BLOCK_N * 4;
  static constexpr uint32_t kBarBytesMax = (2 * 8 + 4 + 8 + 1) * 8 + 32;
  // as many operand stages as fit into the 227 KiB a CTA may use, at most 8
  static constexpr uint32_t kMaxSmem = 232448u;
  static constexpr int kStagesFit =
      (kMaxSmem - kFixedBytes - kBarBytesMax - kAffineBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static_assert(kStages >= 2, "not enough shared memory for the operand pipeline");
  static constexpr uint32_t kBarBytes = (2 * kStages + 4 + 8 + 1) * 8 + 32;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kFixedBytes + kBarBytes + kAffineBytes;
  static_assert(kSmemBytes <= kMaxSmem, "shared memory budget exceeded");
};

__device__ __forceinline__ void tile_coords(const ConvGemmArgs& p, int tile, int& n_blk,
                                            int& sample, int& row0) {
  n_blk = tile % p.n_tiles;
  int m_blk = tile / p.n_tiles;
  if (p.dilated) {
    sample = m_blk / p.tiles_per_sample;
    row0 = (m_blk - sample * p.tiles_per_sample) * kBlockM;
  } else {
    sample = 0;
    row0 = m_blk * kBlockM;
  }
}

__device__ __forceinline__ void group_bar_sync(uint32_t id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");  // the four warps of one epilogue group
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

__device__ __forceinline__ void add_f16x8(float* v, const uint4& u) {
  v[0] += f16_lo_to_f(u.x);
  v[1] += f16_hi_to_f(u.x);
  v[2] += f16_lo_to_f(u.y);
  v[3] += f16_hi_to_f(u.y);
  v[4] += f16_lo_to_f(u.z);
  v[5] += f16_hi_to_f(u.z);
  v[6] += f16_lo_to_f(u.w);
  v[7] += f16_hi_to_f(u.w);
}

__device__ __forceinline__ void add_bf16x8(float* v, const uint4& u) {
  v[0] += bf16_lo_to_f(u.x);
  v[1] += bf16_hi_to_f(u.x);
  v[2] += bf16_lo_to_f(u.y);
  v[3] += bf16_hi_to_f(u.y);
  v[4] += bf16_lo_to_f(u.z);
  v[5] += bf16_hi_to_f(u.z);
  v[6] += bf16_lo_to_f(u.w);
  v[7] += bf16_hi_to_f(u.w);
}

// Lean inference epilogue, 32 accumulator columns of one row:
//   plain:    affine (paired FMAs) -> 16-bit pairs -> ReLU on the pairs
//   residual: affine -> ReLU -> + residual (32 16-bit values in rres[0..3]) -> 16-bit pairs
template <bool F16>
__device__ __forceinline__ void lean_chunk(const uint32_t* a, const float* s_scale,
                                           const float* s_shift, uint32_t* hi) {
  float4 sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = *reinterpret_cast<const float4*>(s_scale + 4 * j);
    sh[j] = *reinterpret_cast<const float4*>(s_shift + 4 * j);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float r0, r1, r2, r3;
    ffma2(r0, r1, __uint_as_float(a[4 * j + 0]), __uint_as_float(a[4 * j + 1]), sc[j].x, sc[j].y,
          sh[j].x, sh[j].y);
    ffma2(r2, r3, __uint_as_float(a[4 * j + 2]), __uint_as_float(a[4 * j + 3]), sc[j].z, sc[j].w,
          sh[j].z, sh[j].w);
    if (F16) {
      hi[2 * j] = relu_f16x2(pack_f16x2(r0, r1));
      hi[2 * j + 1] = relu_f16x2(pack_f16x2(r2, r3));
    } else {
      hi[2 * j] = relu_bf16x2(pack_bf16x2(r0, r1));
      hi[2 * j + 1] = relu_bf16x2(pack_bf16x2(r2, r3));
    }
  }
}

template <bool F16>
__device__ __forceinline__ void lean_chunk_res(const uint32_t* a, const float* s_scale,
                                               const float* s_shift, const uint4* rres,
                                               uint32_t* hi) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * j);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + 4 * j);
    float r0, r1, r2, r3;
    ffma2(r0, r1, __uint_as_float(a[4 * j + 0]), __uint_as_float(a[4 * j + 1]), sc.x, sc.y, sh.x,
          sh.y);
    ffma2(r2, r3, __uint_as_float(a[4 * j + 2]), __uint_as_float(a[4 * j + 3]), sc.z, sc.w, sh.z,
          sh.w);
    const uint4 u4 = rres[j >> 1];
    const uint32_t u0 = (j & 1) ? u4.z : u4.x, u1 = (j & 1) ? u4.w : u4.y;
    if (F16) {
      r0 = fmaxf(r0, 0.0f) + f16_lo_to_f(u0);
      r1 = fmaxf(r1, 0.0f) + f16_hi_to_f(u0);
      r2 = fmaxf(r2, 0.0f) + f16_lo_to_f(u1);
      r3 = fmaxf(r3, 0.0f) + f16_hi_to_f(u1);
      hi[2 * j] = pack_f16x2(r0, r1);
      hi[2 * j + 1] = pack_f16x2(r2, r3);
    } else {
      r0 = fmaxf(r0, 0.0f) + bf16_lo_to_f(u0);
      r1 = fmaxf(r1, 0.0f) + bf16_hi_to_f(u0);
      r2 = fmaxf(r2, 0.0f) + bf16_lo_to_f(u1);
      r3 = fmaxf(r3, 0.0f) + bf16_hi_to_f(u1);
      hi[2 * j] = pack_bf16x2(r0, r1);
      hi[2 * j + 1] = pack_bf16x2(r2, r3);
    }
  }
}

#ifdef VP3D_TIMELINE
__device__ __forceinline__ void tl_stamp(const ConvGemmArgs& p, int ev) {
  if (!p.timeline) return;
  int slot;
  if (blockIdx.x == 0) slot = 0;
  else if (blockIdx.x == gridDim.x - 1) slot = 1;
  else return;
  unsigned long long g;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
  p.timeline[(slot * 32 + ev) * 2] = g;
  p.timeline[(slot * 32 + ev) * 2 + 1] = (unsigned long long)clock64();
}
#define TL(ev) tl_stamp(p, ev)
#else
#define TL(ev) ((void)0)
#endif

// TRAIN compiles in the training-only epilogue paths (BatchNorm batch statistics of the stored
// value, fused BatchNorm-backward reductions); eval launches use the leaner TRAIN = false build.
template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool TRAIN, bool PAIR, bool LEAN>
__global__ void __launch_bounds__(384, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ CUtensorMap tmap_res,
                 const __grid_constant__ CUtensorMap tmap_z, const ConvGemmArgs p) {
  static_assert(!LEAN || !TRAIN, "the lean epilogue is inference-only");
  using Cfg = GemmCfg<BLOCK_N, RES, WRES, OUT2, PAIR, LEAN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kBlocksPerTile = BLOCK_N / 64;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);   // SWIZZLE_128B tiles need 1024 B alignment
  if (base & 1023u) __trap();                 // (no slack is reserved for re-aligning)
  uint8_t* smem = smem_raw;

  const uint32_t smem_a = base;
  const uint32_t smem_b = base + kStages * Cfg::kABytes;  // WRES: the resident W slab
  const uint32_t smem_store = base + kStages * Cfg::kStageBytes + Cfg::kWResBytes;
  const uint32_t smem_res = smem_store + Cfg::kStoreTiles * Cfg::kTileBytes;
  const uint32_t bar_base = smem_res + Cfg::kResSlots * Cfg::kTileBytes;
  const uint32_t full_bar = bar_base;
  const uint32_t empty_bar = bar_base + kStages * 8;
  const uint32_t tfull_bar = bar_base + 2 * kStages * 8;
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t rfull_bar = tempty_bar + 16;   // up to 4 auxiliary stages
  const uint32_t rempty_bar = rfull_bar + 32;
  const uint32_t wfull_bar = rempty_bar + 32;   // WRES: the resident W slab has landed
  const uint32_t tmem_slot = wfull_bar + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
  const uint32_t dep_bar = tmem_slot + 8;       // "the dependency wait has returned" (see the producer)
  const uint32_t affine_off = (dep_bar + 8 - base + 15u) & ~15u;   // [scale | shift][BLOCK_N]
  float* s_affine = reinterpret_cast<float*>(smem + affine_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TL(0);

  const int m_tiles = p.dilated ? p.samples * p.tiles_per_sample : p.tiles_per_sample;
  // Work distribution.  A "worker" is a CTA, or a CTA pair; work item w covers N block w % n_tiles of
  // row tile w / n_tiles (pairs: row tiles 2*(w / n_tiles) + rank; the odd row tile of an odd count
  // is out of range for rank 1: its loads are zero-filled and its stores clipped by the tensor maps).
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;
  auto tile_of = [&](int w) -> int {   // work item -> this CTA's (row tile, N block) as m_blk*n_tiles+n_blk
    if (!PAIR) return w;
    return (2 * (w / p.n_tiles) + (int)cta_rank) * p.n_tiles + w % p.n_tiles;
  };
  const int k_iters = p.pairs * p.taps * p.kblocks_per_tap;
  // auxiliary tiles per 64-column store block: the residual plane(s) and, for the fused
  // BatchNorm-backward reductions, the Z tile.  kResSlots / tiles stages are in flight.
  const bool has_res = (p.flags & kEpiResidual) != 0;
  const bool bnb = TRAIN && p.bnb != 0;
  const int aux_tiles = (has_res ? p.res_planes : 0) + (bnb ? 1 : 0);
  const int res_stages =
      (aux_tiles > 0 && Cfg::kResSlots >= aux_tiles) ? Cfg::kResSlots / aux_tiles : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    if (RES) tma_prefetch_desc(&tmap_res);
    if (RES && bnb) tma_prefetch_desc(&tmap_z);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s * 8, 1);
      mbar_init(empty_bar + s * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s * 8, 1);
      // both epilogue groups release an accumulator stage (pairs: of both CTAs, on the leader's)
      mbar_init(tempty_bar + s * 8, PAIR ? 512 : 256);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(rfull_bar + s * 8, 1);
      // one group consumes an auxiliary stage (in place: its four store-issuing lanes hand it back)
      mbar_init(rempty_bar + s * 8, LEAN ? 4 : 128);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(dep_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  __syncwarp();
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TL(1);
  // Everything above touched only this CTA's shared memory / TMEM.  From here on global memory
  // written by the previous kernel of the stream is read: wait for it (no-op without PDL), then let
  // the next kernel start its own prologue on SMs this grid leaves.  The TMA producer thread waits
  // inside its own loop (below): in inference kernels it first streams the W tiles of the first
  // pipeline stages, which no kernel of the forward writes.
  // No lane of the producer's warp may execute griddepcontrol.wait: the instruction stalls the
  // whole warp, diverged lanes included (measured: the producer lane did not issue anything until
  // the wait of its sibling lanes had returned).  The MMA thread tells the producer through an
  // mbarrier that the wait has returned -- the prerequisite grids' writes are visible to the whole
  // grid from then on.
  if (warp != 0) {
    griddep_wait();
    griddep_launch_dependents();
  }
  if (threadIdx.x == 32) {
    mbar_arrive(dep_bar);
    TL(2);
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // Weights ahead of the dependency wait: the packed weights are written at plan build / by the
      // optimiser step, always at least one full kernel boundary (the input pack) before any conv
      // kernel of a forward, so they are safe to read while the previous layer is still running.
      // Besides hiding the cold-HBM latency of the first W tiles this runs the loop body once
      // before the wait: the instruction fetches of the first iterations (measured: 1.5-3 k cycles
      // of instruction-cache misses between the wait and the first load) move off the critical path.
      // Training kernels keep the plain order (their weight packs change every step).
      constexpr bool kEarlyW = !TRAIN;
      if (!kEarlyW) mbar_wait(dep_bar, 0);
      if (WRES) {
        // every (plane, tap, k-block) tile of this CTA's N block, once
        const int n_blk = blockIdx.x % p.n_tiles;
        const int w_tiles = (p.pairs == 3 ? 2 : 1) * p.taps * p.kblocks_per_tap;
        mbar_expect_tx(wfull_bar, w_tiles * Cfg::kBBytes);
        for (int wi = 0; wi < w_tiles; ++wi) {
          const int kb = wi % p.kblocks_per_tap;
          const int slab = wi / p.kblocks_per_tap;  // w_plane * taps + tap
          tma_load_2d(&tmap_w, wfull_bar, smem_b + wi * Cfg::kBBytes, kb * kBlockK,
                      slab * p.n_pad + n_blk * BLOCK_N);
        }
      }
      // mode 0: W only (before the wait), 1: the A tiles of those same stages, 2: steady state
      const int n_pre = (kEarlyW && !WRES) ? (k_iters < kStages ? k_iters : kStages) : 0;
      int mode = n_pre > 0 ? 0 : 2;
      if (kEarlyW && mode == 2) mbar_wait(dep_bar, 0);
      int w = worker;
      if (w < total_tiles) {
        int n_blk, sample, row0;
        tile_coords(p, tile_of(w), n_blk, sample, row0);
        uint32_t stage = 0, phase = 0;
        int it = 0, pair = 0, tap = 0, kb = 0;
        for (;;) {
          const int a_plane = (pair == 1) ? 1 : 0;
          const int w_plane = (pair == 2) ? 1 : 0;
          const uint32_t fb = full_bar + stage * 8;
          // pairs: both CTAs' tiles are credited to the leader's barrier (the MMA issuer waits there)
          const uint32_t lbar = PAIR ? leader_cta_addr(fb) : fb;
          if (mode != 1) {
            mbar_wait(empty_bar + stage * 8, phase ^ 1);
            if (PAIR) {
              if (is_leader) mbar_expect_tx(fb, 2 * Cfg::kStageBytes);
            } else {
              mbar_expect_tx(fb, Cfg::kStageBytes);
            }
            if (!WRES) {
              // pairs: this CTA streams the W rows of its half of the N block
              const int w_row = (w_plane * p.taps + tap) * p.n_pad + n_blk * BLOCK_N +
                                (PAIR ? (int)cta_rank * (BLOCK_N / 2) : 0);
              if (PAIR) tma_load_2d_pair(&tmap_w, lbar, smem_b + stage * Cfg::kBBytes, kb * kBlockK, w_row);
              else tma_load_2d(&tmap_w, fb, smem_b + stage * Cfg::kBBytes, kb * kBlockK, w_row);
#ifdef VP3D_TIMELINE
              if (mode == 0 && it == 0) TL(23);
#endif
            }
          }
          if (mode != 0) {
            const int a_row = row0 + tap * p.tap_row_step;
            const int a_col = tap * p.tap_col_step + kb * kBlockK;
            if (PAIR) tma_load_4d_pair(&tmap_a, lbar, smem_a + stage * Cfg::kABytes, a_col, a_row, sample, a_plane);
            else tma_load_4d(&tmap_a, fb, smem_a + stage * Cfg::kABytes, a_col, a_row, sample, a_plane);
#ifdef VP3D_TIMELINE
            if (w == worker && it == 0) TL(3);
#endif
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          ++it;
          if (++kb == p.kblocks_per_tap) {
            kb = 0;
            if (++tap == p.taps) { tap = 0; ++pair; }
          }
          if (mode == 0 && it == n_pre) {
#ifdef VP3D_TIMELINE
            TL(22);
#endif
            mbar_wait(dep_bar, 0);
#ifdef VP3D_TIMELINE
            TL(21);
#endif
            mode = 1;   // rewind: the A tiles of the stages just primed
            it = 0; pair = 0; tap = 0; kb = 0;
            stage = 0; phase = 0;
          } else if (mode == 1 && it == n_pre) {
            mode = 2;
          }
          if (it == k_iters) {
            w += num_workers;
            if (w >= total_tiles) break;
            tile_coords(p, tile_of(w), n_blk, sample, row0);
            it = 0; pair = 0; tap = 0; kb = 0;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && is_leader) {
      constexpr int kMmaM = PAIR ? 2 * kBlockM : kBlockM;   // pairs: one 256-row MMA over both CTAs
      const uint32_t idesc =
          p.f16 ? make_idesc_f16(kMmaM, BLOCK_N) : make_idesc_bf16(kMmaM, BLOCK_N);
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      if (WRES) mbar_wait(wfull_bar, 0);
      const int per_pair = p.taps * p.kblocks_per_tap;
      for (int w = worker; w < total_tiles; w += num_workers) {
        mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar + stage * 8, phase);
          tc_fence_after();
#ifdef VP3D_TIMELINE
          if (w == worker && it == 0) TL(4);
#endif
          const uint64_t desc_a = make_smem_desc_k_sw128(smem_a + stage * Cfg::kABytes);
          // resident slab index: pairs 0 and 1 read the hi plane of W, pair 2 the lo plane
          const int wi = WRES ? ((it / per_pair == 2 ? per_pair : 0) + it % per_pair) : 0;
          const uint64_t desc_b = make_smem_desc_k_sw128(
              WRES ? smem_b + wi * Cfg::kBBytes : smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
            if (PAIR) umma_bf16_ss_pair(d_tmem, desc_a + 2 * k, desc_b + 2 * k, idesc, (it | k) ? 1u : 0u);
            else umma_bf16_ss(d_tmem, desc_a + 2 * k, desc_b + 2 * k, idesc, (it | k) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if (PAIR) umma_commit_pair(empty_bar + stage * 8);
          else umma_commit(empty_bar + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs of a pair)
        if (PAIR) umma_commit_pair(tfull_bar + acc * 8);
        else umma_commit(tfull_bar + acc * 8);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef VP3D_TIMELINE
        if (w == worker) TL(5);
        TL(6);
#endif
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ auxiliary-tile producer
    if (RES && lane == 0) {
      uint32_t rs = 0, rphase = 0;
      for (int w = worker; w < total_tiles; w += num_workers) {
        int n_blk, sample, row0;
        tile_coords(p, tile_of(w), n_blk, sample, row0);
        for (int sb = 0; sb < kBlocksPerTile; ++sb) {
          const int col = n_blk * BLOCK_N + sb * 64;
          const bool res_here = has_res && col >= p.res_col_begin && col < p.res_col_begin + p.res_cols;
          if (!res_here && !bnb) continue;
          const int n_res = res_here ? p.res_planes : 0;
          mbar_wait(rempty_bar + rs * 8, rphase ^ 1);
          mbar_expect_tx(rfull_bar + rs * 8, (n_res + (bnb ? 1 : 0)) * Cfg::kTileBytes);
          const uint32_t slot0 = smem_res + rs * aux_tiles * Cfg::kTileBytes;
          for (int pl = 0; pl < n_res; ++pl)
            tma_load_4d(&tmap_res, rfull_bar + rs * 8, slot0 + pl * Cfg::kTileBytes,
                        col - p.res_col_begin + p.res_tma_col_off, row0 + p.res_tma_row_off, sample,
                        pl);
          if (bnb)  // the Z tile always sits in the last slot of the stage
            tma_load_4d(&tmap_z, rfull_bar + rs * 8, slot0 + (aux_tiles - 1) * Cfg::kTileBytes, col,
                        row0, sample, 0);
          if (++rs == (uint32_t)res_stages) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && LEAN) {
    // ------------------------------------------------------------ lean inference epilogue
    // Same roles as below (two groups of four warps, alternating 64-column store blocks, per-warp
    // 32-row bulk stores), with affine + ReLU [+ residual] -> one 16-bit plane fixed at compile
    // time.  Residual variant: store block b uses landing slot b % 4 -- always the same group's --
    // and the result is written over the residual bytes this very thread has consumed; the slot
    // goes back to the auxiliary producer when the four bulk stores issued from it have read it.
    const int eg = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int r_in_tile = ew * 32 + lane;
    const uint32_t stage_row = r_in_tile * 128;
    const uint32_t sw = r_in_tile & 7;
    const bool f16 = p.f16 != 0;
    const uint32_t tempty_addr = PAIR ? leader_cta_addr(tempty_bar) : tempty_bar;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t gblock = 0;
    int aff_n_blk = -1;
    uint32_t held_slot = 0;
    bool holding = false;
    for (int w = worker; w < total_tiles; w += num_workers) {
      int n_blk, sample, row0;
      tile_coords(p, tile_of(w), n_blk, sample, row0);
      if (n_blk != aff_n_blk) {   // per-channel affine of this N block -> shared memory
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const int e = (int)threadIdx.x - 128;
        if (e < BLOCK_N) {
          s_affine[e] = __ldg(p.scale + n_blk * BLOCK_N + e);
          s_affine[BLOCK_N + e] = __ldg(p.shift + n_blk * BLOCK_N + e);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        aff_n_blk = n_blk;
      }
      int last_owned = -1;
#pragma unroll
      for (int sb = 0; sb < kBlocksPerTile; ++sb)
        if (((gblock + (uint32_t)sb) & 1u) == (uint32_t)eg) last_owned = sb;
      mbar_wait(tfull_bar + acc * 8, acc_phase);
      tc_fence_after();
#ifdef VP3D_TIMELINE
      if (warp == 4 && lane == 0) { if (w == worker) TL(7); TL(8); }
#endif
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);
      if (last_owned < 0) {
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
      }
#pragma unroll 1
      for (int sb = 0; sb < kBlocksPerTile; ++sb, ++gblock) {
        if ((gblock & 1u) != (uint32_t)eg) continue;
        const int cb = n_blk * BLOCK_N + sb * 64;
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_addr + sb * 64, a0);
        tmem_ld_32x32(t_addr + sb * 64 + 32, a1);
        uint32_t tile_base = smem_store + eg * Cfg::kTileBytes;
        uint4 rres[RES ? 8 : 1];
        if (RES) {
          const uint32_t rs = gblock & 3u;
          if (holding && lane == 0) {
            // this group's previous block: its bulk stores have read the slot -> refill allowed
            tma_store_wait_read<0>();
            mbar_arrive(rempty_bar + held_slot * 8);
          }
          held_slot = rs;
          holding = true;
          mbar_wait(rfull_bar + rs * 8, (gblock >> 2) & 1u);   // residual tile has landed
          tile_base = smem_res + rs * Cfg::kTileBytes;
#pragma unroll
          for (int q = 0; q < (RES ? 8 : 1); ++q)
            rres[q] = ld_shared_v4(tile_base + stage_row + (((uint32_t)q ^ sw) << 4));
        }
        tmem_ld_wait();
#ifdef VP3D_TIMELINE
        if (warp == 4 && lane == 0 && w == worker) TL(13 + 4 * (sb >> 1));
#endif
        if (sb == last_owned) {
          tc_fence_before();
          if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
        }
        uint32_t h0[16], h1[16];
        const float* sc = s_affine + sb * 64;
        const float* sh = s_affine + BLOCK_N + sb * 64;
        if (RES) {
          if (f16) {
            lean_chunk_res<true>(a0, sc, sh, rres, h0);
            lean_chunk_res<true>(a1, sc + 32, sh + 32, rres + (RES ? 4 : 0), h1);
          } else {
            lean_chunk_res<false>(a0, sc, sh, rres, h0);
            lean_chunk_res<false>(a1, sc + 32, sh + 32, rres + (RES ? 4 : 0), h1);
          }
        } else {
          if (f16) {
            lean_chunk<true>(a0, sc, sh, h0);
            lean_chunk<true>(a1, sc + 32, sh + 32, h1);
          } else {
            lean_chunk<false>(a0, sc, sh, h0);
            lean_chunk<false>(a1, sc + 32, sh + 32, h1);
          }
          // the staging slice of this warp must have been read by its previous bulk store
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          st_shared_v4(tile_base + stage_row + (((uint32_t)q ^ sw) << 4), h0[4 * q], h0[4 * q + 1],
                       h0[4 * q + 2], h0[4 * q + 3]);
          st_shared_v4(tile_base + stage_row + (((uint32_t)(4 + q) ^ sw) << 4), h1[4 * q],
                       h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_out, tile_base + (uint32_t)ew * 32u * 128u, cb, row0 + ew * 32, sample, 0);
          tma_store_commit();
#ifdef VP3D_TIMELINE
          if (warp == 4 && w == worker) TL(16 + 4 * (sb >> 1));
#endif
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef VP3D_TIMELINE
      if (warp == 4 && lane == 0) { if (w == worker) TL(9); TL(10); }
#endif
    }
    if (lane == 0) tma_store_wait_all<0>();
#ifdef VP3D_TIMELINE
    if (warp == 4 && lane == 0) TL(11);
#endif
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (two groups)
    const int eg = (warp - 4) >> 2;   // epilogue group 0 / 1
    const int ew = warp & 3;          // TMEM lane quarter this warp may access
    const int r_in_tile = ew * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t gblock = 0;              // store blocks seen so far (both groups count all of them)
    uint32_t ablock = 0;              // auxiliary stages seen so far
    const bool do_relu = p.flags & kEpiRelu;
    const bool do_res = p.flags & kEpiResidual;
    const bool do_stats = TRAIN && (p.flags & kEpiStats);
    const bool f16 = p.f16 != 0;  // IEEE fp16 storage instead of bf16 (eval fp16 mode)
    const bool do_f32 = p.flags & kEpiOutF32;
    const bool do_affine = p.flags & kEpiAffine;
    const bool two_planes = OUT2 && p.out_planes == 2;
    // swizzled tile address of this thread's row: chunk j (16 B) lives at j ^ (row & 7)
    const uint32_t stage_row = r_in_tile * 128;
    const uint32_t sw = r_in_tile & 7;
    // this group's staging tile(s): [hi] or [hi, lo]
    const uint32_t my_store = smem_store + eg * (OUT2 ? 2 : 1) * Cfg::kTileBytes;
    // accumulator release: on this CTA's barrier, or (pairs) on the leader's through the cluster window
    const uint32_t tempty_addr = PAIR ? leader_cta_addr(tempty_bar) : tempty_bar;
    int aff_n_blk = -1;               // N block whose scale / shift sit in shared memory

    for (int w = worker; w < total_tiles; w += num_workers) {
      const int tile = tile_of(w);
      int n_blk, sample, row0;
      tile_coords(p, tile, n_blk, sample, row0);
      const int m_blk = tile / p.n_tiles;  // row-tile index: 4 slabs of 32 rows each
      const int t = row0 + r_in_tile;
      const bool valid = t < p.out_rows;
      const long long out_row = (long long)sample * p.out_rows + t;
      long long res_row = 0;
      bool res_ok = valid;
      if (!RES) {
        int rsmp = sample, rt = t;
        if (!p.dilated && p.res_sample_div > 0) {
          rsmp = t / p.res_sample_div;
          rt = t - rsmp * p.res_sample_div;
        }
        const long long in_sample = (long long)rt * p.res_row_step + p.res_row_off;
        if (p.res_check_rows && (in_sample < 0 || in_sample >= p.res_rows_per_sample)) res_ok = false;
        res_row = (long long)rsmp * p.res_rows_per_sample + in_sample;
      }

      // lo plane only where somebody reads it (tile-uniform)
      const bool two_planes_t =
          two_planes && (p.dilated || (row0 < p.lo_row_end && row0 + kBlockM > p.lo_row_begin));

      // Per-channel affine of this N block: staged once in shared memory (all 32 lanes of a warp
      // need the same 32 values per chunk: broadcast LDS instead of 16 global loads per chunk whose
      // latency sat on the critical path).  Two 256-thread barriers per change of N block (nobody
      // still reads the old values / the new ones are visible), none while it stays the same -- the
      // usual case: the number of workers is a multiple of n_tiles.
      if (do_affine && n_blk != aff_n_blk) {
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const int e = (int)threadIdx.x - 128;
        if (e < BLOCK_N) {
          s_affine[e] = __ldg(p.scale + n_blk * BLOCK_N + e);
          s_affine[BLOCK_N + e] = __ldg(p.shift + n_blk * BLOCK_N + e);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        aff_n_blk = n_blk;
      }
      const float* s_scale = s_affine;
      const float* s_shift = s_affine + BLOCK_N;

      // The last store block of this tile that this group owns: once its accumulator columns are in
      // registers the TMEM stage is released, so the MMAs of the next-but-one tile overlap the math
      // and the store of that block.  (A group that owns no block of the tile releases at once.)
      int last_owned = -1;
#pragma unroll
      for (int sb = 0; sb < kBlocksPerTile; ++sb)
        if (((gblock + (uint32_t)sb) & 1u) == (uint32_t)eg) last_owned = sb;

      // Both groups wait for the accumulator even when a tile holds no block for one of them: the
      // release must not run ahead of the MMAs that refill this TMEM stage.
      mbar_wait(tfull_bar + acc * 8, acc_phase);
      tc_fence_after();
#ifdef VP3D_TIMELINE
      if (warp == 4 && lane == 0) { if (w == worker) TL(7); TL(8); }
#endif
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);
      if (last_owned < 0) {
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
      }

#pragma unroll 1
      for (int sb = 0; sb < kBlocksPerTile; ++sb, ++gblock) {
        const int cb = n_blk * BLOCK_N + sb * 64;  // first column of the store block
        const bool res_here = do_res && cb >= p.res_col_begin && cb < p.res_col_begin + p.res_cols;
        const bool aux_here = RES && (res_here || bnb);
        const uint32_t rs = aux_here ? ablock % (uint32_t)res_stages : 0u;
        const uint32_t ruse = aux_here ? ablock / (uint32_t)res_stages : 0u;  // n-th use of stage rs
        const uint32_t rphase = ruse & 1u;
        if (aux_here) ++ablock;
        if ((gblock & 1u) != (uint32_t)eg) continue;  // the other group owns this store block

        // Inference launches without auxiliary tiles have the registers to fetch BOTH 32-column
        // halves of the store block up front: one exposed TMEM round trip per block instead of two,
        // and the second half's load overlaps the first half's math.
        constexpr bool kPreload = !RES && !TRAIN;
        uint32_t pre[kPreload ? 2 : 1][32];
        if (kPreload) {
          tmem_ld_32x32(t_addr + (sb * 2) * 32, pre[0]);
          tmem_ld_32x32(t_addr + (sb * 2 + 1) * 32, pre[kPreload ? 1 : 0]);
          tmem_ld_wait();
#ifdef VP3D_TIMELINE
          if (warp == 4 && lane == 0 && w == worker) TL(13 + 4 * (sb >> 1));
#endif
          if (sb == last_owned) {
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
          }
        }
#pragma unroll (kPreload ? 2 : 1)
        for (int half = 0; half < 2; ++half) {
          const int chunk = sb * 2 + half;
          const int c0 = cb + half * 32;
          uint32_t raw[32];
          if (kPreload) {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = pre[kPreload ? half : 0][j];
          } else {
            tmem_ld_32x32(t_addr + chunk * 32, raw);
          }
          // while the TMEM load is in flight: make sure the auxiliary tiles of this block have
          // landed and fetch this row's residual bytes (plane 0) from shared memory
          uint4 rres[4];
          if (RES) {
            if (aux_here && half == 0) {
              // The previous use of this stage may belong to the other epilogue group.  A parity
              // wait only tells "one phase ahead" from "done", so first make sure that use has been
              // released (which implies its fill completed); then the fill wait is unambiguous.
              if (ruse > 0) mbar_wait(rempty_bar + rs * 8, (ruse - 1) & 1u);
              mbar_wait(rfull_bar + rs * 8, rphase);  // tiles have landed
            }
            if (res_here) {
              const uint32_t src = smem_res + (rs * aux_tiles) * Cfg::kTileBytes + stage_row;
#pragma unroll
              for (int q = 0; q < 4; ++q) rres[q] = ld_shared_v4(src + (((half * 4 + q) ^ sw) << 4));
            }
          }
          if (!kPreload) {
            tmem_ld_wait();
            if (sb == last_owned && half == 1) {
              // every accumulator column this thread needs from the tile is in registers
              tc_fence_before();
              if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
            }
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

          if (do_affine) {
            const int cl = c0 - n_blk * BLOCK_N;   // column inside the N block
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + cl + j);
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + cl + j);
              v[j + 0] = fmaf(v[j + 0], sc.x, sh.x);
              v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
              v[j + 2] = fmaf(v[j + 2], sc.z, sh.z);
              v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
            }
          }
          if (do_relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (RES) {
            if (res_here) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (f16) add_f16x8(v + q * 8, rres[q]); else add_bf16x8(v + q * 8, rres[q]);
              }
              for (int pl = 1; pl < p.res_planes; ++pl) {   // lo plane of a split-bf16 residual
                const uint32_t src = smem_res + (rs * aux_tiles + pl) * Cfg::kTileBytes + stage_row;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  add_bf16x8(v + q * 8, ld_shared_v4(src + (((half * 4 + q) ^ sw) << 4)));
              }
            }
          } else if (res_here && res_ok) {
            const __nv_bfloat16* rp = p.res + res_row * p.res_ld + (c0 - p.res_col_begin);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              if (pl < p.res_planes) {
                const uint4* r4 = reinterpret_cast<const uint4*>(rp + pl * p.res_plane_stride);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = __ldg(r4 + q);
                  if (f16) add_f16x8(v + q * 8, u); else add_bf16x8(v + q * 8, u);
                }
              }
            }
          }
          if (do_f32) {
            if (valid) {
              float* op = p.out_f32 + out_row * p.out_f32_ld + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < p.n_valid) op[j] = v[j];
            }
          } else {
            uint32_t hi[16];
            if (f16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) hi[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            }
            if (half == 0) {
              // This warp's 32 rows of the group's staging tile(s) must have been read out by the
              // bulk store it issued from them (two store blocks ago).  Every warp stores its own
              // 32 x 64 box, so nothing but the warp itself has to be waited for: no CTA-level
              // barrier sits in the store path.  The wait comes after the TMEM load and the math
              // of this chunk so that store latency overlaps that work.
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(my_store + stage_row + (((half * 4 + q) ^ sw) << 4), hi[4 * q],
                           hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            if (two_planes_t) {
              uint32_t lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float r0 = v[2 * j] - bf16_lo_to_f(hi[j]);
                const float r1 = v[2 * j + 1] - bf16_hi_to_f(hi[j]);
                lo[j] = pack_bf16x2(r0, r1);
              }
              const uint32_t dst_lo = my_store + Cfg::kTileBytes + stage_row;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                st_shared_v4(dst_lo + (((half * 4 + q) ^ sw) << 4), lo[4 * q], lo[4 * q + 1],
                             lo[4 * q + 2], lo[4 * q + 3]);
            }
#ifdef VP3D_TIMELINE
            if (kPreload && warp == 4 && lane == 0 && w == worker) TL(14 + half + 4 * (sb >> 1));
#endif
            if (half == 1) {
              fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
              __syncwarp();
              if (lane == 0) {
                // rows [32 ew, 32 ew + 32) of the staging tile: a 1024-byte aligned slice, so the
                // 128-byte swizzle pattern of the 32-row box matches the one the tile was written in
                const uint32_t src = my_store + (uint32_t)ew * 32u * 128u;
                tma_store_4d(&tmap_out, src, cb, row0 + ew * 32, sample, 0);
                if (two_planes_t)
                  tma_store_4d(&tmap_out, src + Cfg::kTileBytes, cb, row0 + ew * 32, sample, 1);
                tma_store_commit();
#ifdef VP3D_TIMELINE
                if (kPreload && warp == 4 && w == worker) TL(16 + 4 * (sb >> 1));
#endif
              }
            }
          }
          if (RES && bnb) {
            // dY = G(as stored) * dropmask/(1-p) * [Z*scale+shift > 0]; sums over this warp's 32 rows
            float zf[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) zf[j] = 0.0f;
            const uint32_t zsrc =
                smem_res + (rs * aux_tiles + aux_tiles - 1) * Cfg::kTileBytes + stage_row;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              add_bf16x8(zf + q * 8, ld_shared_v4(zsrc + (((half * 4 + q) ^ sw) << 4)));
            const int ch0 = c0 % p.bnb_c;   // channel of the layer below (columns repeat per tap)
            const bool drop = p.bnb_p > 0.0f;
            const uint32_t thresh = (uint32_t)(p.bnb_p * 65536.0f);
            const float inv_keep = drop ? 1.0f / (1.0f - p.bnb_p) : 1.0f;
            const unsigned long long elem0 = (unsigned long long)out_row * p.out_ld + c0;
            const uint32_t key = p.bnb_seed_lo ^ (p.bnb_seed_hi * 0x7F4A7C15u) ^
                                 (p.bnb_layer * 0x632BE5ABu) ^
                                 ((uint32_t)((elem0 >> 1) >> 32) * 0x85EBCA77u);
            const uint32_t pbase = (uint32_t)(elem0 >> 1);
            float s[32], q2[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float m0 = 1.0f, m1 = 1.0f;
              if (drop) {
                uint32_t h = (pbase + (j >> 1)) * 0x9E3779B1u + key;
                h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
                m0 = ((h & 0xFFFFu) >= thresh) ? inv_keep : 0.0f;
                m1 = ((h >> 16) >= thresh) ? inv_keep : 0.0f;
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int jj = j + e;
                // gradient exactly as the next pass will read it back (bf16-rounded)
                const float g = __bfloat162float(__float2bfloat16_rn(v[jj]));
                const float y = fmaf(zf[jj], __ldg(p.bnb_scale + ch0 + jj), __ldg(p.bnb_shift + ch0 + jj));
                float dy = (valid && y > 0.0f) ? g : 0.0f;
                dy *= e ? m1 : m0;
                s[jj] = dy;
                q2[jj] = dy * (zf[jj] - __ldg(p.bnb_mean + ch0 + jj));
              }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = lane & off;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send_s = upper ? s[i] : s[i + off];
                const float keep_s = upper ? s[i + off] : s[i];
                s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                const float send_q = upper ? q2[i] : q2[i + off];
                const float keep_q = upper ? q2[i + off] : q2[i];
                q2[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
              }
            }
            // per-slab partials (slab = 32 rows of this tile), summed in a fixed order by
            // launch_ordered_col_sums: run-to-run reproducible, unlike atomics
            if (m_blk < m_tiles) {   // (a pair's odd row tile has no slab)
              float* bp = p.bnb_sums + ((size_t)(m_blk * 4 + ew) * 2) * p.n_pad + c0 + lane;
              bp[0] = s[0];
              bp[p.n_pad] = q2[0];  // x invstd is applied after the ordered sum
            }
          }
          if (RES && aux_here && half == 1)
            mbar_arrive(rempty_bar + rs * 8);  // this thread is done with the auxiliary stage
          if (do_stats) {
            // Per-channel sum / sum of squares over this warp's 32 rows: butterfly transpose-reduce,
            // 31 shuffles per statistic; afterwards lane j owns channel c0 + j.
            float s[32], q[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = valid ? v[j] : 0.0f;
              s[j] = x;
              q[j] = x * x;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = lane & off;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send_s = upper ? s[i] : s[i + off];
                const float keep_s = upper ? s[i + off] : s[i];
                s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                const float send_q = upper ? q[i] : q[i + off];
                const float keep_q = upper ? q[i + off] : q[i];
                q[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
              }
            }
            // per-slab (32 rows) sum and sum of squares; bn_stats_finalize turns each slab into
            // (count, mean, M2) and merges the slabs in a fixed order (Chan et al.): reproducible,
            // and free of the E[x^2] - E[x]^2 cancellation over the whole batch
            if (m_blk < m_tiles) {   // (a pair's odd row tile has no slab)
              float* sp = p.stats + ((size_t)(m_blk * 4 + ew) * 2) * p.n_pad + c0 + lane;
              sp[0] = s[0];
              sp[p.n_pad] = q[0];
            }
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef VP3D_TIMELINE
      if (warp == 4 && lane == 0) { if (w == worker) TL(9); TL(10); }
#endif
    }
    // the staging tiles must outlive every bulk store that reads them
    if (lane == 0) tma_store_wait_all<0>();
#ifdef VP3D_TIMELINE
    if (warp == 4 && lane == 0) TL(11);
#endif
  }

  __syncwarp();
  tc_fence_before();
  // pairs: neither CTA may leave (or free TMEM) while the other can still read its shared memory
  // through the pair MMAs or signal its barriers
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) TL(12);
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Programmatic dependent launch (on by default, VP3D_PDL=0 turns it off): the next kernel of the
// stream may start its prologue (barrier init, TMEM allocation, descriptor prefetch) on SMs this
// grid has already left; its griddepcontrol.wait still orders every global access behind the
// completion of this grid.
static int g_pdl = -1;   // -1: not decided yet (VP3D_PDL, default on); conv_gemm_set_pdl overrides
static bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("VP3D_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void conv_gemm_set_pdl(int on) { g_pdl = on ? 1 : 0; }
bool conv_gemm_pdl_enabled() { return pdl_enabled(); }

// CTA pairs (cta_group::2) for the 256-wide streamed-W launches: VP3D_PAIR=0 turns them off.
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VP3D_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

bool conv_gemm_pairs_enabled() { return pair_enabled(); }

template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool TRAIN, bool PAIR, bool LEAN = false>
static cudaError_t launch_impl(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w,
                               const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                               const CUtensorMap& tmap_z, const ConvGemmArgs& args, int num_sms,
                               cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, RES, WRES, OUT2, PAIR, LEAN>;
  auto kernel = conv_gemm_kernel<BLOCK_N, RES, WRES, OUT2, TRAIN, PAIR, LEAN>;
  // the dynamic shared memory opt-in is a per-device attribute
  static bool attr_set[kMaxDevices] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int m_tiles = args.dilated ? args.samples * args.tiles_per_sample : args.tiles_per_sample;
  const int total = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * args.n_tiles;
  if (total <= 0) return cudaSuccess;
  const int max_workers = PAIR ? num_sms / 2 : num_sms;
  int workers = total < max_workers ? total : max_workers;
  if (WRES) workers = workers / args.n_tiles * args.n_tiles;  // a CTA keeps its N block for every tile
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(PAIR ? 2 * workers : workers, 1, 1);
  cfg.blockDim = dim3(384, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  if (PAIR) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  return cudaLaunchKernelEx(&cfg, kernel, tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args);
}

// Lean inference epilogue (VP3D_LEAN=0 falls back to the general one): exactly affine + ReLU
// [+ a one-plane TMA residual over every column] into one 16-bit plane.
static bool lean_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VP3D_LEAN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
static bool lean_ok(const ConvGemmArgs& a, bool res) {
  if (!lean_enabled() || a.bnb || a.out_planes != 1 || !a.out) return false;
  if (!res) return a.flags == (kEpiAffine | kEpiRelu) && !a.res_tma;
  return a.flags == (kEpiAffine | kEpiRelu | kEpiResidual) && a.res_tma && a.res_planes == 1 &&
         a.res_col_begin == 0 && a.res_cols >= a.n_pad;
}

template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool PAIR>
static cudaError_t launch_train(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& o,
                                const CUtensorMap& r, const CUtensorMap& z, const ConvGemmArgs& args,
                                int num_sms, cudaStream_t stream) {
  if ((args.flags & kEpiStats) || args.bnb)
    return launch_impl<BLOCK_N, RES, WRES, OUT2, true, PAIR>(a, w, o, r, z, args, num_sms, stream);
  if constexpr (!OUT2) {
    if (lean_ok(args, RES))
      return launch_impl<BLOCK_N, RES, WRES, false, false, PAIR, true>(a, w, o, r, z, args, num_sms, stream);
  }
  return launch_impl<BLOCK_N, RES, WRES, OUT2, false, PAIR>(a, w, o, r, z, args, num_sms, stream);
}

template <int BLOCK_N, bool RES, bool WRES, bool PAIR = false>
static cudaError_t launch_planes(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& o,
                                 const CUtensorMap& r, const CUtensorMap& z, const ConvGemmArgs& args,
                                 int num_sms, cudaStream_t stream) {
  if (args.out_planes == 2 && !(args.flags & kEpiOutF32))
    return launch_train<BLOCK_N, RES, WRES, true, PAIR>(a, w, o, r, z, args, num_sms, stream);
  return launch_train<BLOCK_N, RES, WRES, false, PAIR>(a, w, o, r, z, args, num_sms, stream);
}

// Whether launch_conv_gemm will run this launch on CTA pairs (the W tensor map must then be built
// with 128-row boxes: each CTA of a pair loads half of the N block).
// W-resident variant: whole weight slab of one N block <= 128 KiB and enough row tiles per CTA
bool conv_gemm_uses_wres(const ConvGemmArgs& args, int block_n, int num_sms) {
  const bool res = args.res_tma != 0 || args.bnb != 0;
  const int m_tiles = args.dilated ? args.samples * args.tiles_per_sample : args.tiles_per_sample;
  const long long w_bytes = (long long)(args.pairs == 3 ? 2 : 1) * args.taps * args.kblocks_per_tap *
                            block_n * kBlockK * 2;
  return !res && block_n >= 128 && w_bytes <= 128 * 1024 &&
         (long long)m_tiles * args.n_tiles >= 4LL * num_sms && args.n_tiles <= num_sms;
}

bool conv_gemm_uses_pair(const ConvGemmArgs& args, int block_n, int num_sms) {
  if (!pair_enabled() || block_n != 256 || (num_sms & 1)) return false;
  const bool res = args.res_tma != 0 || args.bnb != 0;
  const int m_tiles = args.dilated ? args.samples * args.tiles_per_sample : args.tiles_per_sample;
  if (conv_gemm_uses_wres(args, block_n, num_sms)) return false;
  if (res && args.out_planes == 2) return false;
  // a pair needs two row tiles to share an N block; tiny launches keep single CTAs
  return m_tiles >= 2;
}

#ifdef VP3D_TIMELINE
static unsigned long long* g_timeline = nullptr;
static int g_timeline_max = 0, g_timeline_next = 0;
void conv_gemm_debug_set_timeline(unsigned long long* buf, int max_launches) {
  g_timeline = buf;
  g_timeline_max = max_launches;
  g_timeline_next = 0;
}
#endif

cudaError_t launch_conv_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w,
                             const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                             const CUtensorMap& tmap_z, const ConvGemmArgs& args_in, int block_n,
                             int num_sms, cudaStream_t stream) {
#ifdef VP3D_TIMELINE
  ConvGemmArgs args = args_in;
  args.timeline = (g_timeline && g_timeline_next < g_timeline_max)
                      ? g_timeline + (size_t)(g_timeline_next++) * 128 : nullptr;
#else
  const ConvGemmArgs& args = args_in;
#endif
  const bool res = args.res_tma != 0 || args.bnb != 0;
  const bool wres = conv_gemm_uses_wres(args, block_n, num_sms);
  switch (block_n) {
    case 256:
      if (conv_gemm_uses_pair(args, block_n, num_sms)) {
        if (res)
          return launch_train<256, true, false, false, true>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
        return launch_planes<256, false, false, true>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
      }
      if (wres) return launch_planes<256, false, true>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
      if (res) {
        // (two output planes next to a TMA residual always run on 128-wide tiles, see run_conv)
        if (args.out_planes == 2) return cudaErrorInvalidConfiguration;
        return launch_train<256, true, false, false, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
      }
      return launch_planes<256, false, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
    case 128:
      if (wres) return launch_planes<128, false, true>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
      return res ? launch_planes<128, true, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream)
                 : launch_planes<128, false, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
    case 64:
      return res ? launch_planes<64, true, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream)
                 : launch_planes<64, false, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace vp3d
