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
//                   into the group's SWIZZLE_128B staging tile(s) -> TMA store (coalesced 128-byte
//                   rows, clipped at the tensor edge by the tensor map)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
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
template <int BLOCK_N, bool RES, bool WRES, bool OUT2>
struct GemmCfg {
  static_assert(!(RES && WRES), "W-resident variant has no residual path");
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = WRES ? kABytes : kABytes + kBBytes;
  static constexpr uint32_t kWResBytes = WRES ? 128u * 1024u : 0u;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
  static constexpr uint32_t kTileBytes = kBlockM * 64 * 2;  // one 128 x 64 bf16 tile (16 KiB)
  static constexpr int kStoreTiles = OUT2 ? 4 : 2;          // one (hi[, lo]) set per epilogue group
  // auxiliary (residual / Z) landing tiles: 3 next to 128x256 tiles, 4 next to narrower ones, so
  // that two-tile store blocks (hi+lo residual, or residual + Z) still get two stages in flight
  // (with two output planes the four staging tiles already take 64 KiB: one auxiliary stage only,
  // the second epilogue group covers the exposed load latency, and the operand pipeline keeps its
  // depth)
  static constexpr int kResSlots = RES ? (BLOCK_N == 256 ? 3 : 4) : 0;
  static constexpr uint32_t kFixedBytes = kWResBytes + (kStoreTiles + kResSlots) * kTileBytes;
  // as many operand stages as fit below 224 KiB, at most 8
  static constexpr int kStagesFit = (224u * 1024u - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static_assert(kStages >= 2, "not enough shared memory for the operand pipeline");
  static constexpr uint32_t kBarBytes = (2 * kStages + 4 + 8 + 1) * 8 + 16;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kFixedBytes + kBarBytes + 1024;
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

// TRAIN compiles in the training-only epilogue paths (BatchNorm batch statistics of the stored
// value, fused BatchNorm-backward reductions); eval launches use the leaner TRAIN = false build.
template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool TRAIN>
__global__ void __launch_bounds__(384, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ CUtensorMap tmap_res,
                 const __grid_constant__ CUtensorMap tmap_z, const ConvGemmArgs p) {
  using Cfg = GemmCfg<BLOCK_N, RES, WRES, OUT2>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kBlocksPerTile = BLOCK_N / 64;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = p.dilated ? p.samples * p.tiles_per_sample : p.tiles_per_sample;
  const int total_tiles = m_tiles * p.n_tiles;
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
      mbar_init(tempty_bar + s * 8, 256);  // both epilogue groups release an accumulator stage
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(rfull_bar + s * 8, 1);
      mbar_init(rempty_bar + s * 8, 128);  // one group consumes an auxiliary stage
    }
    mbar_init(wfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Everything above touched only this CTA's shared memory / TMEM.  From here on global memory
  // written by the previous kernel of the stream is read: wait for it (no-op without PDL), then let
  // the next kernel start its own prologue on SMs this grid leaves.
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n_blk, sample, row0;
        tile_coords(p, tile, n_blk, sample, row0);
        for (int pair = 0; pair < p.pairs; ++pair) {
          const int a_plane = (pair == 1) ? 1 : 0;
          const int w_plane = (pair == 2) ? 1 : 0;
          for (int tap = 0; tap < p.taps; ++tap) {
            const int a_row = row0 + tap * p.tap_row_step;
            const int a_col0 = tap * p.tap_col_step;
            const int w_row = (w_plane * p.taps + tap) * p.n_pad + n_blk * BLOCK_N;
            for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
              mbar_wait(empty_bar + stage * 8, phase ^ 1);
              mbar_expect_tx(full_bar + stage * 8, Cfg::kStageBytes);
              tma_load_4d(&tmap_a, full_bar + stage * 8, smem_a + stage * Cfg::kABytes,
                          a_col0 + kb * kBlockK, a_row, sample, a_plane);
              if (!WRES)
                tma_load_2d(&tmap_w, full_bar + stage * 8, smem_b + stage * Cfg::kBBytes,
                            kb * kBlockK, w_row);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc =
          p.f16 ? make_idesc_f16(kBlockM, BLOCK_N) : make_idesc_bf16(kBlockM, BLOCK_N);
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      if (WRES) mbar_wait(wfull_bar, 0);
      const int per_pair = p.taps * p.kblocks_per_tap;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar + stage * 8, phase);
          tc_fence_after();
          const uint64_t desc_a = make_smem_desc_k_sw128(smem_a + stage * Cfg::kABytes);
          // resident slab index: pairs 0 and 1 read the hi plane of W, pair 2 the lo plane
          const int wi = WRES ? ((it / per_pair == 2 ? per_pair : 0) + it % per_pair) : 0;
          const uint64_t desc_b = make_smem_desc_k_sw128(
              WRES ? smem_b + wi * Cfg::kBBytes : smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
            umma_bf16_ss(d_tmem, desc_a + 2 * k, desc_b + 2 * k, idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(empty_bar + stage * 8);  // frees the smem stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + acc * 8);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ auxiliary-tile producer
    if (RES && lane == 0) {
      uint32_t rs = 0, rphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n_blk, sample, row0;
        tile_coords(p, tile, n_blk, sample, row0);
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
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (two groups)
    const int eg = (warp - 4) >> 2;   // epilogue group 0 / 1
    const int ew = warp & 3;          // TMEM lane quarter this warp may access
    const int r_in_tile = ew * 32 + lane;
    const bool store_leader = (threadIdx.x == 128u + 128u * eg);
    const uint32_t bar_id = 1 + eg;
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

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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

      // Both groups wait for the accumulator even when a tile holds no block for one of them: the
      // release below must not run ahead of the MMAs that refill this TMEM stage.
      mbar_wait(tfull_bar + acc * 8, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);

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

#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int chunk = sb * 2 + half;
          const int c0 = cb + half * 32;
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + chunk * 32, raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

          if (do_affine) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + j));
              const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + j));
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
            if (aux_here && half == 0) {
              // The previous use of this stage may belong to the other epilogue group.  A parity
              // wait only tells "one phase ahead" from "done", so first make sure that use has been
              // released (which implies its fill completed); then the fill wait is unambiguous.
              if (ruse > 0) mbar_wait(rempty_bar + rs * 8, (ruse - 1) & 1u);
              mbar_wait(rfull_bar + rs * 8, rphase);  // tiles have landed
            }
            if (res_here) {
              for (int pl = 0; pl < p.res_planes; ++pl) {
                const uint32_t src = smem_res + (rs * aux_tiles + pl) * Cfg::kTileBytes + stage_row;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = ld_shared_v4(src + (((half * 4 + q) ^ sw) << 4));
                  if (f16) add_f16x8(v + q * 8, u); else add_bf16x8(v + q * 8, u);
                }
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
              // This group's staging tile(s) must have been read out by the bulk store that used
              // them last (two store blocks ago).  The wait sits after the TMEM load and the math
              // of this chunk so that store latency overlaps that work.
              if (store_leader) tma_store_wait_read<0>();
              group_bar_sync(bar_id);
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
            if (half == 1) {
              fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
              group_bar_sync(bar_id);
              if (store_leader) {
                tma_store_4d(&tmap_out, my_store, cb, row0, sample, 0);
                if (two_planes_t)
                  tma_store_4d(&tmap_out, my_store + Cfg::kTileBytes, cb, row0, sample, 1);
                tma_store_commit();
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
            float* bp = p.bnb_sums + ((size_t)(m_blk * 4 + ew) * 2) * p.n_pad + c0 + lane;
            bp[0] = s[0];
            bp[p.n_pad] = q2[0];  // x invstd is applied after the ordered sum
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
            float* sp = p.stats + ((size_t)(m_blk * 4 + ew) * 2) * p.n_pad + c0 + lane;
            sp[0] = s[0];
            sp[p.n_pad] = q[0];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + acc * 8);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    // the staging tiles must outlive every bulk store that reads them
    if (store_leader) tma_store_wait_all<0>();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Programmatic dependent launch (on by default, VP3D_PDL=0 turns it off): the next kernel of the
// stream may start its prologue (barrier init, TMEM allocation, descriptor prefetch) on SMs this
// grid has already left; its griddepcontrol.wait still orders every global access behind the
// completion of this grid.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VP3D_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BLOCK_N, bool RES, bool WRES, bool OUT2, bool TRAIN>
static cudaError_t launch_impl(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w,
                               const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                               const CUtensorMap& tmap_z, const ConvGemmArgs& args, int num_sms,
                               cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, RES, WRES, OUT2>;
  auto kernel = conv_gemm_kernel<BLOCK_N, RES, WRES, OUT2, TRAIN>;
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
  const int total = m_tiles * args.n_tiles;
  if (total <= 0) return cudaSuccess;
  int grid = total < num_sms ? total : num_sms;
  if (WRES) grid = grid / args.n_tiles * args.n_tiles;  // a CTA keeps its N block for every tile
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(384, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args);
}

template <int BLOCK_N, bool RES, bool WRES, bool OUT2>
static cudaError_t launch_train(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& o,
                                const CUtensorMap& r, const CUtensorMap& z, const ConvGemmArgs& args,
                                int num_sms, cudaStream_t stream) {
  if ((args.flags & kEpiStats) || args.bnb)
    return launch_impl<BLOCK_N, RES, WRES, OUT2, true>(a, w, o, r, z, args, num_sms, stream);
  return launch_impl<BLOCK_N, RES, WRES, OUT2, false>(a, w, o, r, z, args, num_sms, stream);
}

template <int BLOCK_N, bool RES, bool WRES>
static cudaError_t launch_planes(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& o,
                                 const CUtensorMap& r, const CUtensorMap& z, const ConvGemmArgs& args,
                                 int num_sms, cudaStream_t stream) {
  if (args.out_planes == 2 && !(args.flags & kEpiOutF32))
    return launch_train<BLOCK_N, RES, WRES, true>(a, w, o, r, z, args, num_sms, stream);
  return launch_train<BLOCK_N, RES, WRES, false>(a, w, o, r, z, args, num_sms, stream);
}

cudaError_t launch_conv_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w,
                             const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                             const CUtensorMap& tmap_z, const ConvGemmArgs& args, int block_n,
                             int num_sms, cudaStream_t stream) {
  const bool res = args.res_tma != 0 || args.bnb != 0;
  // W-resident variant: whole weight slab of one N block <= 128 KiB and enough row tiles per CTA
  const int m_tiles = args.dilated ? args.samples * args.tiles_per_sample : args.tiles_per_sample;
  const long long w_bytes = (long long)(args.pairs == 3 ? 2 : 1) * args.taps * args.kblocks_per_tap *
                            block_n * kBlockK * 2;
  const bool wres = !res && block_n >= 128 && w_bytes <= 128 * 1024 &&
                    (long long)m_tiles * args.n_tiles >= 4LL * num_sms && args.n_tiles <= num_sms;
  switch (block_n) {
    case 256:
      if (wres) return launch_planes<256, false, true>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
      if (res) {
        // (two output planes next to a TMA residual always run on 128-wide tiles, see run_conv)
        if (args.out_planes == 2) return cudaErrorInvalidConfiguration;
        return launch_train<256, true, false, false>(tmap_a, tmap_w, tmap_out, tmap_res, tmap_z, args, num_sms, stream);
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
