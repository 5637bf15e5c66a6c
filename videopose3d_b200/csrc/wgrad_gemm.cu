// tcgen05 weight-gradient GEMM with MN-major operands.  See wgrad_gemm.cuh for the math.
//
// Same warp-specialised skeleton as conv_gemm.cu (TMA producer / MMA issuer / TMEM allocator /
// 4 epilogue warps, 2 accumulator stages), but the operand tiles are [64 reduction rows x 64
// channels] boxes: channels (the M / N index of the GEMM) are contiguous, so the shared-memory
// descriptors are MN-major SWIZZLE_128B (LBO = distance between 64-channel groups = one 8 KiB box,
// SBO = 1024 B between 8-row groups) and a K = 16 step advances the start address by 16 rows.
#include "ptx.cuh"
#include "wgrad_gemm.cuh"

namespace vp3d {

namespace {
constexpr int kWM = 128;
constexpr int kWK = 64;
constexpr uint32_t kBoxBytes = 64 * 64 * 2;  // 64 rows x 64 channels bf16

template <int BLOCK_N>
struct WCfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr uint32_t kABytes = 2 * kBoxBytes;
  static constexpr uint32_t kBBytes = (BLOCK_N / 64) * kBoxBytes;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

struct Item {
  int tap, m_blk, n_blk, split, kb_begin, kb_end;
};

__device__ __forceinline__ Item decode_item(const WgradArgs& p, int item) {
  Item it;
  it.split = item % p.splits;
  int r = item / p.splits;
  it.n_blk = r % p.n_tiles;
  r /= p.n_tiles;
  it.m_blk = r % p.m_tiles;
  it.tap = r / p.m_tiles;
  const int total_kb = p.per_sample ? p.samples * p.kchunks : p.kchunks;
  it.kb_begin = (int)((long long)it.split * total_kb / p.splits);
  it.kb_end = (int)((long long)(it.split + 1) * total_kb / p.splits);
  return it;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmap_dz,
                  const __grid_constant__ CUtensorMap tmap_x, const WgradArgs p) {
  using Cfg = WCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + kStages * Cfg::kABytes;
  const uint32_t bar_base = base + kStages * Cfg::kStageBytes;
  const uint32_t full_bar = bar_base;
  const uint32_t empty_bar = bar_base + kStages * 8;
  const uint32_t tfull_bar = bar_base + 2 * kStages * 8;
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.taps * p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dz);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s * 8, 1);
      mbar_init(empty_bar + s * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s * 8, 1);
      mbar_init(tempty_bar + s * 8, 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          int sample = 0, krow = kb * kWK;
          if (p.per_sample) {
            sample = kb / p.kchunks;
            krow = (kb - sample * p.kchunks) * kWK;
          }
          for (int pair = 0; pair < p.pairs; ++pair) {
            const int a_plane = (pair == 1) ? 1 : 0;
            const int b_plane = (pair == 2) ? 1 : 0;
            mbar_wait(empty_bar + stage * 8, phase ^ 1);
            mbar_expect_tx(full_bar + stage * 8, Cfg::kStageBytes);
            const uint32_t sa = smem_a + stage * Cfg::kABytes;
            const uint32_t sb = smem_b + stage * Cfg::kBBytes;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_4d(&tmap_dz, full_bar + stage * 8, sa + j * kBoxBytes,
                          it.m_blk * kWM + j * 64, krow, sample, a_plane);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_4d(&tmap_x, full_bar + stage * 8, sb + j * kBoxBytes,
                          it.tap * p.tap_col_step + it.n_blk * BLOCK_N + j * 64,
                          krow + it.tap * p.tap_row_step, sample, b_plane);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kWM, BLOCK_N, 1, 1);  // both operands MN-major
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        const int iters = (it.kb_end - it.kb_begin) * p.pairs;
        mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int i = 0; i < iters; ++i) {
          mbar_wait(full_bar + stage * 8, phase);
          tc_fence_after();
          const uint32_t sa = smem_a + stage * Cfg::kABytes;
          const uint32_t sb = smem_b + stage * Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < kWK / 16; ++k) {
            const uint64_t da = make_smem_desc_mn_sw128(sa + k * 2048, kBoxBytes, 1024);
            const uint64_t db = make_smem_desc_mn_sw128(sb + k * 2048, kBoxBytes, 1024);
            umma_bf16_ss(d_tmem, da, db, idesc, (i | k) ? 1u : 0u);
          }
          umma_commit(empty_bar + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + acc * 8);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const int co = it.m_blk * kWM + ew * 32 + lane;
      float* dst = p.partial +
                   (((long long)it.split * p.taps + it.tap) * p.m_pad + co) * p.n_pad +
                   it.n_blk * BLOCK_N;
      mbar_wait(tfull_bar + acc * 8, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        uint32_t raw[32];
        tmem_ld_32x32(t_addr + chunk * 32, raw);
        tmem_ld_wait();
        float4* o4 = reinterpret_cast<float4*>(dst + chunk * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          o4[q] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                              __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + acc * 8);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
cudaError_t launch_wgrad_impl(const CUtensorMap& tmap_dz, const CUtensorMap& tmap_x,
                              const WgradArgs& a, int num_sms, cudaStream_t stream) {
  using Cfg = WCfg<BLOCK_N>;
  // the dynamic shared memory opt-in is a per-device attribute
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    e = cudaFuncSetAttribute(wgrad_gemm_kernel<BLOCK_N>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int total = a.taps * a.m_tiles * a.n_tiles * a.splits;
  if (total <= 0) return cudaSuccess;
  const int grid = total < num_sms ? total : num_sms;
  wgrad_gemm_kernel<BLOCK_N><<<grid, 256, Cfg::kSmemBytes, stream>>>(tmap_dz, tmap_x, a);
  return cudaGetLastError();
}

// One thread per (co, ci): partial reads are coalesced along ci for every (split, tap); the taps of
// one (co, ci) are written as one contiguous run of the (c_out, c_in, taps) gradient layout.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad,
                                    int splits, int taps_p, int m_pad, int n_pad, int c_out, int c_in,
                                    int taps_out, int merged) {
  const long long total = (long long)c_out * c_in;
  const long long split_stride = (long long)taps_p * m_pad * n_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % c_in);
    const int co = (int)(i / c_in);
    for (int tap = 0; tap < taps_out; ++tap) {
      const long long src = merged ? ((long long)co * n_pad + tap * c_in + ci)
                                   : (((long long)tap * m_pad + co) * n_pad + ci);
      float s = 0.0f;
      for (int sp = 0; sp < splits; ++sp) s += __ldg(partial + sp * split_stride + src);
      grad[i * taps_out + tap] = s;
    }
  }
}
}  // namespace

cudaError_t launch_wgrad_gemm(const CUtensorMap& tmap_dz, const CUtensorMap& tmap_x,
                              const WgradArgs& args, int block_n, int num_sms, cudaStream_t stream) {
  switch (block_n) {
    case 256: return launch_wgrad_impl<256>(tmap_dz, tmap_x, args, num_sms, stream);
    case 128: return launch_wgrad_impl<128>(tmap_dz, tmap_x, args, num_sms, stream);
    case 64: return launch_wgrad_impl<64>(tmap_dz, tmap_x, args, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_wgrad_reduce(const float* partial, float* grad, int splits, int taps_p, int m_pad,
                                int n_pad, int c_out, int c_in, int taps_out, int merged,
                                cudaStream_t stream) {
  const long long total = (long long)c_out * c_in;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wgrad_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(partial, grad, splits, taps_p, m_pad, n_pad,
                                                       c_out, c_in, taps_out, merged);
  return cudaGetLastError();
}

}  // namespace vp3d
