// tcgen05 weight-gradient GEMM with MN-major operands.  See wgrad_gemm.cuh for the math.
//
// Same warp-specialised skeleton as conv_gemm.cu (TMA producer / MMA issuer / TMEM allocator /
// 4 epilogue warps, 2 accumulator stages), but the operand tiles are [64 reduction rows x 64
// channels] boxes: channels (the M / N index of the GEMM) are contiguous, so the shared-memory
// descriptors are MN-major SWIZZLE_128B (LBO = distance between 64-channel groups = one 8 KiB box,
// SBO = 1024 B between 8-row groups) and a K = 16 step advances the start address by 16 rows.
#include <stdlib.h>
#include <string.h>

#include "launch.cuh"
#include "ptx.cuh"
#include "wgrad_gemm.cuh"

namespace vp3d {

namespace {
constexpr int kWM = 128;
constexpr int kWK = 64;
constexpr uint32_t kBoxBytes = 64 * 64 * 2;  // 64 rows x 64 channels bf16

// PAIR (BLOCK_N = 256, even number of C_out tiles): clusters of two CTAs compute a 256 x 256 tile
// of one (tap, split) together with one tcgen05.mma.cta_group::2 stream issued by the cluster's
// rank-0 CTA; each CTA loads its own 128 dZ channels and HALF of the X channels (32 instead of
// 48 KiB per k-chunk), which is what the shared-memory-port-bound single-CTA tiles were missing
// (768 -> 512 port cycles per chunk) and deepens the operand pipeline from 4 to 6 stages.
template <int BLOCK_N, bool PAIR = false>
struct WCfg {
  static_assert(!PAIR || BLOCK_N == 256, "CTA pairs run 256-wide tiles");
  static constexpr int kStages = PAIR ? 6 : ((BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8));
  static constexpr uint32_t kABytes = 2 * kBoxBytes;
  static constexpr uint32_t kBBytes = ((PAIR ? BLOCK_N / 2 : BLOCK_N) / 64) * kBoxBytes;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

struct Item {
  int tap, m_blk, n_blk, split, kb_begin, kb_end;
};

// m_units: C_out tiles (single CTAs) or pairs of them (CTA pairs: this CTA takes tile 2*unit + rank)
__device__ __forceinline__ Item decode_item(const WgradArgs& p, int item, int m_units, int m_mul,
                                            int m_add) {
  Item it;
  it.split = item % p.splits;
  int r = item / p.splits;
  it.n_blk = r % p.n_tiles;
  r /= p.n_tiles;
  it.m_blk = (r % m_units) * m_mul + m_add;
  it.tap = r / m_units;
  const int total_kb = p.per_sample ? p.samples * p.kchunks : p.kchunks;
  it.kb_begin = (int)((long long)it.split * total_kb / p.splits);
  it.kb_end = (int)((long long)(it.split + 1) * total_kb / p.splits);
  return it;
}

template <int BLOCK_N, bool PAIR>
__global__ void __launch_bounds__(256, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmap_dz,
                  const __grid_constant__ CUtensorMap tmap_x, const WgradArgs p) {
  using Cfg = WCfg<BLOCK_N, PAIR>;
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
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = PAIR ? p.m_tiles / 2 : p.m_tiles;
  const int m_mul = PAIR ? 2 : 1, m_add = (int)cta_rank;
  const int total_items = p.taps * m_units * p.n_tiles * p.splits;

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
      mbar_init(tempty_bar + s * 8, PAIR ? 256 : 128);   // pairs: both CTAs' epilogues, on the leader's
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // (set-up above overlaps the predecessor's tail under programmatic dependent launch)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = worker; item < total_items; item += num_workers) {
        const Item it = decode_item(p, item, m_units, m_mul, m_add);
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
            const uint32_t sa = smem_a + stage * Cfg::kABytes;
            const uint32_t sb = smem_b + stage * Cfg::kBBytes;
            if (PAIR) {
              // both CTAs' boxes are credited to the leader's barrier (the MMA issuer waits there);
              // this CTA streams the X channels of its half of the N block
              const uint32_t lbar = leader_cta_addr(full_bar + stage * 8);
              if (is_leader) mbar_expect_tx(full_bar + stage * 8, 2 * Cfg::kStageBytes);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                tma_load_4d_pair(&tmap_dz, lbar, sa + j * kBoxBytes, it.m_blk * kWM + j * 64, krow,
                                 sample, a_plane);
#pragma unroll
              for (int j = 0; j < BLOCK_N / 128; ++j)
                tma_load_4d_pair(&tmap_x, lbar, sb + j * kBoxBytes,
                                 it.tap * p.tap_col_step + it.n_blk * BLOCK_N +
                                     (int)cta_rank * (BLOCK_N / 2) + j * 64,
                                 krow + it.tap * p.tap_row_step, sample, b_plane);
            } else {
              mbar_expect_tx(full_bar + stage * 8, Cfg::kStageBytes);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                tma_load_4d(&tmap_dz, full_bar + stage * 8, sa + j * kBoxBytes,
                            it.m_blk * kWM + j * 64, krow, sample, a_plane);
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_4d(&tmap_x, full_bar + stage * 8, sb + j * kBoxBytes,
                            it.tap * p.tap_col_step + it.n_blk * BLOCK_N + j * 64,
                            krow + it.tap * p.tap_row_step, sample, b_plane);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && is_leader) {
      // both operands MN-major; pairs: one 256-row MMA over both CTAs
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kWM : kWM, BLOCK_N, 1, 1);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int item = worker; item < total_items; item += num_workers) {
        const Item it = decode_item(p, item, m_units, m_mul, m_add);
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
            if (PAIR) umma_bf16_ss_pair(d_tmem, da, db, idesc, (i | k) ? 1u : 0u);
            else umma_bf16_ss(d_tmem, da, db, idesc, (i | k) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if (PAIR) umma_commit_pair(empty_bar + stage * 8);
          else umma_commit(empty_bar + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(tfull_bar + acc * 8);
        else umma_commit(tfull_bar + acc * 8);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    const uint32_t tempty_addr = PAIR ? leader_cta_addr(tempty_bar) : tempty_bar;
    for (int item = worker; item < total_items; item += num_workers) {
      const Item it = decode_item(p, item, m_units, m_mul, m_add);
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
      if (PAIR) mbar_arrive_cluster(tempty_addr + acc * 8); else mbar_arrive(tempty_addr + acc * 8);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  __syncwarp();
  tc_fence_before();
  // pairs: neither CTA may leave (or free TMEM) while the other can still read its shared memory
  // through the pair MMAs or signal its barriers
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, bool PAIR>
cudaError_t launch_wgrad_impl(const CUtensorMap& tmap_dz, const CUtensorMap& tmap_x,
                              const WgradArgs& a, int num_sms, cudaStream_t stream) {
  using Cfg = WCfg<BLOCK_N, PAIR>;
  auto kernel = wgrad_gemm_kernel<BLOCK_N, PAIR>;
  // the dynamic shared memory opt-in is a per-device attribute
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int total = a.taps * (PAIR ? a.m_tiles / 2 : a.m_tiles) * a.n_tiles * a.splits;
  if (total <= 0) return cudaSuccess;
  const int max_workers = PAIR ? num_sms / 2 : num_sms;
  const int workers = total < max_workers ? total : max_workers;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(PAIR ? 2 * workers : workers, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (conv_gemm_pdl_enabled()) {
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
  return cudaLaunchKernelEx(&cfg, kernel, tmap_dz, tmap_x, a);
}

// One thread per gradient element (co, ci, tap) in the (c_out, c_in, taps) order of
// Conv1d.weight.grad: the writes are fully coalesced and the split partials of an element are the
// only serial chain (summed left to right: deterministic).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad,
                                    int splits, int taps_p, int m_pad, int n_pad, int c_out, int c_in,
                                    int taps_out, int merged) {
  pdl_entry();
  const long long total = (long long)c_out * c_in * taps_out;
  const long long split_stride = (long long)taps_p * m_pad * n_pad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(e % taps_out);
    const long long i = e / taps_out;
    const int ci = (int)(i % c_in);
    const int co = (int)(i / c_in);
    const long long src = merged ? ((long long)co * n_pad + tap * c_in + ci)
                                 : (((long long)tap * m_pad + co) * n_pad + ci);
    float s = 0.0f;
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {   // four loads in flight, same summation order
      const float a = __ldg(partial + (sp + 0) * split_stride + src);
      const float b = __ldg(partial + (sp + 1) * split_stride + src);
      const float c = __ldg(partial + (sp + 2) * split_stride + src);
      const float d = __ldg(partial + (sp + 3) * split_stride + src);
      s += a; s += b; s += c; s += d;
    }
    for (; sp < splits; ++sp) s += __ldg(partial + sp * split_stride + src);
    grad[e] = s;
  }
}
}  // namespace

cudaError_t launch_wgrad_gemm(const CUtensorMap& tmap_dz, const CUtensorMap& tmap_x,
                              const WgradArgs& args, int block_n, int num_sms, cudaStream_t stream) {
  switch (block_n) {
    case 256:
      if (wgrad_gemm_uses_pair(args, block_n, num_sms))
        return launch_wgrad_impl<256, true>(tmap_dz, tmap_x, args, num_sms, stream);
      return launch_wgrad_impl<256, false>(tmap_dz, tmap_x, args, num_sms, stream);
    case 128: return launch_wgrad_impl<128, false>(tmap_dz, tmap_x, args, num_sms, stream);
    case 64: return launch_wgrad_impl<64, false>(tmap_dz, tmap_x, args, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

// CTA pairs: 256-wide tiles, an even number of C_out tiles, an even number of SMs (VP3D_PAIR=0 off)
bool wgrad_gemm_uses_pair(const WgradArgs& args, int block_n, int num_sms) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VP3D_PAIR");
    on = (e && e[0] == '0') ? 0 : 1;
    const char* w = getenv("VP3D_PAIR_WGRAD");   // measurement knob: wgrad pairs alone
    if (w && w[0] == '0') on = 0;
  }
  return on && block_n == 256 && !(args.m_tiles & 1) && !(num_sms & 1);
}

cudaError_t launch_wgrad_reduce(const float* partial, float* grad, int splits, int taps_p, int m_pad,
                                int n_pad, int c_out, int c_in, int taps_out, int merged,
                                cudaStream_t stream) {
  const long long total = (long long)c_out * c_in * taps_out;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  const cudaError_t le = launch_pdl(wgrad_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, stream,
                                    partial, grad, splits, taps_p, m_pad, n_pad, c_out, c_in, taps_out,
                                    merged);
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace vp3d
