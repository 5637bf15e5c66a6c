// C-ABI implementation: plan construction, parameter packing, eval-mode forward schedules
// (strided / dependency-cone and dilated) built from the conv GEMM kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "internal.cuh"
#include "pack.cuh"

using namespace vp3d;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
namespace vp3d {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace vp3d

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D bf16 map (k, row, sample, plane), box (64, box_rows, 1, 1), 128-byte swizzle.
namespace vp3d {
int make_map_4d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                       uint64_t row_stride, uint64_t samples, uint64_t sample_stride,
                       uint64_t planes, uint64_t plane_stride, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VP3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride * 2) % 16 || (sample_stride * 2) % 16 ||
      (plane_stride * 2) % 16)
    return fail(VP3D_ERR_INVALID, "tensor map operand not 16-byte aligned");
  cuuint64_t dims[4] = {inner, rows, samples, planes};
  cuuint64_t strides[3] = {row_stride * 2, sample_stride * 2, plane_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VP3D_ERR_CUDA,
                "cuTensorMapEncodeTiled(4d) failed: %d (inner=%llu rows=%llu rs=%llu samples=%llu "
                "ss=%llu planes=%llu ps=%llu)",
                (int)r, (unsigned long long)inner, (unsigned long long)rows,
                (unsigned long long)row_stride, (unsigned long long)samples,
                (unsigned long long)sample_stride, (unsigned long long)planes,
                (unsigned long long)plane_stride);
  return VP3D_OK;
}

// 2-D bf16 map (k, row), box (64, box_rows), 128-byte swizzle.
int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                       uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VP3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VP3D_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu rows=%llu)",
                (int)r, (unsigned long long)inner, (unsigned long long)rows);
  return VP3D_OK;
}

int pick_block_n(int n_pad) {
  if (n_pad % 256 == 0) return 256;
  if (n_pad % 128 == 0) return 128;
  return 64;
}

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
static int g_sm_limit = 0;   // vp3d_set_sm_limit: persistent grids leave the other SMs to NCCL
int num_sms() {
  static int cache[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
    if (const char* e = getenv("VP3D_SM_LIMIT")) {
      const int lim = atoi(e);
      if (lim >= 2 && !g_sm_limit) g_sm_limit = lim;
    }
  }
  int n = cache[dev];
  if (g_sm_limit >= 2 && g_sm_limit < n) n = g_sm_limit & ~1;   // even: CTA pairs
  return n;
}

// ------------------------------------------------------------------ operator level
int run_conv(const vp3d_conv_desc* d, cudaStream_t stream) {
  if (!d || !d->a || !d->w) return fail(VP3D_ERR_INVALID, "conv_gemm: null operand");
  if (d->a_ld % 64 || d->k_per_tap % 64 || d->n_pad % 64)
    return fail(VP3D_ERR_INVALID, "conv_gemm: a_ld, k_per_tap and n_pad must be multiples of 64");
  if (d->taps < 1 || d->out_rows < 0 || d->samples < 1)
    return fail(VP3D_ERR_INVALID, "conv_gemm: bad geometry");
  if (d->out_rows == 0) return VP3D_OK;
  const int a_planes = d->a_planes > 0 ? d->a_planes : 1;
  const int pairs = d->precision == VP3D_PRECISION_BF16X3 ? 3 : 1;
  const int f16 = d->precision == VP3D_PRECISION_FP16 ? 1 : 0;
  if (f16 && (a_planes != 1 || d->out_planes > 1 || d->stats || d->bnb_z))
    return fail(VP3D_ERR_INVALID, "conv_gemm: fp16 is a single-plane, inference-only format");
  if (pairs == 3 && a_planes != 2)
    return fail(VP3D_ERR_INVALID, "conv_gemm: bf16x3 needs hi/lo planes of A");
  const int w_planes = pairs == 3 ? 2 : 1;
  // Largest N tile that still yields at least one tile per SM; small layers (few row tiles) fall
  // back to narrower tiles so that more SMs (and more TMEM/TMA pipelines) share the work.
  int block_n = 64;
  {
    const long long rows_total = d->per_sample_tiles ? 0 : d->out_rows;
    const long long m_tiles = d->per_sample_tiles
                                  ? (long long)d->samples * ((d->out_rows + kBlockM - 1) / kBlockM)
                                  : (rows_total + kBlockM - 1) / kBlockM;
    const int cands[3] = {256, 128, 64};
    bool found = false;
    static int wave_num = 0;   // measurement knob: accept a tile width once it fills 1/wave_num wave
    if (!wave_num) {
      const char* e = getenv("VP3D_WAVE_NUM");
      wave_num = (e && atoi(e) > 0) ? atoi(e) : 2;
    }
    for (int c : cands) {
      if (d->n_pad % c) continue;
      // half a wave of full-rate 128x256 tiles beats a full wave of narrower (smem-bound) ones
      if (m_tiles * (d->n_pad / c) * wave_num >= num_sms()) { block_n = c; found = true; break; }
    }
    if (!found) block_n = 64;
    // Store blocks that need two auxiliary tiles (hi+lo residual, or residual + the Z tile of the
    // fused BatchNorm backward) only get a second prefetch stage next to 128-wide tiles; those
    // launches are epilogue-bound, so the narrower MMA costs nothing.
    // (CTA pairs keep 256-wide tiles: their four landing tiles give two-tile blocks two stages)
    const int aux_tiles = (d->res ? (d->res_planes > 0 ? d->res_planes : 1) : 0) + (d->bnb_z ? 1 : 0);
    const bool pair_ok = conv_gemm_pairs_enabled() && !(num_sms() & 1) && m_tiles >= 2 &&
                         (d->out_planes <= 1);
    if (aux_tiles >= 2 && block_n == 256 && d->n_pad % 128 == 0 && !pair_ok) block_n = 128;
    // (the residual kernel variant has no 256-wide, two-output-plane instantiation)
    if (aux_tiles >= 1 && d->out_planes == 2 && block_n == 256) block_n = 128;
  }

  CUtensorMap ma, mw;
  const uint64_t a_rows = d->a_rows, a_ld = d->a_ld;
  const uint64_t plane_stride = (uint64_t)d->samples * a_rows * a_ld;
  VP3D_TRY(make_map_4d(&ma, d->a, a_ld, a_rows, a_ld, d->samples, a_rows * a_ld, a_planes,
                       plane_stride, kBlockM));
  ConvGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.dilated = d->per_sample_tiles ? 1 : 0;
  g.samples = d->samples;
  g.out_rows = d->out_rows;
  g.tiles_per_sample = (d->out_rows + kBlockM - 1) / kBlockM;
  g.taps = d->taps;
  g.kblocks_per_tap = d->k_per_tap / kBlockK;
  g.tap_row_step = d->tap_row_step;
  g.tap_col_step = d->tap_col_step;
  g.n_pad = d->n_pad;
  g.n_tiles = d->n_pad / block_n;
  g.pairs = pairs;
  g.f16 = f16;
  g.flags = 0;
  if (d->scale && d->shift) g.flags |= kEpiAffine;
  if (d->relu) g.flags |= kEpiRelu;
  if (d->res) g.flags |= kEpiResidual;
  if (d->stats) g.flags |= kEpiStats;
  if (d->out_f32) g.flags |= kEpiOutF32;
  g.scale = d->scale;
  g.shift = d->shift;
  g.res = static_cast<const __nv_bfloat16*>(d->res);
  g.res_plane_stride = d->res_plane_stride;
  g.res_planes = d->res ? (d->res_planes > 0 ? d->res_planes : 1) : 0;
  g.res_ld = d->res_ld;
  g.res_rows_per_sample = d->res_rows_per_sample;
  g.res_row_step = d->res_row_step;
  g.res_row_off = d->res_row_off;
  g.res_sample_div = d->res_sample_div;
  g.res_col_begin = d->res_col_begin;
  g.res_cols = d->res_cols > 0 ? d->res_cols : d->n_pad;
  g.res_check_rows = d->res_check_rows;
  g.out = static_cast<__nv_bfloat16*>(d->out);
  g.out_plane_stride = d->out_plane_stride;
  g.out_planes = d->out_planes > 0 ? d->out_planes : 1;
  g.out_ld = d->out_ld;
  g.out_f32 = d->out_f32;
  g.out_f32_ld = d->out_f32_ld;
  g.n_valid = d->n_valid > 0 ? d->n_valid : d->n_pad;
  g.stats = d->stats;
  g.lo_row_begin = d->lo_row_end > 0 ? d->lo_row_begin : 0;
  g.lo_row_end = d->lo_row_end > 0 ? d->lo_row_end : 0x7fffffff;
  if (!d->out && !d->out_f32) return fail(VP3D_ERR_INVALID, "conv_gemm: no output");
  if (d->out && (d->out_ld % 8)) return fail(VP3D_ERR_INVALID, "conv_gemm: out_ld % 8 != 0");
  if (d->res && (d->res_ld % 8)) return fail(VP3D_ERR_INVALID, "conv_gemm: res_ld % 8 != 0");
  CUtensorMap mo = ma;
  if (d->out) {
    const uint64_t o_rows = d->out_rows, o_ld = d->out_ld;
    const uint64_t o_samples = d->per_sample_tiles ? d->samples : 1;
    uint64_t o_plane = (uint64_t)d->out_plane_stride;
    if (g.out_planes == 1 || o_plane == 0) o_plane = o_samples * o_rows * o_ld;
    // (32-row boxes: every epilogue warp stores its own quarter of a tile)
    VP3D_TRY(make_map_4d(&mo, d->out, o_ld, o_rows, o_ld, o_samples, o_rows * o_ld, g.out_planes,
                         o_plane, 32));
  }
  // Residual through TMA (warp 3 prefetches each 128 x 64 residual tile into shared memory) whenever
  // the residual rows of a tile are one box of a strided row view:
  //   row(t) = sample*rows_per_sample + t*step + off  ==  view row (sample, t), column off*ld + c
  // with the view's rows `step*ld` elements long (strided layout: off < step), or rows shifted by
  // `off` (step == 1).  Other maps (flat tiles split over samples, bounds-checked rows) keep the
  // register path.
  CUtensorMap mr = ma;
  g.res_tma = 0;
  // (per-sample maps zero-fill rows outside the sample, which is exactly what res_check_rows asks)
  if (d->res && (!d->res_check_rows || (d->per_sample_tiles && d->res_row_step == 1)) && d->out &&
      (d->per_sample_tiles || d->res_sample_div == 0) &&
      d->res_row_step >= 1 && (d->res_row_step == 1 || d->res_row_off < d->res_row_step)) {
    const uint64_t step = d->res_row_step, ld = d->res_ld;
    const uint64_t r_samples = d->per_sample_tiles ? d->samples : 1;
    uint64_t view_rows, col_off;
    long long row_off;
    if (step == 1) {
      view_rows = d->per_sample_tiles ? (uint64_t)d->res_rows_per_sample
                                      : (uint64_t)d->out_rows + d->res_row_off;
      col_off = 0;
      row_off = d->res_row_off;
    } else {
      view_rows = d->per_sample_tiles ? (uint64_t)d->res_rows_per_sample / step
                                      : (uint64_t)d->out_rows;
      col_off = (uint64_t)d->res_row_off * ld;
      row_off = 0;
    }
    const uint64_t inner = step * ld;
    const uint64_t sample_stride = d->per_sample_tiles ? (uint64_t)d->res_rows_per_sample * ld
                                                       : view_rows * inner;
    uint64_t r_plane = (uint64_t)d->res_plane_stride;
    if (g.res_planes == 1 || r_plane == 0) r_plane = r_samples * sample_stride;
    if (view_rows > 0 &&
        make_map_4d(&mr, d->res, inner, view_rows, inner, r_samples, sample_stride, g.res_planes,
                    r_plane, kBlockM) == VP3D_OK) {
      g.res_tma = 1;
      g.res_tma_col_off = (int)col_off;
      g.res_tma_row_off = (int)row_off;
    }
  }
  // fused BatchNorm-backward reductions: Z rides the auxiliary TMA path with the output's geometry
  CUtensorMap mz = mo;
  g.bnb = 0;
  if (d->bnb_z) {
    if (!d->out || g.out_planes != 1 || (d->res && !g.res_tma) || d->bnb_c <= 0 || d->bnb_c % 64)
      return fail(VP3D_ERR_INVALID, "conv_gemm: fused BN-backward needs a single-plane bf16 output, "
                  "a TMA-loadable residual (if any) and bnb_c % 64 == 0");
    const uint64_t o_rows = d->out_rows, o_ld = d->out_ld;
    const uint64_t o_samples = d->per_sample_tiles ? d->samples : 1;
    VP3D_TRY(make_map_4d(&mz, d->bnb_z, o_ld, o_rows, o_ld, o_samples, o_rows * o_ld, 1,
                         o_samples * o_rows * o_ld, kBlockM));
    g.bnb = 1;
    g.bnb_c = d->bnb_c;
    g.bnb_scale = d->bnb_scale; g.bnb_shift = d->bnb_shift; g.bnb_mean = d->bnb_mean;
    g.bnb_invstd = d->bnb_invstd; g.bnb_sums = d->bnb_sums;
    g.bnb_p = d->bnb_p;
    g.bnb_seed_lo = (unsigned)(d->bnb_seed & 0xFFFFFFFFu);
    g.bnb_seed_hi = (unsigned)(d->bnb_seed >> 32);
    g.bnb_layer = (unsigned)d->bnb_layer;
  }
  // W boxes: the whole N block, or half of it per CTA when the launch runs on CTA pairs
  const bool pair = conv_gemm_uses_pair(g, block_n, num_sms());
  VP3D_TRY(make_map_2d(&mw, d->w, d->k_per_tap, (uint64_t)w_planes * d->taps * d->n_pad,
                       pair ? block_n / 2 : block_n));
  CUDA_TRY(launch_conv_gemm(ma, mw, mo, mr, mz, g, block_n, num_sms(), stream));
  return VP3D_OK;
}

// ------------------------------------------------------------------ plan
int plan_alloc(vp3d_plan* p, void** out, size_t bytes) {
  void* q = nullptr;
  CUDA_TRY(cudaMalloc(&q, bytes));
  p->allocs.push_back(q);
  *out = q;
  return VP3D_OK;
}

}  // namespace vp3d

static int alloc_packed(vp3d_plan* p, PackedConv& pc, int taps, int k_per_tap, int n_pad,
                        int merged) {
  pc.taps = taps;
  pc.k_per_tap = k_per_tap;
  pc.n_pad = n_pad;
  pc.merged = merged;
  const size_t elems = (size_t)p->planes * taps * n_pad * k_per_tap;
  VP3D_TRY(plan_alloc(p, reinterpret_cast<void**>(&pc.w), elems * 2));
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_version(void) { return VP3D_VERSION; }
extern "C" __attribute__((visibility("default"))) int vp3d_set_pdl(int on) {
  vp3d::conv_gemm_set_pdl(on);
  return VP3D_OK;
}
extern "C" __attribute__((visibility("default"))) int vp3d_set_sm_limit(int n) {
  if (n != 0 && n < 2) return fail(VP3D_ERR_INVALID, "set_sm_limit: need 0 (no limit) or >= 2 SMs");
  vp3d::g_sm_limit = n;
  return VP3D_OK;
}
extern "C" __attribute__((visibility("default"))) const char* vp3d_last_error(void) { return g_err; }
#ifdef VP3D_TIMELINE
// debug build only (`make dbg`, tools/timeline.py): device buffer of [max_launches][2][32][2] u64
extern "C" __attribute__((visibility("default"))) int vp3d_debug_set_timeline(unsigned long long* buf,
                                                                                int max_launches) {
  vp3d::conv_gemm_debug_set_timeline(buf, max_launches);
  return VP3D_OK;
}
#endif

extern "C" __attribute__((visibility("default"))) int vp3d_plan_create(const vp3d_config* cfg, vp3d_plan** out_plan) {
  if (!cfg || !out_plan) return fail(VP3D_ERR_INVALID, "plan_create: null argument");
  if (cfg->num_widths < 1 || cfg->num_widths > VP3D_MAX_WIDTHS)
    return fail(VP3D_ERR_INVALID, "plan_create: len(filter_widths) must be in [1, %d]",
                VP3D_MAX_WIDTHS);
  for (int i = 0; i < cfg->num_widths; ++i)
    if (cfg->filter_widths[i] < 1 || cfg->filter_widths[i] % 2 == 0)
      return fail(VP3D_ERR_INVALID, "Only odd filter widths are supported");  // model.py:20-21
  if (cfg->num_joints_in < 1 || cfg->in_features < 1 || cfg->num_joints_out < 1)
    return fail(VP3D_ERR_INVALID, "plan_create: joint / feature counts must be positive");
  if (cfg->channels < 1)
    return fail(VP3D_ERR_INVALID, "channels must be positive (got %d)", cfg->channels);
  if (cfg->precision < VP3D_PRECISION_BF16 || cfg->precision > VP3D_PRECISION_FP16)
    return fail(VP3D_ERR_INVALID, "plan_create: unknown precision %d", cfg->precision);
  if (cfg->variant != VP3D_VARIANT_DILATED && cfg->variant != VP3D_VARIANT_STRIDED)
    return fail(VP3D_ERR_INVALID, "plan_create: unknown variant %d", cfg->variant);
  if (cfg->variant == VP3D_VARIANT_STRIDED && cfg->dense)
    return fail(VP3D_ERR_INVALID, "dense=True only exists for TemporalModel");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaFree(0) != cudaSuccess)
    return fail(VP3D_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));

  vp3d_plan* p = new vp3d_plan();
  p->cfg = *cfg;
  p->nb = cfg->num_widths - 1;
  // any `channels` (the reference takes any -ch, arguments.py:47): activations and packed weights
  // are laid out with the channel count padded to 64; padding channels carry zero weights and a
  // zero affine, so they stay exactly zero through every layer
  p->c_real = cfg->channels;
  p->C = round_up(cfg->channels, 64);
  p->c_in_raw = cfg->num_joints_in * cfg->in_features;
  p->c_out_raw = cfg->num_joints_out * 3;
  p->c_in_pad = round_up(p->c_in_raw, 64);
  p->k0_pad = round_up(p->c_in_raw * cfg->filter_widths[0], 64);
  p->c_out_pad = round_up(p->c_out_raw, 64);
  p->f16 = cfg->precision == VP3D_PRECISION_FP16 ? 1 : 0;
  p->planes = (cfg->precision == VP3D_PRECISION_BF16 || p->f16) ? 1 : 2;
  // model.py:31, 107-121 / :172-184
  p->pad[0] = cfg->filter_widths[0] / 2;
  p->shift_dil[0] = p->shift_str[0] = cfg->causal ? cfg->filter_widths[0] / 2 : 0;
  p->dilation[0] = 1;
  p->taps[0] = cfg->filter_widths[0];
  int next_dilation = cfg->filter_widths[0];
  for (int i = 1; i < cfg->num_widths; ++i) {
    const int w = cfg->filter_widths[i];
    p->pad[i] = (w - 1) * next_dilation / 2;
    p->shift_dil[i] = cfg->causal ? (w / 2) * next_dilation : 0;
    p->shift_str[i] = cfg->causal ? (w / 2) : 0;
    p->dilation[i] = cfg->dense ? 1 : next_dilation;
    p->taps[i] = cfg->dense ? 2 * p->pad[i] + 1 : w;
    next_dilation *= w;
  }

  int st = VP3D_OK;
  do {
    if ((st = alloc_packed(p, p->expand_dil, cfg->filter_widths[0], p->c_in_pad, p->C, 0))) break;
    if ((st = alloc_packed(p, p->expand_flat, 1, p->k0_pad, p->C, 1))) break;
    for (int i = 0; i < p->nb && !st; ++i) {
      st = alloc_packed(p, p->conv[2 * i], p->taps[i + 1], p->C, p->C, 0);
      if (!st) st = alloc_packed(p, p->conv[2 * i + 1], 1, p->C, p->C, 0);
    }
    if (st) break;
    if ((st = alloc_packed(p, p->shrink, 1, p->C, p->c_out_pad, 0))) break;
    // affine vectors: expand + 2*nb layers (C each) + shrink (c_out_pad)
    float* aff = nullptr;
    const size_t n_aff = (size_t)(2 * p->nb + 1) * 2 * p->C + 2 * p->c_out_pad;
    if ((st = plan_alloc(p, reinterpret_cast<void**>(&aff), n_aff * sizeof(float)))) break;
    p->expand_dil.scale = p->expand_flat.scale = aff;
    p->expand_dil.shift = p->expand_flat.shift = aff + p->C;
    for (int l = 0; l < 2 * p->nb; ++l) {
      p->conv[l].scale = aff + (size_t)(l + 1) * 2 * p->C;
      p->conv[l].shift = p->conv[l].scale + p->C;
    }
    p->shrink.scale = aff + (size_t)(2 * p->nb + 1) * 2 * p->C;
    p->shrink.shift = p->shrink.scale + p->c_out_pad;
  } while (0);
  if (st) {
    vp3d_plan_destroy(p);
    return st;
  }
  *out_plan = p;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) void vp3d_plan_destroy(vp3d_plan* p) {
  if (!p) return;
  for (void* q : p->allocs) cudaFree(q);
  if (p->d_x) cudaFree(p->d_x);
  if (p->d_y) cudaFree(p->d_y);
  if (p->d_ws) cudaFree(p->d_ws);
  if (p->stream) cudaStreamDestroy(p->stream);
  for (auto& s : p->slots) {
    if (s.d_x) cudaFree(s.d_x);
    if (s.d_y) cudaFree(s.d_y);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  for (cudaEvent_t e : p->copy_events) cudaEventDestroy(e);
  for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
  if (p->train) train_state_destroy(p->train);
  delete p;
}

extern "C" __attribute__((visibility("default"))) int vp3d_receptive_field(const vp3d_plan* p) {
  if (!p) return fail(VP3D_ERR_INVALID, "null plan");
  int frames = 0;
  for (int i = 0; i < p->cfg.num_widths; ++i) frames += p->pad[i];
  return 1 + 2 * frames;
}

extern "C" __attribute__((visibility("default"))) int vp3d_total_causal_shift(const vp3d_plan* p) {
  if (!p) return fail(VP3D_ERR_INVALID, "null plan");
  const int* cs = p->cfg.variant == VP3D_VARIANT_STRIDED ? p->shift_str : p->shift_dil;
  int frames = cs[0];
  int next_dilation = p->cfg.filter_widths[0];
  for (int i = 1; i < p->cfg.num_widths; ++i) {
    frames += cs[i] * next_dilation;
    next_dilation *= p->cfg.filter_widths[i];
  }
  return frames;
}

extern "C" __attribute__((visibility("default"))) int vp3d_set_weights(vp3d_plan* p, const vp3d_weights* w, int what, void* stream_) {
  if (!p || !w) return fail(VP3D_ERR_INVALID, "set_weights: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int w0 = p->cfg.filter_widths[0];
  if (what & VP3D_PACK_CONV) {
    if (!w->expand_conv_weight || !w->shrink_weight)
      return fail(VP3D_ERR_INVALID, "set_weights: missing conv weights");
    CUDA_TRY(launch_pack_conv_weight(w->expand_conv_weight, p->expand_dil.w, p->planes, p->c_real,
                                     p->c_in_raw, w0, p->C, p->c_in_pad, 0, stream, p->f16));
    CUDA_TRY(launch_pack_conv_weight(w->expand_conv_weight, p->expand_flat.w, p->planes, p->c_real,
                                     p->c_in_raw, w0, p->C, p->k0_pad, 1, stream, p->f16));
    // with VP3D_PACK_CONV_T the transposed-pack kernels below also write these forward packs
    const bool fused = (what & VP3D_PACK_CONV_T) != 0;
    for (int i = 0; i < p->nb; ++i) {
      if (!w->layers_conv_weight[2 * i] || !w->layers_conv_weight[2 * i + 1])
        return fail(VP3D_ERR_INVALID, "set_weights: missing layers_conv.%d", 2 * i);
      if (fused) continue;
      CUDA_TRY(launch_pack_conv_weight(w->layers_conv_weight[2 * i], p->conv[2 * i].w, p->planes,
                                       p->c_real, p->c_real, p->taps[i + 1], p->C, p->C, 0, stream,
                                       p->f16));
      CUDA_TRY(launch_pack_conv_weight(w->layers_conv_weight[2 * i + 1], p->conv[2 * i + 1].w,
                                       p->planes, p->c_real, p->c_real, 1, p->C, p->C, 0, stream,
                                       p->f16));
    }
    if (!fused)
      CUDA_TRY(launch_pack_conv_weight(w->shrink_weight, p->shrink.w, p->planes, p->c_out_raw,
                                       p->c_real, 1, p->c_out_pad, p->C, 0, stream, p->f16));
    p->conv_packed = true;
  }
  if (what & VP3D_PACK_BN_EVAL) {
    const float eps = 1e-5f;  // nn.BatchNorm1d default, model.py:32
    for (int k = 0; k < 4; ++k)
      if (!w->expand_bn[k]) return fail(VP3D_ERR_INVALID, "set_weights: missing expand_bn");
    if (!w->shrink_bias) return fail(VP3D_ERR_INVALID, "set_weights: missing shrink.bias");
    CUDA_TRY(launch_bn_fold(w->expand_bn[0], w->expand_bn[1], w->expand_bn[2], w->expand_bn[3], eps,
                            p->expand_dil.scale, p->expand_dil.shift, p->c_real, p->C, stream));
    for (int l = 0; l < 2 * p->nb; ++l) {
      for (int k = 0; k < 4; ++k)
        if (!w->layers_bn[l][k]) return fail(VP3D_ERR_INVALID, "set_weights: missing layers_bn.%d", l);
      CUDA_TRY(launch_bn_fold(w->layers_bn[l][0], w->layers_bn[l][1], w->layers_bn[l][2],
                              w->layers_bn[l][3], eps, p->conv[l].scale, p->conv[l].shift,
                              p->c_real, p->C, stream));
    }
    CUDA_TRY(launch_bias_affine(w->shrink_bias, p->shrink.scale, p->shrink.shift, p->c_out_raw,
                                p->c_out_pad, stream));
    p->bn_packed = true;
  }
  if ((what & VP3D_PACK_CONV_T) && p->f16)
    return fail(VP3D_ERR_UNSUPPORTED, "fp16 plans are inference-only (train with bf16 / bf16x3)");
  if (what & VP3D_PACK_CONV_T)
    VP3D_TRY(train_pack_transposed(p, w, stream, (what & VP3D_PACK_CONV) != 0));
  return VP3D_OK;
}

// ------------------------------------------------------------------ eval schedules
// The strided ("flat") schedule is used for TemporalModelOptimized1f and for TemporalModel in eval
// mode when the input is exactly one receptive field long: with running statistics every output
// frame depends only on its own dependency cone, whose rows are exactly the stride-w rows
// Optimized1f computes (the reference states the weights are interchangeable, model.py:146-148).
namespace vp3d {
bool use_strided(const vp3d_plan* p, int T) {
  if (p->cfg.variant == VP3D_VARIANT_STRIDED) return true;
  return !p->cfg.dense && T == vp3d_receptive_field(p);
}

// rows per sample after each stage: L[0] = rows out of expand, L[i] = rows out of block i
int layer_rows(const vp3d_plan* p, int T, bool strided, int* L) {
  const int* fw = p->cfg.filter_widths;
  if (strided) {
    L[0] = T / fw[0];  // Conv1d(stride=w, kernel=w): floor((T - w)/w) + 1
    for (int i = 1; i <= p->nb; ++i) L[i] = L[i - 1] / fw[i];
    // (when T is not exactly one receptive field the floors drop trailing frames layer by layer;
    // strided_trim() below tells which rows the output actually depends on)
  } else {
    L[0] = T - (fw[0] - 1);
    for (int i = 1; i <= p->nb; ++i) L[i] = L[i - 1] - 2 * p->pad[i];
  }
  for (int i = 0; i <= p->nb; ++i)
    if (L[i] < 1) return 0;
  return L[p->nb];
}

}  // namespace vp3d

namespace vp3d {
// Strided model on an input whose length is not a multiple of the widths, as Conv1d(stride = w)
// handles it (model.py:167, 178, 191): every conv floors its output length, i.e. ignores trailing
// frames, and the residual slice x[:, :, shift + w//2 :: w] must come out with the conv's length
// (otherwise the reference's `res + x` raises a size mismatch -- reproduced here as an error).
// On success L[] is trimmed to the rows the output depends on: L[i-1] = w_i * L[i], so that the
// row-region (tap-major) schedule applies; in eval mode (running statistics) dropping the unused
// rows changes nothing.  Returns VP3D_OK / an error status with the reference's message.
int strided_trim(const vp3d_plan* p, int* L) {
  const int* fw = p->cfg.filter_widths;
  for (int i = 1; i <= p->nb; ++i) {
    const int first = p->shift_str[i] + fw[i] / 2;
    const int res_len = L[i - 1] > first ? (L[i - 1] - first + fw[i] - 1) / fw[i] : 0;
    if (res_len != L[i])
      return fail(VP3D_ERR_INVALID, "The size of tensor a (%d) must match the size of tensor b (%d) "
                  "at non-singleton dimension 2 (residual slice of block %d on %d frames, width %d)",
                  res_len, L[i], i, L[i - 1], fw[i]);
  }
  for (int i = p->nb; i >= 1; --i) L[i - 1] = fw[i] * L[i];
  return VP3D_OK;
}
}  // namespace vp3d

extern "C" __attribute__((visibility("default"))) int vp3d_output_frames(const vp3d_plan* p, int T) {
  if (!p) return fail(VP3D_ERR_INVALID, "null plan");
  int L[VP3D_MAX_WIDTHS];
  return layer_rows(p, T, p->cfg.variant == VP3D_VARIANT_STRIDED, L);
}

struct WsLayout {
  size_t a0 = 0, x0 = 0, x1 = 0, h = 0, total = 0;
  size_t a0_plane = 0, x_plane = 0, h_plane = 0;  // elements per plane
};

static WsLayout ws_layout(const vp3d_plan* p, int N, int T, bool strided, const int* L) {
  WsLayout w;
  const size_t a0_rows = strided ? (size_t)N * L[0] : (size_t)N * T;
  const size_t a0_ld = strided ? p->k0_pad : p->c_in_pad;
  w.a0_plane = a0_rows * a0_ld;
  w.x_plane = (size_t)N * L[0] * p->C;
  w.h_plane = p->nb > 0 ? (size_t)N * L[1] * p->C : 0;
  size_t off = 0;
  w.a0 = off; off = align_up(off + w.a0_plane * p->planes * 2, 1024);
  w.x0 = off; off = align_up(off + w.x_plane * p->planes * 2, 1024);
  w.x1 = off; off = align_up(off + w.h_plane * p->planes * 2, 1024);  // block outputs are <= L[1] rows
  w.h = off;  off = align_up(off + w.h_plane * p->planes * 2, 1024);
  w.total = off + 1024;
  return w;
}

extern "C" __attribute__((visibility("default"))) size_t vp3d_workspace_bytes(const vp3d_plan* p, int N, int T) {
  if (!p || N < 1) return 0;
  int L[VP3D_MAX_WIDTHS];
  const bool strided = use_strided(p, T);
  if (!layer_rows(p, T, strided, L)) return 0;
  if (strided && strided_trim(p, L) != VP3D_OK) return 0;
  return ws_layout(p, N, T, strided, L).total;
}

extern "C" __attribute__((visibility("default"))) int vp3d_forward_eval(vp3d_plan* p, const float* x, float* y, int N, int T, void* ws,
                                 size_t ws_bytes, void* stream_) {
  if (!p || !x || !y) return fail(VP3D_ERR_INVALID, "forward_eval: null argument");
  if (N < 1) return fail(VP3D_ERR_INVALID, "forward_eval: batch must be >= 1");
  if (!p->conv_packed || !p->bn_packed)
    return fail(VP3D_ERR_STATE, "forward_eval: vp3d_set_weights has not been called");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool strided = use_strided(p, T);
  int L[VP3D_MAX_WIDTHS];
  if (!layer_rows(p, T, strided, L))
    return fail(VP3D_ERR_INVALID, "forward_eval: sequence of %d frames is shorter than the "
                "receptive field (%d)", T, vp3d_receptive_field(p));
  if (strided) VP3D_TRY(strided_trim(p, L));   // trailing frames the strided convs ignore
  const WsLayout wl = ws_layout(p, N, T, strided, L);
  if (!ws || ws_bytes < wl.total) return fail(VP3D_ERR_WORKSPACE, "workspace too small: %zu < %zu",
                                              ws_bytes, wl.total);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  __nv_bfloat16* a0 = reinterpret_cast<__nv_bfloat16*>(base + wl.a0);
  __nv_bfloat16* xb[2] = {reinterpret_cast<__nv_bfloat16*>(base + wl.x0),
                          reinterpret_cast<__nv_bfloat16*>(base + wl.x1)};
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(base + wl.h);
  const int* fw = p->cfg.filter_widths;
  const int C = p->C;
  int launches = 0;
  // measurement hook: record an event pair around launch number p->prof_launch
  auto prof_event = [&](bool begin) -> int {
    if (p->prof_launch < 0 || launches != p->prof_launch) return VP3D_OK;
    const size_t idx = p->prof_used + (begin ? 0 : 1);
    while (p->prof_events.size() <= idx) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreate(&e));
      p->prof_events.push_back(e);
    }
    CUDA_TRY(cudaEventRecord(p->prof_events[idx], stream));
    if (!begin) p->prof_used += 2;
    return VP3D_OK;
  };
#define VP3D_LAUNCH(call)          \
  do {                             \
    VP3D_TRY(prof_event(true));    \
    call;                          \
    VP3D_TRY(prof_event(false));   \
    ++launches;                    \
  } while (0)

  // Per-layer operand precision.  index 0 = expand, 1..nb = residual blocks, nb+1 = shrink.
  //   bf16   : every GEMM single-plane bf16.
  //   bf16x3 : every GEMM split-bf16 (3 MMAs per product).
  //   mixed  : the residual stream X keeps hi+lo planes (skip path exact); expand and shrink run
  //            split-bf16 (they carry most of the bf16 error: profiles/r1_precision_study.txt),
  //            residual blocks run plain bf16 on the hi plane unless they hold < 0.5% of the
  //            forward FLOPs (negligible even at the narrow-tile rate of such layers).
  bool x3[VP3D_MAX_WIDTHS + 1];
  {
    double fl[VP3D_MAX_WIDTHS + 1], total = 0.0;
    fl[0] = (double)N * L[0] * p->c_in_raw * fw[0] * C;
    for (int i = 1; i <= p->nb; ++i) fl[i] = (double)N * L[i] * (p->taps[i] + 1.0) * C * C;
    fl[p->nb + 1] = (double)N * L[p->nb] * C * p->c_out_raw;
    for (int i = 0; i <= p->nb + 1; ++i) total += fl[i];
    for (int i = 0; i <= p->nb + 1; ++i) {
      if (p->cfg.precision == VP3D_PRECISION_BF16 || p->f16) x3[i] = false;
      else if (p->cfg.precision == VP3D_PRECISION_BF16X3) x3[i] = true;
      else x3[i] = (i == 0 || i == p->nb + 1) ? true : (fl[i] < 0.005 * total);
    }
  }

  vp3d_conv_desc d;
  auto common = [&](vp3d_conv_desc& q, bool layer_x3) {
    memset(&q, 0, sizeof(q));
    q.a_planes = p->planes;
    q.precision = p->f16 ? VP3D_PRECISION_FP16
                         : (layer_x3 ? VP3D_PRECISION_BF16X3 : VP3D_PRECISION_BF16);
    q.out_planes = p->planes;
    q.res_planes = p->planes;
  };

  // ---- input packing + expand conv (model.py:127 / :188)
  // Strided schedule: every activation is kept in tap-major row order (pack.cuh), so the w taps
  // of block i are w contiguous row regions of R[i] = N * L[i] rows: tap k of output row j is row
  // k * R[i] + j of the block input, and the residual of the block is its centre (or, causal, last)
  // region.  Only that region of X needs the lo plane in `mixed` mode.
  long long R[VP3D_MAX_WIDTHS];
  for (int i = 0; i <= p->nb; ++i) R[i] = (long long)N * L[i];
  auto lo_rows = [&](int i, vp3d_conv_desc& q) {  // q produces X_i
    q.lo_row_begin = 0;
    q.lo_row_end = 0;  // every row
    if (!strided || p->planes != 2 || i >= p->nb || x3[i + 1]) return;
    const long long c = fw[i + 1] / 2 + p->shift_str[i + 1];
    q.lo_row_begin = (int)(c * R[i + 1]);
    q.lo_row_end = (int)((c + 1) * R[i + 1]);
  };
  if (strided) {
    if ((long long)N * L[0] > 0x7fffffffll)
      return fail(VP3D_ERR_UNSUPPORTED, "forward_eval: too many rows (%lld)", (long long)N * L[0]);
    PackPerm perm;
    memset(&perm, 0, sizeof(perm));
    perm.levels = p->nb;
    perm.last_rows = L[p->nb];
    for (int i = 1; i <= p->nb; ++i) {
      perm.region[i - 1] = (unsigned)R[i];
      perm.width[i - 1] = fw[i];
    }
    VP3D_LAUNCH(CUDA_TRY(launch_pack_input(x, a0, p->planes, N, T, p->c_in_raw, L[0], fw[0], fw[0], p->k0_pad,
                               (long long)wl.a0_plane, stream, &perm, p->f16)));
    common(d, x3[0]);
    d.a = a0; d.samples = 1; d.a_rows = N * L[0]; d.a_ld = p->k0_pad;
    d.w = p->expand_flat.w; d.taps = 1; d.k_per_tap = p->k0_pad; d.n_pad = C;
    d.per_sample_tiles = 0; d.out_rows = N * L[0];
  } else {
    VP3D_LAUNCH(CUDA_TRY(launch_pack_input(x, a0, p->planes, N, T, p->c_in_raw, T, 1, 1, p->c_in_pad,
                               (long long)wl.a0_plane, stream, nullptr, p->f16)));
    common(d, x3[0]);
    d.a = a0; d.samples = N; d.a_rows = T; d.a_ld = p->c_in_pad;
    d.w = p->expand_dil.w; d.taps = fw[0]; d.k_per_tap = p->c_in_pad; d.n_pad = C;
    d.per_sample_tiles = 1; d.tap_row_step = 1; d.out_rows = L[0];
  }
  d.scale = p->expand_dil.scale; d.shift = p->expand_dil.shift; d.relu = 1;
  d.out = xb[0]; d.out_plane_stride = (long long)wl.x_plane; d.out_ld = C;
  lo_rows(0, d);
  VP3D_LAUNCH(VP3D_TRY(run_conv(&d, stream)));

  // ---- residual blocks (model.py:129-135 / :190-194)
  int cur = 0;
  size_t cur_plane = wl.x_plane;
  for (int i = 1; i <= p->nb; ++i) {
    const PackedConv& c0 = p->conv[2 * (i - 1)];
    const PackedConv& c1 = p->conv[2 * (i - 1) + 1];
    const int Lin = L[i - 1], Lout = L[i];
    const size_t h_plane = (size_t)N * Lout * C;
    // first conv of the block: dilated / strided k-tap conv + BN + ReLU
    common(d, x3[i]);
    d.out_planes = x3[i] ? 2 : 1;  // H only needs a lo plane when its consumer is split-bf16
    d.a = xb[cur];
    d.w = c0.w; d.taps = c0.taps; d.k_per_tap = C; d.n_pad = C;
    d.scale = c0.scale; d.shift = c0.shift; d.relu = 1;
    d.out = hb; d.out_plane_stride = (long long)h_plane; d.out_ld = C;
    if (strided) {
      d.tap_col_step = 0; d.tap_row_step = (int)R[i];  // tap k = row region k of the block input
      d.samples = 1; d.a_rows = N * Lin; d.a_ld = C;
      d.per_sample_tiles = 0; d.out_rows = N * Lout;
    } else {
      d.samples = N; d.a_rows = Lin; d.a_ld = C;
      d.per_sample_tiles = 1; d.tap_row_step = p->dilation[i]; d.tap_col_step = 0;
      d.out_rows = Lout;
    }
    // plane stride of A is implied by (samples, a_rows, a_ld) == cur_plane by construction
    if ((size_t)d.samples * d.a_rows * d.a_ld != cur_plane)
      return fail(VP3D_ERR_STATE, "internal: activation plane mismatch in block %d", i);
    VP3D_LAUNCH(VP3D_TRY(run_conv(&d, stream)));

    // second conv: 1x1 + BN + ReLU + sliced residual
    common(d, x3[i]);
    d.a_planes = x3[i] ? 2 : 1;
    d.a = hb; d.samples = 1; d.a_rows = N * Lout; d.a_ld = C;
    d.w = c1.w; d.taps = 1; d.k_per_tap = C; d.n_pad = C;
    d.per_sample_tiles = 0; d.out_rows = N * Lout;
    d.scale = c1.scale; d.shift = c1.shift; d.relu = 1;
    d.res = xb[cur]; d.res_plane_stride = (long long)cur_plane; d.res_ld = C;
    if (strided) {
      d.res_rows_per_sample = 0; d.res_row_step = 1;
      d.res_row_off = (int)((fw[i] / 2 + p->shift_str[i]) * R[i]); d.res_sample_div = 0;
    } else {
      // per-sample tiles: the residual rows of a tile are then one TMA box of the block input
      d.samples = N; d.a_rows = Lout; d.per_sample_tiles = 1; d.out_rows = Lout;
      d.res_rows_per_sample = Lin; d.res_row_step = 1;
      d.res_row_off = p->pad[i] + p->shift_dil[i]; d.res_sample_div = 0;
    }
    d.out = xb[cur ^ 1]; d.out_plane_stride = (long long)h_plane; d.out_ld = C;
    lo_rows(i, d);
    VP3D_LAUNCH(VP3D_TRY(run_conv(&d, stream)));
    cur ^= 1;
    cur_plane = h_plane;
  }

  // ---- shrink (model.py:137 / :196) writing (N, T_out, J_out, 3) directly (fuses :74-75)
  common(d, x3[p->nb + 1]);
  d.a = xb[cur]; d.samples = 1; d.a_rows = N * L[p->nb]; d.a_ld = C;
  d.w = p->shrink.w; d.taps = 1; d.k_per_tap = C; d.n_pad = p->c_out_pad;
  d.per_sample_tiles = 0; d.out_rows = N * L[p->nb];
  d.scale = p->shrink.scale; d.shift = p->shrink.shift; d.relu = 0;
  d.out = nullptr; d.out_f32 = y; d.out_f32_ld = p->c_out_raw; d.n_valid = p->c_out_raw;
  VP3D_LAUNCH(VP3D_TRY(run_conv(&d, stream)));
  p->last_launches = launches;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_forward_eval_host(vp3d_plan* p, const float* x_host, float* y_host, int N,
                                      int T) {
  if (!p || !x_host || !y_host) return fail(VP3D_ERR_INVALID, "forward_eval_host: null argument");
  const int t_out = vp3d_output_frames(p, T);
  if (t_out < 1 || N < 1) return fail(VP3D_ERR_INVALID, "forward_eval_host: bad shape");
  if (!p->stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  const size_t xb = (size_t)N * T * p->c_in_raw * sizeof(float);
  const size_t yb = (size_t)N * t_out * p->c_out_raw * sizeof(float);
  const size_t wb = vp3d_workspace_bytes(p, N, T);
  if (xb > p->d_x_bytes) {
    if (p->d_x) cudaFree(p->d_x);
    p->d_x = nullptr; p->d_x_bytes = 0;
    CUDA_TRY(cudaMalloc(&p->d_x, xb));
    p->d_x_bytes = xb;
  }
  if (yb > p->d_y_bytes) {
    if (p->d_y) cudaFree(p->d_y);
    p->d_y = nullptr; p->d_y_bytes = 0;
    CUDA_TRY(cudaMalloc(&p->d_y, yb));
    p->d_y_bytes = yb;
  }
  if (wb > p->d_ws_bytes) {
    if (p->d_ws) cudaFree(p->d_ws);
    p->d_ws = nullptr; p->d_ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&p->d_ws, wb));
    p->d_ws_bytes = wb;
  }
  // Batch rows are independent in eval mode: split the batch into chunks so that the host->device
  // copy of chunk i+1 (copy stream) overlaps the kernels of chunk i (compute stream).  PCIe moves
  // 33.8 MB per 1024 x 243 batch, which is longer than the whole forward.
  int chunks = N >= 512 ? 2 : 1;
  if (const char* env = getenv("VP3D_HOST_CHUNKS")) {  // measurement knob
    const int c = atoi(env);
    if (c >= 1 && c <= 16 && c <= N) chunks = c;
  }
  if (chunks == 1) {
    CUDA_TRY(cudaMemcpyAsync(p->d_x, x_host, xb, cudaMemcpyHostToDevice, p->stream));
    VP3D_TRY(vp3d_forward_eval(p, p->d_x, p->d_y, N, T, p->d_ws, p->d_ws_bytes, p->stream));
  } else {
    if (!p->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    while ((int)p->copy_events.size() < chunks) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p->copy_events.push_back(e);
    }
    const size_t x_row = (size_t)T * p->c_in_raw, y_row = (size_t)t_out * p->c_out_raw;
    int launches = 0;
    // the previous call's kernels may still read d_x: order the copies behind them
    CUDA_TRY(cudaEventRecord(p->copy_events[0], p->stream));
    CUDA_TRY(cudaStreamWaitEvent(p->copy_stream, p->copy_events[0], 0));
    for (int c = 0; c < chunks; ++c) {
      const int n0 = (int)((long long)N * c / chunks), n1 = (int)((long long)N * (c + 1) / chunks);
      CUDA_TRY(cudaMemcpyAsync(p->d_x + n0 * x_row, x_host + n0 * x_row,
                               (size_t)(n1 - n0) * x_row * sizeof(float), cudaMemcpyHostToDevice,
                               p->copy_stream));
      CUDA_TRY(cudaEventRecord(p->copy_events[c], p->copy_stream));
    }
    for (int c = 0; c < chunks; ++c) {
      const int n0 = (int)((long long)N * c / chunks), n1 = (int)((long long)N * (c + 1) / chunks);
      CUDA_TRY(cudaStreamWaitEvent(p->stream, p->copy_events[c], 0));
      VP3D_TRY(vp3d_forward_eval(p, p->d_x + n0 * x_row, p->d_y + n0 * y_row, n1 - n0, T, p->d_ws,
                                 p->d_ws_bytes, p->stream));
      launches += p->last_launches;
    }
    p->last_launches = launches;
  }
  CUDA_TRY(cudaMemcpyAsync(y_host, p->d_y, yb, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return VP3D_OK;
}

// Pipelined host API: submit() enqueues H2D (copy stream) -> forward (compute stream) -> D2H for one
// batch and returns immediately; wait() blocks until that batch's output is in y_host.  With two
// slots the PCIe copy of batch i+1 overlaps the kernels of batch i, so the sustained rate is
// max(copy, compute) instead of their sum.
extern "C" __attribute__((visibility("default"))) int vp3d_forward_eval_host_submit(
    vp3d_plan* p, const float* x_host, float* y_host, int N, int T, int slot) {
  if (!p || !x_host || !y_host) return fail(VP3D_ERR_INVALID, "host_submit: null argument");
  if (slot < 0 || slot > 1) return fail(VP3D_ERR_INVALID, "host_submit: slot must be 0 or 1");
  const int t_out = vp3d_output_frames(p, T);
  if (t_out < 1 || N < 1) return fail(VP3D_ERR_INVALID, "host_submit: bad shape");
  vp3d_plan::HostSlot& s = p->slots[slot];
  if (s.busy) return fail(VP3D_ERR_STATE, "host_submit: slot %d still in flight (call wait first)", slot);
  if (!p->stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  if (!p->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
  if (!s.copied) {
    CUDA_TRY(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  const size_t xb = (size_t)N * T * p->c_in_raw * sizeof(float);
  const size_t yb = (size_t)N * t_out * p->c_out_raw * sizeof(float);
  const size_t wb = vp3d_workspace_bytes(p, N, T);
  if (xb > s.x_bytes) {
    if (s.d_x) cudaFree(s.d_x);
    s.d_x = nullptr; s.x_bytes = 0;
    CUDA_TRY(cudaMalloc(&s.d_x, xb));
    s.x_bytes = xb;
  }
  if (yb > s.y_bytes) {
    if (s.d_y) cudaFree(s.d_y);
    s.d_y = nullptr; s.y_bytes = 0;
    CUDA_TRY(cudaMalloc(&s.d_y, yb));
    s.y_bytes = yb;
  }
  if (wb > p->d_ws_bytes) {
    CUDA_TRY(cudaStreamSynchronize(p->stream));  // the other slot may be using the old workspace
    if (p->d_ws) cudaFree(p->d_ws);
    p->d_ws = nullptr; p->d_ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&p->d_ws, wb));
    p->d_ws_bytes = wb;
  }
  CUDA_TRY(cudaMemcpyAsync(s.d_x, x_host, xb, cudaMemcpyHostToDevice, p->copy_stream));
  CUDA_TRY(cudaEventRecord(s.copied, p->copy_stream));
  CUDA_TRY(cudaStreamWaitEvent(p->stream, s.copied, 0));
  VP3D_TRY(vp3d_forward_eval(p, s.d_x, s.d_y, N, T, p->d_ws, p->d_ws_bytes, p->stream));
  CUDA_TRY(cudaMemcpyAsync(y_host, s.d_y, yb, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaEventRecord(s.done, p->stream));
  s.busy = true;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_forward_eval_host_wait(vp3d_plan* p,
                                                                                 int slot) {
  if (!p || slot < 0 || slot > 1) return fail(VP3D_ERR_INVALID, "host_wait: bad argument");
  vp3d_plan::HostSlot& s = p->slots[slot];
  if (!s.busy) return fail(VP3D_ERR_STATE, "host_wait: slot %d has nothing in flight", slot);
  CUDA_TRY(cudaEventSynchronize(s.done));
  s.busy = false;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_last_launch_count(const vp3d_plan* p) { return p ? p->last_launches : 0; }

extern "C" __attribute__((visibility("default"))) int vp3d_profile_launch(vp3d_plan* p, int launch_index) {
  if (!p) return fail(VP3D_ERR_INVALID, "null plan");
  p->prof_launch = launch_index;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_profile_read(vp3d_plan* p, float* total_ms, int* count) {
  if (!p || !total_ms || !count) return fail(VP3D_ERR_INVALID, "profile_read: null argument");
  float sum = 0.0f;
  int n = 0;
  for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
    CUDA_TRY(cudaEventSynchronize(p->prof_events[i + 1]));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->prof_events[i], p->prof_events[i + 1]));
    sum += ms;
    ++n;
  }
  p->prof_used = 0;
  *total_ms = sum;
  *count = n;
  return VP3D_OK;
}

extern "C" __attribute__((visibility("default"))) int vp3d_conv_gemm(const vp3d_conv_desc* d, void* stream) {
  return run_conv(d, static_cast<cudaStream_t>(stream));
}
