// Internals shared by the C-ABI translation units (api.cu: plan + eval; train_api.cu: training).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/vp3d_b200.h"
#include "conv_gemm.cuh"
#include "pack.cuh"

namespace vp3d {

int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::vp3d::fail(VP3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                     \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)
#define VP3D_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != VP3D_OK) return _s; \
  } while (0)

// 4-D bf16 map (k, row, sample, plane), box (64, box_rows, 1, 1), 128-byte swizzle.
int make_map_4d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride,
                uint64_t samples, uint64_t sample_stride, uint64_t planes, uint64_t plane_stride,
                uint32_t box_rows);
// 2-D bf16 map (k, row), box (64, box_rows), 128-byte swizzle.
int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint32_t box_rows);

int pick_block_n(int n_pad);
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int num_sms();
int run_conv(const vp3d_conv_desc* d, cudaStream_t stream);


struct PackedConv {
  __nv_bfloat16* w = nullptr;
  int taps = 0;       // taps as stored (1 when merged)
  int k_per_tap = 0;  // padded
  int n_pad = 0;
  int merged = 0;
  float* scale = nullptr;  // eval affine [n_pad]
  float* shift = nullptr;
};

struct TrainState;  // train_api.cu

// step_ops.cu: Adam / AMSGrad update of conv weights that also refreshes their bf16 packs
struct AdamPackItem {
  vp3d_adam_tensor t;
  __nv_bfloat16* fwd;   // forward pack [planes][taps][fwd_n_pad][fwd_k_pad] or null
  __nv_bfloat16* tr;    // transposed pack [planes][taps][tr_n_pad][tr_k_pad] or null
  int c_out, c_in, taps;
  int fwd_n_pad, fwd_k_pad, tr_n_pad, tr_k_pad;
};
int launch_adam_pack(const AdamPackItem* items, int n, int planes, int64_t step, double lr,
                     double beta1, double beta2, double eps, double weight_decay,
                     cudaStream_t stream);

}  // namespace vp3d

struct vp3d_plan {
  vp3d_config cfg;
  int nb = 0;  // residual blocks
  int C = 0;        // channels of the residual stream as laid out in memory: padded to 64
  int c_real = 0;   // the model's `channels` argument (any positive value, model.py:85-86)
  int c_in_raw = 0, c_out_raw = 0, c_in_pad = 0, k0_pad = 0, c_out_pad = 0;
  int planes = 1;
  int f16 = 0;  // VP3D_PRECISION_FP16: 16-bit stores hold IEEE fp16
  int pad[VP3D_MAX_WIDTHS];
  int shift_dil[VP3D_MAX_WIDTHS];  // causal shift in frames (TemporalModel, model.py:111)
  int shift_str[VP3D_MAX_WIDTHS];  // causal shift in strided units (Optimized1f, model.py:176)
  int dilation[VP3D_MAX_WIDTHS];
  int taps[VP3D_MAX_WIDTHS];       // taps of block i's first conv (dense: 2*pad+1)
  vp3d::PackedConv expand_dil, expand_flat, shrink;
  vp3d::PackedConv conv[VP3D_MAX_LAYERS];
  std::vector<void*> allocs;
  bool conv_packed = false, bn_packed = false;
  // host-API staging (owned)
  float* d_x = nullptr;
  float* d_y = nullptr;
  void* d_ws = nullptr;
  size_t d_x_bytes = 0, d_y_bytes = 0, d_ws_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;     // host API: H2D chunks overlap the compute stream
  std::vector<cudaEvent_t> copy_events;
  int last_launches = 0;
  // pipelined host API (vp3d_forward_eval_host_submit / _wait): two independent staging slots
  struct HostSlot {
    float* d_x = nullptr;
    float* d_y = nullptr;
    size_t x_bytes = 0, y_bytes = 0;
    cudaEvent_t copied = nullptr, done = nullptr;
    bool busy = false;
  };
  HostSlot slots[2];
  // measurement hook: event pairs around one chosen launch of each forward
  int prof_launch = -1;
  std::vector<cudaEvent_t> prof_events;  // start/stop pairs
  size_t prof_used = 0;                  // events consumed since the last read
  // training-mode state (transposed weight packs, per-layer BN vectors, dropout config)
  vp3d::TrainState* train = nullptr;
};

namespace vp3d {
int plan_alloc(vp3d_plan* p, void** out, size_t bytes);
bool use_strided(const vp3d_plan* p, int T);
// rows per sample after each stage: L[0] = rows out of expand, L[i] = rows out of block i
int layer_rows(const vp3d_plan* p, int T, bool strided, int* L);
int strided_trim(const vp3d_plan* p, int* L);
void train_state_destroy(TrainState* t);
int train_pack_transposed(vp3d_plan* p, const vp3d_weights* w, cudaStream_t stream,
                          bool also_forward);
}  // namespace vp3d
