// Device-resident pose store -> batch windows.
//
// Replaces the per-chunk Python loop of the reference generators (common/generators.py:99-160
// ChunkedGenerator.next_epoch, :213-240 UnchunkedGenerator.next_epoch): every sequence lives in HBM
// once; a batch is produced by one launch that reads a row table (sequence, first frame, end
// frame, flip) and writes the (window, frame, joint, feature) fp32 tensor the model consumes.
//   * frames outside the sequence replicate the nearest edge frame (np.pad(..., 'edge'), :108-118)
//   * flip negates feature 0 and reads each joint from its mirror partner (:120-123, :137-143)
//   * cameras: entries 2 and 7 change sign under flip (:147-152)
// HBM-bound integer/index work: one coalesced 4-byte store per element, reads served mostly by L2
// (windows of neighbouring frames overlap).  Bit-exact by construction (copy + sign flip).
#include "internal.cuh"

namespace vp3d {
namespace {

constexpr int kGatherThreads = 256;
constexpr int kGatherFrames = 128;  // frames per block (amortises the per-block table set-up)
constexpr int kMaxRowElems = 256;   // joints * features supported by the shared index table

// div_magic = floor(2^32 / row_elems) + 1: floor(i / row_elems) == umulhi(i, div_magic) for every
// i < 2^32 / row_elems, which covers a tile (i < kGatherFrames * kMaxRowElems = 2^15).
__global__ void __launch_bounds__(kGatherThreads)
gather_windows_kernel(const float* __restrict__ src, const long long* __restrict__ seq_first,
                      const int* __restrict__ seq_len, const int* __restrict__ rows,
                      const int* __restrict__ src_joint, float* __restrict__ out, int frames,
                      int joints, int features, int first_offset, int frame_tiles,
                      unsigned div_magic) {
  __shared__ int s_src[kMaxRowElems];      // source element of each output element of a frame
  __shared__ float s_sign[kMaxRowElems];
  __shared__ int s_frame[kGatherFrames];   // clamped source frame of each frame of the tile
  const int w = blockIdx.x / frame_tiles;
  const int tile = blockIdx.x - w * frame_tiles;
  const int* row = rows + 4ll * w;
  const int seq = row[0];
  const bool flip = row[3] != 0;
  const int row_elems = joints * features;
  const int t0 = tile * kGatherFrames;
  const int n_frames = min(frames - t0, kGatherFrames);
  for (int c = threadIdx.x; c < row_elems; c += kGatherThreads) {
    const int j = c / features;
    const int f = c - j * features;
    const int sj = (flip && src_joint != nullptr) ? src_joint[j] : j;
    s_src[c] = sj * features + f;
    s_sign[c] = (flip && f == 0) ? -1.0f : 1.0f;
  }
  if (threadIdx.x < n_frames) {
    const int fr = row[1] + first_offset + t0 + (int)threadIdx.x;
    s_frame[threadIdx.x] = min(max(fr, 0), seq_len[seq] - 1);
  }
  __syncthreads();
  const float* base = src + seq_first[seq] * row_elems;
  float* dst = out + ((long long)w * frames + t0) * row_elems;
  const unsigned n = (unsigned)(n_frames * row_elems);
  for (unsigned i = threadIdx.x; i < n; i += kGatherThreads) {
    const unsigned t = div_magic != 0u ? __umulhi(i, div_magic) : i;  // 0: row_elems == 1
    const unsigned c = i - t * (unsigned)row_elems;
    dst[i] = s_sign[c] * __ldg(base + (long long)s_frame[t] * row_elems + s_src[c]);
  }
}

__global__ void gather_cameras_kernel(const float* __restrict__ cams, int cam_dim,
                                      const int* __restrict__ rows, int n_windows,
                                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_windows * cam_dim) return;
  const int w = i / cam_dim;
  const int k = i - w * cam_dim;
  const int* row = rows + 4ll * w;
  const float v = cams[(long long)row[0] * cam_dim + k];
  out[i] = (row[3] != 0 && (k == 2 || k == 7)) ? -v : v;
}

}  // namespace
}  // namespace vp3d

#define VP3D_EXPORT extern "C" __attribute__((visibility("default")))

VP3D_EXPORT int vp3d_gather_windows(const vp3d_gather_desc* d, void* stream) {
  using namespace vp3d;
  if (d == nullptr) return fail(VP3D_ERR_INVALID, "vp3d_gather_windows: null descriptor");
  if (d->n_windows < 0 || d->frames < 0 || d->joints <= 0 || d->features <= 0)
    return fail(VP3D_ERR_INVALID, "vp3d_gather_windows: bad sizes (windows %d, frames %d, joints %d, "
                "features %d)", d->n_windows, d->frames, d->joints, d->features);
  if (d->joints * d->features > kMaxRowElems)
    return fail(VP3D_ERR_UNSUPPORTED, "vp3d_gather_windows: joints*features = %d > %d",
                d->joints * d->features, kMaxRowElems);
  if (d->n_windows == 0 || d->frames == 0) return VP3D_OK;
  if (d->src == nullptr || d->seq_first == nullptr || d->seq_len == nullptr || d->rows == nullptr ||
      d->out == nullptr)
    return fail(VP3D_ERR_INVALID, "vp3d_gather_windows: null pointer");
  const int tiles = (d->frames + kGatherFrames - 1) / kGatherFrames;
  const long long blocks = (long long)tiles * d->n_windows;
  if (blocks > 0x7fffffffll)
    return fail(VP3D_ERR_UNSUPPORTED, "vp3d_gather_windows: %lld blocks", blocks);
  gather_windows_kernel<<<(unsigned)blocks, kGatherThreads, 0, (cudaStream_t)stream>>>(
      d->src, (const long long*)d->seq_first, d->seq_len, d->rows, d->src_joint, d->out, d->frames,
      d->joints, d->features, d->first_offset, tiles,
      d->joints * d->features == 1 ? 0u : 0xffffffffu / (unsigned)(d->joints * d->features) + 1u);
  CUDA_TRY(cudaGetLastError());
  return VP3D_OK;
}

VP3D_EXPORT int vp3d_gather_cameras(const float* cams, int32_t cam_dim, const int32_t* rows,
                                 int32_t n_windows, float* out, void* stream) {
  using namespace vp3d;
  if (n_windows < 0 || cam_dim <= 0) return fail(VP3D_ERR_INVALID, "vp3d_gather_cameras: bad sizes");
  if (n_windows == 0) return VP3D_OK;
  if (cams == nullptr || rows == nullptr || out == nullptr)
    return fail(VP3D_ERR_INVALID, "vp3d_gather_cameras: null pointer");
  const int n = n_windows * cam_dim;
  gather_cameras_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cams, cam_dim, rows,
                                                                           n_windows, out);
  CUDA_TRY(cudaGetLastError());
  return VP3D_OK;
}
