"""Device-resident batch generators -- drop-ins for `common/generators.py`.

The reference builds every training batch on the host: a Python loop slices, edge-pads and mirrors
one chunk at a time into a float64 buffer (generators.py:99-160), `run.py` casts it to float32 and
copies it to the GPU (run.py:401-403) -- tens of milliseconds per 1024-window batch next to a model
step of a few.  Here all sequences are uploaded ONCE (fp32, back to back); per epoch the host only
draws the permutation of the (sequence, first, end, flip) table exactly as the reference does
(same `RandomState`, same draw, generators.py:89-97) and uploads it; every batch is then one gather
launch per tensor (`vp3d_gather_windows` / `vp3d_gather_cameras`, csrc/gather.cu) that writes the
float32 `(N, T, J, F)` tensor the model consumes.  No host->device copy remains in the step.

Same constructor arguments, methods and batch order as the reference classes; differences:
  * batches are CUDA float32 tensors (fresh tensors, safe to mutate), i.e. what run.py produces
    from the reference's float64 NumPy buffers with `.astype('float32')` + `.cuda()`;
  * optional `rank` / `world_size`: each batch's rows are split contiguously across ranks (§8e);
  * a CUDA device is required -- there is no host fallback.
"""
import ctypes

import numpy as np
import torch

from . import _capi

__all__ = ["ChunkedGenerator", "UnchunkedGenerator", "chunk_table", "mirror_source", "shard_rows"]


# ---------------------------------------------------------------------------------------------
# host-side index logic (pure NumPy; unit-tested on CPU against the oracle and the reference)
# ---------------------------------------------------------------------------------------------

def chunk_table(lengths, chunk_length, augment):
    """(P, 4) int64 table of (sequence, first frame, end frame, flip), ordered as the reference's
    `pairs` list (generators.py:39-48): for each sequence its chunks on a grid centred on the
    sequence, followed -- with `augment` -- by the same chunks flagged for mirroring."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n_chunks = (lengths + chunk_length - 1) // chunk_length
    offsets = (n_chunks * chunk_length - lengths) // 2
    copies = 2 if augment else 1
    total = int(n_chunks.sum()) * copies
    table = np.empty((total, 4), dtype=np.int64)
    at = 0
    for s in range(len(lengths)):
        k = int(n_chunks[s])
        first = np.arange(k, dtype=np.int64) * chunk_length - offsets[s]
        for flip in range(copies):
            blk = table[at:at + k]
            blk[:, 0] = s
            blk[:, 1] = first
            blk[:, 2] = first + chunk_length
            blk[:, 3] = flip
            at += k
    return table


def mirror_source(n_joints, left, right):
    """int32 map `src` with mirrored[..., j, :] = plain[..., src[j], :].  The reference performs
    `x[:, left + right] = x[:, right + left]` (generators.py:123, 142-143): destination i-th of
    left+right takes source i-th of right+left, later assignments overriding earlier ones."""
    if left is None or right is None:
        raise ValueError("augment=True needs the left/right joint lists (generators.py:123)")
    src = np.arange(n_joints, dtype=np.int32)
    for dst, s in zip(list(left) + list(right), list(right) + list(left)):
        src[dst] = s
    return src


def shard_rows(lo, hi, rank, world_size):
    """Contiguous sub-range of batch rows [lo, hi) owned by `rank`."""
    n = hi - lo
    return lo + n * rank // world_size, lo + n * (rank + 1) // world_size


class _EpochPlanner:
    """Epoch order and resumable state of the chunked stream (generators.py:89-97, 154-160)."""

    def __init__(self, lengths, batch_size, chunk_length, shuffle, random_seed, augment, endless):
        self.pairs = chunk_table(lengths, chunk_length, augment)
        self.batch_size = batch_size
        self.num_batches = (len(self.pairs) + batch_size - 1) // batch_size
        self.shuffle = shuffle
        self.endless = endless
        self.random = np.random.RandomState(random_seed)
        self.state = None

    def begin(self):
        """(first batch index, (P, 4) row order).  A new epoch draws ONE permutation of the table
        from `self.random` -- the same call on the same (P, 4) integer array the reference's
        `self.random.permutation(self.pairs)` reduces to, so seeds reproduce its order exactly."""
        if self.state is not None:
            return self.state
        return 0, (self.random.permutation(self.pairs) if self.shuffle else self.pairs)

    def batch_bounds(self, b):
        return b * self.batch_size, min((b + 1) * self.batch_size, len(self.pairs))


# ---------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------

def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("videopose3d_b200.generators needs a CUDA device: the batch gather runs "
                           "in csrc/gather.cu and has no host fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"videopose3d_b200.generators: device must be CUDA, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _PoseStore:
    """All sequences of one kind (2-D keypoints or 3-D joints) back to back in device memory."""

    def __init__(self, sequences, device):
        shapes = {tuple(s.shape[1:]) for s in sequences}
        if len(shapes) != 1:
            raise ValueError(f"sequences disagree on (joints, features): {sorted(shapes)}")
        self.joints, self.features = next(iter(shapes))
        lens = np.array([s.shape[0] for s in sequences], dtype=np.int64)
        if (lens < 1).any():
            raise ValueError("empty sequence (the reference's np.pad(..., 'edge') rejects it too)")
        first = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        flat = np.concatenate([np.asarray(s, dtype=np.float32) for s in sequences], axis=0)
        self.lengths = lens
        self.data = torch.from_numpy(np.ascontiguousarray(flat)).to(device)
        self.seq_first = torch.from_numpy(first).to(device)
        self.seq_len = torch.from_numpy(lens.astype(np.int32)).to(device)
        self.src_joint = None

    def set_mirror(self, left, right, device):
        self.src_joint = torch.from_numpy(mirror_source(self.joints, left, right)).to(device)

    def gather(self, lib, rows, row_lo, n_windows, frames, first_offset, stream):
        """rows: int32 (P, 4) device table; returns (n_windows, frames, joints, features) fp32."""
        out = torch.empty((n_windows, frames, self.joints, self.features), dtype=torch.float32,
                          device=self.data.device)
        d = _capi.GatherDesc()
        d.src = self.data.data_ptr()
        d.seq_first = self.seq_first.data_ptr()
        d.seq_len = self.seq_len.data_ptr()
        d.rows = rows.data_ptr() + 16 * row_lo
        d.src_joint = self.src_joint.data_ptr() if self.src_joint is not None else None
        d.out = out.data_ptr()
        d.n_windows, d.frames = n_windows, frames
        d.joints, d.features, d.first_offset = self.joints, self.features, first_offset
        _capi.check(lib.vp3d_gather_windows(ctypes.byref(d), stream), "vp3d_gather_windows")
        return out


class _DeviceGeneratorBase:
    def _setup(self, cameras, poses_3d, poses_2d, augment, kps_left, kps_right, joints_left,
               joints_right, device):
        assert poses_3d is None or len(poses_3d) == len(poses_2d), (len(poses_3d), len(poses_2d))
        assert cameras is None or len(cameras) == len(poses_2d)
        self.device = _require_cuda(device)
        self._lib = _capi.load()
        with torch.cuda.device(self.device):
            self._p2 = _PoseStore(poses_2d, self.device)
            self._p3 = _PoseStore(poses_3d, self.device) if poses_3d is not None else None
            if poses_3d is not None and not np.array_equal(self._p3.lengths, self._p2.lengths):
                raise ValueError("poses_3d and poses_2d disagree on sequence lengths")
            self._cams = None
            if cameras is not None:
                cams = np.stack([np.asarray(c, dtype=np.float32) for c in cameras])
                self._cams = torch.from_numpy(cams).to(self.device)
            if kps_left is not None and kps_right is not None:
                self._p2.set_mirror(kps_left, kps_right, self.device)
            if self._p3 is not None and joints_left is not None and joints_right is not None:
                self._p3.set_mirror(joints_left, joints_right, self.device)
        self.augment = augment
        self.kps_left, self.kps_right = kps_left, kps_right
        self.joints_left, self.joints_right = joints_left, joints_right

    def _check_mirror(self):
        if self._p2.src_joint is None or (self._p3 is not None and self._p3.src_joint is None):
            raise ValueError("augment=True needs kps_left/kps_right (and joints_left/joints_right "
                             "when 3-D poses are given)")

    def _emit(self, rows, lo, n, frames_2d, offset_2d, frames_3d):
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            b2 = self._p2.gather(self._lib, rows, lo, n, frames_2d, offset_2d, stream)
            b3 = cam = None
            if self._p3 is not None:
                b3 = self._p3.gather(self._lib, rows, lo, n, frames_3d, 0, stream)
            if self._cams is not None:
                cam = torch.empty((n, self._cams.shape[1]), dtype=torch.float32, device=self.device)
                _capi.check(self._lib.vp3d_gather_cameras(
                    self._cams.data_ptr(), self._cams.shape[1], rows.data_ptr() + 16 * lo, n,
                    cam.data_ptr(), stream), "vp3d_gather_cameras")
        return cam, b3, b2

    def augment_enabled(self):
        return self.augment


class ChunkedGenerator(_DeviceGeneratorBase):
    """Training stream (reference: common/generators.py:11-160; used at run.py:277-287, 323, 401).

    Arguments as in the reference: batch_size, cameras, poses_3d, poses_2d, chunk_length, pad,
    causal_shift, shuffle, random_seed, augment, kps_left/right, joints_left/right, endless; plus
    `device`, `rank`, `world_size`.  `next_epoch()` yields `(cam | None, batch_3d | None, batch_2d)`.
    """

    def __init__(self, batch_size, cameras, poses_3d, poses_2d, chunk_length, pad=0, causal_shift=0,
                 shuffle=True, random_seed=1234, augment=False, kps_left=None, kps_right=None,
                 joints_left=None, joints_right=None, endless=False, device=None, rank=0,
                 world_size=1):
        self._setup(cameras, poses_3d, poses_2d, augment, kps_left, kps_right, joints_left,
                    joints_right, device)
        if augment:
            self._check_mirror()
        self._plan = _EpochPlanner(self._p2.lengths, batch_size, chunk_length, shuffle, random_seed,
                                   augment, endless)
        self.batch_size = batch_size
        self.num_batches = self._plan.num_batches
        self.chunk_length = chunk_length
        self.pad, self.causal_shift = pad, causal_shift
        self.shuffle, self.endless = shuffle, endless
        self.rank, self.world_size = rank, world_size
        self.cameras, self.poses_3d, self.poses_2d = cameras, poses_3d, poses_2d
        self._rows_src = None
        self._rows_dev = None
        self.last_shard = (0, 0)

    # -- reference API -------------------------------------------------------------------------
    @property
    def pairs(self):
        return self._plan.pairs

    @property
    def random(self):
        return self._plan.random

    def num_frames(self):
        return self.num_batches * self.batch_size

    def random_state(self):
        return self._plan.random

    def set_random_state(self, random):
        self._plan.random = random

    def _device_rows(self, order):
        if self._rows_src is not order:
            with torch.cuda.device(self.device):
                host = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int32))
                self._rows_dev = host.to(self.device)
            self._rows_src = order
        return self._rows_dev

    def next_epoch(self):
        plan = self._plan
        while True:
            start, order = plan.begin()
            rows = self._device_rows(order)
            for b in range(start, plan.num_batches):
                g_lo, g_hi = plan.batch_bounds(b)
                if plan.endless:
                    plan.state = (b + 1, order)
                # A trailing batch with fewer rows than ranks would leave some ranks without work
                # while the others enter the gradient all-reduce: every rank drops it (the decision
                # depends only on (b, world_size), so it is the same everywhere).
                if self.world_size > 1 and g_hi - g_lo < self.world_size:
                    continue
                lo, hi = shard_rows(g_lo, g_hi, self.rank, self.world_size)
                # rows of this rank / of the whole batch: GradientReducer.set_step_rows() uses
                # them to weight unequal shards (sum_r n_r/n * g_r instead of a plain mean)
                self.last_shard = (hi - lo, g_hi - g_lo)
                yield self._emit(rows, lo, hi - lo, self.chunk_length + 2 * self.pad,
                                 -self.pad - self.causal_shift, self.chunk_length)
            if not plan.endless:
                return
            plan.state = None


class UnchunkedGenerator(_DeviceGeneratorBase):
    """Evaluation stream (reference: common/generators.py:163-240; run.py:224, 280, 289, 735, 842):
    one whole sequence per batch, the 2-D input edge-padded by (pad + causal_shift, pad -
    causal_shift) frames; with augmentation the mirrored copy is appended as batch row 1."""

    def __init__(self, cameras, poses_3d, poses_2d, pad=0, causal_shift=0, augment=False,
                 kps_left=None, kps_right=None, joints_left=None, joints_right=None, device=None):
        self._setup(cameras, poses_3d, poses_2d, augment, kps_left, kps_right, joints_left,
                    joints_right, device)
        self.pad, self.causal_shift = pad, causal_shift
        self.cameras = [] if cameras is None else cameras
        self.poses_3d = [] if poses_3d is None else poses_3d
        self.poses_2d = poses_2d
        lens = self._p2.lengths
        rows = np.zeros((2 * len(lens), 4), dtype=np.int32)  # (s, plain), (s, mirrored), ...
        rows[:, 0] = np.repeat(np.arange(len(lens)), 2)
        rows[:, 2] = np.repeat(lens, 2)
        rows[1::2, 3] = 1
        with torch.cuda.device(self.device):
            self._rows = torch.from_numpy(rows).to(self.device)

    def num_frames(self):
        return int(self._p2.lengths.sum())

    def set_augment(self, augment):
        self.augment = augment

    def next_epoch(self):
        for s, n in enumerate(self._p2.lengths.tolist()):
            if self.augment:
                self._check_mirror()
            yield self._emit(self._rows, 2 * s, 2 if self.augment else 1, n + 2 * self.pad,
                             -self.pad - self.causal_shift, n)
