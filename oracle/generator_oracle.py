"""CPU oracle for the batch generators -- TEST INFRASTRUCTURE ONLY.

Restates, with explicit index arithmetic, what the reference's `common/generators.py` produces:

  * `ChunkedGenerator`   (generators.py:11-160): the training stream of (camera, 3-D chunk, padded
    2-D window) batches, optionally shuffled and mirror-augmented;
  * `UnchunkedGenerator` (generators.py:163-240): whole padded sequences for evaluation, with the
    mirrored copy appended under test-time augmentation.

The restatement replaces the reference's per-chunk slice + `np.pad(..., 'edge')` by a clamped frame
index (identical whenever the slice is non-empty, which the reference requires anyway) and the
in-place left/right swap by a joint permutation.  It is pinned against fixtures generated from the
real reference classes (tests/golden/make_generator_golden.py -> tests/golden/gen_*.npz) in
tests/test_generator_oracle.py.  Only tests/, __graft_entry__.smoke() and the CPU legs of the bench
tools may import this module; the product (videopose3d_b200/generators.py) never does.
"""
import numpy as np


def chunk_pairs(lengths, chunk_length, augment):
    """(P, 4) int64 rows (sequence, first frame, end frame, flip) in the reference's order
    (generators.py:39-48): per sequence all plain chunks, then -- if `augment` -- the same chunks
    again flagged as mirrored.  The chunk grid is centred on the sequence (`offset`, :43)."""
    rows = []
    for s, n in enumerate(lengths):
        n_chunks = (n + chunk_length - 1) // chunk_length
        offset = (n_chunks * chunk_length - n) // 2
        first = np.arange(n_chunks, dtype=np.int64) * chunk_length - offset
        block = np.stack([np.full(n_chunks, s, np.int64), first, first + chunk_length,
                          np.zeros(n_chunks, np.int64)], axis=1)
        rows.append(block)
        if augment:
            mirrored = block.copy()
            mirrored[:, 3] = 1
            rows.append(mirrored)
    if not rows:
        return np.zeros((0, 4), np.int64)
    return np.concatenate(rows, axis=0)


def mirror_permutation(n, left, right):
    """src[j] such that mirrored[:, j] = plain[:, src[j]]: the reference assigns
    `x[:, left + right] = x[:, right + left]` (generators.py:123, 142-143); later entries win."""
    src = np.arange(n, dtype=np.int64)
    dst_list = list(left) + list(right)
    src_list = list(right) + list(left)
    for d, s in zip(dst_list, src_list):
        src[d] = s
    return src


def gather_window(seq, first, frames, flip, src_joint):
    """`frames` rows of `seq` starting at frame `first`, out-of-range frames replicated from the
    nearest edge (generators.py:108-118); `flip` negates feature 0 and swaps left/right joints."""
    idx = np.clip(np.arange(first, first + frames), 0, seq.shape[0] - 1)
    win = np.array(seq[idx], dtype=np.float64)
    if flip:
        win[..., 0] *= -1
        win = win[:, src_joint]
    return win


def mirror_camera(cam):
    """Horizontal flip of the intrinsics vector: entries 2 (c_x) and 7 (tangential p) change sign
    (generators.py:150-152)."""
    cam = np.array(cam, dtype=np.float64)
    cam[2] *= -1
    cam[7] *= -1
    return cam


class ChunkedGeneratorOracle:
    """Same constructor and methods as the reference class; yields fresh float64 arrays."""

    def __init__(self, batch_size, cameras, poses_3d, poses_2d, chunk_length, pad=0, causal_shift=0,
                 shuffle=True, random_seed=1234, augment=False, kps_left=None, kps_right=None,
                 joints_left=None, joints_right=None, endless=False):
        assert poses_3d is None or len(poses_3d) == len(poses_2d)
        assert cameras is None or len(cameras) == len(poses_2d)
        self.pairs = chunk_pairs([p.shape[0] for p in poses_2d], chunk_length, augment)
        self.batch_size = batch_size
        self.num_batches = (len(self.pairs) + batch_size - 1) // batch_size
        self.random = np.random.RandomState(random_seed)
        self.shuffle, self.pad, self.causal_shift, self.endless = shuffle, pad, causal_shift, endless
        self.chunk_length = chunk_length
        self.cameras, self.poses_3d, self.poses_2d = cameras, poses_3d, poses_2d
        self.augment = augment
        self.src_kps = self.src_joints = None
        if augment:
            self.src_kps = mirror_permutation(poses_2d[0].shape[-2], kps_left, kps_right)
            if poses_3d is not None:
                self.src_joints = mirror_permutation(poses_3d[0].shape[-2], joints_left, joints_right)
        self.state = None

    def num_frames(self):
        return self.num_batches * self.batch_size

    def random_state(self):
        return self.random

    def set_random_state(self, random):
        self.random = random

    def augment_enabled(self):
        return self.augment

    def epoch_order(self):
        """Row order of one epoch (generators.py:89-97): a fresh permutation per epoch when
        shuffling (one `RandomState.permutation` draw over the (P, 4) table), resumable state for
        `endless` streams."""
        if self.state is not None:
            return self.state
        return 0, (self.random.permutation(self.pairs) if self.shuffle else self.pairs)

    def batch(self, rows):
        """One batch for (n, 4) table rows, all rows at once: clamped frame indices into the
        concatenated sequences (= per-chunk slice + 'edge' padding, generators.py:103-118), then
        sign flip and joint swap of the mirrored rows (:120-123, :137-143, :150-152)."""
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 4)
        seq, first, flip = rows[:, 0], rows[:, 1], rows[:, 3] != 0
        frames = int(rows[0, 2] - rows[0, 1]) if len(rows) else self.chunk_length

        def windows(store, lead, count, src_joint):
            t = np.arange(count, dtype=np.int64)[None, :] + (first - lead)[:, None]
            idx = np.clip(t, 0, (store["len"][seq] - 1)[:, None]) + store["first"][seq][:, None]
            flat = store["flat"]
            out = np.empty(idx.shape + flat.shape[1:], dtype=np.float64)  # float64 as in :52-54
            for lo in range(0, len(idx), 16):  # cache-sized blocks of rows
                blk = idx[lo:lo + 16]
                win = flat.take(blk.ravel(), axis=0).reshape(blk.shape + flat.shape[1:])
                f = flip[lo:lo + 16]
                if f.any():
                    mirrored = win[f].take(src_joint, axis=2)
                    mirrored[..., 0] *= -1
                    win[f] = mirrored
                out[lo:lo + 16] = win
            return out

        p2 = windows(self._flat(self.poses_2d, "_s2"), self.pad + self.causal_shift,
                     frames + 2 * self.pad, self.src_kps)
        p3 = cams = None
        if self.poses_3d is not None:
            p3 = windows(self._flat(self.poses_3d, "_s3"), 0, frames, self.src_joints)
        if self.cameras is not None:
            cams = np.stack([np.asarray(c, dtype=np.float64) for c in self.cameras])[seq]
            cams[flip, 2] *= -1
            cams[flip, 7] *= -1
        return cams, p3, p2

    def _flat(self, sequences, key):
        if not hasattr(self, key):
            lens = np.array([s.shape[0] for s in sequences], dtype=np.int64)
            setattr(self, key, dict(len=lens, first=np.concatenate([[0], np.cumsum(lens)[:-1]]),
                                    flat=np.concatenate(sequences, axis=0)))
        return getattr(self, key)

    def next_epoch(self):
        while True:
            start, order = self.epoch_order()
            for b in range(start, self.num_batches):
                rows = order[b * self.batch_size:(b + 1) * self.batch_size]
                if self.endless:
                    self.state = (b + 1, order)
                yield self.batch(rows)
            if not self.endless:
                return
            self.state = None


class UnchunkedGeneratorOracle:
    """Whole sequences, one per batch; 2-D input padded by (pad + shift, pad - shift) edge frames;
    with augmentation the mirrored copy is row 1 of the batch (generators.py:213-240)."""

    def __init__(self, cameras, poses_3d, poses_2d, pad=0, causal_shift=0, augment=False,
                 kps_left=None, kps_right=None, joints_left=None, joints_right=None):
        assert poses_3d is None or len(poses_3d) == len(poses_2d)
        assert cameras is None or len(cameras) == len(poses_2d)
        self.cameras, self.poses_3d, self.poses_2d = cameras, poses_3d, poses_2d
        self.pad, self.causal_shift, self.augment = pad, causal_shift, augment
        self.kps = (kps_left, kps_right)
        self.joints = (joints_left, joints_right)

    def num_frames(self):
        return sum(p.shape[0] for p in self.poses_2d)

    def augment_enabled(self):
        return self.augment

    def set_augment(self, augment):
        self.augment = augment

    def next_epoch(self):
        for s, seq_2d in enumerate(self.poses_2d):
            n = seq_2d.shape[0]
            flips = (False, True) if self.augment else (False,)
            src_kps = mirror_permutation(seq_2d.shape[-2], *self.kps) if self.augment else None
            p2 = np.stack([gather_window(seq_2d, -self.pad - self.causal_shift, n + 2 * self.pad, f,
                                         src_kps) for f in flips])
            p3 = cam = None
            if self.poses_3d is not None:
                src_j = (mirror_permutation(self.poses_3d[s].shape[-2], *self.joints)
                         if self.augment else None)
                p3 = np.stack([gather_window(self.poses_3d[s], 0, n, f, src_j) for f in flips])
            if self.cameras is not None:
                cam = np.stack([mirror_camera(self.cameras[s]) if f
                                else np.array(self.cameras[s], dtype=np.float64) for f in flips])
            yield cam, p3, p2
