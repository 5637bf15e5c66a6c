"""CPU: host-side logic of videopose3d_b200/generators.py.

  * index tables (`chunk_table`, `mirror_source`, `_EpochPlanner`, `shard_rows`) against the oracle;
  * the classes end to end with the two C-ABI gather entry points replaced by a NumPy emulation
    that dereferences the same descriptor fields (pointer offsets, row stride, first_offset, ...),
    so that everything except the CUDA kernels themselves is checked here against the reference
    fixtures.  The kernels are checked on the GPU in tests/test_gpu_generators.py.
  * no silent fallback: without a CUDA device the real classes refuse to construct.
"""
import contextlib
import ctypes

import numpy as np
import pytest
import torch

from conftest import generator_golden_names, generator_kwargs, load_generator_golden
from oracle import generator_oracle as gorc
from videopose3d_b200 import _capi, generators as G


def test_chunk_table_matches_oracle():
    rng = np.random.RandomState(0)
    for _ in range(50):
        lengths = rng.randint(1, 60, size=rng.randint(1, 6)).tolist()
        chunk = int(rng.randint(1, 9))
        for aug in (False, True):
            assert np.array_equal(G.chunk_table(lengths, chunk, aug), gorc.chunk_pairs(lengths, chunk, aug))


def test_mirror_source_matches_oracle():
    left, right = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]
    assert np.array_equal(G.mirror_source(17, left, right), gorc.mirror_permutation(17, left, right))
    # overlapping lists: later assignments win, exactly like NumPy fancy assignment
    left, right = [0, 1], [1, 2]
    x = np.arange(6.0)
    y = x.copy()
    y[left + right] = x[right + left]
    assert np.array_equal(x[G.mirror_source(6, left, right)], y)
    with pytest.raises(ValueError):
        G.mirror_source(17, None, None)


def test_epoch_planner_reproduces_reference_permutation():
    lengths = [17, 40, 9]
    plan = G._EpochPlanner(lengths, 16, 1, True, 1234, True, False)
    orc = gorc.ChunkedGeneratorOracle(16, None, None, [np.zeros((n, 17, 2)) for n in lengths], 1,
                                      shuffle=True, random_seed=1234, augment=True,
                                      kps_left=[1], kps_right=[2])
    for _ in range(3):  # successive epochs keep drawing from the same stream
        assert np.array_equal(plan.begin()[1], orc.epoch_order()[1])
    assert plan.num_batches == orc.num_batches


def test_shard_rows_partitions_every_batch():
    for lo, hi in ((0, 1024), (1024, 1500), (7, 8), (5, 5)):
        for world in (1, 2, 3, 8):
            parts = [G.shard_rows(lo, hi, r, world) for r in range(world)]
            assert parts[0][0] == lo and parts[-1][1] == hi
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA"):
        G.ChunkedGenerator(4, None, None, [np.zeros((5, 17, 2), np.float32)], 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        G.UnchunkedGenerator(None, None, [np.zeros((5, 17, 2), np.float32)])


# ---- NumPy emulation of the two gather entry points (host-logic testing only) -------------------

def _view(ptr, n, ctype, dtype):
    return np.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


class _FakeLib:
    """Same argument contract as csrc/gather.cu, evaluated with NumPy on host memory."""

    def vp3d_gather_windows(self, dref, stream):
        d = dref._obj
        je = d.joints * d.features
        rows = _view(d.rows, 4 * d.n_windows, ctypes.c_int32, np.int32).reshape(-1, 4)
        n_seq = int(rows[:, 0].max()) + 1
        first = _view(d.seq_first, n_seq, ctypes.c_int64, np.int64)
        lens = _view(d.seq_len, n_seq, ctypes.c_int32, np.int32)
        total = int(first[-1] + lens[-1]) if n_seq else 0
        src = _view(d.src, max(total, int((first + lens).max())) * je, ctypes.c_float, np.float32)
        src = src.reshape(-1, d.joints, d.features)
        out = _view(d.out, d.n_windows * d.frames * je, ctypes.c_float, np.float32)
        out = out.reshape(d.n_windows, d.frames, d.joints, d.features)
        sj = (_view(d.src_joint, d.joints, ctypes.c_int32, np.int32) if d.src_joint
              else np.arange(d.joints))
        for w, (s, f0, _, flip) in enumerate(rows):
            fr = np.clip(np.arange(d.frames) + f0 + d.first_offset, 0, lens[s] - 1) + first[s]
            win = src[fr]
            if flip:
                win = win[:, sj].copy()
                win[..., 0] *= -1
            out[w] = win
        return 0

    def vp3d_gather_cameras(self, cams, cam_dim, rows, n, out, stream):
        rows = _view(rows, 4 * n, ctypes.c_int32, np.int32).reshape(-1, 4)
        n_seq = int(rows[:, 0].max()) + 1
        cams = _view(cams, n_seq * cam_dim, ctypes.c_float, np.float32).reshape(n_seq, cam_dim)
        out = _view(out, n * cam_dim, ctypes.c_float, np.float32).reshape(n, cam_dim)
        out[:] = cams[rows[:, 0]]
        flipped = rows[:, 3] != 0
        out[flipped, 2] *= -1
        out[flipped, 7] *= -1
        return 0


@pytest.fixture
def host_emulation(monkeypatch):
    monkeypatch.setattr(G, "_require_cuda", lambda device: torch.device("cpu"))
    monkeypatch.setattr(_capi, "load", lambda: _FakeLib())
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    class _Stream:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _Stream())


def _collect(cfg, gen):
    if cfg.get("endless"):
        it = gen.next_epoch()
        return [next(it) for _ in range(cfg["take_batches"])]
    out = []
    for _ in range(cfg.get("epochs", 1)):
        out.extend(gen.next_epoch())
    return out


def _np(t):
    return None if t is None else t.numpy()


@pytest.mark.parametrize("name", generator_golden_names())
def test_classes_reproduce_reference_batches_with_emulated_kernels(name, host_emulation):
    cfg, cams, p3, p2, ref = load_generator_golden(name)
    kw = generator_kwargs(cfg)
    if cfg["kind"] == "chunked":
        gen = G.ChunkedGenerator(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"], **kw)
    else:
        gen = G.UnchunkedGenerator(cams, p3, p2, **kw)
    assert gen.num_frames() == cfg["num_frames"]
    assert gen.augment_enabled() == cfg["augment"]
    got = _collect(cfg, gen)
    assert len(got) == len(ref)
    for i, ((c, b3, b2), (rc, r3, r2)) in enumerate(zip(got, ref)):
        assert b2.dtype == torch.float32 and np.array_equal(_np(b2), r2), (name, i)
        assert (b3 is None) == (r3 is None) and (b3 is None or np.array_equal(_np(b3), r3)), (name, i)
        assert (c is None) == (rc is None) and (c is None or np.array_equal(_np(c), rc)), (name, i)


def test_ranks_partition_each_batch(host_emulation):
    cfg, cams, p3, p2, ref = load_generator_golden("gen_sup_shuffle_aug")
    kw = generator_kwargs(cfg)
    gens = [G.ChunkedGenerator(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"], rank=r,
                               world_size=3, **kw) for r in range(3)]
    streams = [list(g.next_epoch()) for g in gens]
    for b, (_, r3, r2) in enumerate(ref[:len(streams[0])]):
        assert np.array_equal(np.concatenate([_np(s[b][2]) for s in streams]), r2)
        assert np.array_equal(np.concatenate([_np(s[b][1]) for s in streams]), r3)


def test_tail_batch_smaller_than_world_is_dropped_on_every_rank(host_emulation):
    """len(pairs) % batch_size < world_size: the trailing batch cannot give every rank a row, so
    all ranks skip it together (a rank entering the gradient all-reduce alone would hang)."""
    rng = np.random.RandomState(3)
    lens = [10, 9]                      # 19 windows, batch 8 -> batches of 8, 8, 3
    p2 = [rng.uniform(-1, 1, (n, 17, 2)).astype(np.float32) for n in lens]
    p3 = [rng.normal(0, 1, (n, 17, 3)).astype(np.float32) for n in lens]
    world = 4
    gens = [G.ChunkedGenerator(8, None, p3, p2, 1, pad=2, shuffle=True, random_seed=5, rank=r,
                               world_size=world) for r in range(world)]
    streams = [list(g.next_epoch()) for g in gens]
    assert [len(s) for s in streams] == [2] * world            # the 3-row tail is gone everywhere
    assert all(s[b][2].shape[0] == 2 for s in streams for b in range(2))
    assert gens[0].last_shard == (2, 8)
    # a tail that still feeds every rank is kept, with unequal shards reported for weighting
    lens = [10, 11]                     # 21 windows -> 8, 8, 5
    p2 = [rng.uniform(-1, 1, (n, 17, 2)).astype(np.float32) for n in lens]
    p3 = [rng.normal(0, 1, (n, 17, 3)).astype(np.float32) for n in lens]
    gens = [G.ChunkedGenerator(8, None, p3, p2, 1, pad=2, shuffle=False, rank=r, world_size=world)
            for r in range(world)]
    streams = [list(g.next_epoch()) for g in gens]
    assert [len(s) for s in streams] == [3] * world
    assert sorted(s[2][2].shape[0] for s in streams) == [1, 1, 1, 2]
    assert sum(g.last_shard[0] for g in gens) == 5 and all(g.last_shard[1] == 5 for g in gens)


def test_set_random_state_and_unchunked_toggle(host_emulation):
    cfg, cams, p3, p2, ref = load_generator_golden("gen_unchunked_aug")
    gen = G.UnchunkedGenerator(cams, p3, p2, **generator_kwargs(cfg))
    gen.set_augment(False)
    for (c, b3, b2), (rc, r3, r2) in zip(gen.next_epoch(), ref):
        assert b2.shape[0] == 1 and np.array_equal(_np(b2)[0], r2[0])
        assert np.array_equal(_np(b3)[0], r3[0]) and np.array_equal(_np(c)[0], rc[0])
    cfg, cams, p3, p2, ref = load_generator_golden("gen_2d_only")
    gen = G.ChunkedGenerator(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"], **generator_kwargs(cfg))
    gen.set_random_state(np.random.RandomState(cfg["random_seed"]))
    assert gen.random_state() is gen.random
    first = next(gen.next_epoch())
    assert np.array_equal(_np(first[2]), ref[0][2])
