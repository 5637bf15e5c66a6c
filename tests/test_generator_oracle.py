"""CPU: pin oracle/generator_oracle.py against the batches the real reference generators yielded
(tests/golden/make_generator_golden.py).  Index / sign / permutation work: comparisons are exact."""
import numpy as np
import pytest

from conftest import generator_golden_names, generator_kwargs, load_generator_golden
from oracle import generator_oracle as gorc


def _collect(cfg, gen):
    if cfg.get("endless"):
        it = gen.next_epoch()
        return [next(it) for _ in range(cfg["take_batches"])]
    out = []
    for _ in range(cfg.get("epochs", 1)):
        out.extend(gen.next_epoch())
    return out


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return a.shape == b.shape and np.array_equal(np.asarray(a, dtype=np.float64),
                                                 np.asarray(b, dtype=np.float64))


@pytest.mark.parametrize("name", generator_golden_names())
def test_oracle_matches_reference_batches(name):
    cfg, cams, p3, p2, ref = load_generator_golden(name)
    kw = generator_kwargs(cfg)
    if cfg["kind"] == "chunked":
        gen = gorc.ChunkedGeneratorOracle(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"], **kw)
    else:
        gen = gorc.UnchunkedGeneratorOracle(cams, p3, p2, **kw)
    assert gen.num_frames() == cfg["num_frames"]
    got = _collect(cfg, gen)
    assert len(got) == len(ref)
    for i, ((c, b3, b2), (rc, r3, r2)) in enumerate(zip(got, ref)):
        assert _same(b2, r2), (name, i, "2d")
        assert _same(b3, r3), (name, i, "3d")
        assert _same(c, rc), (name, i, "cam")


def test_chunk_grid_is_centred():
    # generators.py:41-44: 10 frames in chunks of 4 -> 3 chunks starting at -1, 3, 7
    pairs = gorc.chunk_pairs([10], 4, augment=True)
    assert pairs[:, 1].tolist() == [-1, 3, 7, -1, 3, 7]
    assert pairs[:, 3].tolist() == [0, 0, 0, 1, 1, 1]


def test_mirror_permutation_is_an_involution_for_disjoint_lists():
    src = gorc.mirror_permutation(17, [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16])
    assert sorted(src.tolist()) == list(range(17))
    assert np.array_equal(src[src], np.arange(17))
