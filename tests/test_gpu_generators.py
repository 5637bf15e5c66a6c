"""GPU: the device-resident generators (videopose3d_b200/generators.py -> vp3d_gather_windows /
vp3d_gather_cameras) against the reference's batches (tests/golden/gen_*.npz) and the oracle.
Copy / sign / index work: every comparison is bit-exact."""
import numpy as np
import pytest
import torch

from conftest import generator_golden_names, generator_kwargs, load_generator_golden
from oracle import generator_oracle as gorc
from videopose3d_b200 import generators as G

pytestmark = pytest.mark.gpu

KPS_L, KPS_R = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]


def _collect(cfg, gen):
    if cfg.get("endless"):
        it = gen.next_epoch()
        return [next(it) for _ in range(cfg["take_batches"])]
    out = []
    for _ in range(cfg.get("epochs", 1)):
        out.extend(gen.next_epoch())
    return out


def _np(t):
    return None if t is None else t.cpu().numpy()


def _same(t, ref):
    if t is None or ref is None:
        return t is None and ref is None
    a = _np(t)
    return t.is_cuda and t.dtype == torch.float32 and a.shape == ref.shape and \
        np.array_equal(a, np.asarray(ref, dtype=np.float32))


@pytest.mark.parametrize("name", generator_golden_names())
def test_batches_equal_reference(name, cuda_device):
    cfg, cams, p3, p2, ref = load_generator_golden(name)
    kw = generator_kwargs(cfg)
    if cfg["kind"] == "chunked":
        gen = G.ChunkedGenerator(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"],
                                 device=cuda_device, **kw)
    else:
        gen = G.UnchunkedGenerator(cams, p3, p2, device=cuda_device, **kw)
    got = _collect(cfg, gen)
    assert len(got) == len(ref)
    for i, ((c, b3, b2), (rc, r3, r2)) in enumerate(zip(got, ref)):
        assert _same(b2, r2), (name, i, "2d")
        assert _same(b3, r3), (name, i, "3d")
        assert _same(c, rc), (name, i, "cam")


def _h36m_like(rng, n_seq, lo, hi):
    lens = rng.randint(lo, hi, size=n_seq)
    p2 = [rng.uniform(-1, 1, (n, 17, 2)).astype(np.float32) for n in lens]
    p3 = [rng.normal(0, 0.5, (n, 17, 3)).astype(np.float32) for n in lens]
    cams = [rng.uniform(-1, 1, 9).astype(np.float32) for _ in lens]
    return cams, p3, p2


def test_training_shape_stream_equals_oracle(cuda_device):
    """BASELINE cfg3 batch shape: 1024 windows of 243 frames (pad 121), shuffled + mirrored, from
    40 sequences of 300-900 frames; two full batches against the oracle, bit for bit."""
    cams, p3, p2 = _h36m_like(np.random.RandomState(3), 40, 300, 900)
    kw = dict(pad=121, causal_shift=0, shuffle=True, random_seed=1234, augment=True, kps_left=KPS_L,
              kps_right=KPS_R, joints_left=KPS_L, joints_right=KPS_R)
    dev = G.ChunkedGenerator(1024, cams, p3, p2, 1, device=cuda_device, **kw)
    orc = gorc.ChunkedGeneratorOracle(1024, cams, p3, p2, 1, **kw)
    assert dev.num_batches == orc.num_batches and dev.num_frames() == orc.num_frames()
    for b, ((c, b3, b2), (oc, o3, o2)) in enumerate(zip(dev.next_epoch(), orc.next_epoch())):
        assert b2.shape == (1024, 243, 17, 2)
        assert _same(b2, o2) and _same(b3, o3) and _same(c, oc), b
        if b == 1:
            break


def test_mirroring_twice_is_identity_and_edges_replicate(cuda_device):
    """Size-independent properties on a long sequence: (1) mirrored(mirrored(x)) == x through the
    unchunked generator fed with its own mirrored output; (2) every padded frame equals the edge
    frame; (3) the interior equals the source."""
    rng = np.random.RandomState(5)
    seq = rng.uniform(-1, 1, (5000, 17, 2)).astype(np.float32)
    kw = dict(pad=121, causal_shift=0, augment=True, kps_left=KPS_L, kps_right=KPS_R,
              joints_left=KPS_L, joints_right=KPS_R)
    g1 = G.UnchunkedGenerator(None, None, [seq], device=cuda_device, **kw)
    _, _, b2 = next(g1.next_epoch())
    assert b2.shape == (2, 5000 + 242, 17, 2)
    plain, mirrored = _np(b2[0]), _np(b2[1])
    assert np.array_equal(plain[121:-121], seq)
    assert np.array_equal(plain[:121], np.broadcast_to(seq[0], (121, 17, 2)))
    assert np.array_equal(plain[-121:], np.broadcast_to(seq[-1], (121, 17, 2)))
    g2 = G.UnchunkedGenerator(None, None, [mirrored[121:-121]], device=cuda_device, **kw)
    _, _, c2 = next(g2.next_epoch())
    assert np.array_equal(_np(c2[1]), plain)


def test_causal_padding_and_ranks(cuda_device):
    cams, p3, p2 = _h36m_like(np.random.RandomState(9), 6, 20, 90)
    kw = dict(pad=13, causal_shift=13, shuffle=True, random_seed=7, augment=True, kps_left=KPS_L,
              kps_right=KPS_R, joints_left=KPS_L, joints_right=KPS_R)
    full = list(gorc.ChunkedGeneratorOracle(64, cams, p3, p2, 1, **kw).next_epoch())
    parts = [list(G.ChunkedGenerator(64, cams, p3, p2, 1, device=cuda_device, rank=r, world_size=2,
                                     **kw).next_epoch()) for r in range(2)]
    assert len(parts[0]) == len(full)
    for b, (oc, o3, o2) in enumerate(full):
        for k, ref in ((0, oc), (1, o3), (2, o2)):
            got = np.concatenate([_np(p[b][k]) for p in parts])
            assert np.array_equal(got, ref.astype(np.float32)), (b, k)


def test_feeds_the_model(cuda_device):
    """A gathered batch goes straight into the model (no host round trip) and matches feeding the
    same windows from the host."""
    import videopose3d_b200 as vp
    cams, p3, p2 = _h36m_like(np.random.RandomState(11), 4, 60, 120)
    gen = G.ChunkedGenerator(32, None, p3, p2, 1, pad=13, shuffle=True, random_seed=1, augment=False,
                             device=cuda_device)
    orc = gorc.ChunkedGeneratorOracle(32, None, p3, p2, 1, pad=13, shuffle=True, random_seed=1)
    m = vp.TemporalModel(17, 2, 17, filter_widths=[3, 3, 3], channels=64).to(cuda_device).eval()
    (_, b3, b2), (_, o3, o2) = next(gen.next_epoch()), next(orc.next_epoch())
    with torch.no_grad():
        y_dev = m(b2)
        y_host = m(torch.from_numpy(o2.astype(np.float32)).to(cuda_device))
    assert np.array_equal(_np(b2), o2.astype(np.float32))
    assert y_dev.shape == b3.shape and torch.allclose(y_dev, y_host, rtol=0, atol=1e-6)
