"""GPU: the training-step companions against plain PyTorch fp32:
FusedAdam (vp3d_adam_step) vs torch.optim.Adam, fused mpjpe / weighted_mpjpe (vp3d_mpjpe_fwd_bwd)
vs the reference formulas of common/loss.py:11-25.  Floating point: tolerances stated per test."""
import numpy as np
import pytest
import torch

import videopose3d_b200 as vp
from videopose3d_b200 import loss as vloss
from videopose3d_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu


def _params(dev, seed, n_extra=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1024, 34, 3), (257,), (8193,), (3, 5, 7), (1,), (64, 64, 1), (16385,)] + [(33,)] * n_extra
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    # an unaligned (4-byte aligned only) parameter carved out of a larger buffer
    buf = torch.randn(1003, generator=g).to(dev)
    ps.append(torch.nn.Parameter(buf[1:1000]))
    return ps


def _close(a, b, rtol, atol):
    return torch.allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("amsgrad,weight_decay,n_extra", [(True, 0.0, 0), (False, 0.0, 0),
                                                          (True, 0.01, 70)])
def test_fused_adam_matches_torch(cuda_device, amsgrad, weight_decay, n_extra):
    """8 steps with fresh random gradients and an lr decay in between (run.py:583-586).
    Tolerances (fp32 round-off of a different operation order): parameters 2e-6 relative + 2e-7
    absolute (one Adam step moves a parameter by ~lr = 1e-3); first moment 2e-6 + 2e-6 absolute
    (it crosses zero, gradients are O(1..8)); second moments 5e-6 relative."""
    ours_p = _params(cuda_device, 1, n_extra)
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p in ours_p]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay, amsgrad=amsgrad)
    ours, ref = FusedAdam(ours_p, **kw), torch.optim.Adam(ref_p, **kw)
    g = torch.Generator().manual_seed(2)
    for step in range(8):
        for a, b in zip(ours_p, ref_p):
            grad = (torch.randn(a.shape, generator=g) * (0.1 + step)).to(cuda_device)
            a.grad, b.grad = grad.clone(), grad.clone()
        v0 = ours_p[0]._version
        ours.step()
        ref.step()
        assert ours_p[0]._version > v0  # the weight-pack cache keys on the version counter
        if step == 3:
            for opt in (ours, ref):
                for group in opt.param_groups:
                    group["lr"] *= 0.95
    for a, b in zip(ours_p, ref_p):
        assert _close(a, b, 2e-6, 2e-7)
        assert _close(ours.state[a]["exp_avg"], ref.state[b]["exp_avg"], 2e-6, 2e-6)
        for key in ("exp_avg_sq",) + (("max_exp_avg_sq",) if amsgrad else ()):
            assert _close(ours.state[a][key], ref.state[b][key], 5e-6, 1e-9), key
        assert float(ours.state[a]["step"]) == float(ref.state[b]["step"]) == 8.0


def test_fused_adam_checkpoint_round_trip(cuda_device):
    ours_p = _params(cuda_device, 3)
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p in ours_p]
    ours, ref = FusedAdam(ours_p, lr=1e-3, amsgrad=True), torch.optim.Adam(ref_p, lr=1e-3, amsgrad=True)
    g = torch.Generator().manual_seed(4)

    def grads():
        for a, b in zip(ours_p, ref_p):
            grad = torch.randn(a.shape, generator=g).to(cuda_device)
            a.grad, b.grad = grad.clone(), grad.clone()
    grads(); ours.step(); ref.step()
    # swap optimiser states through state_dict() and continue
    sd_ours, sd_ref = ours.state_dict(), ref.state_dict()
    ours.load_state_dict(sd_ref)
    ref.load_state_dict(sd_ours)
    grads(); ours.step(); ref.step()
    for a, b in zip(ours_p, ref_p):
        assert _close(a, b, 2e-6, 2e-7)


def test_fused_adam_drives_the_model(cuda_device):
    """FusedAdam on the real model (gradients are views of the model's flat gradient buffer; the
    weight packs are refreshed through the version counter): at every step a twin parameter set
    receives the SAME gradients and is stepped by torch.optim.Adam -> parameters agree to fp32
    round-off (2e-6 relative + 1e-6 absolute, lr = 1e-3); afterwards the trained model and a fresh
    model loaded with the twin's parameters produce the same training-mode output (<= 2e-5), which
    only holds if the kernels really run on the updated weights."""
    torch.manual_seed(0)
    kw = dict(filter_widths=[3, 3, 3], dropout=0.0, channels=128)
    a = vp.TemporalModelOptimized1f(17, 2, 17, **kw).to(cuda_device).train()
    a.set_train_precision("bf16x3")
    twin = [torch.nn.Parameter(p.detach().clone()) for p in a.parameters()]
    oa = FusedAdam(a.parameters(), lr=1e-3, amsgrad=True)
    ob = torch.optim.Adam(twin, lr=1e-3, amsgrad=True)
    g = torch.Generator().manual_seed(1)
    losses = []
    for _ in range(5):
        x = (torch.rand(64, 27, 17, 2, generator=g) * 2 - 1).to(cuda_device)
        y = (torch.randn(64, 1, 17, 3, generator=g) * 0.3).to(cuda_device)
        oa.zero_grad()
        loss = vloss.mpjpe(a(x), y)
        loss.backward()
        for pa, pb in zip(a.parameters(), twin):
            pb.grad = pa.grad.detach().clone()
        oa.step()
        ob.step()
        losses.append(loss.item())
        for (name, pa), pb in zip(a.named_parameters(), twin):
            assert _close(pa, pb, 2e-6, 1e-6), name
    assert losses[-1] < losses[0]  # it trains
    b = vp.TemporalModelOptimized1f(17, 2, 17, **kw)
    b.load_state_dict(a.state_dict())
    with torch.no_grad():
        for pb, pt in zip(b.parameters(), twin):
            pb.copy_(pt)
    b = b.to(cuda_device).train()
    b.set_train_precision("bf16x3")
    ya, yb = a(x).detach(), b(x).detach()
    assert float((ya - yb).abs().max()) <= 2e-5 * float(yb.abs().max())


def test_mpjpe_matches_reference_formula(cuda_device):
    g = torch.Generator().manual_seed(5)
    pred = (torch.randn(1024, 1, 17, 3, generator=g) * 0.5).to(cuda_device).requires_grad_(True)
    tgt = (torch.randn(1024, 1, 17, 3, generator=g) * 0.5).to(cuda_device)
    with torch.no_grad():
        tgt[3, 0, 5] = pred[3, 0, 5]  # a zero-length error vector: gradient must be 0, not NaN
    ref_in = pred.detach().clone().requires_grad_(True)
    ref = torch.mean(torch.norm(ref_in - tgt, dim=len(tgt.shape) - 1))   # loss.py:17
    ref.backward()
    ours = vloss.mpjpe(pred, tgt)
    (ours * 3.0).backward()
    assert ours.shape == () and abs(ours.item() - ref.item()) <= 2e-6 * ref.item()
    assert torch.isfinite(pred.grad).all() and float(pred.grad[3, 0, 5].abs().max()) == 0.0
    assert _close(pred.grad, 3.0 * ref_in.grad, 1e-5, 1e-10)
    with torch.no_grad():  # evaluation use (run.py:452): no gradient buffer, same value
        assert abs(vloss.mpjpe(pred, tgt).item() - ref.item()) <= 2e-6 * ref.item()


def test_weighted_mpjpe_matches_reference_formula(cuda_device):
    g = torch.Generator().manual_seed(6)
    pred = (torch.randn(512, 1, 1, 3, generator=g)).to(cuda_device).requires_grad_(True)
    tgt = (torch.randn(512, 1, 1, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])).to(cuda_device)
    w = 1 / tgt[:, :, :, 2]                                   # run.py:358
    ref_in = pred.detach().clone().requires_grad_(True)
    ref = torch.mean(w * torch.norm(ref_in - tgt, dim=len(tgt.shape) - 1))  # loss.py:25
    ref.backward()
    ours = vloss.weighted_mpjpe(pred, tgt, w)
    ours.backward()
    assert abs(ours.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert _close(pred.grad, ref_in.grad, 1e-5, 1e-10)


def _project_reference(X, cam, linear):
    """Plain torch restatement of the H3.6M projection (common/camera.py:37-88)."""
    cp = cam[:, None, None, :]
    f, c, k, p = cp[..., :2], cp[..., 2:4], cp[..., 4:7], cp[..., 7:]
    XX = torch.clamp(X[..., :2] / X[..., 2:], min=-1, max=1)
    if linear:
        return f * XX + c
    r2 = torch.sum(XX ** 2, dim=-1, keepdim=True)
    radial = 1 + torch.sum(k * torch.cat((r2, r2 ** 2, r2 ** 3), dim=-1), dim=-1, keepdim=True)
    tan = torch.sum(p * XX, dim=-1, keepdim=True)
    return f * (XX * (radial + tan) + p * r2) + c


@pytest.mark.parametrize("linear", [False, True])
def test_projected_mpjpe_matches_reference_formula(cuda_device, linear):
    """Semi-supervised reconstruction loss (run.py:374-379) and both gradients against autograd
    through the plain torch formula: loss 2e-6 relative, gradients 1e-5 relative + 1e-8 absolute (typical gradient 3e-5)."""
    g = torch.Generator().manual_seed(7)
    n, t, j = 512, 1, 17
    pos = (torch.randn(n, t, j, 3, generator=g) * 0.3).to(cuda_device)
    traj = (torch.randn(n, t, 1, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 4.5])).to(cuda_device)
    with torch.no_grad():
        pos[0, 0, 0, 0] = 9.0    # beyond the [-1, 1] clamp: no gradient through x
        pos[1, 0, 3, 1] = -9.0
    cam = torch.cat([torch.rand(n, 2, generator=g) + 1.0, torch.rand(n, 2, generator=g) * 0.1,
                     torch.randn(n, 3, generator=g) * 0.1, torch.randn(n, 2, generator=g) * 0.01],
                    dim=1).to(cuda_device)
    target = (torch.randn(n, t, j, 2, generator=g) * 0.3).to(cuda_device)
    pa, ta = pos.clone().requires_grad_(True), traj.clone().requires_grad_(True)
    pb, tb = pos.clone().requires_grad_(True), traj.clone().requires_grad_(True)
    ref = torch.mean(torch.norm(_project_reference(pb + tb, cam, linear) - target, dim=-1))
    ref.backward()
    ours = vloss.projected_mpjpe(pa, ta, cam, target, linear=linear)
    ours.backward()
    assert abs(ours.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert float(pa.grad[0, 0, 0, 0]) == 0.0 and float(pa.grad[1, 0, 3, 1]) == 0.0
    assert _close(pa.grad, pb.grad, 1e-5, 1e-8)
    assert _close(ta.grad, tb.grad, 1e-5, 1e-8)
    with torch.no_grad():
        assert abs(vloss.projected_mpjpe(pos, traj, cam, target, linear=linear).item() - ref.item()) \
            <= 2e-6 * abs(ref.item())


# ---------------------------------------------------------------------------------------------
# Semi-supervised loss head (BASELINE configs[4]) against the golden produced by the REAL
# reference: run.py:329-390 executed literally with common/loss.py, common/camera.py and the
# Human3.6M skeleton (tests/golden/make_semi_golden.py -> semi_333_c64.npz).
# ---------------------------------------------------------------------------------------------
def _load_semi():
    import json
    import os
    import numpy as np
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "semi_333_c64.npz"))
    return json.loads(str(z["meta"])), z


@pytest.mark.parametrize("tag,linear", [("full/", False), ("lin/", True)])
def test_semi_supervised_loss_head_matches_real_reference(cuda_device, tag, linear):
    """The fused kernel on the reference's own model outputs: every loss term and
    d loss_total / d (both model outputs) as autograd computed them through loss.py / camera.py."""
    meta, z = _load_semi()
    dev = cuda_device
    pad = meta["pad"]
    pos = torch.from_numpy(z[tag + "pred_pos"]).to(dev).requires_grad_(True)
    traj = torch.from_numpy(z[tag + "pred_traj"]).to(dev).requires_grad_(True)
    in3 = torch.from_numpy(z["inputs_3d"]).to(dev)
    cam = torch.from_numpy(z["cam_semi"]).to(dev)
    tgt2 = torch.from_numpy(z["inputs_2d_semi"]).to(dev)[:, pad:-pad, :, :2].contiguous()  # run.py:368
    total, terms = vloss.semi_supervised_loss(pos, traj, in3, cam, tgt2, meta["parents"],
                                              linear_projection=linear)
    total.backward()
    ref = z[tag + "losses"]
    got = terms.cpu().numpy()
    print("loss terms", got, "reference", ref)
    for i in range(5):
        assert abs(got[i] - ref[i]) <= 5e-6 * abs(ref[i]) + 1e-7, (i, got, ref)
    assert abs(total.item() - ref[4]) <= 5e-6 * ref[4]
    for g, name in ((pos.grad, "d_pred_pos"), (traj.grad, "d_pred_traj")):
        r = torch.from_numpy(z[tag + name]).to(dev)
        assert _close(g, r, 2e-5, 1e-8), (name, float((g - r).abs().max()), float(r.abs().max()))
    # unselected terms carry no gradient: --no-proj / --no-bone-length of run.py:380, 382
    p2, t2 = pos.detach().clone().requires_grad_(True), traj.detach().clone().requires_grad_(True)
    tot2, terms2 = vloss.semi_supervised_loss(p2, t2, in3, cam, tgt2, meta["parents"], linear_projection=linear,
                                              no_proj=True, bone_length_term=False)
    tot2.backward()
    assert abs(tot2.item() - (ref[0] + ref[1])) <= 5e-6 * (ref[0] + ref[1])
    assert abs(terms2[2].item() - ref[2]) <= 5e-6 * ref[2]          # still reported (run.py:377)
    n_lab = meta["n_labeled"]
    assert float(p2.grad[n_lab:].abs().max()) == 0.0 and float(t2.grad[n_lab:].abs().max()) == 0.0


def test_bone_length_penalty_matches_reference_formula(cuda_device):
    """run.py:383-387 restated with torch ops on the GPU vs the kernel (penalty term alone)."""
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(12, 3, 17, 3, generator=g).to(cuda_device)
    parents = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]   # h36m 17-joint skeleton
    split = 7
    a = pred.clone().requires_grad_(True)
    dists = a[:, :, 1:] - a[:, :, parents[1:]]
    lengths = torch.mean(torch.norm(dists, dim=3), dim=1)
    ref = torch.mean(torch.abs(torch.mean(lengths[:split], dim=0) - torch.mean(lengths[split:], dim=0)))
    ref.backward()
    b = pred.clone().requires_grad_(True)
    ours = vloss.bone_length_penalty(b, split, parents)
    (2.0 * ours).backward()
    assert abs(ours.item() - ref.item()) <= 5e-6 * ref.item()
    assert _close(b.grad, 2.0 * a.grad, 2e-5, 1e-9)


def test_semi_supervised_step_end_to_end_matches_real_reference(cuda_device):
    """Both models (position J_out = 17, trajectory J_out = 1; run.py:238-246) in train mode through
    the fp32-faithful kernels + the fused loss head: model outputs, loss terms and EVERY parameter
    gradient of both models against the reference's autograd (<= 1e-3 of each tensor's scale)."""
    import numpy as np
    import videopose3d_b200 as vp
    meta, z = _load_semi()
    dev = cuda_device
    arc, C, J, pad = meta["arc"], meta["C"], meta["J"], meta["pad"]
    models = {}
    for name, jout in (("pos", J), ("traj", 1)):
        sd = {k[len(f"sd_{name}/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"sd_{name}/")}
        m = vp.TemporalModelOptimized1f(J, 2, jout, filter_widths=arc, dropout=0.0, channels=C)
        m.load_state_dict(sd)
        models[name] = m.to(dev).train().set_train_precision("bf16x3")
    x_cat = torch.cat((torch.from_numpy(z["inputs_2d"]), torch.from_numpy(z["inputs_2d_semi"]))).to(dev)
    pos = models["pos"](x_cat)
    traj = models["traj"](x_cat)
    assert float((pos.detach().cpu() - torch.from_numpy(z["full/pred_pos"])).abs().max()) <= \
        1e-3 * float(np.abs(z["full/pred_pos"]).max())
    tgt2 = torch.from_numpy(z["inputs_2d_semi"]).to(dev)[:, pad:-pad, :, :2].contiguous()
    total, terms = vloss.semi_supervised_loss(pos, traj, torch.from_numpy(z["inputs_3d"]).to(dev),
                                              torch.from_numpy(z["cam_semi"]).to(dev), tgt2, meta["parents"])
    total.backward()
    ref = z["full/losses"]
    assert abs(total.item() - ref[4]) <= 1e-3 * ref[4]
    worst = {}
    for name, m in models.items():
        for k, prm in m.named_parameters():
            r = z[f"full/grad_{name}/{k}"]
            worst[f"{name}.{k}"] = float(np.abs(prm.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-30))
    print(f"semi-supervised step: worst parameter-gradient deviation {max(worst.values()):.2e}")
    bad = {k: v for k, v in worst.items() if not v <= 1e-3}
    assert not bad, bad


def test_fused_adam_repack_keeps_packed_weights_current(cuda_device):
    """the optimizer update that also writes the bf16 forward / transposed packs gives
    bit-identical parameters, optimizer state and next-step outputs as the plain update followed by
    the separate re-pack kernels, with fewer launches."""
    import videopose3d_b200 as vp
    from videopose3d_b200.optim import FusedAdam
    from oracle import temporal_model_oracle as orc
    arc, C, N = [3, 3, 3], 128, 48
    sd = orc.make_state_dict(17, 2, 17, arc, C, seed=9)
    xs = [orc.make_input(N, 27, 17, 2, seed=20 + i).to(cuda_device) for i in range(4)]
    tgt = torch.randn(N, 1, 17, 3, generator=torch.Generator().manual_seed(3)).to(cuda_device) * 0.3
    runs = {}
    for fused in (False, True):
        for prec in ("bf16", "bf16x3"):
            m = vp.TemporalModelOptimized1f(17, 2, 17, filter_widths=arc, dropout=0.0, channels=C)
            m.load_state_dict(sd)
            m = m.to(cuda_device).train().set_train_precision(prec)
            opt = FusedAdam(m.parameters(), lr=1e-3, amsgrad=True)
            opt.fuse_repack = fused
            outs, launches = [], []
            for x in xs:
                opt.zero_grad()
                y = m(x)
                vloss.mpjpe(y, tgt).backward()
                opt.step()
                outs.append(y.detach().clone())
                launches.append(opt.last_launches)
            runs[(fused, prec)] = (outs, {k: v.detach().clone() for k, v in m.state_dict().items()},
                                   opt.state_dict(), launches)
    for prec in ("bf16", "bf16x3"):
        a, b = runs[(False, prec)], runs[(True, prec)]
        for ya, yb in zip(a[0], b[0]):
            assert torch.equal(ya, yb)                     # same packs -> same forward, bit for bit
        for k in a[1]:
            assert torch.equal(a[1][k], b[1][k]), k
        for sa, sb in zip(a[2]["state"].values(), b[2]["state"].values()):
            for key in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                assert torch.equal(sa[key], sb[key])
        # step 1 runs before any training plan is packed? no: the forward packed it -> fused from
        # the first step on: plain kernel + fused conv kernel + 2 tiny expand packs
        assert a[3][-1] == 1 and b[3][-1] == 4, (a[3], b[3])
    # the eval plan still re-packs from the fp32 masters after training (separate cache entry)
    m.eval()
    with torch.no_grad():
        y_eval = m.set_precision("bf16x3")(xs[0])
    sd_now = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    y_o = orc.forward_numpy(sd_now, xs[0].cpu().numpy(), arc, strided=True)
    assert float((y_eval.cpu() - torch.from_numpy(y_o).float()).abs().max() / abs(y_o).max()) <= 1e-3
