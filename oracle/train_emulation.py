"""Quantisation-aware CPU restatement of one TemporalModelOptimized1f training step.

TEST INFRASTRUCTURE ONLY (see oracle/temporal_model_oracle.py for the rules: only tests/, smoke()
and bench.py's CPU legs may import oracle/).

`train_step(..., planes=0)` is the reference algorithm in float64: forward of common/model.py:187-197
in train() mode (BatchNorm batch statistics, dropout = 0) followed by the analytic backward that
autograd performs (conv weight / data gradients, BatchNorm backward with the sum(dY) and
sum(dY * xhat) reductions, ReLU mask, residual fan-in).  It is pinned against gradients produced by
the real reference (tests/test_oracle_golden.py::test_train_emulation_matches_reference).

`planes = 1 / 2` additionally rounds every tensor the CUDA path stores in bf16 (1 plane) or split
bf16 (hi + lo, 2 planes) at exactly the points where the kernels round: packed input and weights,
pre-BN conv outputs Z, activations X / H, incoming gradients G and dZ.  ReLU networks are not
smooth: a pre-activation that sits within the rounding error of zero flips its mask and moves a
gradient by O(1/rows), so a bf16 path cannot be compared with an fp32 reference at tight tolerance
on small batches — but it can be compared tightly with this emulation, which shares its rounding
points.  The fp32-faithful mode (planes = 2) is compared with the reference directly, on fixtures
chosen away from ReLU kinks (tests/golden/make_golden.py).
"""
import numpy as np
import torch

EPS = 1e-5


def _q(t, planes):
    if planes == 0:
        return t
    hi = t.float().to(torch.bfloat16).double()
    if planes == 1:
        return hi
    return hi + (t - hi).float().to(torch.bfloat16).double()


def train_step(sd, x, gy, filter_widths, causal=False, planes=0, momentum=0.1):
    """sd: state_dict (torch tensors), x: (N, T, J, F), gy: upstream gradient of the output.
    Returns dict(y=..., grads={name: tensor}, new_stats={name: tensor}, min_abs_preact=float)."""
    q = lambda t: _q(t, planes)
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    fw = list(filter_widths)
    C = sd["expand_conv.weight"].shape[0]
    x = x.double()
    N, T = x.shape[0], x.shape[1]
    c_in = x.shape[2] * x.shape[3]
    saved, new_stats = {}, {}
    min_pre = float("inf")

    def bn(z_exact, prefix, layer):
        # the kernels take the batch statistics from the fp32 accumulators (before rounding) and
        # normalise the stored (rounded) Z with them
        nonlocal min_pre
        n = z_exact.shape[0]
        mu = z_exact.mean(0)
        var = z_exact.var(0, unbiased=False)
        z = q(z_exact)
        inv = 1.0 / torch.sqrt(var + EPS)
        sc = sd[prefix + ".weight"] * inv
        sh = sd[prefix + ".bias"] - mu * sc
        new_stats[prefix + ".running_mean"] = (1 - momentum) * sd[prefix + ".running_mean"] + momentum * mu
        new_stats[prefix + ".running_var"] = ((1 - momentum) * sd[prefix + ".running_var"] +
                                              momentum * var * n / max(n - 1, 1))
        saved[layer] = (z, mu, inv, sc, sh)
        y = z * sc + sh
        min_pre = min(min_pre, float(y.abs().min()))
        return torch.relu(y)

    # ---- forward (strided layout: every conv is a GEMM on [rows, w*C] views)
    L0 = T // fw[0]
    a0 = q(x.reshape(N, T, c_in)[:, :L0 * fw[0]].reshape(N * L0, fw[0] * c_in))
    w0 = sd["expand_conv.weight"].permute(0, 2, 1).reshape(C, -1)       # [co][tap*c_in + ci]
    X = q(bn(a0 @ q(w0).T, "expand_bn", 0))
    Xs, Hs = [X], [None]
    nb = len(fw) - 1
    offs = [None]
    for i in range(1, nb + 1):
        w = fw[i]
        rows = X.shape[0] // w
        A = X.reshape(rows, w * C)
        w1 = sd[f"layers_conv.{2 * (i - 1)}.weight"].permute(0, 2, 1).reshape(C, w * C)
        H = q(bn(A @ q(w1).T, f"layers_bn.{2 * (i - 1)}", 2 * i - 1))
        w2 = sd[f"layers_conv.{2 * (i - 1) + 1}.weight"][:, :, 0]
        Y2 = bn(H @ q(w2).T, f"layers_bn.{2 * (i - 1) + 1}", 2 * i)
        off = w // 2 + (w // 2 if causal else 0)                           # model.py:191
        X = q(X.reshape(rows, w, C)[:, off] + Y2)
        Xs.append(X)
        Hs.append(H)
        offs.append(off)
    wsh = sd["shrink.weight"][:, :, 0]
    y = X @ q(wsh).T + sd["shrink.bias"]

    # ---- backward
    gy = gy.double().reshape(-1, y.shape[1])
    grads = {"shrink.bias": gy.sum(0)}
    gyq = q(gy)
    grads["shrink.weight"] = (gyq.T @ Xs[nb])[:, :, None]
    G = q(gyq @ q(wsh))

    def bn_bwd(G, layer):
        z, mu, inv, sc, sh = saved[layer]
        dy = G * ((z * sc + sh) > 0)
        xh = (z - mu) * inv
        s1, s2, n = dy.sum(0), (dy * xh).sum(0), z.shape[0]
        return q(sc * (dy - s1 / n - xh * s2 / n)), s2, s1

    for i in range(nb, 0, -1):
        w = fw[i]
        c1, c2 = 2 * (i - 1), 2 * (i - 1) + 1
        dz2, dg, db = bn_bwd(G, 2 * i)
        grads[f"layers_bn.{c2}.weight"], grads[f"layers_bn.{c2}.bias"] = dg, db
        grads[f"layers_conv.{c2}.weight"] = (dz2.T @ Hs[i])[:, :, None]
        GH = q(dz2 @ q(sd[f"layers_conv.{c2}.weight"][:, :, 0]))
        dz1, dg, db = bn_bwd(GH, 2 * i - 1)
        grads[f"layers_bn.{c1}.weight"], grads[f"layers_bn.{c1}.bias"] = dg, db
        rows = dz1.shape[0]
        A = Xs[i - 1].reshape(rows, w, C)
        grads[f"layers_conv.{c1}.weight"] = torch.einsum("ro,rkc->ock", dz1, A)
        Gn = torch.einsum("ro,ock->rkc", dz1, q(sd[f"layers_conv.{c1}.weight"]))
        Gn[:, offs[i]] += G                                                # skip-connection gradient
        G = q(Gn.reshape(rows * w, C))
    dz0, dg, db = bn_bwd(G, 0)
    grads["expand_bn.weight"], grads["expand_bn.bias"] = dg, db
    grads["expand_conv.weight"] = (dz0.T @ a0).reshape(C, fw[0], c_in).permute(0, 2, 1)
    return dict(y=y.reshape(N, -1, y.shape[1] // 3, 3), grads=grads, new_stats=new_stats,
                min_abs_preact=min_pre)


def rel_l2(a, b):
    a = a.detach().cpu().double().numpy() if hasattr(a, "detach") else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if hasattr(b, "detach") else np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rel_max(a, b):
    a = a.detach().cpu().double().numpy() if hasattr(a, "detach") else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if hasattr(b, "detach") else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
