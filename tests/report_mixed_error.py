"""Not a test: prints the relative error of the eval precision modes on the C = 1024 goldens
(run on the GPU box: `python tests/report_mixed_error.py`)."""
import numpy as np
import torch

from conftest import golden_names, load_golden
import videopose3d_b200 as vp

dev = torch.device("cuda:0")
for name in golden_names():
    meta, sd, x, y_ref, _ = load_golden(name)
    if meta.get("train") or meta["C"] != 1024:
        continue
    cls = getattr(vp, meta["cls"])
    kw = dict(filter_widths=meta["fw"], causal=meta["causal"], channels=meta["C"])
    if meta["cls"] == "TemporalModel":
        kw["dense"] = meta["dense"]
    m = cls(meta["J"], meta["F"], meta["Jout"], **kw)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    out = []
    for prec in ("fp16", "bf16x3", "mixed", "bf16"):
        with torch.no_grad():
            y = m.set_precision(prec)(x.to(dev)).cpu().numpy()
        out.append(f"{prec} {np.abs(y - y_ref).max() / np.abs(y_ref).max():.2e}")
    print(name, " | ".join(out), flush=True)
