#!/usr/bin/env python
"""In-kernel timeline of the conv GEMM launches of one eval forward (BASELINE configs[1]).

Needs the time-stamping build (`make dbg` -> videopose3d_b200/_lib/dbg/libvp3d_b200.so).  Every
conv_gemm launch records, for its first and its last CTA, globaltimer / clock64 at: kernel entry,
set-up done, dependency wait done, first TMA issued, first operands landed, first / last tile
committed by the MMA issuer, first / last accumulator seen by the epilogue, last store issued,
stores drained, CTA exit.  Prints one line per launch relative to the first launch's entry.

    python tools/timeline.py [precision]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("VP3D_LIB_PATH",
                      os.path.join(ROOT, "videopose3d_b200", "_lib", "dbg", "libvp3d_b200.so"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import videopose3d_b200 as vp  # noqa: E402
from videopose3d_b200 import _capi  # noqa: E402

ARC, C, J, F, N, T = [3, 3, 3, 3, 3], 1024, 17, 2, 1024, 243
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = torch.device("cuda:0")
lib = _capi.load()
lib.vp3d_debug_set_timeline.restype = ctypes.c_int
lib.vp3d_debug_set_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
g = torch.Generator().manual_seed(0)
x = (torch.rand(N, T, J, F, generator=g) * 2 - 1).to(dev)
m = vp.TemporalModel(J, F, J, filter_widths=ARC, channels=C).to(dev).eval().set_precision(prec)
L = 32
buf = torch.zeros(L, 2, 32, 2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
EV = ["entry", "setup", "depwait", "tma0", "land0", "mma_t0", "mma_tN", "epi_t0", "epi_tN_in",
      "epi_t0_out", "epi_tN_out", "drained", "exit",
      # first tile, first / second store block of epilogue warp 4: accumulator in registers, first /
      # second half staged, TMA store issued
      "b0_ld", "b0_h0", "b0_h1", "b0_st", "b1_ld", "b1_h0", "b1_h1", "b1_st",
      # producer thread: its own dependency wait done (after priming the W tiles)
      "p_wait", "p_primed", "p_w0"]
NE = len(EV)
with torch.no_grad():
    for _ in range(3):
        m(x)
    for rep in range(2):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        buf.zero_()
        lib.vp3d_debug_set_timeline(buf.data_ptr(), L)
        e0.record()
        m(x)
        e1.record()
        torch.cuda.synchronize()
        lib.vp3d_debug_set_timeline(None, 0)
        t = buf.cpu().numpy()
        print(f"# rep {rep}: forward {e0.elapsed_time(e1) * 1e3:.1f} us (events) ")
        t0 = None
        prev_exit = None
        for li in range(L):
            if t[li, 0, 0, 0] == 0:
                continue
            if t0 is None:
                t0 = int(t[li, 0, 0, 0])
            for cta in (0, 1):
                if t[li, cta, 0, 0] == 0:
                    continue
                gt = t[li, cta, :NE, 0].astype("int64")
                ck = t[li, cta, :NE, 1].astype("int64")
                rel = [(int(v) - t0) / 1e3 if v else float("nan") for v in gt]
                cyc = [(int(v) - int(ck[0])) if v else -1 for v in ck]
                print(f"L{li:02d} cta{'0' if cta == 0 else 'N'} us: " +
                      " ".join(f"{n}={v:.2f}" for n, v in zip(EV, rel)))
                print(f"          cyc: " + " ".join(f"{n}={v}" for n, v in zip(EV, cyc)))
