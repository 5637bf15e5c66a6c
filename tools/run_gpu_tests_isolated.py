"""Run every `-m gpu` test in its own forked process with a timeout.

A device-side trap poisons the CUDA context of the process that hit it; isolating tests keeps one
bad kernel from hiding the results of the others during bring-up.  The parent never initialises
CUDA (fork-safe); each child runs exactly one test id.

    python tools/run_gpu_tests_isolated.py [-k expr] [--timeout 120] [paths...]
"""
import argparse
import os
import signal
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("paths", nargs="*", default=["tests"])
    ap.add_argument("-k", default=None)
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    os.chdir(ROOT)
    cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + args.paths
    if args.k:
        cmd += ["-k", args.k]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    ids = [l.strip() for l in out.splitlines() if "::" in l]
    print(f"collected {len(ids)} gpu tests", flush=True)

    import pytest  # noqa: F401  (import in the parent so children do not pay for it)
    import torch   # noqa: F401  (no CUDA initialisation here)

    results = {}
    for tid in ids:
        t0 = time.time()
        pid = os.fork()
        if pid == 0:
            import pytest as _pt
            rc = _pt.main(["-q", "-x", "--no-header", "-p", "no:cacheprovider", tid])
            sys.stdout.flush()
            os._exit(int(rc))
        status = None
        while time.time() - t0 < args.timeout:
            done, st = os.waitpid(pid, os.WNOHANG)
            if done:
                status = st
                break
            time.sleep(0.05)
        if status is None:
            os.kill(pid, signal.SIGKILL)
            os.waitpid(pid, 0)
            results[tid] = "TIMEOUT"
        elif os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0:
            results[tid] = "ok"
        else:
            results[tid] = f"FAIL({status})"
        print(f"[{results[tid]:>10}] {time.time() - t0:6.1f}s {tid}", flush=True)
    bad = {k: v for k, v in results.items() if v != "ok"}
    print(f"\n{len(results) - len(bad)} passed, {len(bad)} failed")
    for k, v in bad.items():
        print("  ", v, k)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
