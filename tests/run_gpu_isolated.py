"""Runs every selected GPU test in its own process (a device-side trap poisons the CUDA context of the process that
hit it, so isolation keeps one bad kernel from failing everything after it).  Development aid, not part of the suite:

    python tests/run_gpu_isolated.py [-k EXPR] [--timeout S] [files...]
"""
import argparse
import subprocess
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default=None)
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("files", nargs="*", default=["tests"])
    a = ap.parse_args()
    cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + a.files + (["-k", a.k] if a.k else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    ids = [l.strip() for l in out.splitlines() if "::" in l]
    print("collected %d tests" % len(ids), flush=True)
    bad = 0
    for nid in ids:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "--no-header", "-p", "no:cacheprovider", nid],
                               capture_output=True, text=True, timeout=a.timeout)
            ok, tail = r.returncode == 0, (r.stdout + r.stderr)
        except subprocess.TimeoutExpired as e:
            ok, tail = False, "TIMEOUT\n" + ((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
        print("%s %6.1fs %s" % ("PASS" if ok else "FAIL", time.time() - t0, nid), flush=True)
        if not ok:
            bad += 1
            lines = [l for l in tail.splitlines() if l.strip()]
            keep = [l for l in lines if ("Error" in l or "assert" in l or "sga_b200" in l or "E  " in l)][:25]
            print("    " + "\n    ".join(keep or lines[-15:]), flush=True)
    print("%d failed of %d" % (bad, len(ids)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
