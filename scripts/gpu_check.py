"""Developer helper for the GPU box: runs every `-m gpu` test in its own process (a faulting kernel poisons a CUDA
context, and a lost mbarrier arrival must not take the whole run down) and writes gpurun_out/check.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    sel = sys.argv[1:] or ["tests"]
    r = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + sel, cwd=ROOT,
                       capture_output=True, text=True)
    ids = [l.strip() for l in r.stdout.splitlines() if "::" in l]
    log = open(os.path.join(OUT, "check.log"), "w")
    summary = []
    for tid in ids:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", tid], cwd=ROOT,
                               capture_output=True, text=True, timeout=240)
            status = "PASS" if p.returncode == 0 else "FAIL"
            tail = (p.stdout + p.stderr)[-3000:]
        except subprocess.TimeoutExpired as e:
            status, tail = "TIMEOUT", ((e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace"))[-3000:]
        dt = time.time() - t0
        summary.append(f"{status:8s} {dt:6.1f}s {tid}")
        print(summary[-1], flush=True)
        log.write(f"==== {status} {tid} ({dt:.1f}s)\n")
        if status != "PASS":
            log.write(tail + "\n")
        log.flush()
    log.write("\n".join(summary) + "\n")
    log.close()
    print(sum(s.startswith("PASS") for s in summary), "/", len(summary), "passed")


if __name__ == "__main__":
    main()
