"""Runs each GPU test function in its own process (a device trap poisons the CUDA context, so one bad kernel
must not hide the results of the others) with a hard timeout, and writes a compact report under gpurun_out/.

Usage (on the GPU box):  python tools/gpu_probe.py [tests/test_ops_gpu.py] [-k expr]
"""
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    files = [a for a in args if a.endswith(".py")] or ["tests/test_ops_gpu.py"]
    kexpr = args[args.index("-k") + 1] if "-k" in args else None
    timeout = int(os.environ.get("PROBE_TIMEOUT", "240"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    report = open(os.path.join(ROOT, "gpurun_out", "probe_report.txt"), "a")
    funcs = []
    for f in files:
        src = open(os.path.join(ROOT, f)).read()
        for m in re.finditer(r"^def (test_\w+)", src, re.M):
            if kexpr is None or re.search(kexpr, m.group(1)):
                funcs.append((f, m.group(1)))
    ok = 0
    for f, fn in funcs:
        t = time.time()
        cmd = [sys.executable, "-m", "pytest", f"{f}::{fn}", "-q", "-s", "-m", "gpu", "-p", "no:cacheprovider"]
        try:
            pr = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
            out, rc = pr.stdout, pr.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), -999
        status = "PASS" if rc == 0 else ("TIMEOUT" if rc == -999 else f"FAIL({rc})")
        ok += rc == 0
        line = f"{status:10s} {fn}  ({time.time() - t:.1f}s)"
        print(line, flush=True)
        report.write(line + "\n")
        info = [l for l in out.splitlines() if l.startswith("[")]
        if rc == 0 and info:
            print("\n".join(info[:20]), flush=True)
            report.write("\n".join(info[:20]) + "\n")
        if rc != 0:
            tail = "\n".join(out.splitlines()[-60:])
            keep = [l for l in out.splitlines() if l.startswith("[") or "Error" in l or "error" in l or "assert" in l or "wfl:" in l]
            report.write("\n".join(keep[:80]) + "\n--- tail ---\n" + tail + "\n\n")
            print("\n".join(keep[:30]), flush=True)
        report.flush()
    print(f"{ok}/{len(funcs)} test functions passed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
