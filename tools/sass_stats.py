#!/usr/bin/env python
"""Build libackb.so and print register / spill / SASS-size statistics of the step kernels (no GPU needed).

    python tools/sass_stats.py [--no-build] [--kernel step_kernelIfLi4] [--lines]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-build", action="store_true")
    ap.add_argument("--kernel", default="step_kernelIfLi4")
    ap.add_argument("--lines", action="store_true", help="attribute SASS instructions to source lines")
    ap.add_argument("--defs", default="", help="extra -D flags, comma separated")
    ap.add_argument("--out", default=os.path.join(ROOT, "mujoco_playground_b200", "libackb.so"))
    a = ap.parse_args()
    csrc = os.path.join(ROOT, "mujoco_playground_b200", "csrc")
    if not a.no_build:
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-diag-suppress", "177", "-Xptxas", "-v",
               "-I" + os.path.join(ROOT, "include"), "-I" + csrc, "-shared", "-Xcompiler", "-fPIC", "-o", a.out,
               os.path.join(csrc, "ackb_kernels.cu")] + ["-D" + d for d in a.defs.split(",") if d]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(r.stdout, r.stderr)
            sys.exit(1)
        cur = None
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = re.sub(r"^_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}\d+", "", m.group(1))[:24]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                stack = m.groups()
            m = re.search(r"Used (\d+) registers", line)
            if m and cur and "step_kernel" in cur:
                print(f"{cur:26s} regs {m.group(1):>3s}  stack {stack[0]:>5s}  spill st/ld {stack[1]}/{stack[2]}")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", a.out], cwd=td, capture_output=True)
        cubins = [f for f in os.listdir(td) if f.endswith(".cubin")]
        sass = subprocess.run(["nvdisasm", "--print-line-info"] + cubins, cwd=td, capture_output=True, text=True).stdout
    fn, cur = None, None
    size = collections.Counter()
    ops = collections.Counter()
    lines = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            size[fn] += 1
            if a.kernel in fn:
                ops[m.group(2).split(".")[0]] += 1
                lines[cur] += 1
    for k, v in sorted(size.items(), key=lambda x: x[1]):
        if "step_kernel" in k:
            print(f"{v:6d} SASS instructions ({v * 16 / 1024:.0f} KiB)  {re.sub(r'^_ZN.*?(step_kernel)', 'step_kernel', k)[:20]}")
    print(a.kernel, "opcode mix:", ops.most_common(24))
    if a.lines:
        for k, v in sorted(lines.items(), key=lambda x: -x[1])[:40]:
            print(v, k)


if __name__ == "__main__":
    main()
