#!/usr/bin/env python
"""List the backward branches (loops) of a kernel with their body size in SASS instructions."""
import re, subprocess, sys, tempfile, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(ROOT, "mujoco_playground_b200", "libackb.so")
kern = sys.argv[1] if len(sys.argv) > 1 else "step_kernelIfLi4"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")]
    sass = subprocess.run(["nvdisasm", "--print-line-info"] + cub, cwd=td, capture_output=True, text=True).stdout
fn = None; labels = {}; instrs = []; cur=None
for line in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
    if m: fn = m.group(1); continue
    if not fn or kern not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur=(os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"^(\.L_x_\d+):", line)
    if m: labels[m.group(1)] = len(instrs); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m: instrs.append((int(m.group(1), 16), m.group(2), cur))
print("total", len(instrs))
for i, (addr, txt, cur) in enumerate(instrs):
    m = re.search(r"\bBRA\b.*?(\.L_x_\d+)", txt)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
        j = labels[m.group(1)]
        print(f"loop: body {i - j + 1:5d} instrs  [{j}..{i}]  back-edge at src {cur}  head src {instrs[j][2]}")
