"""Summarise `nvcc -Xptxas -v` for one TU of the CUDA library: python tools/ptxas_summary.py [name-filter] [file.cu]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(ROOT, "neorl-industrial-gym_b200", "csrc")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xptxas", "-v",
       "-c", os.path.join(csrc, sys.argv[2] if len(sys.argv) > 2 else "nig_step.cu"), "-o", "/tmp/nig_ptxas.o"]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
out = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout
pat = sys.argv[1] if len(sys.argv) > 1 else ""
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(.*)' for", line)
    if m:
        name = m.group(1).replace("nig::", "").replace("void ", "")
        name = re.sub(r"\(.*\)$", "", name)
        spill = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = m.groups()
    m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
    if m and name and (pat in name):
        sm = re.search(r"(\d+) bytes smem", line)
        print(f"{name:60s} regs {m.group(1):>3s} stack/spill {spill} smem {sm.group(1) if sm else 0}")
