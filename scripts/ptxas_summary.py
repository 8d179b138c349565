"""Summarise `nvcc -Xptxas -v` output: kernel (demangled), registers, spill bytes, static smem.
    nvcc ... -Xptxas -v -c x.cu -o x.o 2> log ; python scripts/ptxas_summary.py log [filter]"""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt):
    rows.append(m.groups())
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for nm, r in zip(names, rows):
    short = re.sub(r"\(.*", "", nm).replace("void ", "").replace("dq::", "")
    short = short.replace("true", "1").replace("false", "0").replace("(bool)", "")
    if flt in short:
        print("%-60s regs %3s  stack %4s  spill st/ld %4s/%4s  smem %s" % (short, r[4], r[1], r[2], r[3], r[5] or 0))
