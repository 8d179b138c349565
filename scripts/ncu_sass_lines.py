"""Attribute ncu's per-SASS-instruction counters (ncu -i X.ncu-rep --page source --csv, kernel number `sel`) to CUDA source
lines using nvdisasm -g line info of the same cubin (function-relative addresses match when the kernel's code is unchanged).
    python scripts/ncu_sass_lines.py <source.csv> <sel> <dis.txt> <mangled-substring> [top]
Prints, per source line (file:line incl. inlined-at chain's innermost line): share of warp instructions, average active
threads, share of lost lane-slots ((32 - threads) x instructions), share of stall samples."""
import collections
import csv
import re
import sys

src, sel, dis, mang = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
# --- line info per address from nvdisasm
addr_line = {}
infn = False
cur = None
for ln in open(dis):
    if ln.startswith("//---------------------"):
        infn = mang in ln
        cur = None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
    if m and cur:
        addr_line[int(m.group(1), 16)] = cur
# --- counters from the ncu CSV (kernels are concatenated: split where the address restarts)
rows = list(csv.reader(open(src)))
hdr = rows[1] if rows[0][0] != "Address" else rows[0]
kernels, curk = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curk = []; kernels.append(curk)
    elif r and curk is not None and re.match(r"^0x[0-9a-f]+$", r[0]):
        curk.append(r)
k = kernels[sel]
base = int(k[0][0], 16)
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0])
ti = tt = ts = 0
for r in k:
    a = int(r[0], 16) - base
    ins = int(r[iI] or 0); thr = int(r[iT] or 0); smp = int(r[iS] or 0)
    ln = addr_line.get(a, ("?", 0))
    g = agg[ln]; g[0] += ins; g[1] += thr; g[2] += smp
    ti += ins; tt += thr; ts += smp
print("kernel %d: warp instructions %.3e, avg threads %.2f, samples %d" % (sel, ti, tt / max(ti, 1), ts))
lost_tot = 32 * ti - tt
srcs = {}
for (f, l), (ins, thr, smp) in sorted(agg.items(), key=lambda kv: -(32 * kv[1][0] - kv[1][1]))[:top]:
    if f not in srcs:
        try:
            srcs[f] = open([p for p in ("diffiqult_b200/csrc/" + f, f) if __import__("os").path.exists(p)][0]).read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print("%5.2f%% inst  thr %5.1f  lost %5.2f%%  stall %5.2f%%  %s:%d  %s" % (100.0 * ins / ti, thr / max(ins, 1), 100.0 * (32 * ins - thr) / max(lost_tot, 1),
                                                                      100.0 * smp / max(ts, 1), f, l, text))
