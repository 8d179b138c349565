"""Summarise an ncu source page (SASS view): instruction mix and top stall addresses."""
import csv, sys, subprocess, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.split("\n")
kern = None; rows = []; hdr = None
kernels = []
for ln in out:
    if ln.startswith('"Kernel Name"'):
        kern = ln; hdr = None; rows = []; kernels.append((kern, rows)); continue
    if ln.startswith('"Address"'):
        hdr = next(csv.reader([ln])); continue
    if hdr and ln.startswith('"0x'):
        rows.append(dict(zip(hdr, next(csv.reader([ln])))))
sel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern, rows = kernels[sel]
print(kern[:160])
tot = sum(int(r["Instructions Executed"]) for r in rows)
samples = sum(int(r["# Samples"]) for r in rows)
mix = collections.Counter()
for r in rows:
    op = r["Source"].split()[0]
    if op.startswith("@"): op = r["Source"].split()[1]
    op = op.split(".")[0]
    mix[op] += int(r["Instructions Executed"])
print("total warp instructions", tot, "samples", samples, "static instrs", len(rows))
for op, c in mix.most_common(28):
    print("  %-10s %6.2f%%" % (op, 100.0 * c / tot))
print("top stall sample addresses:")
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:25]:
    st = {k: int(v) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("  %5.2f%% %-60s %s" % (100.0 * int(r["# Samples"]) / samples, r["Source"][:60], top))
