"""Per CUDA source line: executed warp instructions and stall samples (ncu --print-source cuda,sass)."""
import csv, sys, subprocess, collections
rep, sel = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.split("\n")
blocks = []; cur = None; fname = None
for ln in out:
    if ln.startswith('"File Path"'):
        fname = next(csv.reader([ln]))[1]
    elif ln.startswith('"Function Name"'):
        cur = {"fn": next(csv.reader([ln]))[1], "rows": [], "file": fname, "hdr": None}; blocks.append(cur)
    elif ln.startswith('"Line No"') and cur is not None:
        cur["hdr"] = next(csv.reader([ln]))
    elif cur is not None and cur["hdr"] and ln.startswith('"'):
        cur["rows"].append(next(csv.reader([ln])))
fns = []
for b in blocks:
    if b["fn"] not in fns: fns.append(b["fn"])
target = fns[sel]
print(target[:150])
agg = collections.defaultdict(lambda: [0, 0, ""])
tot_i = tot_s = 0
for b in blocks:
    if b["fn"] != target: continue
    h = b["hdr"]; iL = 0; iSrc = 1; iI = h.index("Instructions Executed"); iS = h.index("# Samples")
    line = None; src = ""
    for r in b["rows"]:
        if r[0] != "":
            line = (b["file"].split("/")[-1], int(r[0])); src = r[1]
        if len(r) > iI and r[iI].isdigit():
            a = agg[line]; a[0] += int(r[iI]); a[1] += int(r[iS]) if r[iS].isdigit() else 0; a[2] = src
            tot_i += int(r[iI]); tot_s += int(r[iS]) if r[iS].isdigit() else 0
print("total instr", tot_i, "samples", tot_s)
for k, (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%5.2f%% instr %5.2f%% stall  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), k[0], k[1], src.strip()[:100]))
