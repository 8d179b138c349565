"""Build profiles/r2_summary.md from the round-2 ncu artefacts under profiles/ (launch list, FP64 metric pass, full-set raw pages).
    python scripts/profile_summary_r2.py
"""
import collections
import csv
import json
import re

out = []
w = out.append
w("# Round 2 profile summary (B200, dense (H2O)_32 / STO-3G, `python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --profile-range`)\n")
w("All passes use `ncu --profile-from-start off` with `bench.py --profile-range` (cudaProfilerStart/Stop around the timed step), so the 32 "
  "monomer guess SCFs and the warm-up step are NOT in the lists; every pass ran after the same command had exited 0 without ncu.  "
  "Per-launch times under ncu are cold-cache and serialised (the 7 class launches of a pass overlap on side streams in a real run): "
  "compare SHARES.  Code state: the commit that adds this file.\n")
# ---- launch list
lines = [l for l in open("profiles/r2_launches.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dq::", "")
    v = float(row["Metric Value"].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
w("## Launch list of one timed step (`profiles/r2_launches.csv`, `--metrics gpu__time_duration.sum --clock-control none`)\n")
w("%d launches, %.1f ms in total.\n" % (sum(a[0] for a in agg.values()), tot))
w("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
fam = collections.OrderedDict()
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot > 0.002:
        w("| `%s` | %d | %.2f | %.1f %% | %.3f |" % (k, n, t, 100 * t / tot, t / n))
    f = re.sub(r"<.*", "", k); fam[f] = fam.get(f, 0) + t
w("\nBy kernel family: " + ", ".join("`%s` %.1f %%" % (k, 100 * v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:8]) + "\n")
# ---- FP64 pass
d = json.load(open("profiles/fp64_executed_dense.json"))
w("## Executed FP64 work (`profiles/r2_fp64_executed_dense.csv` -> `profiles/fp64_executed_dense.json`, the file `bench.py` reads)\n")
w("Walked primitive-quartet slots per pass: %.4g (evaluated unique quartets: 2.304e10; the slots also hold the diagonal-tile redundancy "
  "and the lanes whose Gaussian product underflowed).\n" % d["primitive_quartets_evaluated_per_pass"])
w("| kernel family | launches | ms under ncu | FP64 thread instructions / slot | flop / slot | executed TFLOP/s | FP64 pipe busy (warp level) | lanes on per instruction | DRAM bytes |\n|---|---|---|---|---|---|---|---|---|")
for k, e in d.items():
    if not isinstance(e, dict):
        continue
    w("| `%s` | %d | %.2f | %s | %s | %.2f | %.1f %% | %.1f | %s |" % (
        k, e["launches"], e["ms_under_ncu"],
        "%.1f" % e["fp64_instructions_per_primitive_quartet"] if "fp64_instructions_per_primitive_quartet" in e else "-",
        "%.1f" % e["flop_per_primitive_quartet"] if "flop_per_primitive_quartet" in e else "-",
        e["executed_tflops_under_ncu"], 100 * e["pipe_fp64_active_frac"], e.get("lanes_per_instruction", 0.0),
        ("%.3g per step" % e["dram_bytes_per_step"]) if "dram_bytes_per_step" in e else ("%.4g per launch" % e["dram_bytes_per_launch"])))
w("\nFor comparison, the same pass at the START of the round (round-1 Boys code on this geometry): `k_eri_tiles` 281.7 ms, 58.9 FP64 "
  "instructions / slot, 17.7 lanes; `k_eri_tiles_b` (warp-batched Boys, then the default) 302.3 ms, 78.1, 23.5 lanes; "
  "`k_eri_grad_chunks` 655.1 ms, 149.5, 19.1 lanes.\n")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out_rows = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        e = dict(zip(hdr, r))
        u = dict(zip(hdr, units)).get("gpu__time_duration.sum", "ms")       # ncu picks the unit per report
        try:
            t = float(e["gpu__time_duration.sum"].replace(",", ""))
            e["gpu__time_duration.sum"] = str(t * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(u, 1.0))
        except Exception:
            pass
        out_rows.append(e)
    return out_rows


def stall_table(title, entries):
    w("### %s\n" % title)
    keys = ["wait", "no_instruction", "barrier", "math_pipe_throttle", "short_scoreboard", "not_selected", "branch_resolving",
            "long_scoreboard", "dispatch_stall", "lg_throttle", "mio_throttle"]
    w("| capture | ms | regs | warps active % | issue active % | FP64 pipe % | lanes/inst | " + " | ".join(keys) + " |")
    w("|---|---|---|---|---|---|---|" + "---|" * len(keys))
    for label, e in entries:
        def g(k):
            try:
                return float(e[k].replace(",", ""))
            except Exception:
                return float("nan")
        w("| %s | %.1f | %d | %.1f | %.1f | %.1f | %.1f | " % (
            label, g("gpu__time_duration.sum"), g("launch__registers_per_thread"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            g("smsp__thread_inst_executed_per_inst_executed.ratio")) +
          " | ".join("%.2f" % g("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k) for k in keys) + " |")
    w("\n(stall columns: warps stalled for that reason per issued instruction, `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`)\n")


w("## Full-set captures (`ncu --set full --clock-control none --import-source on`, raw / details pages under `profiles/r2_ncu_full_*.csv`)\n")
before = raw("profiles/r2_ncu_full_grad_dense_before.raw.csv")
after = raw("profiles/r2_ncu_full_grad.raw.csv")
ent = []
for e in before[:1]:
    ent.append(("BEFORE (round start): " + re.sub(r"\(.*", "", e["Kernel Name"]).replace("void ", ""), e))
for e in after[:1]:
    ent.append(("AFTER: " + re.sub(r"\(.*", "", e["Kernel Name"]).replace("void ", ""), e))
stall_table("Gradient kernel, class (sp|pp) - the largest launch of the dominant family (AFTER = the first kernel of the deferred form)", ent)
_g = raw("profiles/r2_ncu_full_gdef.raw.csv")
stall_table("Second kernel of the deferred gradient form (loop-regime quartets, one launch)", [(re.sub(r"\(.*", "", e["Kernel Name"]).replace("void ", "").replace("dq::", ""), e) for e in _g][:1])
w("In between (Boys series fully unrolled in the gradient kernels, 58 KB of code): `no_instruction` 2.63, 195 ms for this launch - the "
  "instruction-cache episode of DESIGN.md 3.\n")
_v = raw("profiles/r2_ncu_full_eri.raw.csv")
stall_table("Value kernels (the two launches captured)", [(re.sub(r"\(.*", "", e["Kernel Name"]).replace("void ", "").replace("dq::", ""), e) for e in _v if e.get("smsp__inst_executed.sum", "nan") not in ("nan", "-nan", "")][:2])
jk = raw("profiles/r2_ncu_full_jk.raw.csv")[0]
stall_table("J/K contraction (one Fock build)", [("k_jk_stream", jk)])


def gj(k):
    try:
        return float(jk[k].replace(",", ""))
    except Exception:
        return float("nan")


w("`k_jk_stream`: DRAM read %s %s + write %s %s per launch against 2.61 GB algorithmic; `dram__throughput` %.1f %% of peak; shared-memory "
  "bank conflicts %s; the compare-and-swap loops of the K-strip adds (`ATOMS.CAST.SPIN.64`) are what keeps the issue slots busy - see "
  "`profiles/r2_sass_k_jk_stream.txt` for the TMA / mbarrier instructions (UBLKCP, SYNCS) and DESIGN.md 3.\n" % (
      jk.get("dram__bytes_read.sum"), "GB", jk.get("dram__bytes_write.sum"), "MB", gj("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
      jk.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")))
w("## A/B runs of the round (bench.py phases, ms; each row one process on a fresh box)\n")
w("| variant | ERI pass | gradient pass | J/K per build | step |\n|---|---|---|---|---|")
for row in [("round start (round-1 kernels, warp-batched Boys default)", "305", "644", "0.742", "982"),
            ("plain value kernel (`DQ_ERI_KERNEL=plain`)", "282", "644", "0.742", "958"),
            ("+ per-class asymptotic thresholds", "274", "628", "0.742", "934"),
            ("+ loop-free mid range, Horner series (loops unrolled everywhere)", "236-258", "621", "0.74", "889"),
            ("loops rolled everywhere", "286", "568", "0.74", "887"),
            ("loops unrolled in value kernels, rolled in gradient kernels (final)", "230-237", "565", "0.74", "827-835"),
            ("row-compacted loop regime in the value kernel (reverted)", "234-249", "566", "0.74", "832-849"),
            ("private K strips per warp (`DQ_JK_PRIVATE=1`)", "-", "-", "0.96", "-"),
            ("K-run J/K kernel (no strip atomics, one CTA barrier per run; reverted)", "-", "-", "0.94", "-"),
            ("contiguous stages per warp + K-role sums in registers (final J/K kernel)", "229.5", "565", "0.678", "826"),
            ("deferred value kernel, list step-major / by argument bin / by bin and only for classes with >= 2 p functions (final)", "232-242 / 219.5-226 / 213-221", "565", "0.678", "829 / 816 / 810"),
            ("+ argument bins packed in registers at marking time (no second evaluation of t when listing): six processes", "208.2 208.2 208.2 209.2 219.0 222.1", "565-566", "=", "805-819"),
            ("(pp|pp) gradient class at two CTAs per SM (128 registers)", "208", "556", "=", "796"),
            ("gradient pass in two kernels for classes with >= 2 p functions (k_eri_grad_chunks<DEFER> + k_eri_grad_deferred): bins recomputed / packed in registers of the first kernel / one byte per marked step in global memory (final)", "208", "529 / 537 / 516", "=", "769 / 776 / 755"),
            ("... also for the classes with one p function", "=", "518", "=", "-"),
            ("timing experiment: loop-regime quartets sent through the loop-free branch (wrong numbers)", "133.7", "324.2", "-", "-"),
            ("pair records pinned in L2 during the ERI pass: six processes", "229.5 229.5 229.5 229.6 230.9 238.8 251.6", "565-567", "=", "826-848"),
            ("`DQ_JK_LDG=1` (register-path loads instead of the TMA ring)", "-", "-", "0.79", "-"),
            ("`DQ_ACCUM=float` (floating-point atomics)", "=", "611-618", "0.735", "-"),
            ("`DQ_SORT=0` (caller's function order instead of exponent bucket + Morton)", "=", "675", "=", "-"),
            ("`DQ_GRAD_CHUNK` = 6 / 12 (default) / 24 / 48 / 96", "=", "617 / 620 / 627 / 637 / 659 (before the rolled loops)", "=", "-")]:
    w("| " + " | ".join(row) + " |")
w("\nMulti-GPU (torchrun, NCCL, final code): N=2 397 ms per step (ERI 109, gradient 259, SCF 17.1, reverse 4.4); "
  "N=8 121 ms (30.3 / 64.9 / 15.5 / 3.2; J/K 0.19 ms per build).  One code state earlier (gradient pass in one kernel): N=2 414, N=4 225, N=8 125 ms; "
  "at the start of the round's GPU work: N=2 462 ms, N=8 128 ms.\n")
open("profiles/r2_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
