"""profiles/launches_r02.csv (ncu --metrics gpu__time_duration.sum --csv) -> profiles/launches_r02_summary.md."""
import collections, csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "launches_r02.csv")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "launches_r02_summary.md")
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki].split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
out = ["# One headline step (tf32+refine, binary16 panel operands), n=4096, d=20, 2^20 candidates -- ncu launch list, "
       "round 2", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 2 "
       "--warmup 1 --no-fit --no-cpu --no-e2e --no-other-mode` right after the same command had exited 0 without ncu "
       "(`profiles/launches_r02.csv`). The window covers the state set-up (one factorisation: the k_chol_* / k_trtri_* "
       "/ k_syrk_* rows), the warm-up step and most of the timed steps. Per-launch times are cold-cache and "
       "serialised: compare shares.", "",
       "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    out.append(f"| {k} | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f}% |")
keys = ("vnorm_tf32", "kstar_umma", "acq_eval", "acq_filter", "filter_heavy", "merge_survivors", "prescale_cand", "refine")
sc = [k for k in agg if any(x in k for x in keys)]
st = sum(agg[k][1] for k in sc)
panel = [k for k in agg if "vnorm_tf32_2sm" in k][0]
out += ["", f"Scoring kernels only: {st / 1e6:.2f} ms, of which the panel {100 * agg[panel][1] / st:.1f} % (bench.py's "
        "`roofline.share_of_step`, from CUDA events inside the step, reads 0.77-0.80).",
        "Earlier list of this round (TF32 operands, same command): panel 1009.5 us per 37888-candidate launch, K* "
        "122.0 us, selection 51 + 59 us (`k_acq_topk` + `k_topk_merge_running`)."]
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
