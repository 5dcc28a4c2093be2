"""Turn the ncu artefacts under gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py gpurun_out/launches_r1.csv gpurun_out/prof_r1_final.ncu-rep r1
"""
import csv, io, json, subprocess, sys, collections
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = "profiles"

# ---- launch list: per-kernel totals and shares (cold-cache, serialised: compare SHARES)
rows = [r for r in csv.reader(open(launch_csv)) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200splat::", "")
    unit = r[idx["Metric Unit"]]
    val = float(r[idx["Metric Value"]].replace(",", ""))
    val_us = val / 1000.0 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1000.0)
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += val_us
total = sum(v[1] for v in tot.values())
with open(f"{out_dir}/{tag}_ncu_launch_list.md", "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` on "
            "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n"
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n"
            "torch kernels (elementwise, copy, reduce) belong to the bench harness (loss, e2e), not to libb200splat.\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f} % |\n")
print("launch list:", len(tot), "kernels, total us", round(total))

# ---- full capture: one row per kernel instance
# rep: an .ncu-rep, or the CSV of its raw page (`ncu -i rep --page raw --csv`, exported on the GPU box when the report
# itself is too large to bring back)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); hdr = rr[0]; units = rr[1]; idx = {h: i for i, h in enumerate(hdr)}
keys = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"),
        ("sm__inst_executed_pipe_fma.sum", "fma_inst"), ("sm__inst_executed_pipe_xu.sum", "xu_inst")]
def num(x):
    try: return float(x.replace(",", ""))
    except: return None
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
traffic = {}
with open(f"{out_dir}/{tag}_ncu_full_summary.md", "w") as f:
    f.write(f"# ncu --set full summary ({tag}), batched step of 4 views, headline workload (1M Gaussians, SH3, 512x512)\n\n"
            "`ncu --set full --clock-control none --profile-from-start off` on `scripts/profile_batch.py headline_1m_512_sh3 3 4` (last step captured).\n"
            "One launch covers the 4 views of the step.  dram = dram__bytes_{read,write}.sum per launch.\n\n"
            "| kernel | us | DRAM read MB | DRAM write MB | DRAM % | SM % | issue % | warps % | regs | warp inst | top stalls |\n"
            "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for r in rr[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200splat::", "")
        g = lambda k: num(r[idx[k]]) if k in idx else None
        rd = to_bytes(g("dram__bytes_read.sum") or 0, units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(g("dram__bytes_write.sum") or 0, units[idx["dram__bytes_write.sum"]])
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), num(r[i]) or 0)
              for h, i in idx.items() if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
        st.sort(key=lambda x: -x[1])
        f.write(f"| `{name}` | {g('gpu__time_duration.sum'):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(g('launch__registers_per_thread'))} | "
                f"{g('smsp__inst_executed.sum') / 1e6:.1f} M | {', '.join(f'{a} {b:.1f}' for a, b in st[:3])} |\n")
        traffic.setdefault(name, []).append(rd + wr)
# family traffic per launch group (one step of the view batch): all kernels of the family summed
fam_of = [("preprocess_backward", "preprocess_bwd"), ("preprocess_kernel", "preprocess"), ("pair_count_kernel", "preprocess"),
          ("scan_lookback_kernel", "scan"), ("scan_duplicate_kernel", "duplicate"), ("duplicate_kernel", "duplicate"),
          ("radix_histogram_kernel", "sort"), ("onesweep_pass_kernel", "sort"), ("tile_partition", "sort"),
          ("partition_offsets_kernel", "sort"), ("tile_ranges", "ranges"),
          ("tile_order_kernel", "ranges"), ("render_forward_kernel", "render_fwd"),
          ("block_scatter_kernel", "render_bwd"), ("render_backward_kernel", "render_bwd")]
steps_captured = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tj = {}
for k, v in traffic.items():
    f_ = next((f for pre, f in fam_of if k.startswith(pre)), None)
    if f_:
        tj[f_] = tj.get(f_, 0.0) + sum(v) / steps_captured
json.dump({"headline_1m_512_sh3": tj, "_note": "dram__bytes_read.sum + dram__bytes_write.sum per LAUNCH GROUP (one step of "
           "the 4-view batch; all kernels of the family summed), from profiles/%s_ncu_full_summary.md" % tag},
          open(f"{out_dir}/ncu_traffic.json", "w"), indent=1)
print(tj)
