"""Print per-kernel summary + hottest SASS lines of an ncu report: python scripts/ncu_hot.py rep.ncu-rep regex [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
def num(x):
    try: return float(x.replace(',', ''))
    except: return 0.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_active.min', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:3]:
    print(r[idx['Kernel Name']][:80])
    for k in keys:
        if k in idx: print(f"   {k} = {r[idx[k]]} {rows[1][idx[k]]}")
    st = [(h, num(r[i])) for h, i in idx.items() if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    st.sort(key=lambda x: -x[1])
    print("   stalls:", [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2)) for h, v in st[:6]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) > 10]
seen = set(); uniq = []
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
S = idx['# Samples']
tot = sum(num(r[S]) for r in uniq)
print("total samples", tot, "instructions", len(uniq))
for r in sorted(uniq, key=lambda r: -num(r[S]))[:topn]:
    st = {k: num(r[idx[k]]) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k and idx[k] < len(r)}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{int(num(r[S])):6d} {100*num(r[S])/tot:5.1f}% exec={r[idx['Instructions Executed']]:>8} thr={r[idx['Avg. Threads Executed']][:4]:>4} {r[idx['Source']][:66]:66s} {[(a.replace('stall_',''),int(b)) for a,b in top]}")
