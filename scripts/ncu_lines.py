"""Stall samples of one kernel of an ncu report aggregated per CUDA source line (needs -lineinfo + --import-source on):
python scripts/ncu_lines.py rep.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
def num(x):
    try: return float(x.replace(',', ''))
    except: return 0.0
agg = collections.OrderedDict()
cur_file = None
hdr = None
first_fn = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        if first_fn is None: first_fn = r[1]
        elif r[1] != first_fn and False: break
        continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8: continue
    # the source text may contain commas that the csv writer did not quote: locate the numeric columns from the right
    # columns (from header): Line No, Source, Address, Source, WarpStall(all), WarpStall(not issued), # Samples, Instr Executed, ...
    extra = len(r) - len(hdr)
    line = r[0]
    try:
        samples = num(r[6 + extra]); execd = num(r[7 + extra])
    except Exception:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, [0.0, 0.0, ','.join(r[1:2 + extra])[:90]])
    a[0] += samples; a[1] += execd
tot = sum(v[0] for v in agg.values())
print(first_fn, "total samples", tot)
for (f, line), (s, e, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{int(s):7d} {100*s/tot:5.1f}%  inst={int(e):>10}  {f}:{line}  {text}")
