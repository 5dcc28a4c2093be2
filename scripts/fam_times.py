import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(round(d["value"],1), {k["kernel"]: round(k.get("us_per_view", k["avg_ms"]*1000),1) for k in d["kernels"]})
