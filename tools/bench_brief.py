import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["value"], d["ms_per_step"]); print(d["host_call_us"]); print([(k["name"][5:],round(k["avg_us"]),round(k["launches_per_step"],1)) for k in d["host_segments"]])
print([(k["name"], round(k["avg_us"],1)) for k in d["kernels"][:12]])
