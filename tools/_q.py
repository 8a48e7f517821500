import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["roofline"]["per_kernel_ms"])
