import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "both %.4g"%d["e2e"]["both_surfaces_serial"]["value"])
