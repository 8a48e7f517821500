# usage: TAG=r02h bash tools/batch_env.sh "NAME1:ENV=VAL ENV2=VAL" "NAME2:" ...   -- one bench line per environment setting
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG:-r02x}_bench_$name.json 2> gpurun_out/${TAG:-r02x}_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG:-r02x}_bench_$name.json"))
    print("$name", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()}, "e2e", "%.3g"%d["e2e"]["value"])
except Exception as e:
    print("$name", "FAILED", e)
PY
done
