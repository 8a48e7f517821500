python -m pytest tests/test_gpu_parity.py -x -q -k "division or first_steps or alternative or bitwise or chaotic" > gpurun_out/${TAG:-r02x}_pytest.log 2>&1; tail -3 gpurun_out/${TAG:-r02x}_pytest.log
for v in ${VARIANTS:-base}; do
  if [ $v = base ]; then unset WCFE_LIB; else export WCFE_LIB=$PWD/wholecellfe_b200/libwcfe_$v.so; fi
  python bench.py --steps 300 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG:-r02x}_bench_$v.json 2> gpurun_out/${TAG:-r02x}_bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG:-r02x}_bench_$v.json"))
print("$v", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()}, "e2e", "%.3g"%d["e2e"]["value"])
PY
done
