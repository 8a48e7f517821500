set -x
python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02z_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; tail -c 600 gpurun_out/r02z_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; tail -c 400 gpurun_out/r02z_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02z_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cell_patch|k_finish_reduce|k_subs_ring" -s 12 -c 3 -o gpurun_out/r02z_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02z_full.log 2>&1; tail -2 gpurun_out/r02z_full.log
