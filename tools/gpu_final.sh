set -x
cd $GRAFT_REPO_ROOT
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo rc=${PIPESTATUS[0]}) > gpurun_out/pytest_gpu_final2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo rc=$? >> gpurun_out/smoke_final.log
timeout 300 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_lazy.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-other > gpurun_out/ncu_bench_lazy.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:update_tc -c 16 -f -o /tmp/upd_lazy python tools/solve_probe.py 296 1024 1280 3 > gpurun_out/ncu_upd_lazy.log 2>&1; echo rc=$?
ncu -i /tmp/upd_lazy.ncu-rep --page raw --csv > gpurun_out/r02_ncu_update_tc_lazy_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu_final2.log; cat gpurun_out/smoke_final.log | tail -2; head -c 400 gpurun_out/r02_bench_final.json
