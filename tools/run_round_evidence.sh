set -x
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo smoke_rc=$?
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2_pytest_final.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/r2_pytest_final.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench_rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo ref_rc=$?
timeout 300 python bench.py --workload config4 --steps 30 --warmup 5 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo cfg4_rc=$?
timeout 300 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo cfg3_rc=$?
timeout 400 python tools/sweep_bench.py --out gpurun_out/r2_config45_sweep.json > gpurun_out/r2_sweep.log 2>&1; echo sweep_rc=$?
timeout 200 python tools/fill_bench.py > gpurun_out/r2_fill_bench.log 2>&1; echo fillbench_rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt-layout > gpurun_out/r2_plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:estimate_kernel -c 40 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt-layout > gpurun_out/r2_ncu_l.log 2>&1; echo ncu_l_rc=$?
python tools/quick_time.py 16384 > gpurun_out/r2_plain_q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:estimate_kernel -s 3 -c 1 -o gpurun_out/r2_prof_cfg2 python tools/quick_time.py 16384 > gpurun_out/r2_ncu_q.log 2>&1; echo ncu_q_rc=$?
python tools/fill_bench.py --points comb4:4 --reps 2 > gpurun_out/r2_plain_f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fill_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_fill_tc python tools/fill_bench.py --points comb4:4 --reps 2 > gpurun_out/r2_ncu_f.log 2>&1; echo ncu_f_rc=$?
cat gpurun_out/r2_smoke.log | tail -2
