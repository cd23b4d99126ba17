#!/bin/bash
# the commands behind profiles/r02_*: final validation of the round on one B200 (run through gpurun from the repo root)
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final/tests.log 2>&1; tail -3 gpurun_out/final/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
timeout 600 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; echo "bench 20/5 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?"
S="python bench.py --steps 2 --warmup 6 --no-c3 --no-fp --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final/launches.csv $S > gpurun_out/final/ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_history -c 400 --csv --log-file gpurun_out/final/traffic.csv $S > gpurun_out/final/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -s 40 -c 1 -o gpurun_out/final/prof_history -f $S > gpurun_out/final/ncu_hist.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split3_attempts -s 6 -c 1 -o gpurun_out/final/prof_attempts -f $S > gpurun_out/final/ncu_att.log 2>&1
ls -la gpurun_out/final
