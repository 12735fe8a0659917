set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest_gpu.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 64 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --sims 20 --no-graph > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_conv -s 13 -c 10 -o gpurun_out/r1_tc_conv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --sims 8 --no-graph > gpurun_out/r1_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 2 -o gpurun_out/r1_k_step python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --sims 8 --no-graph > gpurun_out/r1_ncu_full2.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu.log; cat gpurun_out/r1_bench.json
