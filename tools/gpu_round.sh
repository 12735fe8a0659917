# usage: bash tools/gpu_round.sh TAG   (run under gpurun; writes gpurun_out/TAG_*)
T=${1:-r01b}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python tools/layer_times.py 11 4096 20 > gpurun_out/${T}_layers.txt 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 64 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --sims 20 --no-graph > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_conv -s 13 -c 10 -o gpurun_out/${T}_tc_conv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --sims 8 --no-graph > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_bench.json; cat gpurun_out/${T}_layers.txt
