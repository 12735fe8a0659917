for mi in 1 2 3 4 8; do
  A5_MAX_INNER=$mi python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/inner_$mi.json 2> gpurun_out/inner_$mi.err
done
A5_NVCC_EXTRA="-DA5_STEP_MINB=7" python -m alphafive_b200.build --force > /dev/null 2>&1
for mi in 1 4; do
  A5_MAX_INNER=$mi python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/inner7_$mi.json 2> gpurun_out/inner7_$mi.err
done
