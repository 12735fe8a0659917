for mb in 1 4 5 6 7; do
  A5_NVCC_EXTRA="-DA5_STEP_MINB=$mb" python -m alphafive_b200.build --force > /dev/null 2>&1
  python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/step_mb$mb.json 2> gpurun_out/step_mb$mb.err
done
