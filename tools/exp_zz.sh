timeout 200 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -4 > gpurun_out/t_net.log
for z in 1 0 1 0; do
  A5_TC_ZIGZAG=$z timeout 100 python tools/layer_times.py 11 4096 20 > gpurun_out/zz_layers_$z.txt 2>&1
  A5_TC_ZIGZAG=$z python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline >> gpurun_out/zz_bench_$z.json 2> gpurun_out/zz_$z.err
done
