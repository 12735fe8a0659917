# compute-sanitizer passes over the latency kernel tests and smoke() (run under gpurun; writes gpurun_out/san_*.txt)
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "small" 2>&1 | tail -15 > gpurun_out/san_small.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke 2>&1 | tail -15 > gpurun_out/san_smoke.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "small and 11" 2>&1 | tail -15 > gpurun_out/san_race.txt
