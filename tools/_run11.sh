timeout -k 5 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
NFPB200_BENCH_NO_HINT=1 timeout -k 5 300 python bench.py --steps 300 --no-train --no-cpu-baseline > gpurun_out/bench_nohint.json 2> gpurun_out/bench_nohint.err; echo "rc=$?"
timeout -k 5 300 python bench.py --steps 300 --no-train --no-cpu-baseline > gpurun_out/bench_hint.json 2> gpurun_out/bench_hint.err; echo "rc=$?"
