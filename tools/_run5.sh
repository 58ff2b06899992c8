python tools/_diag.py two > gpurun_out/diag4.log 2>&1; cat gpurun_out/diag4.log
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | head -2
python bench.py --no-prove --no-cpu-baseline > gpurun_out/bench_fix.json 2> gpurun_out/bench_fix.err
for c in 16 17 20; do python tools/prove_bench.py --k 20 --steps 4 --precompute $c > gpurun_out/prove_c$c.json 2> gpurun_out/prove_c$c.err; done
