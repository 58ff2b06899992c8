(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/prove_bench.py --k 20 --steps 4 > gpurun_out/prove_agg.json 2> gpurun_out/prove_agg.err
H2A_MSM_ROUNDS_BIAS=1 python tools/prove_bench.py --k 20 --steps 4 > gpurun_out/prove_agg_r1.json 2> gpurun_out/prove_agg_r1.err
H2A_MSM_ROUNDS_BIAS=-1 python tools/prove_bench.py --k 20 --steps 4 > gpurun_out/prove_agg_rm1.json 2> gpurun_out/prove_agg_rm1.err
H2A_PROVE_NTT_PRIO=1 python tools/prove_bench.py --k 20 --steps 4 > gpurun_out/prove_agg_prio.json 2> gpurun_out/prove_agg_prio.err
python bench.py --no-prove --no-cpu-baseline > gpurun_out/bench_agg.json 2> gpurun_out/bench_agg.err
H2A_MSM_ROUNDS_BIAS=1 python bench.py --no-prove --no-cpu-baseline > gpurun_out/bench_agg_r1.json 2> gpurun_out/bench_agg_r1.err
python tools/sweep.py --msm 20 --ntt 20 --precompute 20 > gpurun_out/sweep20.jsonl 2>&1
H2A_MSM_ROUNDS_BIAS=1 python tools/sweep.py --msm 20 --ntt 20 --precompute 20 > gpurun_out/sweep20_r1.jsonl 2>&1
