set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_v17_bench.json 2> gpurun_out/r02_v17_bench.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r02_v17_bench_reference_arm.json 2> gpurun_out/r02_v17_bench_reference_arm.err; echo ref rc=$?
python tools/small_batch_latency.py > gpurun_out/r02_v17_small_batch_latency.txt 2>&1
python tools/single_chain_mh.py --iterations 3000 --burn-in 1000 --json gpurun_out/r02_v17_single_chain_mh.json > gpurun_out/r02_v17_single_chain_mh.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v17_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1; echo launches rc=$?
ncu --set full --import-source on --clock-control none -k regex:sepaihrd_batch_kernel -s 1 -c 1 -o gpurun_out/r02_v17_full -f python tools/prof_run.py --B 1048576 --launches 1 > gpurun_out/ncu_full_run.log 2>&1; echo full rc=$?
tail -2 gpurun_out/ncu_full_run.log
