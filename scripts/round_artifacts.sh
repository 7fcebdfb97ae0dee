# Everything the round's profiles/ entries come from, in one GPU call (each command under its own timeout).
set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2g_tests.log
timeout 300 python bench.py --config 2 > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-pipelined --no-cpu-baseline --no-graph > gpurun_out/r2g_pre_ncu.json 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_config2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-pipelined --no-cpu-baseline --no-graph > gpurun_out/r2g_ncu_launches.log 2>&1
timeout 100 python scripts/one_step.py > gpurun_out/r2g_one_step.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -s 8 -c 8 -f -o gpurun_out/r2g_step_full python scripts/one_step.py > gpurun_out/r2g_ncu_full.log 2>&1
timeout 60 python scripts/layer_times.py --tag final > gpurun_out/r2g_layers.txt 2>&1
timeout 60 python scripts/layer_times.py --batch 1 --nc 2 --cfg 1 --tag final-config1 >> gpurun_out/r2g_layers.txt 2>&1
timeout 120 python scripts/time_proposal.py 1 2 4 8 16 32 64 > gpurun_out/r2g_proposal_vs_batch.txt 2>&1
