set -x
python bench.py > gpurun_out/bench_r4d.json 2> gpurun_out/bench_r4d.err
python tests/gpu_stage_profile.py 1024 > gpurun_out/stage_r4d.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-extra > gpurun_out/bench_short_r4d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r4d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-extra > gpurun_out/ncu_launch_r4d.log 2>&1
SNACB_NO_TRIM=1 python tests/gpu_one.py 1024 fp16 1 > gpurun_out/one_r4d.log 2>&1 && \
SNACB_NO_TRIM=1 ncu --set full --clock-control none --import-source on -k regex:k_chain -c 3 -f -o gpurun_out/prof_chain_r4d python tests/gpu_one.py 1024 fp16 1 > gpurun_out/ncu_chain_r4d.log 2>&1
SNACB_NO_TRIM=1 ncu --set full --clock-control none -k regex:'k_vq_stem|k_gemm_tc|k_resunit2|k_convt|k_tail' -c 11 -f -o gpurun_out/prof_mem_r4d python tests/gpu_one.py 1024 fp16 1 > gpurun_out/ncu_mem_r4d.log 2>&1
tail -2 gpurun_out/ncu_mem_r4d.log
cut -c1-300 gpurun_out/bench_r4d.json
