#!/bin/bash
# The measurement suite of a round on one B200 box (run through gpurun): GPU tests, smoke, the bench lines of every
# workload, both arms, and the ncu evidence.  Everything lands in gpurun_out/ (copied under profiles/ afterwards).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
python bench.py > gpurun_out/r2_bench_config3.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/ref.err
python bench.py --workload config2 > gpurun_out/r2_bench_config2.json 2>> gpurun_out/bench.err
python bench.py --workload config1 --steps 50 > gpurun_out/r2_bench_config1.json 2>> gpurun_out/bench.err
python bench.py --workload config5 --steps 5 > gpurun_out/r2_bench_config5.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_config3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --profile-range > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --profile-from-start off -c 29 -o gpurun_out/r2_top_kernels -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --profile-range > gpurun_out/ncu_f.log 2>&1
# the 29-launch capture is ~70 MB: summarise it here (gpurun brings back at most 64 MiB) and drop the report
python tools/ncu_summary.py full gpurun_out/r2_top_kernels.ncu-rep gpurun_out/r2_top_kernels.txt 29
python tools/ncu_summary.py dram gpurun_out/r2_ncu_dram.json 8192 gpurun_out/r2_top_kernels.ncu-rep
rm -f gpurun_out/r2_top_kernels.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:resum_kernel -c 1 -o gpurun_out/r2_resum_kernel -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --profile-range > gpurun_out/ncu_r.log 2>&1
ls -la gpurun_out
