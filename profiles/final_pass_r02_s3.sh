set -x
python -m pytest tests -m gpu -q > gpurun_out/gputest_r02_s3.log 2>&1; tail -2 gpurun_out/gputest_r02_s3.log
python bench.py > gpurun_out/bench_r02_s3.json 2> gpurun_out/bench_r02_s3.err; tail -c 300 gpurun_out/bench_r02_s3.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_s3_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/bench_r02_s3_reference_arm.json
python bench.py --workload synthetic-radial-10k-homes-x96 --no-split --steps 2 --warmup 3 --no-exact --no-convergence > gpurun_out/bench_r02_s3_radial10k.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r02_s3_radial10k.json
python bench.py --workload synthetic-multifeeder-125k-homes-per-gpu-x96 --steps 3 --warmup 3 --no-exact --no-convergence --no-cpu-baseline > gpurun_out/bench_r02_s3_round1_workload.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r02_s3_round1_workload.json
ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_s3.csv python bench.py --steps 1 --warmup 1 --pipelines 1 --no-cpu-baseline --no-exact --no-convergence > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches_r02_s3.csv
ncu --set full --clock-control none --import-source on --launch-skip 172 --launch-count 14 -o gpurun_out/r02_s3_final python profiles/ncu_target.py > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
