#!/bin/sh
# One measurement pass for profiles/: full bench line, ncu launch list of a short bench, ncu --set full capture of the
# factorization kernel, task traces. Usage (on the GPU box): sh tools/measure_round.sh <tag>
tag=${1:-r02}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_${tag}_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
PROFILE_OUT=gpurun_out/${tag} timeout 300 python tools/profile_factor.py > gpurun_out/profile_factor_${tag}.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_factor_tasks -s 3 -c 1 -o gpurun_out/prof_factor_${tag} -f \
    python tools/profile_factor.py > gpurun_out/ncu_factor_${tag}.log 2>&1
ncu -i gpurun_out/prof_factor_${tag}.ncu-rep --page raw --csv > gpurun_out/ncu_factor_tasks_${tag}_raw.csv 2>/dev/null
tail -2 gpurun_out/profile_factor_${tag}.log
tail -c 400 gpurun_out/bench_${tag}.json
