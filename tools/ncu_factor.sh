set -x
timeout 200 python tools/profile_factor.py > gpurun_out/ncu_plain_factor.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_factor_tasks -s 3 -c 1 -o gpurun_out/prof_factor_r02 -f python tools/profile_factor.py > gpurun_out/ncu_factor_r02.log 2>&1
tail -3 gpurun_out/ncu_factor_r02.log
ls -la gpurun_out/prof_factor_r02.ncu-rep
