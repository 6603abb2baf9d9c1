ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_bench.log 2>&1
tail -c 600 gpurun_out/r01e_ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:transpose_tiles -s 8 -c 3 -f -o gpurun_out/prof_r01e_transpose python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_full.log 2>&1
tail -c 300 gpurun_out/r01e_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:rows_copy -s 3 -c 2 -f -o gpurun_out/prof_r01e_rows python tools/ncu_rows.py > gpurun_out/r01e_ncu_rows.log 2>&1
tail -c 300 gpurun_out/r01e_ncu_rows.log
python tools/ncu_rows.py
ls -la gpurun_out/*.ncu-rep
