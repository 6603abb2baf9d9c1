# Round 2, final 1-GPU call: the driver's sequence (suite, smoke, bench both arms) + ncu evidence of the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/r02_pytest_n1.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; python tools/show_bench.py gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err | cut -c 1-200
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c 1-400 gpurun_out/r02_bench_ref.json
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; python tools/summarize_launches.py gpurun_out/r02_launches.csv | tee gpurun_out/r02_launches.md | head -20
timeout 200 python tools/kbench.py --quick > gpurun_out/r02_kbench_quick.txt 2>&1; grep -E "forward|unpack8" gpurun_out/r02_kbench_quick.txt | cut -c 1-120
# full capture of the dominant kernel (and of the row-copy family) with source correlation
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transpose_tiles -c 1 -f -o gpurun_out/prof_r02_transpose python tools/kbench.py --quick > gpurun_out/r02_ncu_transpose.log 2>&1; tail -2 gpurun_out/r02_ncu_transpose.log | cut -c 1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_copy -c 1 -f -o gpurun_out/prof_r02_rows python tools/kbench.py --quick > gpurun_out/r02_ncu_rows.log 2>&1; tail -2 gpurun_out/r02_ncu_rows.log | cut -c 1-200
