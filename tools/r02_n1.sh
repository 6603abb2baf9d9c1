# Round 2, first 1-GPU call:  gpurun --timeout 1500 -- 'bash tools/r02_n1.sh'
mkdir -p gpurun_out
nvidia-smi -L
# 1. the whole GPU suite (includes everything written CPU-only at the end of round 1); no -x: list every failure
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/r02a_pytest_n1.log | tail -25
# 2. the bench line
timeout 300 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; cut -c 1-900 gpurun_out/r02a_bench_n1.json; tail -3 gpurun_out/r02a_bench_n1.err
# 3. cache-policy variants of both kernel families against the default (6.75 TB/s)
for h in 0 1 2; do DTFFTB_CACHE_HINT=$h timeout 200 python tools/kbench.py --quick > gpurun_out/r02a_kbench_hint$h.txt 2>&1; tail -15 gpurun_out/r02a_kbench_hint$h.txt; done
# 4. sanitizers on the kernels through smoke() (small shapes: permutes, multi-peer unpack, plan execute)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_memcheck.txt 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/r02a_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_racecheck.txt 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r02a_racecheck.txt
