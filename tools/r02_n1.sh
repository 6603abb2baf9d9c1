# Round 2, first 1-GPU call (~8 min of box time):  gpurun --timeout 900 -- 'bash tools/r02_n1.sh'
mkdir -p gpurun_out
# 1. the whole GPU suite (includes everything written CPU-only at the end of round 1)
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
# 2. the bench line + its launch list
timeout 300 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; cut -c 1-700 gpurun_out/r02a_bench_n1.json; tail -3 gpurun_out/r02a_bench_n1.err
# 3. cache-policy variants of both kernel families against the default (6.75 TB/s)
for h in 0 1 2; do DTFFTB_CACHE_HINT=$h timeout 200 python tools/kbench.py --quick > gpurun_out/r02a_kbench_hint$h.txt 2>&1; tail -20 gpurun_out/r02a_kbench_hint$h.txt; done
# 4. sanitizers on the kernels through smoke() (small shapes: permutes, multi-peer unpack, plan execute)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_memcheck.txt 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/r02a_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_racecheck.txt 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r02a_racecheck.txt
