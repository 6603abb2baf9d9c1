# Round 2, 8-GPU call:  gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_n8.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
# 1. first run of the process-grid search on real GPUs: 512^3 at 8 ranks, 1x8 vs 8x1 vs 2x4 vs 4x2
DTFFTB_LOG=1 timeout 300 $TR --master-port 29551 tools/grid_search_probe.py > gpurun_out/r02a_grid_search_n8.txt 2>&1; tail -25 gpurun_out/r02a_grid_search_n8.txt
# 2. multi-GPU suite (fused backend only keeps it short) and the bench line
DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 $TR --master-port 29552 bench.py --gpus 8 > gpurun_out/r02a_bench_n8.json 2> gpurun_out/r02a_bench_n8.err; cut -c 1-700 gpurun_out/r02a_bench_n8.json; tail -3 gpurun_out/r02a_bench_n8.err
# 3. (only if the 2-GPU run of DTFFTB_FUSED_SYNC=1 was green) the exchange transposition with folded barriers
DTFFTB_FUSED_SYNC=1 timeout 300 $TR --master-port 29553 bench.py --gpus 8 > gpurun_out/r02a_bench_n8_fusedsync.json 2> gpurun_out/r02a_bench_n8_fusedsync.err; cut -c 1-700 gpurun_out/r02a_bench_n8_fusedsync.json; tail -3 gpurun_out/r02a_bench_n8_fusedsync.err
# 4. (only if green at 2 GPUs) local transposition pipelined with the exchange next to it
for n in 2 4; do DTFFTB_TRANSPOSE_OVERLAP=$n timeout 300 $TR --master-port 2956$n bench.py --gpus 8 > gpurun_out/r02a_bench_n8_pair$n.json 2> gpurun_out/r02a_bench_n8_pair$n.err; cut -c 1-330 gpurun_out/r02a_bench_n8_pair$n.json; tail -2 gpurun_out/r02a_bench_n8_pair$n.err; done
