# Round 2, first 2-GPU call:  gpurun --gpus 2 --timeout 900 -- 'bash tools/r02_n2.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
# 1. every backend through the plan API, brick reshapes with and without the pack-free / unpack-free shortcuts
timeout 700 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -30
# 1b. the same worker with the checks of the features written after the round-1 GPU budget was spent
DTFFTB_TEST_EXPERIMENTAL=1 DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 400 $TR --master-port 29520 tests/_gpu_worker.py 2>&1 | tail -8
# 2. config 5 (bricks, NCCL backends): shortcuts on / off
for s in 1 0; do DTFFTB_RESHAPE_SHORTCUTS=$s timeout 300 $TR --master-port 2952$s tools/configs_bench.py --configs c5 --backends nccl,nccl_pipe --overlap 1 > gpurun_out/r02a_c5_shortcuts${s}_n2.jsonl 2> gpurun_out/r02a_c5_shortcuts${s}_n2.err; cut -c 1-330 gpurun_out/r02a_c5_shortcuts${s}_n2.jsonl; tail -3 gpurun_out/r02a_c5_shortcuts${s}_n2.err; done
# 3. CUDA-graph replay of the NCCL backends on the launch-bound half-size configs
for g in 0 1; do DTFFTB_GRAPHS_NCCL=$g timeout 300 $TR --master-port 2953$g tools/configs_bench.py --configs c2fft,c4 --backends nccl,nccl_pipe --overlap 1 --scale 0.5 > gpurun_out/r02a_half_ncclgraphs${g}_n2.jsonl 2> gpurun_out/r02a_half_ncclgraphs${g}_n2.err; cut -c 1-330 gpurun_out/r02a_half_ncclgraphs${g}_n2.jsonl; tail -3 gpurun_out/r02a_half_ncclgraphs${g}_n2.err; done
# 4. the bench line
timeout 300 $TR --master-port 29540 bench.py --gpus 2 > gpurun_out/r02a_bench_n2.json 2> gpurun_out/r02a_bench_n2.err; cut -c 1-600 gpurun_out/r02a_bench_n2.json; tail -3 gpurun_out/r02a_bench_n2.err
# 5. barriers folded into the fused kernel (opt-in, first run ever: keep it under its own timeout): parity, then A/B
DTFFTB_FUSED_SYNC=1 DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 300 $TR --master-port 29541 tests/_gpu_worker.py 2>&1 | tail -5
DTFFTB_FUSED_SYNC=1 timeout 300 $TR --master-port 29542 bench.py --gpus 2 > gpurun_out/r02a_bench_n2_fusedsync.json 2> gpurun_out/r02a_bench_n2_fusedsync.err; cut -c 1-600 gpurun_out/r02a_bench_n2_fusedsync.json; tail -3 gpurun_out/r02a_bench_n2_fusedsync.err
# 6. local transposition pipelined with the exchange next to it (opt-in, first run ever): A/B of the cycle
for n in 2 4 8; do DTFFTB_TRANSPOSE_OVERLAP=$n timeout 300 $TR --master-port 2955$n bench.py --gpus 2 --backend nvlink > gpurun_out/r02a_bench_n2_pair$n.json 2> gpurun_out/r02a_bench_n2_pair$n.err; cut -c 1-330 gpurun_out/r02a_bench_n2_pair$n.json; tail -2 gpurun_out/r02a_bench_n2_pair$n.err; done
