# Round 2, first 2-GPU call:  gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_n2.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L
# 0. what one GPU pushes through NVLink from a kernel, by access form (single process, 2 devices)
timeout 300 ./tools/nvlink_probe --mb 256 --iters 10 > gpurun_out/r02b_nvlink_probe_n2.jsonl 2> gpurun_out/r02b_nvlink_probe_n2.err; tail -3 gpurun_out/r02b_nvlink_probe_n2.err; python tools/summarize_probe.py gpurun_out/r02b_nvlink_probe_n2.jsonl
# 1. every backend through the plan API (incl. any-pointer publication, NCCL stand-in, brick reshapes +/- shortcuts)
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02b_pytest_n2.log | tail -30
# 1b. the same worker with the checks of the opt-in features (pair overlap, EXHAUSTIVE reshape backend)
DTFFTB_TEST_EXPERIMENTAL=1 DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 400 $TR --master-port 29520 tests/_gpu_worker.py 2>&1 | tee gpurun_out/r02b_worker_experimental_n2.log | tail -8
# 1c. several ranks time-slicing ONE device (what the driver's 1-GPU box will run)
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_zz_shared_device_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02b_pytest_shared_device.log | tail -30
# 2. the bench line: all three backends, parity block first
timeout 400 $TR --master-port 29540 bench.py --gpus 2 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; cut -c 1-1500 gpurun_out/r02b_bench_n2.json; tail -3 gpurun_out/r02b_bench_n2.err
# 3. barriers folded into the fused kernel (opt-in, first run ever: keep it under its own timeout): parity, then A/B
DTFFTB_FUSED_SYNC=1 DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 300 $TR --master-port 29541 tests/_gpu_worker.py 2>&1 | tee gpurun_out/r02b_worker_fusedsync_n2.log | tail -5
DTFFTB_FUSED_SYNC=1 timeout 300 $TR --master-port 29542 bench.py --gpus 2 --backend nvlink > gpurun_out/r02b_bench_n2_fusedsync.json 2> gpurun_out/r02b_bench_n2_fusedsync.err; cut -c 1-700 gpurun_out/r02b_bench_n2_fusedsync.json; tail -3 gpurun_out/r02b_bench_n2_fusedsync.err
# 4. local transposition pipelined with the exchange next to it (opt-in, first run ever): A/B of the cycle
for n in 2 4 8; do DTFFTB_TRANSPOSE_OVERLAP=$n timeout 300 $TR --master-port 2955$n bench.py --gpus 2 --backend nvlink > gpurun_out/r02b_bench_n2_pair$n.json 2> gpurun_out/r02b_bench_n2_pair$n.err; cut -c 1-700 gpurun_out/r02b_bench_n2_pair$n.json; tail -2 gpurun_out/r02b_bench_n2_pair$n.err; done
# 5. config 5 (bricks, NCCL backends): shortcuts on / off
for s in 1 0; do DTFFTB_RESHAPE_SHORTCUTS=$s timeout 300 $TR --master-port 2952$s tools/configs_bench.py --configs c5 --backends nccl,nccl_pipe --overlap 1 > gpurun_out/r02b_c5_shortcuts${s}_n2.jsonl 2> gpurun_out/r02b_c5_shortcuts${s}_n2.err; cut -c 1-330 gpurun_out/r02b_c5_shortcuts${s}_n2.jsonl; tail -3 gpurun_out/r02b_c5_shortcuts${s}_n2.err; done
# 6. CUDA-graph replay of the NCCL backends on the launch-bound half-size configs
for g in 0 1; do DTFFTB_GRAPHS_NCCL=$g timeout 300 $TR --master-port 2953$g tools/configs_bench.py --configs c2fft,c4 --backends nccl,nccl_pipe --overlap 1 --scale 0.5 > gpurun_out/r02b_half_ncclgraphs${g}_n2.jsonl 2> gpurun_out/r02b_half_ncclgraphs${g}_n2.err; cut -c 1-330 gpurun_out/r02b_half_ncclgraphs${g}_n2.jsonl; tail -3 gpurun_out/r02b_half_ncclgraphs${g}_n2.err; done
