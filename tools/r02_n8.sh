# Round 2, 8-GPU call:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/r02_n8.sh'   (every minute costs 8 GPU-minutes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L | head -2
# 1. the bench line: every backend, parity block first (bit-exact multi-rank parity in the record)
timeout 400 $TR --master-port 29700 bench.py --gpus 8 > gpurun_out/r02e_bench_n8.json 2> gpurun_out/r02e_bench_n8.err; python tools/show_bench.py gpurun_out/r02e_bench_n8.json; tail -3 gpurun_out/r02e_bench_n8.err
# 2. A/B of the fused exchange forms
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port 29701 bench.py --gpus 8 --backend nvlink > gpurun_out/r02e_bench_n8_$name.json 2> gpurun_out/r02e_bench_n8_$name.err; python tools/show_bench.py gpurun_out/r02e_bench_n8_$name.json 2>&1 | head -4; tail -2 gpurun_out/r02e_bench_n8_$name.err; }
run store DTFFTB_FUSED_MODE=store
run dma_nopair DTFFTB_FUSED_MODE=dma DTFFTB_PAIR_OVERLAP=0
run dma_nosub DTFFTB_FUSED_MODE=dma DTFFTB_DMA_SUB_BYTES=100000000000
run dma_sub8m DTFFTB_FUSED_MODE=dma DTFFTB_DMA_SUB_BYTES=8388608
# 3. configs C3 / C4 / C5 at full size, per stage against the roofline
timeout 600 $TR --master-port 29702 tools/configs_profile.py --configs c3,c4,c5 --backends nvlink,nccl > gpurun_out/r02e_configs_profile_n8.jsonl 2> gpurun_out/r02e_configs_profile_n8.err; python tools/show_profile.py gpurun_out/r02e_configs_profile_n8.jsonl; tail -3 gpurun_out/r02e_configs_profile_n8.err
# 4. the multi-rank suite (every backend through the plan API; 8 ranks)
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k all_backends 2>&1 | tee gpurun_out/r02e_pytest_n8.log | tail -8
# 5. host link ceiling at 8 GPUs (what bounds e2e), process-grid search on real GPUs
timeout 200 $TR --master-port 29703 tools/e2e_probe.py > gpurun_out/r02e_e2e_probe_n8.jsonl 2> gpurun_out/r02e_e2e_probe_n8.err; cut -c 1-260 gpurun_out/r02e_e2e_probe_n8.jsonl
DTFFTB_LOG=1 timeout 300 $TR --master-port 29704 tools/grid_search_probe.py > gpurun_out/r02e_grid_search_n8.txt 2>&1; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02e_grid_search_n8.txt | tail -12
