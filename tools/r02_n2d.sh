# Round 2, fourth 2-GPU call: sliced copy-engine exchange (parity, A/B)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02f_pytest_n2.log | tail -30
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_zz_shared_device_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02f_pytest_shared_device.log | tail -30
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port 29600 bench.py --gpus 2 --backend nvlink > gpurun_out/r02f_bench_n2_$name.json 2> gpurun_out/r02f_bench_n2_$name.err; python tools/show_bench.py gpurun_out/r02f_bench_n2_$name.json 2>&1 | head -4; tail -2 gpurun_out/r02f_bench_n2_$name.err; }
run dma X=1
run dma_nograph DTFFTB_GRAPHS=0
run dma_nograph_sub4m DTFFTB_GRAPHS=0 DTFFTB_DMA_SUB_BYTES=4194304
run dma_nograph_nosub DTFFTB_GRAPHS=0 DTFFTB_DMA_SUB_BYTES=100000000000
run dma_nopair_nograph DTFFTB_GRAPHS=0 DTFFTB_PAIR_OVERLAP=0
