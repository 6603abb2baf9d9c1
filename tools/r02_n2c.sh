# Round 2, third 2-GPU call: the copy-engine form of the exchange (parity, then A/B against the direct-store kernel)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02d_pytest_n2.log | tail -30
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_zz_shared_device_gpu.py -m gpu -x -q 2>&1 | tee gpurun_out/r02d_pytest_shared_device.log | tail -30
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port 29600 bench.py --gpus 2 --backend nvlink > gpurun_out/r02d_bench_n2_$name.json 2> gpurun_out/r02d_bench_n2_$name.err; python tools/show_bench.py gpurun_out/r02d_bench_n2_$name.json 2>&1 | head -4; tail -2 gpurun_out/r02d_bench_n2_$name.err; }
run store DTFFTB_FUSED_MODE=store
run dma DTFFTB_FUSED_MODE=dma
run dma_nopair DTFFTB_FUSED_MODE=dma DTFFTB_PAIR_OVERLAP=0
run dma_nograph DTFFTB_FUSED_MODE=dma DTFFTB_GRAPHS=0
run auto X=1
# the exchange kernel alone, one process over both devices; then its NVLink / DRAM counters
timeout 300 python tools/exchange_kbench.py --devices 2 --check > gpurun_out/r02d_exchange_kbench_n2.jsonl 2> gpurun_out/r02d_exchange_kbench_n2.err; cat gpurun_out/r02d_exchange_kbench_n2.jsonl; tail -3 gpurun_out/r02d_exchange_kbench_n2.err
timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:transpose_tiles -c 4 --csv --log-file gpurun_out/r02d_ncu_exchange_nvlink.csv python tools/exchange_kbench.py --devices 2 --iters 1 --warmup 0 --tiles 1,2,8 > gpurun_out/r02d_ncu_exchange.log 2>&1; tail -12 gpurun_out/r02d_ncu_exchange_nvlink.csv | cut -c 1-300
