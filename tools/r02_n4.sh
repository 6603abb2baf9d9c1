# Round 2, 4-GPU call: where does the copy-engine pair pipeline stop paying?  (every minute costs 4 GPU-minutes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port 29800 bench.py --gpus 4 --backend nvlink > gpurun_out/r02g_bench_n4_$name.json 2> gpurun_out/r02g_bench_n4_$name.err; python tools/show_bench.py gpurun_out/r02g_bench_n4_$name.json 2>&1 | head -4; tail -2 gpurun_out/r02g_bench_n4_$name.err; }
run store DTFFTB_FUSED_MODE=store
run pairdma DTFFTB_DMA_MAX_GROUP=4
run pairdma_whole DTFFTB_DMA_MAX_GROUP=4 DTFFTB_DMA_SUB_BYTES=100000000000
