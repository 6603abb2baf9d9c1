# Round 2, second 8-GPU call (every minute costs 8 GPU-minutes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port 29701 bench.py --gpus 8 --backend nvlink > gpurun_out/r02h_bench_n8_$name.json 2> gpurun_out/r02h_bench_n8_$name.err; python tools/show_bench.py gpurun_out/r02h_bench_n8_$name.json 2>&1 | head -4; tail -2 gpurun_out/r02h_bench_n8_$name.err; }
# does a copy stream per peer hide the fixed cost of the copies?
run pairdma_streams7 DTFFTB_DMA_MAX_GROUP=8 DTFFTB_DMA_STREAMS=7
run pairdma_streams4 DTFFTB_DMA_MAX_GROUP=8 DTFFTB_DMA_STREAMS=4
# configs C3 / C4 / C5 at full size with the default rule (direct-store kernel for lone transpositions at these group sizes > 4, pair pipelines where groups are <= 4)
timeout 600 $TR --master-port 29702 tools/configs_profile.py --configs c3,c4,c5 --backends nvlink > gpurun_out/r02h_configs_profile_n8.jsonl 2> gpurun_out/r02h_configs_profile_n8.err; python tools/show_profile.py gpurun_out/r02h_configs_profile_n8.jsonl; tail -3 gpurun_out/r02h_configs_profile_n8.err
