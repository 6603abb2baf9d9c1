# Round 2, second 4-GPU call: race hunt on the SHIPPED pair pipeline at 4 GPUs (poisoned intermediates) + the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29920 tools/stress_pair.py --iters 40 2> gpurun_out/r02j_stress_n4.err | tee gpurun_out/r02j_stress_n4.json | cut -c 1-600; tail -2 gpurun_out/r02j_stress_n4.err | cut -c 1-200
DTFFTB_DMA_SUB_BYTES=262144 timeout 300 $TR --master-port 29921 tools/stress_pair.py --iters 40 --size 192 2> gpurun_out/r02j_stress_n4_small.err | tee gpurun_out/r02j_stress_n4_small.json | cut -c 1-600
timeout 400 $TR --master-port 29922 bench.py --gpus 4 > gpurun_out/r02j_bench_n4.json 2> gpurun_out/r02j_bench_n4.err; python tools/show_bench.py gpurun_out/r02j_bench_n4.json; tail -2 gpurun_out/r02j_bench_n4.err | cut -c 1-200
