TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 python tools/p2p_probe.py > gpurun_out/r01c_p2p_n8.json 2>gpurun_out/r01c_p2p_n8.err; cat gpurun_out/r01c_p2p_n8.json
timeout 300 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r01c_bench_n8.json 2> gpurun_out/r01c_bench_n8.err; cat gpurun_out/r01c_bench_n8.json; tail -5 gpurun_out/r01c_bench_n8.err
DTFFTB_NO_SHUFFLE=1 timeout 200 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --backend nvlink > gpurun_out/r01c_bench_n8_noshuffle.json 2> gpurun_out/r01c_bench_n8_noshuffle.err; cat gpurun_out/r01c_bench_n8_noshuffle.json
DTFFTB_GRID_MULT=2 timeout 200 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --backend nvlink > gpurun_out/r01c_bench_n8_gm2.json 2> gpurun_out/r01c_bench_n8_gm2.err; cat gpurun_out/r01c_bench_n8_gm2.json
nvidia-smi topo -m > gpurun_out/r01c_topo.txt 2>&1
