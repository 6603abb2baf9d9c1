TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -12
timeout 200 $TR --master-port 29541 bench.py --gpus 4 > gpurun_out/r01f_bench_n4.json 2> gpurun_out/r01f_bench_n4.err; cat gpurun_out/r01f_bench_n4.json; tail -3 gpurun_out/r01f_bench_n4.err
