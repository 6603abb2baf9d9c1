TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/configs_bench.py --scale 0.25 --iters 3 > gpurun_out/r01d_configs_small_n2.jsonl 2> gpurun_out/r01d_configs_small_n2.err; tail -n 20 gpurun_out/r01d_configs_small_n2.jsonl | cut -c 1-400; tail -5 gpurun_out/r01d_configs_small_n2.err
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -30
timeout 600 $TR --master-port 29522 tools/configs_bench.py --overlap 1,2,4,8 > gpurun_out/r01d_configs_n2.jsonl 2> gpurun_out/r01d_configs_n2.err; cut -c 1-420 gpurun_out/r01d_configs_n2.jsonl; tail -5 gpurun_out/r01d_configs_n2.err
