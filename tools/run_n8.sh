TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29531 tools/configs_bench.py --overlap 1,4,8 > gpurun_out/r01d_configs_n8.jsonl 2> gpurun_out/r01d_configs_n8.err; cut -c 1-330 gpurun_out/r01d_configs_n8.jsonl; tail -5 gpurun_out/r01d_configs_n8.err
DTFFTB_OVERLAP_CTAS=296 timeout 200 $TR --master-port 29532 tools/configs_bench.py --configs c2fft,c3,c4 --backends nvlink --overlap 8 > gpurun_out/r01d_configs_n8_ctas296.jsonl 2> gpurun_out/r01d_configs_n8_ctas296.err; cut -c 1-330 gpurun_out/r01d_configs_n8_ctas296.jsonl
DTFFTB_OVERLAP_CTAS=74 timeout 200 $TR --master-port 29533 tools/configs_bench.py --configs c2fft,c3,c4 --backends nvlink --overlap 8 > gpurun_out/r01d_configs_n8_ctas74.jsonl 2> gpurun_out/r01d_configs_n8_ctas74.err; cut -c 1-330 gpurun_out/r01d_configs_n8_ctas74.jsonl
DTFFTB_TEST_BACKENDS=NVLINK_FUSED timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 $TR --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r01d_bench_n8.json 2> gpurun_out/r01d_bench_n8.err; cat gpurun_out/r01d_bench_n8.json; tail -5 gpurun_out/r01d_bench_n8.err
