TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -30
DTFFTB_LOG=1 timeout 300 $TR --master-port 29521 tools/configs_bench.py --configs c4,c2fft --backends nvlink --overlap 0,1 > gpurun_out/r01f_configs_auto_n2.jsonl 2> gpurun_out/r01f_configs_auto_n2.err; cut -c 1-330 gpurun_out/r01f_configs_auto_n2.jsonl; tail -3 gpurun_out/r01f_configs_auto_n2.err
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --backend nvlink > gpurun_out/r01f_bench_n2.json 2> gpurun_out/r01f_bench_n2.err; cut -c 1-600 gpurun_out/r01f_bench_n2.json; tail -3 gpurun_out/r01f_bench_n2.err
