TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 300 $TR --master-port 29521 tools/configs_bench.py --scale 0.5 --overlap 1,4,8 > gpurun_out/r01e_configs_half_n2.jsonl 2> gpurun_out/r01e_configs_half_n2.err; cut -c 1-300 gpurun_out/r01e_configs_half_n2.jsonl; tail -5 gpurun_out/r01e_configs_half_n2.err
DTFFTB_GRAPHS=0 timeout 300 $TR --master-port 29523 tools/configs_bench.py --scale 0.5 --overlap 1,4,8 --backends nvlink > gpurun_out/r01e_configs_half_n2_nograph.jsonl 2> gpurun_out/r01e_configs_half_n2_nograph.err; cut -c 1-300 gpurun_out/r01e_configs_half_n2_nograph.jsonl
timeout 300 $TR --master-port 29522 bench.py --gpus 2 > gpurun_out/r01e_bench_n2.json 2> gpurun_out/r01e_bench_n2.err; cut -c 1-1200 gpurun_out/r01e_bench_n2.json; tail -5 gpurun_out/r01e_bench_n2.err
timeout 300 python bench.py > gpurun_out/r01e_bench_n1.json 2> gpurun_out/r01e_bench_n1.err; cat gpurun_out/r01e_bench_n1.json; tail -5 gpurun_out/r01e_bench_n1.err
