# Round 2, last 8-GPU call: final code -- bench line (all backends, parity with poisoned intermediates) and configs C3 / C5 per stage
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29700 bench.py --gpus 8 > gpurun_out/r02m_bench_n8.json 2> gpurun_out/r02m_bench_n8.err; python tools/show_bench.py gpurun_out/r02m_bench_n8.json; tail -2 gpurun_out/r02m_bench_n8.err | cut -c 1-200
timeout 300 $TR --master-port 29702 tools/configs_profile.py --configs c3,c5 --backends nvlink --iters 5 > gpurun_out/r02m_configs_profile_n8.jsonl 2> gpurun_out/r02m_configs_profile_n8.err; python tools/show_profile.py gpurun_out/r02m_configs_profile_n8.jsonl; tail -2 gpurun_out/r02m_configs_profile_n8.err | cut -c 1-200
