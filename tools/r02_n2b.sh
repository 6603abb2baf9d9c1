# Round 2, second 2-GPU call: copy-engine forms of the exchange, tile A/B of the fused exchange, tool smoke runs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 ./tools/nvlink_probe --mb 256 --iters 10 > gpurun_out/r02c_nvlink_probe_n2.jsonl 2> gpurun_out/r02c_nvlink_probe_n2.err; tail -3 gpurun_out/r02c_nvlink_probe_n2.err; python tools/summarize_probe.py gpurun_out/r02c_nvlink_probe_n2.jsonl | grep -v "^bulk\|^tma2tma\|^pull"
for t in 1,4,16 2,2,8 2,1,8; do DTFFTB_TILE=$t timeout 300 $TR --master-port 2957${t:0:1} bench.py --gpus 2 --backend nvlink > gpurun_out/r02c_bench_n2_tile_${t//,/_}.json 2> gpurun_out/r02c_bench_n2_tile_${t//,/_}.err; python tools/show_bench.py gpurun_out/r02c_bench_n2_tile_${t//,/_}.json | head -3; done
timeout 400 $TR --master-port 29580 tools/configs_profile.py --scale 0.25 --iters 5 > gpurun_out/r02c_configs_profile_quarter_n2.jsonl 2> gpurun_out/r02c_configs_profile_quarter_n2.err; cut -c 1-600 gpurun_out/r02c_configs_profile_quarter_n2.jsonl; tail -5 gpurun_out/r02c_configs_profile_quarter_n2.err
timeout 300 $TR --master-port 29581 tools/e2e_probe.py > gpurun_out/r02c_e2e_probe_n2.jsonl 2> gpurun_out/r02c_e2e_probe_n2.err; cat gpurun_out/r02c_e2e_probe_n2.jsonl; tail -3 gpurun_out/r02c_e2e_probe_n2.err
