mkdir -p gpurun_out
for z in 250,250,262,262 256,256,256,256; do timeout 100 python tools/exchange_kbench.py --uneven $z --local --tiles "2,2,16" --iters 5 2>&1 | grep "^{" | head -2 | tee -a gpurun_out/r02n_c5_local_pattern.jsonl | cut -c 1-260; done
