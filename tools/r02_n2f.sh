# Round 2, sixth 2-GPU call: line-aligned tile grid (bshift) -- A/B on the config-5-like exchange, then the whole GPU suite
mkdir -p gpurun_out
for al in 0 1; do for z in 250,262 256,256; do DTFFTB_ALIGN_TILES=$al timeout 200 python tools/exchange_kbench.py --uneven $z --tiles "2,2,16" 2>> gpurun_out/r02k_uneven.err | sed "s/^{/{\"align_tiles\": $al, /" | tee -a gpurun_out/r02k_uneven_kbench_n2.jsonl | cut -c 1-330; done; done
tail -2 gpurun_out/r02k_uneven.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/r02k_pytest_2gpu_box.log | tail -8
timeout 200 python tools/kbench.py --quick > gpurun_out/r02k_kbench_quick.txt 2>&1; tail -15 gpurun_out/r02k_kbench_quick.txt | cut -c 1-200
