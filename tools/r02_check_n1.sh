mkdir -p gpurun_out
timeout 200 python tools/kbench.py --quick > gpurun_out/r02l_kbench_quick.txt 2>&1; grep -E "forward|backward\"|unpack8" gpurun_out/r02l_kbench_quick.txt | cut -c 1-110
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/r02l_pytest_n1.log | tail -4
