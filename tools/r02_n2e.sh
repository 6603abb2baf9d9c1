# Round 2, fifth 2-GPU call: race hunt on the copy-engine pair pipeline (poisoned intermediates), sanitizers
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for st in 1 2 7; do DTFFTB_DMA_STREAMS=$st timeout 300 $TR --master-port 2990$st tools/stress_pair.py --iters 40 2> gpurun_out/r02i_stress_n2_streams$st.err | tee gpurun_out/r02i_stress_n2_streams$st.json | cut -c 1-600; tail -2 gpurun_out/r02i_stress_n2_streams$st.err | cut -c 1-200; done
DTFFTB_DMA_SUB_BYTES=1048576 timeout 300 $TR --master-port 29909 tools/stress_pair.py --iters 40 --n 256 2> gpurun_out/r02i_stress_n2_small.err | tee gpurun_out/r02i_stress_n2_small.json | cut -c 1-600
# 4 ranks time-slicing one GPU: most hostile timing
CUDA_VISIBLE_DEVICES=0 DTFFTB_ALLOW_SHARED_DEVICE=1 DTFFTB_FUSED_MODE=dma DTFFTB_DMA_SUB_BYTES=65536 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29910 tools/stress_pair.py --iters 20 --n 128 2> gpurun_out/r02i_stress_shared4.err | tee gpurun_out/r02i_stress_shared4.json | cut -c 1-600
# sanitizers over a 2-rank fused run in both forms
timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 $TR --master-port 29911 tools/sanitize_fused.py > gpurun_out/r02i_memcheck_fused_n2.txt 2>&1; echo "memcheck exit $?"; grep -c "sanitize_fused OK" gpurun_out/r02i_memcheck_fused_n2.txt; grep "ERROR SUMMARY" gpurun_out/r02i_memcheck_fused_n2.txt | sort | uniq -c
timeout 600 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 9 $TR --master-port 29912 tools/sanitize_fused.py > gpurun_out/r02i_racecheck_fused_n2.txt 2>&1; echo "racecheck exit $?"; grep -c "sanitize_fused OK" gpurun_out/r02i_racecheck_fused_n2.txt; grep "RACECHECK SUMMARY" gpurun_out/r02i_racecheck_fused_n2.txt | sort | uniq -c
