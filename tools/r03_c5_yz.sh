# PREPARED, NOT RUN (round 2 ended with its GPU budget spent): isolate why config 5's Y->Z exchange stays at 0.24 of
# 900 GB/s at 8 GPUs.  One process over 4 devices, the exact geometry of one Y<->Z group.
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/r03_c5_yz.sh'
mkdir -p gpurun_out
# 1. the kernel alone on 4 devices, uneven and even split, default tile and the wider-output tiles (32x128: 1 KB chunks at 8 B)
for z in 250,250,262,262 256,256,256,256; do
  timeout 200 python tools/exchange_kbench.py --uneven $z --tiles "2,2,16;1,4,16;2,1,8;1,2,8" | tee -a gpurun_out/r03_c5_yz_kbench_n4.jsonl | cut -c 1-300
done
# 2. NVLink / DRAM counters of the slow case
timeout 400 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum \
  --clock-control none -k regex:transpose_tiles -c 4 --csv --log-file gpurun_out/r03_c5_yz_ncu.csv \
  python tools/exchange_kbench.py --uneven 250,250,262,262 --tiles "2,2,16" --iters 1 --warmup 0 > gpurun_out/r03_c5_yz_ncu.log 2>&1
tail -8 gpurun_out/r03_c5_yz_ncu.csv | cut -c 1-260
