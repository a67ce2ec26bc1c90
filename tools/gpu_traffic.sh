#!/bin/bash
# DRAM traffic and duration of the band kernel at the bench batch (unsplit), quick ncu metrics pass
mkdir -p gpurun_out
TB_LARGE_SPLIT=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_band2 -s 4 -c 2 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 3 --warmup 3 --headline-only > /dev/null 2>&1
grep -E "k_band2" gpurun_out/traffic.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -12
TB_LARGE_SPLIT=0 python tools/quick_time.py 2>&1 | grep -E "bar-942 x(1024|8192):"
