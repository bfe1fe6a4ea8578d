#!/bin/bash
# One GPU: ncu --set full of the largest launch of k_fourier<direct> with the direct record stores and with the
# slot + push path forced (ECT_FFT_PUSH=2; on one rank the "consumer" buffer is local, so this shows the cost of the
# extra pass through L2, not the NVLink side).
cd $GRAFT_REPO_ROOT
P="python tools/profile_step.py --steps 1 --warmup 1"
for m in 0 2; do
  ECT_FFT_PUSH=$m ncu --set full --clock-control none --import-source on -k regex:'^k_fourier$' -s 12 -c 1 -f -o gpurun_out/r02b_ftdir_push$m $P > gpurun_out/r02b_ftdir_push$m.log 2>&1
  ncu -i gpurun_out/r02b_ftdir_push$m.ncu-rep --page raw --csv > gpurun_out/r02b_ftdir_push${m}_raw.csv 2>/dev/null
  python tools/ncu_summary.py raw gpurun_out/r02b_ftdir_push${m}_raw.csv > gpurun_out/r02b_ftdir_push${m}_summary.txt 2>&1
  rm -f gpurun_out/r02b_ftdir_push$m.ncu-rep
done
tail -n 3 gpurun_out/r02b_ftdir_push2.log
grep -h "=== \|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|lts__t_sector_hit_rate.pct\|registers_per_thread\|sm__warps_active" gpurun_out/r02b_ftdir_push0_summary.txt gpurun_out/r02b_ftdir_push2_summary.txt
