set -x
cd $GRAFT_REPO_ROOT
P="python tools/profile_step.py --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv $P > gpurun_out/r01b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_leinv|k_ledir' -s 2 -c 2 -f -o gpurun_out/r01b_leg $P > gpurun_out/r01b_leg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fourier -s 24 -c 1 -f -o gpurun_out/r01b_ftinv $P > gpurun_out/r01b_ftinv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fourier -s 36 -c 1 -f -o gpurun_out/r01b_ftdir $P > gpurun_out/r01b_ftdir.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_ledir -s 1 -c 1 --csv --log-file gpurun_out/r01b_ledir_metrics.csv $P > gpurun_out/r01b_ledir_metrics.log 2>&1
for f in leg ftinv ftdir; do ncu -i gpurun_out/r01b_$f.ncu-rep --page raw --csv > gpurun_out/r01b_${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/ | tail -12
