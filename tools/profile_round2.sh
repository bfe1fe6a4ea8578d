# ncu evidence of round 2 (run under gpurun, one GPU): launch list of one inverse+direct step and full-set captures
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
P="python tools/profile_step.py --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $P > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fourier_cz -s 8 -c 1 -f -o gpurun_out/${TAG}_czinv $P > gpurun_out/${TAG}_czinv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fourier_cz -s 12 -c 1 -f -o gpurun_out/${TAG}_czdir $P > gpurun_out/${TAG}_czdir.log 2>&1
for f in czinv czdir; do ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null; ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null; done
ls -la gpurun_out/ | tail -12
