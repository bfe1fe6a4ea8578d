# Round-2 closing measurements on one GPU (run under gpurun): tests, sanitizer, bench lines, launch list, box probe.
set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
bash tools/sanitize.sh > gpurun_out/sanitize_run.log 2>&1
python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench_n1.err | grep "^{" > gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/r02_bench_ref_n1.json
python bench.py --config TCo399_O400_L137 --steps 20 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r02_bench_tco399sp.json
python bench.py --config T159_O160_L137 --steps 20 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r02_bench_t159.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/r02_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_tco399sp.csv python tools/profile_step.py --config TCo399_O400_L137 --precision sp --steps 1 --warmup 1 > /dev/null 2>&1
( echo "== compilers / libraries on the GPU box"; for c in gfortran nvfortran flang ifort ifx mpirun mpif90 cmake; do printf "%s: " $c; command -v $c || echo "not found"; done; echo "== fftw / blas / lapack / fiat / ecbuild"; ldconfig -p | grep -iE "fftw|openblas|libblas|lapack|fiat" || echo "none in ldconfig"; find / -xdev \( -iname "*fftw3*" -o -iname "*fiat*" -o -iname "ecbuild*" \) 2>/dev/null | grep -v -E "site-packages|/proc/" | head -5; echo "== cpu"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA"; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv ) > gpurun_out/r02_gpu_box_probe.txt 2>&1
ls -la gpurun_out | tail -15
