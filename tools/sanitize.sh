# compute-sanitizer over every kernel family (run under gpurun, one GPU; two GPUs for the last step when available).
# Logs: gpurun_out/sanitize_*.log -- copy the summaries to profiles/.
set -x
cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/sanitize_${tool}.log 2>&1
  echo "exit code $tool: $?" >> gpurun_out/sanitize_${tool}.log
  tail -4 gpurun_out/sanitize_${tool}.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  ECT_SELFCHECK_T=79 timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/selfcheck_run.py > gpurun_out/sanitize_memcheck_2gpu.log 2>&1
  echo "exit code: $?" >> gpurun_out/sanitize_memcheck_2gpu.log
  tail -4 gpurun_out/sanitize_memcheck_2gpu.log
fi
