# round-2 evidence run (under gpurun): gpu test suite, then bench line / reference arm / configs records / ncu passes
# TAG=r2i bash scripts/run_r2_evidence.sh
T=${TAG:-r2i}
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
TAG=$T bash scripts/run_r2h.sh
