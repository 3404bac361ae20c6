set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/dense_ab.py 26 tma=0 tma=1 > gpurun_out/ab26.log 2>&1; echo "rc26=$?"
cat gpurun_out/ab26.log
timeout 600 python tools/dense_ab.py 30 tma=0 tma=1 tma=1,prefetch_ahead=1 tma=1,prefetch_ahead=2 > gpurun_out/ab30.log 2>&1; echo "rc30=$?"
cat gpurun_out/ab30.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests1.log 2>&1; echo "rctests=$?"
tail -15 gpurun_out/gputests1.log
