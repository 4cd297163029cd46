mkdir -p gpurun_out
python tests/gpu_scripts/dbg6.py > gpurun_out/r2_dbg6.log 2>&1; cat gpurun_out/r2_dbg6.log | tail -15
