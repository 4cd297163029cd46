python tests/gpu_scripts/dbg15.py 2>&1 | tail -32
