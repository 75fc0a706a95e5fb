python tools/ab_rollout.py 65536 > gpurun_out/r2_lean_epilogue.txt 2>&1
python tools/wrapper_cost.py >> gpurun_out/r2_lean_epilogue.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/r2_lean_epilogue.txt
