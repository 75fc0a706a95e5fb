python -m pytest tests -m gpu -x -q -k "robot" 2>&1 | tail -5 > gpurun_out/r2_pytest_robot.log
python tools/grid_ab.py 2 1048576 131072 > gpurun_out/r2_robot_ab2.txt 2>&1
python tools/step_pipe_all_ab.py >> gpurun_out/r2_robot_ab2.txt 2>&1
