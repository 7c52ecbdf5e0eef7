# A/B of the pose backward kernel's chain walk (lane per element vs lane per joint, -DSMPLK_POSE_BWD_ELEMWALK=1 build as libsmplk_elemwalk.so)
set -x
python tools/fit_ab.py 1024 pdl > gpurun_out/r02_pose_bwd_walk_ab.txt 2>&1
SMPLK_LIB=$PWD/3d-human-body-reconstruction_b200/libsmplk_elemwalk.so python tools/fit_ab.py 1024 pdl >> gpurun_out/r02_pose_bwd_walk_ab.txt 2>&1
python tools/fit_profile.py 1024 >> gpurun_out/r02_pose_bwd_walk_ab.txt 2>&1
SMPLK_LIB=$PWD/3d-human-body-reconstruction_b200/libsmplk_elemwalk.so python tools/fit_profile.py 1024 >> gpurun_out/r02_pose_bwd_walk_ab.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/r02_pose_bwd_walk_ab.txt; tail -3 gpurun_out/pytest_gpu.log
