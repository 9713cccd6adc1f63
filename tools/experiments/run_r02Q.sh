#!/bin/bash
# r02Q batch (one gpurun call): lane -> record mapping of the ring kernels (build knob RTR_LANE_MAP), reduction cost per
# lane / sector / address, occupancy of the image kernels.  Experiment builds are made on the CPU box first (see README).
PKG=real-time-neural-rendering-of-lidar-point-clouds_b200
mkdir -p gpurun_out
RTR_B200_LIB=$PWD/$PKG/librtr_b200_lm1.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02Q_pytest_gpu_lane_map1.txt 2>&1
tail -3 gpurun_out/r02Q_pytest_gpu_lane_map1.txt
python tools/experiments/fused_ab.py --workload c3 --frames 400 --out gpurun_out/r02Q_exp_lane_map_c3.json base lm1:lib=_lm1 lm1r1:lib=_lm1r1 lm1r4:lib=_lm1r4 lm1r32:lib=_lm1r32 base2 2>&1 | cut -c1-400
python tools/experiments/red_coalescing.py --out gpurun_out/r02Q_exp_red_coalescing.json 2>&1 | cut -c1-300
python tools/experiments/fused_ab.py --workload c3 --frames 400 --out gpurun_out/r02Q_exp_image_occupancy_c3.json base rs6:lib=_rs6 up5:lib=_up5 up6:lib=_up6 up3:lib=_up3 2>&1 | cut -c1-300
python tools/experiments/fused_ab.py --workload c2 --frames 400 --out gpurun_out/r02Q_exp_lane_map_c2.json base lm1:lib=_lm1 lm1r32:lib=_lm1r32 up5:lib=_up5 up3:lib=_up3 rs6:lib=_rs6 2>&1 | cut -c1-300
