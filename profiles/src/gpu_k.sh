#!/bin/bash
python profiles/src/train_prof.py 256 2>&1 | grep -E "ms_per_step|Self CUDA time total"
python profiles/src/train_prof.py 256 bench 2>&1 | grep -E "ms_per_step|Self CUDA time total|nchwToNhwc|nhwcToNchw|convolution_backward |cudnn_convolution_transpose " | cut -c 1-60,130-200
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/k_ncu1.log 2>&1
