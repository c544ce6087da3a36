#!/bin/bash
# ncu --set full of the three streaming scale-space kernels (final versions), octave-0 launches of a 32-frame chunk
bash scripts/gpu_ncu_kernel.sh r02z_fed4 "k_fed4" 0 2 32
bash scripts/gpu_ncu_kernel.sh r02z_blur4 "k_blur4" 0 2 32
bash scripts/gpu_ncu_kernel.sh r02z_deriv4 "k_deriv4" 1 2 32
