// Stand-in for <cuda_runtime.h> when the kernels of xgboost_b200/csrc are compiled for the CPU emulator
// (tests/emu/cuda_emu.h).  Only what common.cuh and the kernel sources mention; host-side runtime calls are stubs that
// are never reached because the host launchers are compiled out (B200_KERNELS_ONLY).
#pragma once
#include "../cuda_emu.h"
