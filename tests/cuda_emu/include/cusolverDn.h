// Stand-in for the CUDA runtime header when the library's sources are compiled for the host by
// tests/cuda_emu/build_emu.py (test infrastructure only).
#pragma once
#include "../cuda_emu.h"
