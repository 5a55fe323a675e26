// shim: the emulation build (tests/cuemu) resolves <cuda_runtime.h> here
#pragma once
#include "../cuemu.h"
