// b2z_dec.h -- decoder-side launchers (internal to libb200z.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
