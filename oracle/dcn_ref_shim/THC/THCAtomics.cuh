// TEST INFRASTRUCTURE (oracle/): the reference includes this header for atomicAdd overloads on half; the fp32
// instantiation the oracle builds needs only CUDA's own atomicAdd(float*, float).
#pragma once
#include <cuda_runtime.h>
