// Deformable convolution v1/v2 forward - see dcn_impl below.
#pragma once
#include "common.cuh"
namespace cfen {
}  // namespace cfen
