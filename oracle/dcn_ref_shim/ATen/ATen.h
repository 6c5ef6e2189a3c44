// TEST INFRASTRUCTURE (oracle/): the few ATen names the reference's dcn kernel file uses
// (dcn/src/deform_conv_cuda_kernel.cu: at::Tensor::data<T>(), ::scalar_type(), AT_DISPATCH_FLOATING_TYPES_AND_HALF),
// restated over raw device pointers so that the UNMODIFIED reference file compiles with plain nvcc, without the
// torch-0.4-era ATen/THC headers it was written against (they no longer exist in torch 2.11).  See oracle/build_dcn_ref.py.
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
namespace at {
enum class ScalarType { Float };
struct Tensor {
  void* ptr = nullptr;
  Tensor() {}
  explicit Tensor(const void* p) : ptr(const_cast<void*>(p)) {}
  template <typename T> T* data() const { return static_cast<T*>(ptr); }
  ScalarType scalar_type() const { return ScalarType::Float; }
};
}  // namespace at
// the oracle runs the reference kernels in fp32 only
#define AT_DISPATCH_FLOATING_TYPES_AND_HALF(TYPE, NAME, ...) \
  do { (void)(TYPE); using scalar_t = float; __VA_ARGS__(); } while (0)
