// Shared device helpers for libcfen_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcfen_b200 targets sm_100a (B200) only"
#endif

namespace cfen {

// ---------------------------------------------------------------- small math / cvt
__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float(f32_to_tf32(x)); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two floats -> packed fp16x2 (round-to-nearest-even, saturating to +-65504 instead of inf); lo in bits [0,16)
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- NHWC maps stored as fp32 or fp16
// Four consecutive channels of a map: `i4` counts 4-element groups from the map base, whatever the storage width
// (fp16 maps are addressed through the same float* the plan hands out).
__device__ __forceinline__ float4 ld_map4(const float* base, size_t i4, bool h16) {
  if (h16) {
    const uint2 u = reinterpret_cast<const uint2*>(base)[i4];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return reinterpret_cast<const float4*>(base)[i4];
}
__device__ __forceinline__ void st_map4(float* base, size_t i4, float4 v, bool h16) {
  if (h16) reinterpret_cast<uint2*>(base)[i4] = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
  else reinterpret_cast<float4*>(base)[i4] = v;
}

// ---------------------------------------------------------------- cp.async (LDGSTS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy; when !pred the destination is zero-filled (src-size 0) and src is not read.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(sz));
}
// 8-byte async copy (.ca: 16 bytes is the only size .cg takes); zero-fill when !pred.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, bool pred) {
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------- warp-level TF32 MMA
// D(16x8) += A(16x8,row) * B(8x8,col), fp32 accumulate.  Fragment ownership (g = lane>>2,
// t = lane&3):  a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4);  b0=(k=t,n=g) b1=(k=t+4,n=g);
// d0=(g,2t) d1=(g,2t+1) d2=(g+8,2t) d3=(g+8,2t+1).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit global access (read-once / write-once data)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------- bilinear x2 applied twice
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) applied twice (v3.py:1323)
// equals, on an edge-replicated signal, a 3-tap stencil over source samples k-1, k, k+1 for output
// 4k+r.  w[0..2] are the weights of k-1, k, k+1.
__device__ __forceinline__ void up4_taps(int r, float (&w)[3]) {
  switch (r) {
    case 0: w[0] = 0.375f;  w[1] = 0.625f; w[2] = 0.0f;    break;
    case 1: w[0] = 0.1875f; w[1] = 0.75f;  w[2] = 0.0625f; break;
    case 2: w[0] = 0.0625f; w[1] = 0.75f;  w[2] = 0.1875f; break;
    default: w[0] = 0.0f;   w[1] = 0.625f; w[2] = 0.375f;  break;
  }
}

}  // namespace cfen
