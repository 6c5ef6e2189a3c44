// TEST INFRASTRUCTURE (oracle/): C entry points over the reference's OWN deformable-convolution CUDA kernels.
//
// REF_KERNEL_FILE is /root/reference/dcn/src/deform_conv_cuda_kernel.cu, included UNMODIFIED from where it lies
// (nothing of it is copied into this repository); oracle/dcn_ref_shim/ supplies the three ATen names it uses.  The
// result, oracle/_ref/libdcn_ref.so (git-ignored, travels to the GPU box), lets tests/test_gpu_dcn.py pin the product's
// fused dcn kernel to the reference's sampling kernels:
//   deformable_im2col                 dcn/src/deform_conv_cuda_kernel.cu:244-276  (kernel :190-242)
//   modulated_deformable_im2col_cuda  :768-798                                    (kernel :570-632)
// followed by the same weight x columns product the reference's host code performs (deform_conv_cuda.cpp:220-234,
// 535-548), done by the test in fp64.
#include REF_KERNEL_FILE

extern "C" int dcn_ref_deformable_im2col(const float* im, const float* offset, int channels, int height, int width,
                                         int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int dil_h,
                                         int dil_w, int parallel_imgs, int deformable_group, float* col) {
  deformable_im2col(at::Tensor(im), at::Tensor(offset), channels, height, width, kh, kw, pad_h, pad_w, stride_h,
                    stride_w, dil_h, dil_w, parallel_imgs, deformable_group, at::Tensor(col));
  return (int)cudaDeviceSynchronize();
}

extern "C" int dcn_ref_modulated_deformable_im2col(const float* im, const float* offset, const float* mask,
                                                   int batch, int channels, int height, int width, int height_col,
                                                   int width_col, int kh, int kw, int pad_h, int pad_w, int stride_h,
                                                   int stride_w, int dil_h, int dil_w, int deformable_group,
                                                   float* col) {
  modulated_deformable_im2col_cuda(at::Tensor(im), at::Tensor(offset), at::Tensor(mask), batch, channels, height, width,
                                   height_col, width_col, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w,
                                   deformable_group, at::Tensor(col));
  return (int)cudaDeviceSynchronize();
}

// ---- backward: the reference's col2im (grad_input) and col2im_coord (grad_offset / grad_mask) kernels, fed with the
// column-gradient matrix W^T x grad_output that the test forms in fp64 (deform_conv_cuda.cpp:320-347, :611-640).
//   deformable_col2im                        :336-386  (kernel :278-334)
//   deformable_col2im_coord                  :437-464  (kernel :388-435)
//   modulated_deformable_col2im_cuda         :800-831  (kernel :634-691)
//   modulated_deformable_col2im_coord_cuda   :833-867  (kernel :693-766)
extern "C" int dcn_ref_deformable_col2im(const float* col, const float* offset, int channels, int height, int width,
                                         int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int dil_h,
                                         int dil_w, int parallel_imgs, int deformable_group, float* grad_im) {
  deformable_col2im(at::Tensor(col), at::Tensor(offset), channels, height, width, kh, kw, pad_h, pad_w, stride_h,
                    stride_w, dil_h, dil_w, parallel_imgs, deformable_group, at::Tensor(grad_im));
  return (int)cudaDeviceSynchronize();
}

extern "C" int dcn_ref_deformable_col2im_coord(const float* col, const float* im, const float* offset, int channels,
                                               int height, int width, int kh, int kw, int pad_h, int pad_w,
                                               int stride_h, int stride_w, int dil_h, int dil_w, int parallel_imgs,
                                               int deformable_group, float* grad_offset) {
  deformable_col2im_coord(at::Tensor(col), at::Tensor(im), at::Tensor(offset), channels, height, width, kh, kw, pad_h,
                          pad_w, stride_h, stride_w, dil_h, dil_w, parallel_imgs, deformable_group,
                          at::Tensor(grad_offset));
  return (int)cudaDeviceSynchronize();
}

extern "C" int dcn_ref_modulated_deformable_col2im(const float* col, const float* offset, const float* mask, int batch,
                                                   int channels, int height, int width, int height_col, int width_col,
                                                   int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                                                   int dil_h, int dil_w, int deformable_group, float* grad_im) {
  modulated_deformable_col2im_cuda(at::Tensor(col), at::Tensor(offset), at::Tensor(mask), batch, channels, height,
                                   width, height_col, width_col, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h,
                                   dil_w, deformable_group, at::Tensor(grad_im));
  return (int)cudaDeviceSynchronize();
}

extern "C" int dcn_ref_modulated_deformable_col2im_coord(const float* col, const float* im, const float* offset,
                                                         const float* mask, int batch, int channels, int height,
                                                         int width, int height_col, int width_col, int kh, int kw,
                                                         int pad_h, int pad_w, int stride_h, int stride_w, int dil_h,
                                                         int dil_w, int deformable_group, float* grad_offset,
                                                         float* grad_mask) {
  modulated_deformable_col2im_coord_cuda(at::Tensor(col), at::Tensor(im), at::Tensor(offset), at::Tensor(mask), batch,
                                         channels, height, width, height_col, width_col, kh, kw, pad_h, pad_w,
                                         stride_h, stride_w, dil_h, dil_w, deformable_group, at::Tensor(grad_offset),
                                         at::Tensor(grad_mask));
  return (int)cudaDeviceSynchronize();
}
