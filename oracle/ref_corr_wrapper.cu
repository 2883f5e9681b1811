// Thin non-THC host wrapper around the REFERENCE correlation CUDA kernels -- TEST INFRASTRUCTURE ONLY.
//
// The reference's own host wrapper (correlation_package/src/corr_cuda.c:7-82) needs THC and
// torch.utils.ffi, which no longer exist; its kernels (src/corr_cuda_kernel.cu) compile unchanged
// for sm_100a.  oracle/Makefile compiles that .cu *from /root/reference where it lies* together
// with this file into oracle/_ref/libcorr_reference.so.  This wrapper reproduces corr_cuda.c:23-78:
// shape math, zero-fill of output and both padded scratch buffers, 2x blob_rearrange,
// CorrelateData, free of the scratch.  It is the "reference GPU" second oracle and the kernel the
// product's correlation must beat; nothing in premvos_b200/ links or loads it.
#include <cuda_runtime.h>
#include <math.h>
#include "corr_cuda_kernel.h"  // from the reference tree (-I on the command line)

extern "C" int ref_corr_cuda_forward(const float* input1, const float* input2, float* output,
                                      int batchSize, int nInputPlane, int nInputRows, int nInputCols,
                                      int pad_size, int kernel_size, int max_displacement,
                                      int stride1, int stride2, int corr_type_multiply,
                                      cudaStream_t stream) {
  long kernel_radius_ = (kernel_size - 1) / 2;
  long border_size_ = max_displacement + kernel_radius_;
  long paddedbottomheight = nInputRows + 2 * pad_size;
  long paddedbottomwidth = nInputCols + 2 * pad_size;
  long nOutputCols = ceil((float)(paddedbottomwidth - border_size_ * 2) / (float)stride1);
  long nOutputRows = ceil((float)(paddedbottomheight - border_size_ * 2) / (float)stride1);
  long neighborhood_grid_radius_ = max_displacement / stride2;
  long neighborhood_grid_width_ = neighborhood_grid_radius_ * 2 + 1;
  int nOutputPlane = neighborhood_grid_width_ * neighborhood_grid_width_;
  size_t out_bytes = (size_t)batchSize * nOutputPlane * nOutputRows * nOutputCols * sizeof(float);
  size_t rbot_bytes = (size_t)batchSize * nInputPlane * paddedbottomheight * paddedbottomwidth * sizeof(float);
  float *rbot1 = nullptr, *rbot2 = nullptr;
  if (cudaMallocAsync(&rbot1, rbot_bytes, stream) != cudaSuccess) return -1;
  if (cudaMallocAsync(&rbot2, rbot_bytes, stream) != cudaSuccess) return -1;
  cudaMemsetAsync(output, 0, out_bytes, stream);
  cudaMemsetAsync(rbot1, 0, rbot_bytes, stream);
  cudaMemsetAsync(rbot2, 0, rbot_bytes, stream);
  int pwidthheight = paddedbottomwidth * paddedbottomheight;
  long inputWidthHeight = (long)nInputRows * nInputCols;
  blob_rearrange_ongpu(input1, rbot1, batchSize, nInputPlane, nInputCols, nInputRows, inputWidthHeight, pad_size, pwidthheight, stream);
  blob_rearrange_ongpu(input2, rbot2, batchSize, nInputPlane, nInputCols, nInputRows, inputWidthHeight, pad_size, pwidthheight, stream);
  CorrelateData_ongpu(rbot1, rbot2, output, batchSize, nOutputCols, nOutputRows, nOutputPlane, max_displacement,
                      neighborhood_grid_radius_, neighborhood_grid_width_, kernel_radius_, kernel_size, stride1, stride2,
                      paddedbottomwidth, paddedbottomheight, nInputPlane, corr_type_multiply, stream);
  cudaFreeAsync(rbot1, stream);
  cudaFreeAsync(rbot2, stream);
  return (int)cudaGetLastError();
}
