// cnn_internal.h -- launchers shared between conv.cu (fp32 direct path) and conv_tc.cu (tensor path)
#pragma once
#include <cuda_runtime.h>

// fnx_conv2d plus an optional device word receiving max|y| (atomicMax on the fp32 bit pattern)
int fnx_conv_direct(const float* x, const float* weight, const float* bias, float* y, int N, int Cin, int H, int W,
                    int Cout, int ksize, int relu, int y_channels_total, int y_channel_offset, unsigned* amax_bits,
                    cudaStream_t st);
