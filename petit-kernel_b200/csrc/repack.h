// Internal interface to the repack / dequant-hook kernels (repack.cu).
#pragma once

#include <cuda_runtime.h>

namespace petit::repack {

bool shape_ok(unsigned size_k, unsigned size_n);

// 0 ok, 1 bad shape, 3 CUDA error
// native16: the fp16-native variant of the packed layout (words keep the native nibble order)
int weights(void *out, const void *in, unsigned size_k, unsigned size_n, bool unpack,
            cudaStream_t stream, bool native16 = false);
int scales(void *out, const void *in, unsigned size_k, unsigned size_n, bool mx, bool unpack,
           cudaStream_t stream);
// 0 ok, -1 on bad shape / type / launch failure (quantization_utils.cu:619-621)
int dequant_dense(void *out, const void *w, const void *sc, float global_scale, int mode,
                  bool packed, unsigned size_k, unsigned size_n, cudaStream_t stream);

} // namespace petit::repack
