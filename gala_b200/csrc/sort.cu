// sort.cu -- (key, value) radix sort used to order the DOP853 orbit queue by dynamical time.
// cub::DeviceRadixSort is library plumbing (ships with the CUDA toolkit); the hot path is in
// dop853.cuh.
#include <cub/device/device_radix_sort.cuh>
#include "kernels.h"

cudaError_t gb_sort_pairs_bytes(size_t n, size_t* temp_bytes) {
    *temp_bytes = 0;
    return cub::DeviceRadixSort::SortPairs(nullptr, *temp_bytes, (const float*)nullptr, (float*)nullptr,
                                           (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
}
cudaError_t gb_sort_pairs(const float* keys_in, float* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                          size_t n, void* temp, size_t temp_bytes, cudaStream_t s) {
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, s);
}
