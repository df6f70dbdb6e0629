// Device-side sample-type adapters (SURVEY "next" row N3): the front end delivers interleaved 16-bit
// or 8-bit I/Q; the reference converts them to gr_complex on the CPU before the channels
// (src/algorithms/data_type_adapter/gnuradio_blocks/cshort_to_gr_complex.cc:48 ->
//  volk_gnsssdr_16ic_convert_32fc, VG kernels/volk_gnsssdr/volk_gnsssdr_16ic_convert_32fc.h:
//  out = (float)re, (float)im; ibyte_to_complex.cc -> volk_8i_s32f_convert_32f with scale 1).
// Here the raw integers cross PCIe (4 or 2 bytes per sample instead of 8) and one streaming kernel
// writes float2 into the band ring.  int -> float conversion is exact, so parity is bit-exact.
#include "common.cuh"

namespace b200
{
namespace
{
template <typename T>
__global__ void convert_to_ring_kernel(const T* __restrict__ raw, float2* __restrict__ ring, unsigned long long mask,
    unsigned long long dst_off, unsigned long long n)
{
    // each thread converts 4 complex samples (vector load of the raw pairs)
    const unsigned long long i0 = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4ULL;
    if (i0 >= n) return;
    if (i0 + 4 <= n)
        {
            T v[8];
            if (sizeof(T) == 2)
                *reinterpret_cast<int4*>(v) = __ldg(reinterpret_cast<const int4*>(raw + 2 * i0));
            else
                *reinterpret_cast<int2*>(v) = __ldg(reinterpret_cast<const int2*>(raw + 2 * i0));
#pragma unroll
            for (int k = 0; k < 4; k++)
                ring[(dst_off + i0 + k) & mask] = make_float2(static_cast<float>(v[2 * k]), static_cast<float>(v[2 * k + 1]));
        }
    else
        {
            for (unsigned long long i = i0; i < n; i++)
                ring[(dst_off + i) & mask] = make_float2(static_cast<float>(raw[2 * i]), static_cast<float>(raw[2 * i + 1]));
        }
}
}  // namespace

int launch_convert_i16(const short* raw, float2* ring, unsigned long long mask, unsigned long long dst_off, unsigned long long n, cudaStream_t st)
{
    if (n == 0) return B200_OK;
    const unsigned long long threads = (n + 3) / 4;
    convert_to_ring_kernel<short><<<static_cast<unsigned int>((threads + 255) / 256), 256, 0, st>>>(raw, ring, mask, dst_off, n);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

int launch_convert_i8(const signed char* raw, float2* ring, unsigned long long mask, unsigned long long dst_off, unsigned long long n, cudaStream_t st)
{
    if (n == 0) return B200_OK;
    const unsigned long long threads = (n + 3) / 4;
    convert_to_ring_kernel<signed char><<<static_cast<unsigned int>((threads + 255) / 256), 256, 0, st>>>(raw, ring, mask, dst_off, n);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
}  // namespace b200
