/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <volk/volk.h> (upstream VOLK, a third-party
 * dependency that is not under /root/reference): the four element-wise kernels pcps_acquisition.cc calls
 * (:250,:531,:538,:547,:551-552), written from their documented semantics as plain float32 loops
 * (the compiler vectorises them with the oracle's -mavx2). */
#pragma once
#include "volk_complex.h"
#include <complex>
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n)
{
    const float* af = reinterpret_cast<const float*>(a);
    const float* bf = reinterpret_cast<const float*>(b);
    float* cf = reinterpret_cast<float*>(c);
    for (unsigned int i = 0; i < n; i++)
        {
            const float ar = af[2 * i], ai = af[2 * i + 1], br = bf[2 * i], bi = bf[2 * i + 1];
            cf[2 * i] = ar * br - ai * bi;
            cf[2 * i + 1] = ar * bi + ai * br;
        }
}
static inline void volk_32fc_conjugate_32fc(lv_32fc_t* c, const lv_32fc_t* a, unsigned int n)
{
    const float* af = reinterpret_cast<const float*>(a);
    float* cf = reinterpret_cast<float*>(c);
    for (unsigned int i = 0; i < n; i++)
        {
            cf[2 * i] = af[2 * i];
            cf[2 * i + 1] = -af[2 * i + 1];
        }
}
static inline void volk_32fc_magnitude_squared_32f(float* m, const lv_32fc_t* a, unsigned int n)
{
    const float* af = reinterpret_cast<const float*>(a);
    for (unsigned int i = 0; i < n; i++) m[i] = af[2 * i] * af[2 * i] + af[2 * i + 1] * af[2 * i + 1];
}
static inline void volk_32f_x2_add_32f(float* c, const float* a, const float* b, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) c[i] = a[i] + b[i];
}
/* sample-type conversions used by src/algorithms/libs/item_type_helpers.cc (documented VOLK semantics:
 * 8i->16i scales by 256, s32f variants divide / multiply by the scalar, float->int rounds to nearest and
 * saturates) */
#include <cmath>
#include <cstdint>
static inline void volk_8i_convert_16i(int16_t* o, const int8_t* in, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) o[i] = static_cast<int16_t>(static_cast<int16_t>(in[i]) * 256);
}
static inline void volk_16i_convert_8i(int8_t* o, const int16_t* in, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) o[i] = static_cast<int8_t>(in[i] >> 8);
}
static inline void volk_8i_s32f_convert_32f(float* o, const int8_t* in, const float scalar, unsigned int n)
{
    const float inv = 1.0f / scalar;
    for (unsigned int i = 0; i < n; i++) o[i] = static_cast<float>(in[i]) * inv;
}
static inline void volk_16i_s32f_convert_32f(float* o, const int16_t* in, const float scalar, unsigned int n)
{
    const float inv = 1.0f / scalar;
    for (unsigned int i = 0; i < n; i++) o[i] = static_cast<float>(in[i]) * inv;
}
static inline void volk_32f_s32f_convert_16i(int16_t* o, const float* in, const float scalar, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++)
        {
            float r = in[i] * scalar;
            if (r > 32767.f) r = 32767.f;
            if (r < -32768.f) r = -32768.f;
            o[i] = static_cast<int16_t>(std::lrintf(r));
        }
}
static inline void volk_32f_s32f_convert_8i(int8_t* o, const float* in, const float scalar, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++)
        {
            float r = in[i] * scalar;
            if (r > 127.f) r = 127.f;
            if (r < -128.f) r = -128.f;
            o[i] = static_cast<int8_t>(std::lrintf(r));
        }
}
/* used by src/algorithms/libs/complex_byte_to_float_x2.cc (cbyte acquisition input, compiled but never run here) */
#include <cstddef>
static inline size_t volk_get_alignment(void) { return 32; }
static inline void volk_8ic_s32f_deinterleave_32f_x2(float* i_out, float* q_out, const lv_8sc_t* in, const float scalar, unsigned int n)
{
    const int8_t* p = reinterpret_cast<const int8_t*>(in);
    const float inv = 1.0f / scalar;
    for (unsigned int k = 0; k < n; k++)
        {
            i_out[k] = static_cast<float>(p[2 * k]) * inv;
            q_out[k] = static_cast<float>(p[2 * k + 1]) * inv;
        }
}
