/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- builds into oracle/liboracle_port.so.
 *
 * Plain-C restatement of the volk_gnsssdr kernels the PCPS acquisition uses
 * (pcps_acquisition.cc:275-281 update_local_carrier -> volk_gnsssdr_s32f_sincos_32fc,
 *  pcps_acquisition.cc:417,464,514 -> volk_gnsssdr_32f_index_max_32u).
 * The grid search itself (FFT-based; FFTW is not in this image) is restated in numpy in
 * oracle/acq_np.py.  Pinned bit-exact against oracle/_ref by tests/test_oracle_acq.py.
 *
 * VG = /root/reference/src/algorithms/libs/volk_gnsssdr_module/volk_gnsssdr
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* VG kernels/volk_gnsssdr/volk_gnsssdr_s32f_sincos_32fc.h:390-400 (generic): the phase is
 * ACCUMULATED IN FLOAT32, out[i] = (cosf(phase), sinf(phase)). */
int port_sincos_generic(float* out_iq, float phase_inc, float* phase, unsigned int n)
{
    float p = *phase;
    for (unsigned int i = 0; i < n; i++)
        {
            out_iq[2 * i] = cosf(p);
            out_iq[2 * i + 1] = sinf(p);
            p += phase_inc;
        }
    *phase = p;
    return 0;
}

static inline float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* One lane of the Cephes-style polynomial of the AVX2 kernel (same file :448-633), every
 * operation a separately rounded float32 op in the upstream order. */
static inline void cephes_sincos_lane(float p, float* s_out, float* c_out)
{
    const float FOPI = 1.27323954473516;
    const float DP1 = -0.78515625, DP2 = -2.4187564849853515625e-4, DP3 = -3.77489497744594108e-8;
    const float cc0 = 2.443315711809948E-005, cc1 = -1.388731625493765E-003, cc2 = 4.166664568298827E-002;
    const float sc0 = -1.9515295891E-4, sc1 = 8.3321608736E-3, sc2 = -1.6666654611E-1;
    float x = u2f(f2u(p) & 0x7fffffffu);
    uint32_t sign_bit_sin = f2u(p) & 0x80000000u;
    float y = x * FOPI;
    int32_t emm2 = (int32_t)y; /* cvttps */
    emm2 = emm2 + 1;
    emm2 = emm2 & ~1;
    y = (float)emm2;
    int32_t emm4 = emm2;
    const uint32_t swap_sign_bit_sin = ((uint32_t)(emm2 & 4)) << 29;
    const uint32_t poly_mask = ((emm2 & 2) == 0) ? 0xffffffffu : 0u;
    const float xmm1 = y * DP1, xmm2 = y * DP2, xmm3 = y * DP3;
    x = x + xmm1;
    x = x + xmm2;
    x = x + xmm3;
    emm4 = emm4 - 2;
    const uint32_t sign_bit_cos = ((uint32_t)((~emm4) & 4)) << 29;
    sign_bit_sin ^= swap_sign_bit_sin;
    const float z = x * x;
    y = cc0;
    y = y * z;
    y = y + cc1;
    y = y * z;
    y = y + cc2;
    y = y * z;
    y = y * z;
    const float tmp = z * 0.5f;
    y = y - tmp;
    y = y + 1.0f;
    float y2 = sc0;
    y2 = y2 * z;
    y2 = y2 + sc1;
    y2 = y2 * z;
    y2 = y2 + sc2;
    y2 = y2 * z;
    y2 = y2 * x;
    y2 = y2 + x;
    const float ysin2 = u2f(poly_mask & f2u(y2));
    const float ysin1 = u2f(~poly_mask & f2u(y));
    y2 = y2 - ysin2;
    y = y - ysin1;
    const float sm = ysin1 + ysin2;
    const float cm = y + y2;
    *s_out = u2f(f2u(sm) ^ sign_bit_sin);
    *c_out = u2f(f2u(cm) ^ sign_bit_cos);
}

/* VG ..._s32f_sincos_32fc.h:448-633 (a_avx2) / :635-820 (u_avx2): eight lane phases
 * p_l = phase + l*inc (l*inc rounded first), each advanced by fl(8*inc) per iteration;
 * scalar tail restarts from fl(phase + fl(inc * (float)(8*iters))) with cosf/sinf. */
int port_sincos_avx2(float* out_iq, float phase_inc, float* phase, unsigned int n)
{
    const unsigned int iters = n / 8;
    float _phase = *phase;
    float lane[8];
    lane[0] = _phase;
    lane[1] = _phase + phase_inc;
    for (int l = 2; l < 8; l++) lane[l] = _phase + (float)l * phase_inc;
    const float inc8 = 8 * phase_inc;
    for (unsigned int it = 0; it < iters; it++)
        {
            for (int l = 0; l < 8; l++)
                {
                    float s, c;
                    cephes_sincos_lane(lane[l], &s, &c);
                    out_iq[2 * (8 * it + l)] = c;
                    out_iq[2 * (8 * it + l) + 1] = s;
                    lane[l] = lane[l] + inc8;
                }
        }
    _phase = _phase + phase_inc * (iters * 8);
    for (unsigned int i = iters * 8; i < n; i++)
        {
            out_iq[2 * i] = cosf(_phase);
            out_iq[2 * i + 1] = sinf(_phase);
            _phase += phase_inc;
        }
    *phase = _phase;
    return 0;
}

/* VG ..._32f_index_max_32u.h:446-466 (generic): strict '>' so the FIRST maximum wins.
 * (The SIMD variants :49-113 resolve ties differently only when equal maxima fall in the
 * same vector lane pass; on distinct-valued data all variants agree.) */
int port_index_max_32u(uint32_t* target, const float* src, uint32_t n)
{
    if (n > 0)
        {
            float max = src[0];
            uint32_t index = 0;
            for (uint32_t i = 1; i < n; ++i)
                {
                    if (src[i] > max)
                        {
                            index = i;
                            max = src[i];
                        }
                }
            *target = index;
        }
    return 0;
}

/* pcps_acquisition.cc:284-291 + :275-281: the Doppler wipe-off grid.
 *   doppler(d) = -doppler_max + doppler_center + doppler_step*d      (int32)
 *   phase_step = (float)TWO_PI * (float)(doppler_bias + doppler) / (float)fs_in
 *   wipe[d][:] = sincos(-phase_step), phase starting at 0
 * variant 0: generic sincos; 1: avx2 sincos (what an x86-64 VOLK dispatch runs). out: bins x n cf32. */
int port_acq_wipeoff_grid(int variant, float* out_iq, unsigned int n, unsigned int bins, int32_t doppler_max,
    int32_t doppler_center, int32_t doppler_step, int32_t doppler_bias, int64_t fs_in)
{
    for (unsigned int d = 0; d < bins; d++)
        {
            const int32_t doppler = -doppler_max + doppler_center + doppler_step * (int32_t)d;
            const float freq = (float)(doppler_bias + doppler);
            const float phase_step_rad = (float)6.283185307179586 * freq / (float)fs_in;
            float ph = 0.0f;
            float* row = out_iq + (size_t)2 * n * d;
            if (variant == 0)
                port_sincos_generic(row, -phase_step_rad, &ph, n);
            else
                port_sincos_avx2(row, -phase_step_rad, &ph, n);
        }
    return 0;
}

/* std::accumulate(begin, end, 0.0f): strictly sequential float32 sum (pcps_acquisition.cc:431) */
float port_seq_sum_f32(const float* src, unsigned int n)
{
    float acc = 0.0f;
    for (unsigned int i = 0; i < n; i++) acc += src[i];
    return acc;
}

/* pcps_acquisition.cc:294-301 update_grid_doppler_wipeoffs_step2: float Doppler offsets around the
 * step-one estimate, doppler = (float(idx) - float(floor(bins2/2.0))) * step2, carrier frequency
 * center2 + doppler (no FDMA bias in this branch upstream). */
int port_acq_wipeoff_grid_step2(int variant, float* out_iq, unsigned int n, unsigned int bins2, float center2, float step2, int64_t fs_in)
{
    for (unsigned int d = 0; d < bins2; d++)
        {
            const float doppler = ((float)d - (float)floor(bins2 / 2.0)) * step2;
            const float freq = center2 + doppler;
            const float phase_step_rad = (float)6.283185307179586 * freq / (float)fs_in;
            float ph = 0.0f;
            float* row = out_iq + (size_t)2 * n * d;
            if (variant == 0)
                port_sincos_generic(row, -phase_step_rad, &ph, n);
            else
                port_sincos_avx2(row, -phase_step_rad, &ph, n);
        }
    return 0;
}
