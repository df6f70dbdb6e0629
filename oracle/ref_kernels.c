/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- builds into oracle/_ref/liboracle_ref.so.
 *
 * This file contains NO algorithm.  It #includes the reference's own
 * volk_gnsssdr kernel headers where they lie under /root/reference
 * (src/algorithms/libs/volk_gnsssdr_module/volk_gnsssdr/kernels/volk_gnsssdr/)
 * and exports thin flat-array wrappers around every architecture variant
 * ("generic", "a_avx", "u_avx", ...) so that tests/ and bench.py's cpu_baseline
 * leg can call the REAL reference arithmetic through ctypes.
 *
 * It also defines the dispatcher function pointers that upstream generates with
 * Mako (tmpl/volk_gnsssdr.tmpl.c:146-184) so that the reference's
 * cpu_multicorrelator_real_codes.cc links unmodified (see ref_engine.cc).
 */
#include <volk_gnsssdr/volk_gnsssdr.h>

#include "volk_gnsssdr_32f_xn_resampler_32f_xn.h"
#include "volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn.h"
#include "volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn.h"
#include "volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn.h"
#include "volk_gnsssdr_32f_index_max_32u.h"
#include "volk_gnsssdr_s32f_sincos_32fc.h"

#include <stdlib.h>
#include <string.h>

/* ---- dispatcher (what upstream's generated volk_gnsssdr.c provides) ------------- */
p_32f_xn_resampler_32f_xn volk_gnsssdr_32f_xn_resampler_32f_xn = volk_gnsssdr_32f_xn_resampler_32f_xn_a_avx;
p_32f_xn_high_dynamics_resampler_32f_xn volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn = volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_a_avx;
p_32fc_32f_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_a_avx;
p_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn_generic;

size_t volk_gnsssdr_get_alignment(void) { return 32; }

/* Select which implementation the dispatcher pointers use: "generic", "a_avx", "u_avx".
 * Returns 0 on success. */
int ref_select_arch(const char* arch)
{
    if (strcmp(arch, "generic") == 0)
        {
            volk_gnsssdr_32f_xn_resampler_32f_xn = volk_gnsssdr_32f_xn_resampler_32f_xn_generic;
            volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn = volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_generic;
            volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_generic;
        }
    else if (strcmp(arch, "a_avx") == 0)
        {
            volk_gnsssdr_32f_xn_resampler_32f_xn = volk_gnsssdr_32f_xn_resampler_32f_xn_a_avx;
            volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn = volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_a_avx;
            volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_a_avx;
        }
    else if (strcmp(arch, "u_avx") == 0)
        {
            volk_gnsssdr_32f_xn_resampler_32f_xn = volk_gnsssdr_32f_xn_resampler_32f_xn_u_avx;
            volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn = volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_u_avx;
            volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_u_avx;
        }
    else
        {
            return -1;
        }
    /* only a generic implementation of the HD rotator exists upstream */
    volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn_generic;
    return 0;
}

/* ---- helpers --------------------------------------------------------------------- */
static void* amalloc(size_t bytes)
{
    void* p = NULL;
    if (bytes == 0) bytes = 32;
    if (posix_memalign(&p, 32, (bytes + 31) & ~(size_t)31) != 0) return NULL;
    return p;
}

/* ---- resampler: out is taps x n, row-major ------------------------------------------ */
/* variant: 0 generic, 1 a_avx, 2 u_avx, 3 a_sse3, 4 a_sse4_1 */
int ref_resampler_32f(int variant, float* out, const float* code, float rem, float step,
    const float* shifts, unsigned int code_len, int taps, unsigned int n)
{
    float** rows = (float**)malloc(sizeof(float*) * (size_t)taps);
    float* sh = (float*)amalloc(sizeof(float) * (size_t)taps);
    float* codea = (float*)amalloc(sizeof(float) * code_len);
    int t;
    memcpy(sh, shifts, sizeof(float) * (size_t)taps);
    memcpy(codea, code, sizeof(float) * code_len);
    for (t = 0; t < taps; t++) rows[t] = (float*)amalloc(sizeof(float) * (n + 8));
    switch (variant)
        {
        case 0: volk_gnsssdr_32f_xn_resampler_32f_xn_generic(rows, codea, rem, step, sh, code_len, taps, n); break;
        case 1: volk_gnsssdr_32f_xn_resampler_32f_xn_a_avx(rows, codea, rem, step, sh, code_len, taps, n); break;
        case 2: volk_gnsssdr_32f_xn_resampler_32f_xn_u_avx(rows, codea, rem, step, sh, code_len, taps, n); break;
        case 3: volk_gnsssdr_32f_xn_resampler_32f_xn_a_sse3(rows, codea, rem, step, sh, code_len, taps, n); break;
        case 4: volk_gnsssdr_32f_xn_resampler_32f_xn_a_sse4_1(rows, codea, rem, step, sh, code_len, taps, n); break;
        default: return -1;
        }
    for (t = 0; t < taps; t++)
        {
            memcpy(out + (size_t)t * n, rows[t], sizeof(float) * n);
            free(rows[t]);
        }
    free(rows);
    free(sh);
    free(codea);
    return 0;
}

/* variant: 0 generic, 1 a_avx, 2 u_avx */
int ref_hd_resampler_32f(int variant, float* out, const float* code, float rem, float step, float rate,
    const float* shifts, unsigned int code_len, int taps, unsigned int n)
{
    float** rows = (float**)malloc(sizeof(float*) * (size_t)taps);
    float* sh = (float*)amalloc(sizeof(float) * (size_t)taps);
    float* codea = (float*)amalloc(sizeof(float) * code_len);
    int t;
    memcpy(sh, shifts, sizeof(float) * (size_t)taps);
    memcpy(codea, code, sizeof(float) * code_len);
    for (t = 0; t < taps; t++) rows[t] = (float*)amalloc(sizeof(float) * (n + 8));
    switch (variant)
        {
        case 0: volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_generic(rows, codea, rem, step, rate, sh, code_len, taps, n); break;
        case 1: volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_a_avx(rows, codea, rem, step, rate, sh, code_len, taps, n); break;
        case 2: volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn_u_avx(rows, codea, rem, step, rate, sh, code_len, taps, n); break;
        default: return -1;
        }
    for (t = 0; t < taps; t++)
        {
            memcpy(out + (size_t)t * n, rows[t], sizeof(float) * n);
            free(rows[t]);
        }
    free(rows);
    free(sh);
    free(codea);
    return 0;
}

/* ---- rotator + dot product: codes is taps x n row-major; complex = interleaved floats ---- */
/* variant: 0 generic, 1 generic_reload, 2 u_avx, 3 a_avx.  phase[2] is in/out. */
int ref_rotator_dot_prod_32fc_32f(int variant, float* result, const float* in_iq, const float* phase_inc,
    float* phase, const float* codes, int taps, unsigned int n)
{
    const float** rows = (const float**)malloc(sizeof(float*) * (size_t)taps);
    lv_32fc_t* res = (lv_32fc_t*)amalloc(sizeof(lv_32fc_t) * (size_t)taps);
    lv_32fc_t* in = (lv_32fc_t*)amalloc(sizeof(lv_32fc_t) * n);
    float* codea = (float*)amalloc(sizeof(float) * (size_t)taps * (n + 8));
    lv_32fc_t ph = lv_cmake(phase[0], phase[1]);
    const lv_32fc_t inc = lv_cmake(phase_inc[0], phase_inc[1]);
    int t;
    memcpy(in, in_iq, sizeof(lv_32fc_t) * n);
    for (t = 0; t < taps; t++)
        {
            float* r = codea + (size_t)t * (n + 8);
            memcpy(r, codes + (size_t)t * n, sizeof(float) * n);
            rows[t] = r;
        }
    switch (variant)
        {
        case 0: volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_generic(res, in, inc, &ph, rows, taps, n); break;
        case 1: volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_generic_reload(res, in, inc, &ph, rows, taps, n); break;
        case 2: volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_u_avx(res, in, inc, &ph, rows, taps, n); break;
        case 3: volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn_a_avx(res, in, inc, &ph, rows, taps, n); break;
        default: return -1;
        }
    for (t = 0; t < taps; t++)
        {
            result[2 * t] = lv_creal(res[t]);
            result[2 * t + 1] = lv_cimag(res[t]);
        }
    phase[0] = lv_creal(ph);
    phase[1] = lv_cimag(ph);
    free(rows);
    free(res);
    free(in);
    free(codea);
    return 0;
}

/* only the generic HD rotator exists upstream (variant 0 generic, 1 generic_arg) */
int ref_hd_rotator_dot_prod_32fc_32f(int variant, float* result, const float* in_iq, const float* phase_inc,
    const float* phase_inc_rate, float* phase, const float* codes, int taps, unsigned int n)
{
    const float** rows = (const float**)malloc(sizeof(float*) * (size_t)taps);
    lv_32fc_t* res = (lv_32fc_t*)amalloc(sizeof(lv_32fc_t) * (size_t)taps);
    lv_32fc_t ph = lv_cmake(phase[0], phase[1]);
    const lv_32fc_t inc = lv_cmake(phase_inc[0], phase_inc[1]);
    const lv_32fc_t rate = lv_cmake(phase_inc_rate[0], phase_inc_rate[1]);
    int t;
    for (t = 0; t < taps; t++) rows[t] = codes + (size_t)t * n;
    switch (variant)
        {
        case 0: volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn_generic(res, (const lv_32fc_t*)in_iq, inc, rate, &ph, rows, taps, n); break;
        case 1: volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn_generic_arg(res, (const lv_32fc_t*)in_iq, inc, rate, &ph, rows, taps, n); break;
        default: return -1;
        }
    for (t = 0; t < taps; t++)
        {
            result[2 * t] = lv_creal(res[t]);
            result[2 * t + 1] = lv_cimag(res[t]);
        }
    phase[0] = lv_creal(ph);
    phase[1] = lv_cimag(ph);
    free(rows);
    free(res);
    return 0;
}

/* ---- acquisition helpers ------------------------------------------------------------- */
/* variant: 0 generic, 1 generic_fxpt, 2 a_sse2, 3 u_sse2, 4 a_avx2, 5 u_avx2.  out: n interleaved cf32 */
int ref_sincos_32fc(int variant, float* out, float phase_inc, float* phase, unsigned int n)
{
    lv_32fc_t* o = (lv_32fc_t*)amalloc(sizeof(lv_32fc_t) * (n + 8));
    switch (variant)
        {
        case 0: volk_gnsssdr_s32f_sincos_32fc_generic(o, phase_inc, phase, n); break;
        case 1: volk_gnsssdr_s32f_sincos_32fc_generic_fxpt(o, phase_inc, phase, n); break;
        case 2: volk_gnsssdr_s32f_sincos_32fc_a_sse2(o, phase_inc, phase, n); break;
        case 3: volk_gnsssdr_s32f_sincos_32fc_u_sse2(o, phase_inc, phase, n); break;
#ifdef LV_HAVE_AVX2
        case 4: volk_gnsssdr_s32f_sincos_32fc_a_avx2(o, phase_inc, phase, n); break;
        case 5: volk_gnsssdr_s32f_sincos_32fc_u_avx2(o, phase_inc, phase, n); break;
#endif
        default: free(o); return -1;
        }
    memcpy(out, o, sizeof(lv_32fc_t) * n);
    free(o);
    return 0;
}

/* variant: 0 generic, 1 a_avx, 2 u_avx, 3 a_sse4_1, 4 a_sse */
int ref_index_max_32u(int variant, unsigned int* target, const float* src, unsigned int n)
{
    float* s = (float*)amalloc(sizeof(float) * (n + 8));
    uint32_t idx = 0;
    memcpy(s, src, sizeof(float) * n);
    switch (variant)
        {
        case 0: volk_gnsssdr_32f_index_max_32u_generic(&idx, s, n); break;
        case 1: volk_gnsssdr_32f_index_max_32u_a_avx(&idx, s, n); break;
        case 2: volk_gnsssdr_32f_index_max_32u_u_avx(&idx, s, n); break;
        case 3: volk_gnsssdr_32f_index_max_32u_a_sse4_1(&idx, s, n); break;
        case 4: volk_gnsssdr_32f_index_max_32u_a_sse(&idx, s, n); break;
        default: free(s); return -1;
        }
    *target = idx;
    free(s);
    return 0;
}
