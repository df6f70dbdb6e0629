/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- builds into oracle/liboracle_port.so.
 *
 * Plain-C restatement of the reference's tracking-correlator arithmetic.  It is
 * the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline leg); the product path never loads it.
 *
 * PINNED: tests/test_oracle_port_vs_ref.py checks every function here BIT-EXACT
 * against the reference's own kernels (oracle/_ref/liboracle_ref.so, compiled
 * from /root/reference in place) on the reference QA's shapes (vlen 8111,
 * puppet parameters, lib/kernel_tests.h:81-89) and on the BASELINE shapes.
 *
 * VG = /root/reference/src/algorithms/libs/volk_gnsssdr_module/volk_gnsssdr
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
    float re, im;
} cf32;

/* -------------------------------------------------------------------------------------
 * a1  code resampler.
 * assoc 0: VG kernels/volk_gnsssdr/volk_gnsssdr_32f_xn_resampler_32f_xn.h:63-82 (generic):
 *          idx = floor(step*n + shift - rem)   [float: (step*n + shift) - rem]
 * assoc 1: same file :362-435 (a_avx / u_avx): for n < 8*(N/8)
 *          idx = floor(step*n + (shift - rem)), fmod by float division (:399-403),
 *          +L if negative (:406-410); tail n >= 8*(N/8) uses the generic formula (:423-433).
 * out: taps x n row-major.  idx_out (optional): the chip indices, same shape.
 * ------------------------------------------------------------------------------------- */
static inline int generic_index(float step, unsigned int n, float shift, float rem, unsigned int L)
{
    int idx = (int)floor(step * (float)n + shift - rem);
    if (idx < 0) idx += (int)L * (abs(idx) / L + 1);
    idx = idx % L;
    return idx;
}

static inline int avx_index(float step, float nf, float aux2, float Lf)
{
    float aux = step * nf;
    aux = aux + aux2;
    aux = floorf(aux);
    /* fmod */
    float c = aux / Lf;
    int i = (int)c; /* cvttps: truncate */
    float cTrunc = (float)i;
    float base = cTrunc * Lf;
    int idx = (int)(aux - base);
    /* no negatives */
    c = (float)idx;
    if (c < 0.0f) c = c + Lf;
    return (int)c;
}

int port_resampler_32f(int assoc, float* out, int* idx_out, const float* code, float rem, float step,
    const float* shifts, unsigned int L, int taps, unsigned int n)
{
    const unsigned int body = (assoc == 1) ? (n / 8) * 8 : 0;
    const float Lf = (float)L;
    for (int t = 0; t < taps; t++)
        {
            const float aux2 = shifts[t] - rem;
            float nf = 0.0f; /* indexn accumulates +8.0f per iteration upstream; exact below 2^24 */
            for (unsigned int k = 0; k < body; k++)
                {
                    const int idx = avx_index(step, nf, aux2, Lf);
                    out[(size_t)t * n + k] = code[idx];
                    if (idx_out) idx_out[(size_t)t * n + k] = idx;
                    nf += 1.0f;
                }
            for (unsigned int k = body; k < n; k++)
                {
                    const int idx = generic_index(step, k, shifts[t], rem, L);
                    out[(size_t)t * n + k] = code[idx];
                    if (idx_out) idx_out[(size_t)t * n + k] = idx;
                }
        }
    return 0;
}

/* -------------------------------------------------------------------------------------
 * a4  high-dynamics resampler, VG ..._32f_xn_high_dynamics_resampler_32f_xn.h:67-91 (generic).
 * Tap 0 gets the quadratic code phase; taps k>0 are circular integer-sample shifts of
 * tap 0 (:84-90).  NOTE upstream computes (float)(n*n) with UNSIGNED 32-bit n (:77), so the
 * square wraps for n >= 65536; restated as is.
 * ------------------------------------------------------------------------------------- */
int port_hd_resampler_32f(float* out, const float* code, float rem, float step, float rate,
    const float* shifts, unsigned int L, int taps, unsigned int n)
{
    for (unsigned int k = 0; k < n; k++)
        {
            int idx = (int)floor(step * (float)k + rate * (float)(k * k) + shifts[0] - rem);
            if (idx < 0) idx += (int)L * (abs(idx) / L + 1);
            idx = idx % L;
            out[k] = code[idx];
        }
    unsigned int shift_samples = 0;
    for (int t = 1; t < taps; t++)
        {
            shift_samples += (int)round((shifts[t] - shifts[t - 1]) / step);
            memcpy(&out[(size_t)t * n], &out[shift_samples], (n - shift_samples) * sizeof(float));
            memcpy(&out[(size_t)t * n + n - shift_samples], &out[0], shift_samples * sizeof(float));
        }
    return 0;
}

/* a4  high-dynamics resampler with the a_avx / u_avx association (same file :433-513):
 *   aux = fl(fl(step*n) + fl(rate*fl(n*n))) ; aux = fl(aux + fl(shift0 - rem)) ; floor
 * with n*n a FLOAT product (no uint32 wrap, unlike the generic kernel), the "+1 then truncate"
 * negative-index correction (:463-468), scalar generic tail for n >= 8*(N/8), then the same
 * circular sample shifts for taps > 0.  idx_out (optional): tap-0 chip indices. */
int port_hd_resampler_avx_32f(float* out, int* idx_out, const float* code, float rem, float step, float rate,
    const float* shifts, unsigned int L, int taps, unsigned int n)
{
    const unsigned int body = (n / 8) * 8;
    const float Lf = (float)L;
    const float aux2 = shifts[0] - rem;
    float nf = 0.0f;
    for (unsigned int k = 0; k < body; k++)
        {
            float aux = step * nf;
            const float nn = nf * nf;
            const float aux3 = rate * nn;
            aux = aux + aux3;
            aux = aux + aux2;
            aux = floorf(aux);
            float c = aux / Lf;
            const float c1 = c + 1.0f;
            const int i = (int)c1;
            const float base = (float)i * Lf;
            int idx = (int)(aux - base);
            c = (float)idx;
            if (c < 0.0f) c = c + Lf;
            idx = (int)c;
            out[k] = code[idx];
            if (idx_out) idx_out[k] = idx;
            nf += 1.0f;
        }
    for (unsigned int k = body; k < n; k++)
        {
            int idx = (int)floor(step * (float)k + rate * (float)(k * k) + shifts[0] - rem);
            if (idx < 0) idx += (int)L * (abs(idx) / L + 1);
            idx = idx % L;
            out[k] = code[idx];
            if (idx_out) idx_out[k] = idx;
        }
    unsigned int shift_samples = 0;
    for (int t = 1; t < taps; t++)
        {
            shift_samples += (int)round((shifts[t] - shifts[t - 1]) / step);
            memcpy(&out[(size_t)t * n], &out[shift_samples], (n - shift_samples) * sizeof(float));
            memcpy(&out[(size_t)t * n + n - shift_samples], &out[0], shift_samples * sizeof(float));
        }
    return 0;
}

/* complex helpers with the exact operation order of C99 `a * b` (no FMA: -ffp-contract=off) */
static inline cf32 cmul(cf32 a, cf32 b)
{
    cf32 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

/* -------------------------------------------------------------------------------------
 * a2  rotator + dot product, generic.
 * VG ..._32fc_32f_rotator_dot_prod_32fc_xn.h:66-98.  Renormalise phase every 256 samples
 * (including n == 0) by hypotf (:80-89).
 * ------------------------------------------------------------------------------------- */
int port_rotator_generic(cf32* result, const cf32* in, cf32 phase_inc, cf32* phase,
    const float* codes, int taps, unsigned int n)
{
    cf32 ph = *phase;
    for (int t = 0; t < taps; t++) result[t].re = result[t].im = 0.0f;
    for (unsigned int k = 0; k < n; k++)
        {
            const cf32 w = cmul(in[k], ph);
            if (k % 256 == 0)
                {
                    const float m = hypotf(ph.re, ph.im);
                    /* C99 complex / real: both parts divided */
                    ph.re = ph.re / m;
                    ph.im = ph.im / m;
                }
            ph = cmul(ph, phase_inc);
            for (int t = 0; t < taps; t++)
                {
                    const float c = codes[(size_t)t * n + k];
                    result[t].re += w.re * c;
                    result[t].im += w.im * c;
                }
        }
    *phase = ph;
    return 0;
}

/* -------------------------------------------------------------------------------------
 * a2  rotator + dot product with the summation order and phase handling of the AVX
 * implementations (u_avx :155-314, a_avx :322-484): 16 running phasors (4 registers x
 * 4 complex lanes) advanced by normalise(phase_inc^16); 4 partial accumulators x 4 lanes per
 * tap; phasors renormalised when (iteration % 64) == 0; scalar tail.
 * VG include/volk_gnsssdr/volk_gnsssdr_avx_intrinsics.h:20-31 (_mm256_complexmul_ps:
 * re = xr*yr - xi*yi, im = xi*yr + xr*yi) and :58-66 (_mm256_complexnormalise_ps).
 * ------------------------------------------------------------------------------------- */
static inline cf32 avx_cmul(cf32 x, cf32 y)
{
    cf32 r;
    r.re = x.re * y.re - x.im * y.im;
    r.im = x.im * y.re + x.re * y.im;
    return r;
}

static inline cf32 avx_normalise(cf32 z)
{
    const float m = sqrtf(z.re * z.re + z.im * z.im);
    cf32 r;
    r.re = z.re / m;
    r.im = z.im / m;
    return r;
}

int port_rotator_avx(cf32* result, const cf32* in, cf32 phase_inc, cf32* phase,
    const float* codes, int taps, unsigned int n)
{
    const unsigned int iters = n / 16;
    cf32 ph = *phase;
    cf32 z[16];
    cf32* acc = (cf32*)calloc((size_t)taps * 16, sizeof(cf32));
    for (int i = 0; i < 16; i++)
        {
            z[i] = ph;
            ph = cmul(ph, phase_inc);
        }
    cf32 dz = phase_inc;
    dz = cmul(dz, dz);
    dz = cmul(dz, dz);
    dz = cmul(dz, dz);
    dz = cmul(dz, dz);
    dz = avx_normalise(dz);

    for (unsigned int it = 0; it < iters; it++)
        {
            cf32 a[16];
            for (int i = 0; i < 16; i++) a[i] = avx_cmul(in[16 * it + i], z[i]);
            for (int i = 0; i < 16; i++) z[i] = avx_cmul(z[i], dz);
            for (int t = 0; t < taps; t++)
                {
                    const float* c = codes + (size_t)t * n + 16 * it;
                    for (int i = 0; i < 16; i++)
                        {
                            acc[t * 16 + i].re += a[i].re * c[i];
                            acc[t * 16 + i].im += a[i].im * c[i];
                        }
                }
            if ((it % 64) == 0)
                {
                    for (int i = 0; i < 16; i++) z[i] = avx_normalise(z[i]);
                }
        }
    for (int t = 0; t < taps; t++)
        {
            cf32 lane[4];
            for (int l = 0; l < 4; l++)
                {
                    /* dotProdVal0 + dotProdVal1, + dotProdVal2, + dotProdVal3 (register j holds samples 4j..4j+3) */
                    float re = acc[t * 16 + l].re + acc[t * 16 + 4 + l].re;
                    float im = acc[t * 16 + l].im + acc[t * 16 + 4 + l].im;
                    re = re + acc[t * 16 + 8 + l].re;
                    im = im + acc[t * 16 + 8 + l].im;
                    re = re + acc[t * 16 + 12 + l].re;
                    im = im + acc[t * 16 + 12 + l].im;
                    lane[l].re = re;
                    lane[l].im = im;
                }
            result[t].re = 0.0f;
            result[t].im = 0.0f;
            for (int l = 0; l < 4; l++)
                {
                    result[t].re += lane[l].re;
                    result[t].im += lane[l].im;
                }
        }
    ph = avx_normalise(z[0]);
    for (unsigned int k = iters * 16; k < n; k++)
        {
            const cf32 wo = cmul(in[k], ph);
            ph = cmul(ph, phase_inc);
            for (int t = 0; t < taps; t++)
                {
                    const float c = codes[(size_t)t * n + k];
                    result[t].re += wo.re * c;
                    result[t].im += wo.im * c;
                }
        }
    *phase = ph;
    free(acc);
    return 0;
}

/* -------------------------------------------------------------------------------------
 * a4  high-dynamics rotator, VG ..._32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn.h:68-110.
 * Sample n is rotated by phase0 * inc^n * rate^((n-1)^2) for n >= 1 (the rate term lags one
 * sample, :92-103), rate power via cpowf + hypotf renormalisation, (float)(n*n) with
 * unsigned wrap as upstream.
 * ------------------------------------------------------------------------------------- */
int port_hd_rotator_generic(cf32* result, const cf32* in, cf32 phase_inc, cf32 phase_inc_rate, cf32* phase,
    const float* codes, int taps, unsigned int n)
{
    float complex ph = CMPLXF(phase->re, phase->im);
    float complex phase_doppler = ph;
    float complex phase_doppler_rate;
    const float complex inc = CMPLXF(phase_inc.re, phase_inc.im);
    const float complex rate = CMPLXF(phase_inc_rate.re, phase_inc_rate.im);
    for (int t = 0; t < taps; t++) result[t].re = result[t].im = 0.0f;
    for (unsigned int k = 0; k < n; k++)
        {
            if (k % 256 == 0)
                {
                    ph /= hypotf(crealf(ph), cimagf(ph));
                }
            const float complex x = CMPLXF(in[k].re, in[k].im);
            const float complex w = x * ph;
            phase_doppler *= inc;
            phase_doppler_rate = cpowf(rate, CMPLXF((float)(k * k), 0.0f));
            phase_doppler_rate /= hypotf(crealf(phase_doppler_rate), cimagf(phase_doppler_rate));
            ph = phase_doppler * phase_doppler_rate;
            for (int t = 0; t < taps; t++)
                {
                    const float c = codes[(size_t)t * n + k];
                    result[t].re += crealf(w) * c;
                    result[t].im += cimagf(w) * c;
                }
        }
    phase->re = crealf(ph);
    phase->im = cimagf(ph);
    return 0;
}

/* -------------------------------------------------------------------------------------
 * a3  Cpu_Multicorrelator_Real_Codes::Carrier_wipeoff_multicorrelator_resampler,
 * src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.cc:103-127:
 *   update_local_code (a1) ; phase0 = (cos(rem), -sin(rem)) ; inc = exp(-j*step) ; a2.
 * arch 0: generic resampler + generic rotator; arch 1: AVX association + AVX order.
 * scratch: taps*n floats (the d_local_codes_resampled buffers).
 * ------------------------------------------------------------------------------------- */
int port_multicorrelator(int arch, cf32* out_taps, const cf32* in, const float* code, unsigned int L,
    const float* shifts, int taps, float rem_carr_rad, float phase_step_rad, float rem_code_chips,
    float code_step_chips, unsigned int n, float* scratch)
{
    port_resampler_32f(arch, scratch, NULL, code, rem_code_chips, code_step_chips, shifts, L, taps, n);
    cf32 ph;
    ph.re = cosf(rem_carr_rad);
    ph.im = -sinf(rem_carr_rad);
    const float complex e = cexpf(CMPLXF(0.0f, -phase_step_rad));
    cf32 inc;
    inc.re = crealf(e);
    inc.im = cimagf(e);
    if (arch == 0) return port_rotator_generic(out_taps, in, inc, &ph, scratch, taps, n);
    return port_rotator_avx(out_taps, in, inc, &ph, scratch, taps, n);
}

/* -------------------------------------------------------------------------------------
 * Float64 ground truth for the same epoch: chip indices from the float32 formula selected
 * by `assoc` (they define WHICH chip multiplies WHICH sample and must not move), carrier
 * exp(-j(rem + n*step)) and accumulation in double.  Bounds the error of both CPU and GPU.
 * ------------------------------------------------------------------------------------- */
int port_multicorrelator_f64(int assoc, double* out_taps_re_im, const cf32* in, const float* code, unsigned int L,
    const float* shifts, int taps, float rem_carr_rad, float phase_step_rad, float rem_code_chips,
    float code_step_chips, unsigned int n)
{
    const unsigned int body = (assoc == 1) ? (n / 8) * 8 : 0;
    const float Lf = (float)L;
    for (int t = 0; t < taps; t++) out_taps_re_im[2 * t] = out_taps_re_im[2 * t + 1] = 0.0;
    for (unsigned int k = 0; k < n; k++)
        {
            const double ph = -((double)rem_carr_rad + (double)k * (double)phase_step_rad);
            const double c = cos(ph), s = sin(ph);
            const double wr = (double)in[k].re * c - (double)in[k].im * s;
            const double wi = (double)in[k].re * s + (double)in[k].im * c;
            for (int t = 0; t < taps; t++)
                {
                    const int idx = (k < body) ? avx_index(code_step_chips, (float)k, shifts[t] - rem_code_chips, Lf)
                                               : generic_index(code_step_chips, k, shifts[t], rem_code_chips, L);
                    out_taps_re_im[2 * t] += wr * (double)code[idx];
                    out_taps_re_im[2 * t + 1] += wi * (double)code[idx];
                }
        }
    return 0;
}

/* Batch driver for bench.py's cpu_baseline when oracle/_ref is unavailable ("port" kind) and
 * for bulk parity checks: `threads` pthreads, item i handled by thread i % threads. */
#include <pthread.h>
typedef struct
{
    int arch, tid, threads, taps, n_items;
    cf32* out_taps;
    const cf32* in;
    size_t in_stride;
    const float* code;
    unsigned int L, n;
    const float* shifts;
    const float* params;
} batch_job;

static void* batch_worker(void* pv)
{
    batch_job* j = (batch_job*)pv;
    float* scratch = (float*)malloc(sizeof(float) * (size_t)j->taps * j->n);
    for (int i = j->tid; i < j->n_items; i += j->threads)
        {
            const float* p = j->params + 4 * (size_t)i;
            port_multicorrelator(j->arch, j->out_taps + (size_t)i * j->taps, j->in + (size_t)i * j->in_stride, j->code, j->L,
                j->shifts, j->taps, p[0], p[1], p[2], p[3], j->n, scratch);
        }
    free(scratch);
    return NULL;
}

/* params: 4 floats per item = rem_carr_rad, phase_step_rad, rem_code_chips, code_step_chips */
int port_multicorrelator_batch(int arch, int threads, cf32* out_taps, const cf32* in, size_t in_stride,
    const float* code, unsigned int L, const float* shifts, int taps, const float* params,
    unsigned int n, int n_items)
{
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * (size_t)threads);
    for (int t = 0; t < threads; t++)
        {
            batch_job j = {arch, t, threads, taps, n_items, out_taps, in, in_stride, code, L, n, shifts, params};
            jobs[t] = j;
            pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
        }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
