/* TEST INFRASTRUCTURE ONLY - CPU restatement (oracle "port") of the per-epoch DLL/PLL cycle of
 * dll_pll_veml_tracking in its tracking state.  Nothing in the product path links or loads this file.
 *
 * Follows, statement by statement and type by type (float where the reference is float, double where it is
 * double, host libm):
 *   start_tracking                      src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:791-1078
 *   general_work, head + cases 1 and 2  :1898-2015, tail :2292-2294
 *   cn0_and_tracking_lock_status        :1167-1224
 *   run_dll_pll                         :1260-1347   (enable_doppler_correction branch not restated: default false)
 *   update_tracking_vars                :1409-1483   (high_dyn branches not restated: default false)
 *   clear_tracking_vars                 :1359-1383
 *   log_data                            :1599-1694
 *   tracking_discriminators.cc          :26-39 phase_unwrap, :69-77 fll_diff_atan, :86-89 pll_four_quadrant_atan,
 *                                       :100-107 pll_cloop_two_quadrant_atan, :119-129 dll_nc_e_minus_l_normalized,
 *                                       :142-153 dll_nc_vemlp_normalized
 *   tracking_FLL_PLL_filter.cc          :23-54 set_params, :57-69 initialize, :72-104 get_carrier_error
 *   tracking_loop_filter.cc             :62-96 apply, :99-196 update_coefficients, :258-263 initialize
 *   lock_detectors.cc                   :99-147 cn0_m2m4_estimator, :160-181 carrier_lock_detector
 *   exponential_smoother.cc             :29-42 set_alpha, :64-69 reset, :86-115 smooth
 *
 * Pinned against the reference's own object code for every library function above by
 * tests/test_oracle_loop.py (oracle/_ref/liboracle_ref_loop.so = those .cc files compiled where they lie).
 * pll_four_quadrant_atan calls gr::fast_atan2f (GNU Radio, absent): restated with atan2f, parity unpinned.
 */
#include "b200gnss.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TWO_PI 6.283185307179586 /* MATH_CONSTANTS.h */
#define GNSS_PI 3.1415926535898
#define HALF_PI 1.570796326794897
#define MAX_LOOP_HISTORY_LENGTH 4
#define MAX_CN0_SAMPLES 64

typedef struct
{
    float re, im;
} cf;

typedef struct
{
    float alpha, one_minus_alpha, old_value, min_value, offset;
    int samples_for_initialization, init_counter, initializing;
    float init_sum; /* std::accumulate(init_buffer_, 0.0F): sequential float sum, same order as a running sum */
    int init_n;
} smoother_t;

static void smoother_defaults(smoother_t* s)
{
    s->alpha = 0.001f;
    s->one_minus_alpha = 0.999f;
    s->old_value = 0.0f;
    s->min_value = 25.0f;
    s->offset = 12.0f;
    s->samples_for_initialization = 200;
    s->init_counter = 0;
    s->initializing = 1;
    s->init_sum = 0.0f;
    s->init_n = 0;
}

static void smoother_set_alpha(smoother_t* s, float alpha)
{
    s->alpha = alpha;
    if (s->alpha < 0) s->alpha = 0;
    if (s->alpha > 1) s->alpha = 1;
    s->one_minus_alpha = 1.0f - s->alpha;
}

static void smoother_set_samples(smoother_t* s, int n) { s->samples_for_initialization = n <= 0 ? 1 : n; }

static void smoother_reset(smoother_t* s)
{
    s->initializing = 1;
    s->init_counter = 0;
    s->init_sum = 0.0f;
    s->init_n = 0;
}

static float smoother_smooth(smoother_t* s, float raw)
{
    float smoothed_value;
    if (s->initializing)
        {
            s->init_counter++;
            smoothed_value = raw;
            s->init_sum = s->init_sum + smoothed_value;
            s->init_n++;
            if (s->init_counter == s->samples_for_initialization)
                {
                    s->old_value = s->init_sum / (float)s->init_n;
                    if (s->old_value < (s->min_value + s->offset))
                        {
                            s->init_counter = 0;
                            s->init_sum = 0.0f;
                            s->init_n = 0;
                        }
                    else
                        {
                            s->initializing = 0;
                        }
                }
        }
    else
        {
            smoothed_value = s->alpha * raw + s->one_minus_alpha * s->old_value;
            s->old_value = smoothed_value;
        }
    return smoothed_value;
}

typedef struct
{
    float in_c[4], out_c[3];
    int n_in, n_out;
    float inputs[MAX_LOOP_HISTORY_LENGTH], outputs[MAX_LOOP_HISTORY_LENGTH];
    int current_index;
    float noise_bandwidth, update_interval;
    int order, include_last_integrator;
} loop_filter_t;

static void loop_filter_update_coefficients(loop_filter_t* f)
{
    float g1, g2, g3, wn;
    const float T = f->update_interval;
    const float zeta = 1.0f / sqrtf(2.0f);
    switch (f->order)
        {
        case 1:
            wn = f->noise_bandwidth * 4.0f;
            g1 = wn;
            if (f->include_last_integrator)
                {
                    f->n_in = 2;
                    f->in_c[0] = (float)(g1 * T / 2.0);
                    f->in_c[1] = (float)(g1 * T / 2.0);
                    f->n_out = 1;
                    f->out_c[0] = 1.0f;
                }
            else
                {
                    f->n_in = 1;
                    f->in_c[0] = g1;
                    f->n_out = 0;
                }
            break;
        case 2:
            wn = f->noise_bandwidth * (8.0f * zeta) / (4.0f * zeta * zeta + 1.0f);
            g1 = wn * wn;
            g2 = wn * 2.0f * zeta;
            if (f->include_last_integrator)
                {
                    f->n_in = 3;
                    f->in_c[0] = (float)(T / 2.0 * (g1 * T / 2.0 + g2));
                    f->in_c[1] = (float)(T * T / 2.0 * g1);
                    f->in_c[2] = (float)(T / 2.0 * (g1 * T / 2.0 - g2));
                    f->n_out = 2;
                    f->out_c[0] = 2.0f;
                    f->out_c[1] = -1.0f;
                }
            else
                {
                    f->n_in = 2;
                    f->in_c[0] = (float)(g1 * T / 2.0 + g2);
                    f->in_c[1] = (float)(g1 * T / 2.0 - g2);
                    f->n_out = 1;
                    f->out_c[0] = 1.0f;
                }
            break;
        default:
            {
                wn = f->noise_bandwidth / 0.7845f;
                const float a3 = 1.1;
                const float b3 = 2.4;
                g1 = wn * wn * wn;
                g2 = a3 * wn * wn;
                g3 = b3 * wn;
                if (f->include_last_integrator)
                    {
                        f->n_in = 4;
                        f->in_c[0] = (float)(T / 2.0 * (g3 + T / 2.0 * (g2 + T / 2.0 * g1)));
                        f->in_c[1] = (float)(T / 2.0 * (-g3 + T / 2.0 * (g2 + 3.0 * T / 2.0 * g1)));
                        f->in_c[2] = (float)(T / 2.0 * (-g3 - T / 2.0 * (g2 - 3.0 * T / 2.0 * g1)));
                        f->in_c[3] = (float)(T / 2.0 * (g3 - T / 2.0 * (g2 - T / 2.0 * g1)));
                        f->n_out = 3;
                        f->out_c[0] = 3.0f;
                        f->out_c[1] = -3.0f;
                        f->out_c[2] = 1.0f;
                    }
                else
                    {
                        f->n_in = 3;
                        f->in_c[0] = (float)(g3 + T / 2.0 * (g2 + T / 2.0 * g1));
                        f->in_c[1] = (float)(g1 * T * T / 2.0 - 2.0 * g3);
                        f->in_c[2] = (float)(g3 + T / 2.0 * (-g2 + T / 2.0 * g1));
                        f->n_out = 2;
                        f->out_c[0] = 2.0f;
                        f->out_c[1] = -1.0f;
                    }
            }
        }
}

static void loop_filter_initialize(loop_filter_t* f, float initial_output)
{
    for (int i = 0; i < MAX_LOOP_HISTORY_LENGTH; i++)
        {
            f->inputs[i] = 0.0f;
            f->outputs[i] = initial_output;
        }
    f->current_index = MAX_LOOP_HISTORY_LENGTH - 1;
}

static float loop_filter_apply(loop_filter_t* f, float current_input)
{
    float result = 0.0f;
    for (int ii = 0; ii < f->n_out; ++ii) result += f->out_c[ii] * f->outputs[(f->current_index + ii) % MAX_LOOP_HISTORY_LENGTH];
    f->current_index--;
    if (f->current_index < 0) f->current_index += MAX_LOOP_HISTORY_LENGTH;
    f->inputs[f->current_index] = current_input;
    for (int ii = 0; ii < f->n_in; ++ii) result += f->in_c[ii] * f->inputs[(f->current_index + ii) % MAX_LOOP_HISTORY_LENGTH];
    f->outputs[f->current_index] = result;
    return result;
}

typedef struct
{
    float pll_w, pll_w0p3, pll_w0f2, pll_x, pll_a2, pll_w0f, pll_a3, pll_w0p2, pll_b3, pll_w0p;
    int order;
} fll_pll_t;

static void fll_pll_set_params(fll_pll_t* f, float fll_bw_hz, float pll_bw_hz, int order)
{
    f->order = order;
    if (order == 3)
        {
            f->pll_b3 = 2.400;
            f->pll_a3 = 1.100;
            f->pll_a2 = 1.414;
            f->pll_w0p = pll_bw_hz / 0.7845f;
            f->pll_w0p2 = f->pll_w0p * f->pll_w0p;
            f->pll_w0p3 = f->pll_w0p2 * f->pll_w0p;
            f->pll_w0f = fll_bw_hz / 0.53f;
            f->pll_w0f2 = f->pll_w0f * f->pll_w0f;
        }
    else
        {
            f->pll_a2 = 1.414;
            f->pll_w0p = pll_bw_hz / 0.53f;
            f->pll_w0p2 = f->pll_w0p * f->pll_w0p;
            f->pll_w0f = fll_bw_hz / 0.25f;
        }
}

static void fll_pll_initialize(fll_pll_t* f, float acq_carrier_doppler_hz)
{
    if (f->order == 3)
        {
            f->pll_x = 2.0f * acq_carrier_doppler_hz;
            f->pll_w = 0;
        }
    else
        {
            f->pll_w = acq_carrier_doppler_hz;
            f->pll_x = 0;
        }
}

static float fll_pll_get_carrier_error(fll_pll_t* f, float FLL_discriminator, float PLL_discriminator, float correlation_time_s)
{
    float carrier_error_hz;
    if (f->order == 3)
        {
            f->pll_w = f->pll_w + correlation_time_s * (f->pll_w0p3 * PLL_discriminator + f->pll_w0f2 * FLL_discriminator);
            f->pll_x = f->pll_x + correlation_time_s * (0.5f * f->pll_w + f->pll_a2 * f->pll_w0f * FLL_discriminator + f->pll_a3 * f->pll_w0p2 * PLL_discriminator);
            carrier_error_hz = 0.5f * f->pll_x + f->pll_b3 * f->pll_w0p * PLL_discriminator;
        }
    else
        {
            const float pll_w_new = f->pll_w + PLL_discriminator * f->pll_w0p2 * correlation_time_s + FLL_discriminator * f->pll_w0f * correlation_time_s;
            carrier_error_hz = 0.5f * (pll_w_new + f->pll_w) + f->pll_a2 * f->pll_w0p * PLL_discriminator;
            f->pll_w = pll_w_new;
        }
    return carrier_error_hz;
}

/* ---- discriminators ------------------------------------------------------------------------- */
static double phase_unwrap(double phase_rad)
{
    if (phase_rad >= HALF_PI) return phase_rad - GNSS_PI;
    if (phase_rad <= -HALF_PI) return phase_rad + GNSS_PI;
    return phase_rad;
}

static double fll_diff_atan(cf s1, cf s2, double t1, double t2)
{
    double diff_atan = atanf(s2.im / s2.re) - atanf(s1.im / s1.re);
    if (isnan(diff_atan)) diff_atan = 0;
    return phase_unwrap(diff_atan) / (t2 - t1);
}

static double pll_four_quadrant_atan(cf s) { return atan2f(s.im, s.re); } /* gr::fast_atan2f in the reference */

static double pll_cloop_two_quadrant_atan(cf s)
{
    if (s.re != 0.0) return (double)atanf(s.im / s.re);
    return 0.0;
}

static double dll_nc_e_minus_l_normalized(cf e, cf l, float spc, float slope, float y_intercept)
{
    const double P_early = hypotf(e.re, e.im);
    const double P_late = hypotf(l.re, l.im);
    const double E_plus_L = P_early + P_late;
    if (E_plus_L == 0.0) return 0.0;
    return ((y_intercept - slope * spc) / slope) * (P_early - P_late) / E_plus_L;
}

static double dll_nc_vemlp_normalized(cf ve, cf e, cf l, cf vl)
{
    const double Early = sqrtf(ve.re * ve.re + ve.im * ve.im + e.re * e.re + e.im * e.im);
    const double Late = sqrtf(l.re * l.re + l.im * l.im + vl.re * vl.re + vl.im * vl.im);
    const double E_plus_L = Early + Late;
    if (E_plus_L == 0.0) return 0.0;
    return (Early - Late) / E_plus_L;
}

/* ---- lock detectors ------------------------------------------------------------------------- */
static float cn0_m2m4_estimator(const cf* buf, int length, float coh_integration_time_s)
{
    float SNR_aux = 0.0f, SNR_dB_Hz = 0.0f, Psig = 0.0f, m_2 = 0.0f, m_4 = 0.0f, aux;
    const float n = (float)length;
    if (length == 0 || coh_integration_time_s == 0.0) return -100.0f;
    for (int i = 0; i < length; i++)
        {
            Psig += fabsf(buf[i].re);
            aux = buf[i].im * buf[i].im + buf[i].re * buf[i].re;
            m_2 += aux;
            m_4 += (aux * aux);
        }
    Psig /= n;
    Psig = Psig * Psig;
    m_2 /= n;
    m_4 /= n;
    aux = sqrtf(2.0f * m_2 * m_2 - m_4);
    float denominator;
    if (isnan(aux))
        {
            denominator = m_2 - Psig;
            if (denominator == 0) return -100.0f;
            SNR_aux = Psig / denominator;
        }
    else
        {
            denominator = m_2 - aux;
            if (denominator == 0) return -100.0f;
            SNR_aux = aux / denominator;
        }
    if (SNR_aux == 0) return -100.0f;
    SNR_dB_Hz = 10.0f * log10f(SNR_aux) - 10.0f * log10f(coh_integration_time_s);
    return SNR_dB_Hz;
}

static float carrier_lock_detector(const cf* buf, int length)
{
    float tmp_sum_I = 0.0f, tmp_sum_Q = 0.0f, NBD, NBP;
    for (int i = 0; i < length; i++)
        {
            tmp_sum_I += buf[i].re;
            tmp_sum_Q += buf[i].im;
        }
    NBP = tmp_sum_I * tmp_sum_I + tmp_sum_Q * tmp_sum_Q;
    NBD = tmp_sum_I * tmp_sum_I - tmp_sum_Q * tmp_sum_Q;
    if (NBP == 0) return 0.0f;
    return NBD / NBP;
}

/* ---- the block ------------------------------------------------------------------------------ */
typedef struct port_loop
{
    b200_trk_loop_conf c;
    smoother_t cn0_smoother, carrier_lock_test_smoother;
    loop_filter_t code_loop_filter;
    fll_pll_t carrier_loop_filter;
    cf Prompt_buffer[MAX_CN0_SAMPLES];
    double acq_code_phase_samples, acq_carrier_doppler_hz, current_correlation_time_s;
    double carr_phase_error_hz, carr_freq_error_hz, carr_error_filt_hz, code_error_chips, code_error_filt_chips;
    double code_freq_chips, carrier_doppler_hz, acc_carrier_phase_rad, rem_code_phase_chips;
    double T_chip_seconds, T_prn_seconds, T_prn_samples, K_blk_samples;
    double carrier_lock_test, CN0_SNV_dB_Hz, carrier_lock_threshold;
    double carrier_phase_step_rad, carrier_phase_rate_step_rad, code_phase_step_chips, code_phase_rate_step_chips;
    double rem_code_phase_samples;
    cf VE_accu, E_accu, P_accu, P_accu_old, L_accu, VL_accu, Prompt;
    uint64_t acq_sample_stamp, nitems_read;
    float rem_carr_phase_rad, spc;
    int32_t state, current_prn_length_samples, cn0_estimation_counter, carrier_lock_fail_counter, code_lock_fail_counter;
    int pull_in_transitory, cloop, loss_of_lock;
    uint64_t epochs;
} port_loop;

static void clear_tracking_vars(port_loop* L)
{
    L->P_accu_old.re = L->P_accu_old.im = 0.0f;
    L->carr_phase_error_hz = 0.0;
    L->carr_freq_error_hz = 0.0;
    L->carr_error_filt_hz = 0.0;
    L->code_error_chips = 0.0;
    L->code_error_filt_chips = 0.0;
    L->carrier_phase_rate_step_rad = 0.0;
    L->code_phase_rate_step_chips = 0.0;
}

port_loop* port_loop_create(const b200_trk_loop_conf* conf)
{
    if (conf->cn0_samples > MAX_CN0_SAMPLES || conf->cn0_samples < 1) return NULL;
    port_loop* L = (port_loop*)calloc(1, sizeof(port_loop));
    L->c = *conf;
    /* constructor :97-146, :601-605, :680-694 */
    L->carrier_lock_threshold = conf->carrier_lock_th;
    L->code_freq_chips = conf->code_chip_rate;
    L->code_loop_filter.update_interval = (float)conf->code_period;
    L->code_loop_filter.noise_bandwidth = conf->dll_bw_hz;
    L->code_loop_filter.order = conf->dll_filter_order;
    L->code_loop_filter.include_last_integrator = 0;
    loop_filter_update_coefficients(&L->code_loop_filter);
    fll_pll_set_params(&L->carrier_loop_filter, conf->fll_bw_hz, conf->pll_bw_hz, conf->pll_filter_order);
    smoother_defaults(&L->cn0_smoother);
    smoother_set_alpha(&L->cn0_smoother, conf->cn0_smoother_alpha);
    if (conf->code_period > 0.0) smoother_set_samples(&L->cn0_smoother, conf->cn0_smoother_samples / (int)(conf->code_period * 1000.0));
    smoother_defaults(&L->carrier_lock_test_smoother);
    smoother_set_alpha(&L->carrier_lock_test_smoother, conf->carrier_lock_test_smoother_alpha);
    L->carrier_lock_test_smoother.min_value = -1.0f;
    L->carrier_lock_test_smoother.offset = 0.0f;
    smoother_set_samples(&L->carrier_lock_test_smoother, conf->carrier_lock_test_smoother_samples);
    L->spc = conf->early_late_space_chips;
    L->pull_in_transitory = 1;
    L->cloop = 1;
    clear_tracking_vars(L);
    return L;
}

void port_loop_destroy(port_loop* L) { free(L); }

void port_loop_start(port_loop* L, double acq_delay_samples, double acq_doppler_hz, uint64_t acq_samplestamp, uint64_t nitems_read)
{
    const b200_trk_loop_conf* c = &L->c;
    L->acq_code_phase_samples = acq_delay_samples;
    L->acq_carrier_doppler_hz = acq_doppler_hz;
    L->acq_sample_stamp = acq_samplestamp;
    L->nitems_read = nitems_read;
    L->carrier_doppler_hz = L->acq_carrier_doppler_hz;
    L->carrier_phase_step_rad = TWO_PI * L->carrier_doppler_hz / c->fs_in;
    L->carrier_phase_rate_step_rad = 0.0;
    L->carrier_lock_fail_counter = 0;
    L->code_lock_fail_counter = 0;
    L->rem_code_phase_samples = 0.0;
    L->rem_carr_phase_rad = 0.0f;
    L->rem_code_phase_chips = 0.0;
    L->acc_carrier_phase_rad = 0.0;
    L->cn0_estimation_counter = 0;
    L->carrier_lock_test = 1.0;
    L->CN0_SNV_dB_Hz = 0.0;
    L->current_correlation_time_s = c->code_period;
    fll_pll_set_params(&L->carrier_loop_filter, c->fll_bw_hz, c->pll_bw_hz, c->pll_filter_order);
    L->code_loop_filter.noise_bandwidth = c->dll_bw_hz;
    loop_filter_update_coefficients(&L->code_loop_filter);
    L->code_loop_filter.update_interval = (float)c->code_period;
    loop_filter_update_coefficients(&L->code_loop_filter);
    fll_pll_initialize(&L->carrier_loop_filter, (float)L->acq_carrier_doppler_hz);
    loop_filter_initialize(&L->code_loop_filter, 0.0f);
    L->state = 1;
    L->cloop = c->cloop ? 1 : 0; /* the block sets d_cloop = true here; pilot tracking reaches false via state 4 */
    L->pull_in_transitory = 1;
    L->loss_of_lock = 0;
    L->epochs = 0;
}

static void pull_in_check(port_loop* L)
{
    if (L->pull_in_transitory)
        {
            if (L->c.pull_in_time_s < (L->nitems_read - L->acq_sample_stamp) / (uint64_t)(int)L->c.fs_in)
                {
                    L->pull_in_transitory = 0;
                    L->carrier_lock_fail_counter = 0;
                    L->code_lock_fail_counter = 0;
                }
        }
}

/* The part of general_work that precedes do_correlation_step.  Returns 0 in standby, else 1 with the epoch's
 * first sample and the six scalars do_correlation_step casts to float (:1237-1244). */
int port_loop_prepare(port_loop* L, uint64_t* sample_index, int32_t* n, float* p6)
{
    const b200_trk_loop_conf* c = &L->c;
    if (L->state == 0) return 0;
    pull_in_check(L);
    if (L->state == 1)
        {
            const int64_t acq_trk_diff_samples = (int64_t)L->nitems_read - (int64_t)L->acq_sample_stamp;
            const double delta_trk_to_acq_prn_start_samples = (double)acq_trk_diff_samples - L->acq_code_phase_samples;
            L->code_freq_chips = c->code_chip_rate;
            L->code_phase_step_chips = L->code_freq_chips / c->fs_in;
            L->code_phase_rate_step_chips = 0.0;
            const double T_chip_mod_seconds = 1.0 / L->code_freq_chips;
            const double T_prn_mod_seconds = T_chip_mod_seconds * (double)c->code_length_chips;
            const double T_prn_mod_samples = T_prn_mod_seconds * c->fs_in;
            L->acq_code_phase_samples = T_prn_mod_samples - fmod(delta_trk_to_acq_prn_start_samples, T_prn_mod_samples);
            L->current_prn_length_samples = (int32_t)round(T_prn_mod_samples);
            const int32_t samples_offset = (int32_t)round(L->acq_code_phase_samples);
            L->acc_carrier_phase_rad -= L->carrier_phase_step_rad * (double)samples_offset;
            L->state = 2;
            smoother_reset(&L->cn0_smoother);
            smoother_reset(&L->carrier_lock_test_smoother);
            L->nitems_read += (uint64_t)(int64_t)samples_offset; /* consume_each(samples_offset) */
            pull_in_check(L);                                      /* head of the next general_work call */
        }
    *sample_index = L->nitems_read;
    *n = (int32_t)c->vector_length;
    p6[0] = L->rem_carr_phase_rad;
    p6[1] = (float)L->carrier_phase_step_rad;
    p6[2] = (float)L->carrier_phase_rate_step_rad;
    p6[3] = (float)L->rem_code_phase_chips * (float)c->code_samples_per_chip;
    p6[4] = (float)L->code_phase_step_chips * (float)c->code_samples_per_chip;
    p6[5] = (float)L->code_phase_rate_step_chips * (float)c->code_samples_per_chip;
    return 1;
}

static int cn0_and_tracking_lock_status(port_loop* L, double coh_integration_time_s)
{
    const b200_trk_loop_conf* c = &L->c;
    if (L->cn0_estimation_counter < c->cn0_samples)
        {
            L->Prompt_buffer[L->cn0_estimation_counter] = L->P_accu;
            L->cn0_estimation_counter++;
            return 1;
        }
    L->Prompt_buffer[L->cn0_estimation_counter % c->cn0_samples] = L->P_accu;
    L->cn0_estimation_counter++;
    const float CN0_raw = cn0_m2m4_estimator(L->Prompt_buffer, c->cn0_samples, (float)coh_integration_time_s);
    L->CN0_SNV_dB_Hz = smoother_smooth(&L->cn0_smoother, CN0_raw);
    L->carrier_lock_test = smoother_smooth(&L->carrier_lock_test_smoother, carrier_lock_detector(L->Prompt_buffer, 1));
    if (!L->pull_in_transitory)
        {
            if (L->carrier_lock_test < L->carrier_lock_threshold)
                L->carrier_lock_fail_counter++;
            else if (L->carrier_lock_fail_counter > 0)
                L->carrier_lock_fail_counter--;
            if (L->CN0_SNV_dB_Hz < c->cn0_min)
                L->code_lock_fail_counter++;
            else if (L->code_lock_fail_counter > 0)
                L->code_lock_fail_counter--;
        }
    if (L->carrier_lock_fail_counter > c->max_carrier_lock_fail || L->code_lock_fail_counter > c->max_code_lock_fail)
        {
            L->carrier_lock_fail_counter = 0;
            L->code_lock_fail_counter = 0;
            return 0;
        }
    return 1;
}

static void run_dll_pll(port_loop* L)
{
    const b200_trk_loop_conf* c = &L->c;
    if (L->cloop)
        L->carr_phase_error_hz = pll_cloop_two_quadrant_atan(L->P_accu) / TWO_PI;
    else
        L->carr_phase_error_hz = pll_four_quadrant_atan(L->P_accu) / TWO_PI;
    if ((L->pull_in_transitory && c->enable_fll_pull_in) || c->enable_fll_steady_state)
        {
            L->carr_freq_error_hz = fll_diff_atan(L->P_accu_old, L->P_accu, 0, L->current_correlation_time_s) / TWO_PI;
            L->P_accu_old = L->P_accu;
            if (L->pull_in_transitory && c->enable_fll_pull_in)
                L->carr_error_filt_hz = fll_pll_get_carrier_error(&L->carrier_loop_filter, (float)L->carr_freq_error_hz, 0.0f, (float)L->current_correlation_time_s);
            else
                L->carr_error_filt_hz = fll_pll_get_carrier_error(&L->carrier_loop_filter, (float)L->carr_freq_error_hz, (float)L->carr_phase_error_hz, (float)L->current_correlation_time_s);
        }
    else
        {
            L->carr_error_filt_hz = fll_pll_get_carrier_error(&L->carrier_loop_filter, 0, (float)L->carr_phase_error_hz, (float)L->current_correlation_time_s);
        }
    L->carrier_doppler_hz = L->carr_error_filt_hz;
    if (c->veml)
        L->code_error_chips = dll_nc_vemlp_normalized(L->VE_accu, L->E_accu, L->L_accu, L->VL_accu);
    else
        L->code_error_chips = dll_nc_e_minus_l_normalized(L->E_accu, L->L_accu, L->spc, c->slope, c->y_intercept);
    L->code_error_filt_chips = loop_filter_apply(&L->code_loop_filter, (float)L->code_error_chips);
    L->code_freq_chips = c->code_chip_rate - L->code_error_filt_chips;
    if (c->carrier_aiding) L->code_freq_chips += L->carrier_doppler_hz * c->code_chip_rate / c->signal_carrier_freq;
}

static void update_tracking_vars(port_loop* L)
{
    const b200_trk_loop_conf* c = &L->c;
    L->T_chip_seconds = 1.0 / L->code_freq_chips;
    L->T_prn_seconds = L->T_chip_seconds * (double)(int32_t)c->code_length_chips;
    L->T_prn_samples = L->T_prn_seconds * c->fs_in;
    L->K_blk_samples = L->T_prn_samples + L->rem_code_phase_samples;
    L->current_prn_length_samples = (int32_t)floor(L->K_blk_samples);
    L->carrier_phase_step_rad = TWO_PI * (L->carrier_doppler_hz + 0.0 /* d_cfo_frequency_hz */) / c->fs_in;
    L->rem_carr_phase_rad += (float)(L->carrier_phase_step_rad * (double)L->current_prn_length_samples + 0.5 * L->carrier_phase_rate_step_rad * (double)L->current_prn_length_samples * (double)L->current_prn_length_samples);
    L->rem_carr_phase_rad = (float)fmod(L->rem_carr_phase_rad, TWO_PI);
    L->acc_carrier_phase_rad -= (L->carrier_phase_step_rad * (double)L->current_prn_length_samples + 0.5 * L->carrier_phase_rate_step_rad * (double)L->current_prn_length_samples * (double)L->current_prn_length_samples);
    L->code_phase_step_chips = L->code_freq_chips / c->fs_in;
    L->rem_code_phase_samples = L->K_blk_samples - (double)L->current_prn_length_samples;
    L->rem_code_phase_chips = L->code_freq_chips * L->rem_code_phase_samples / c->fs_in;
}

static void log_data(const port_loop* L, b200_trk_dump_record* r)
{
    const b200_trk_loop_conf* c = &L->c;
    r->abs_VE = c->veml ? hypotf(L->VE_accu.re, L->VE_accu.im) : 0.0f;
    r->abs_E = hypotf(L->E_accu.re, L->E_accu.im);
    r->abs_P = hypotf(L->P_accu.re, L->P_accu.im);
    r->abs_L = hypotf(L->L_accu.re, L->L_accu.im);
    r->abs_VL = c->veml ? hypotf(L->VL_accu.re, L->VL_accu.im) : 0.0f;
    r->prompt_I = L->Prompt.re;
    r->prompt_Q = L->Prompt.im;
    r->PRN_start_sample_count = L->nitems_read + (uint64_t)L->current_prn_length_samples;
    r->acc_carrier_phase_rad = (float)L->acc_carrier_phase_rad;
    r->carrier_doppler_hz = (float)L->carrier_doppler_hz;
    r->carrier_doppler_rate_hz_s = (float)(L->carrier_phase_rate_step_rad * c->fs_in * c->fs_in / TWO_PI);
    r->code_freq_chips = (float)L->code_freq_chips;
    r->code_freq_rate_chips = (float)(L->code_phase_rate_step_chips * c->fs_in * c->fs_in);
    r->carr_error_hz = (float)L->carr_phase_error_hz;
    r->carr_error_filt_hz = (float)L->carr_error_filt_hz;
    r->code_error_chips = (float)L->code_error_chips;
    r->code_error_filt_chips = (float)L->code_error_filt_chips;
    r->CN0_SNV_dB_Hz = (float)L->CN0_SNV_dB_Hz;
    r->carrier_lock_test = (float)L->carrier_lock_test;
    r->aux1 = (float)L->rem_code_phase_samples;
    r->aux2 = (double)(L->nitems_read + (uint64_t)(int64_t)L->current_prn_length_samples);
    r->PRN = c->prn;
    r->TOW_ms = 0;
    r->WN = 0;
}

/* The part of general_work case 2 that follows do_correlation_step, then consume_each.  taps: E,P,L or
 * VE,E,P,L,VL as (re,im) float pairs.  Returns 1 when a record was logged, 0 on loss of lock. */
int port_loop_update(port_loop* L, const float* taps, b200_trk_dump_record* rec)
{
    const b200_trk_loop_conf* c = &L->c;
    const cf* t = (const cf*)taps;
    if (L->state != 2) return 0;
    if (c->veml)
        {
            L->VE_accu = t[0];
            L->E_accu = t[1];
            L->P_accu = t[2];
            L->L_accu = t[3];
            L->VL_accu = t[4];
        }
    else
        {
            L->E_accu = t[0];
            L->P_accu = t[1];
            L->L_accu = t[2];
        }
    L->Prompt = L->P_accu;
    L->spc = c->early_late_space_chips;
    int logged = 0;
    if (c->bit_synchronization_time_limit_s < (L->nitems_read - L->acq_sample_stamp) / (uint64_t)(int)c->fs_in) L->carrier_lock_fail_counter = 300000;
    if (!cn0_and_tracking_lock_status(L, c->code_period))
        {
            clear_tracking_vars(L);
            L->state = 0;
            L->loss_of_lock = 1;
        }
    else
        {
            run_dll_pll(L);
            update_tracking_vars(L);
            if (rec) log_data(L, rec);
            logged = 1;
            L->epochs++;
        }
    L->nitems_read += (uint64_t)(int64_t)L->current_prn_length_samples; /* consume_each (:2292) */
    return logged;
}

void port_loop_status(const port_loop* L, b200_trk_loop_status* s)
{
    s->state = L->state;
    s->loss_of_lock = L->loss_of_lock;
    s->sample_counter = L->nitems_read;
    s->epochs = L->epochs;
    s->carrier_doppler_hz = L->carrier_doppler_hz;
    s->code_freq_chips = L->code_freq_chips;
    s->rem_code_phase_samples = L->rem_code_phase_samples;
    s->acc_carrier_phase_rad = L->acc_carrier_phase_rad;
    s->CN0_SNV_dB_Hz = L->CN0_SNV_dB_Hz;
    s->carrier_lock_test = L->carrier_lock_test;
}

/* individual library functions, exported so that tests can pin them against the reference's object code */
double port_disc_pll_cloop(float re, float im)
{
    cf s = {re, im};
    return pll_cloop_two_quadrant_atan(s);
}
double port_disc_fll_diff_atan(float re1, float im1, float re2, float im2, double t1, double t2)
{
    cf a = {re1, im1}, b = {re2, im2};
    return fll_diff_atan(a, b, t1, t2);
}
double port_disc_dll_e_minus_l(float er, float ei, float lr, float li, float spc, float slope, float y_intercept)
{
    cf e = {er, ei}, l = {lr, li};
    return dll_nc_e_minus_l_normalized(e, l, spc, slope, y_intercept);
}
double port_disc_dll_vemlp(const float* t8)
{
    const cf* t = (const cf*)t8;
    return dll_nc_vemlp_normalized(t[0], t[1], t[2], t[3]);
}
float port_cn0_m2m4(const float* buf, int length, float T) { return cn0_m2m4_estimator((const cf*)buf, length, T); }
float port_carrier_lock_detector(const float* buf, int length) { return carrier_lock_detector((const cf*)buf, length); }
