/*
 * b200gnss.h -- C ABI of libb200gnss.so: the B200 (sm_100a) implementation of gnss-sdr's two
 * data-parallel hot paths, the multi-tap tracking correlator and the PCPS acquisition grid
 * search.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * gnss-sdr tree; VG = src/algorithms/libs/volk_gnsssdr_module/volk_gnsssdr).
 *
 * Conventions
 *  - All functions return int: B200_OK (0) or a negative B200_ERR_* code; nothing calls
 *    exit() (contrast src/algorithms/tracking/libs/cuda_multicorrelator.cu:294-301).
 *    b200_last_error() returns a thread-local message for the last failure.
 *  - Complex samples are interleaved (re, im) float32 pairs == lv_32fc_t == gr_complex
 *    (VG include/volk_gnsssdr/volk_gnsssdr_complex.h).
 *  - "host" pointers are ordinary (optionally pinned) host memory; "dev" pointers are CUDA
 *    device pointers on the engine's device.  `stream` arguments are cudaStream_t passed as
 *    void* (NULL = the engine's own stream).
 *  - Handles are opaque.  Distinct handles may be used concurrently from different threads
 *    (one tracking block thread per channel, as in the reference's thread-per-block
 *    scheduler); calls on ONE handle must be serialised by the caller (the reference blocks
 *    do that with gr::block::d_setlock).
 */
#ifndef B200GNSS_H
#define B200GNSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)    /* bad argument / handle */
#define B200_ERR_CUDA (-2)   /* a CUDA call failed (message in b200_last_error) */
#define B200_ERR_NOMEM (-3)  /* host or device allocation failed */
#define B200_ERR_STATE (-4)  /* call sequence error (e.g. correlate before set_local_code) */
#define B200_ERR_RANGE (-5)  /* size outside what the engine was created for */
#define B200_ERR_NODEV (-6)  /* no usable sm_100 device */

#define B200_MAX_TAPS 8

    typedef struct b200_cf32
    {
        float re, im;
    } b200_cf32;

    typedef struct b200_engine b200_engine;
    typedef struct b200_trk b200_trk;
    typedef struct b200_acq b200_acq;

    /* ---- library / engine ------------------------------------------------------------- */
    int b200_version(void);
    const char* b200_last_error(void);
    int b200_device_count(int* count);

    /* One engine per (process, GPU): owns the device context, a compute stream, a copy
     * stream, the IQ band stores and the channel registry.  `stream` (cudaStream_t as void*)
     * makes the engine launch on a caller-owned stream instead of its own.
     * Precedent: the FPGA path's device-owned sample stream,
     * src/algorithms/tracking/libs/fpga_multicorrelator.cc:152-192. */
    int b200_engine_create(b200_engine** out, int device, void* stream);
    int b200_engine_destroy(b200_engine* e);
    int b200_engine_sync(b200_engine* e);
    /* device-side timing of work submitted to the engine stream (CUDA events) */
    int b200_engine_timer_start(b200_engine* e);
    int b200_engine_timer_stop_ms(b200_engine* e, float* ms);

    /* ---- IQ band store ---------------------------------------------------------------- */
    /* A band is one conditioned IQ stream (what gnss_flowgraph.cc:1227-1231 fans out to every
     * channel).  Samples are addressed by their absolute index in the stream
     * (== the reference's sample counter, Gnss_Synchro::Acq_samplestamp_samples).
     * capacity_samples is rounded up to a power of two; the store is a ring. */
    int b200_iq_create(b200_engine* e, int band, uint64_t capacity_samples);
    /* Append n host samples (async H2D on the copy stream, ordered before later launches).
     * *first_index receives the absolute index of host[0]. */
    int b200_iq_push(b200_engine* e, int band, const b200_cf32* host, uint64_t n, uint64_t* first_index);
    /* Use caller-owned device memory as the band (no copy): dev[0] is absolute index
     * first_index; n_samples need not be a power of two (no wrap). */
    int b200_iq_attach_dev(b200_engine* e, int band, const b200_cf32* dev, uint64_t n_samples, uint64_t first_index);

    /* ---- tracking: single correlator, 1:1 with Cpu_Multicorrelator_Real_Codes ----------- */
    /* src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.h:37-61 */
    /* init(max_signal_length_samples, n_correlators)                       (.cc:35-50) */
    int b200_trk_create(b200_engine* e, b200_trk** out, int max_signal_length_samples, int n_correlators);
    /* set_high_dynamics_resampler(bool)                                    (.cc:27-31 in .h:41) */
    int b200_trk_set_high_dynamics_resampler(b200_trk* t, int use_high_dynamics_resampler);
    /* set_local_code_and_taps(code_length_chips, local_code_in, shifts_chips) (.cc:53-63).
     * The table and shifts are COPIED (the reference keeps the caller's pointers). */
    int b200_trk_set_local_code_and_taps(b200_trk* t, int code_length_chips, const float* local_code_in, const float* shifts_chips);
    /* set_input_output_vectors + Carrier_wipeoff_multicorrelator_resampler (.cc:66-72,103-127):
     * corr_out[k] = sum_n sig_in[n] * exp(-j(rem_carrier + n*phase_step [+ rate term])) *
     *               code[ floor(step*n + shift_k - rem_code) mod L ].
     * Synchronous: copies sig_in to the device, correlates, copies n_correlators taps back.
     * Arithmetic contract: chip indices bit-exact with VG's a_avx/u_avx resampler
     * (..._32f_xn_resampler_32f_xn.h:362-435); taps within 1e-5*|prompt| of float64. */
    int b200_trk_correlate(b200_trk* t, const b200_cf32* sig_in_host,
        float rem_carrier_phase_in_rad, float phase_step_rad, float phase_rate_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips,
        int signal_length_samples, b200_cf32* corr_out_host);
    /* free()                                                               (.cc:147-160) */
    int b200_trk_destroy(b200_trk* t);

    /* ---- tracking: batched form (all channels in lock, many epochs, one launch) ---------- */
    /* A channel = one (band, code table, tap shifts) registration; the per-epoch scalars are
     * the seven arguments do_correlation_step passes
     * (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:1232-1257). */
    typedef struct b200_trk_item
    {
        int32_t channel;       /* id returned by b200_trk_channel_create */
        int32_t n;             /* signal_length_samples (vector_length) */
        uint64_t sample_index; /* absolute index of the epoch's first sample in the band */
        float rem_carrier_phase_rad;
        float phase_step_rad;
        float phase_rate_step_rad;
        float rem_code_phase_chips;
        float code_phase_step_chips;
        float code_phase_rate_step_chips;
    } b200_trk_item; /* 40 bytes */

    int b200_trk_channel_create(b200_engine* e, int band, int n_correlators, int* channel_id);
    int b200_trk_channel_set_code(b200_engine* e, int channel_id, int code_length_chips, const float* local_code_in, const float* shifts_chips, int high_dynamics);
    /* Correlate n_items (channel, epoch) work items.  out: n_items x out_stride complex taps
     * (out_stride >= the channel's n_correlators).  Host variant is synchronous (items H2D,
     * launch, taps D2H); the _dev variant takes device pointers and is asynchronous on the
     * engine stream. */
    int b200_trk_batch(b200_engine* e, const b200_trk_item* items_host, int n_items, b200_cf32* out_host, int out_stride);
    int b200_trk_batch_dev(b200_engine* e, const b200_trk_item* items_dev, int n_items, b200_cf32* out_dev, int out_stride, int slices);
    /* number of kernel launches issued by this engine so far (bench.py's gpu_launches) */
    int b200_engine_launch_count(b200_engine* e, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* B200GNSS_H */
