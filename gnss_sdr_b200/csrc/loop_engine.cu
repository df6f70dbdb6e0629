// Host side of the free-running DLL/PLL loops (include/b200gnss.h, "tracking: free-running DLL/PLL loops").
// Configuration-time arithmetic only: filter coefficients and start-of-tracking state; the per-epoch cycle is
// loop_kernels_nofma.cu.
//
//   Tracking_loop_filter::update_coefficients   src/algorithms/tracking/libs/tracking_loop_filter.cc:99-196
//   Tracking_FLL_PLL_filter::set_params         src/algorithms/tracking/libs/tracking_FLL_PLL_filter.cc:23-54
//   Tracking_FLL_PLL_filter::initialize         :57-69
//   dll_pll_veml_tracking ctor / start_tracking src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:97-146,601-605,680-694,791-1078
//   Exponential_Smoother defaults               src/algorithms/tracking/libs/exponential_smoother.h:54-62
#include "engine.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace b200;

static_assert(sizeof(b200_trk_dump_record) == 108, "dump record must match log_data()");
static_assert(sizeof(b200_trk_loop_conf) == 152, "b200_trk_loop_conf layout");

namespace
{
constexpr double kTwoPi = 6.283185307179586;

void code_filter_coefficients(LoopDev& L, float update_interval, float noise_bandwidth, int order)
{
    // include_last_integrator = false (dll_pll_veml_tracking.cc:604)
    float g1, g2, g3, wn;
    const float T = update_interval;
    const float zeta = 1.0F / std::sqrt(2.0F);
    switch (order)
        {
        case 1:
            wn = noise_bandwidth * 4.0F;
            g1 = wn;
            L.dll_n_in = 1;
            L.dll_in_c[0] = g1;
            L.dll_n_out = 0;
            break;
        case 2:
            wn = noise_bandwidth * (8.0F * zeta) / (4.0F * zeta * zeta + 1.0F);
            g1 = wn * wn;
            g2 = wn * 2.0F * zeta;
            L.dll_n_in = 2;
            L.dll_in_c[0] = static_cast<float>(g1 * T / 2.0 + g2);
            L.dll_in_c[1] = static_cast<float>(g1 * T / 2.0 - g2);
            L.dll_n_out = 1;
            L.dll_out_c[0] = 1.0F;
            break;
        default:
            {
                wn = noise_bandwidth / 0.7845F;
                const float a3 = 1.1;
                const float b3 = 2.4;
                g1 = wn * wn * wn;
                g2 = a3 * wn * wn;
                g3 = b3 * wn;
                L.dll_n_in = 3;
                L.dll_in_c[0] = static_cast<float>(g3 + T / 2.0 * (g2 + T / 2.0 * g1));
                L.dll_in_c[1] = static_cast<float>(g1 * T * T / 2.0 - 2.0 * g3);
                L.dll_in_c[2] = static_cast<float>(g3 + T / 2.0 * (-g2 + T / 2.0 * g1));
                L.dll_n_out = 2;
                L.dll_out_c[0] = 2.0F;
                L.dll_out_c[1] = -1.0F;
            }
        }
}

void carrier_filter_params(LoopDev& L, float fll_bw_hz, float pll_bw_hz, int order)
{
    L.pll_order = order;
    if (order == 3)
        {
            L.pll_b3 = 2.400;
            L.pll_a3 = 1.100;
            L.pll_a2 = 1.414;
            L.pll_w0p = pll_bw_hz / 0.7845F;
            L.pll_w0p2 = L.pll_w0p * L.pll_w0p;
            L.pll_w0p3 = L.pll_w0p2 * L.pll_w0p;
            L.pll_w0f = fll_bw_hz / 0.53F;
            L.pll_w0f2 = L.pll_w0f * L.pll_w0f;
        }
    else
        {
            L.pll_a2 = 1.414;
            L.pll_w0p = pll_bw_hz / 0.53F;
            L.pll_w0p2 = L.pll_w0p * L.pll_w0p;
            L.pll_w0f = fll_bw_hz / 0.25F;
        }
}

void smoother_init(LoopSmoother& s, float alpha, float min_value, float offset, int samples)
{
    s.alpha = alpha;
    if (s.alpha < 0) s.alpha = 0;
    if (s.alpha > 1) s.alpha = 1;
    s.one_minus_alpha = 1.0F - s.alpha;
    s.old_value = 0.0F;
    s.min_value = min_value;
    s.offset = offset;
    s.init_sum = 0.0F;
    s.samples_for_initialization = samples <= 0 ? 1 : samples;
    s.init_counter = 0;
    s.init_n = 0;
    s.initializing = 1;
}

int ensure_loop_buffers(b200_engine* e, int rec_capacity)
{
    const int n = static_cast<int>(e->loops.size());
    if (n > e->loops_dev_cap)
        {
            const int cap = n + 32;
            LoopDev* nd = nullptr;
            B200_CUDA_TRY(cudaMalloc(&nd, sizeof(LoopDev) * cap));
            B200_CUDA_TRY(cudaMemsetAsync(nd, 0, sizeof(LoopDev) * cap, e->stream));
            if (e->loops_dev && e->loops_dev_cap > 0)
                B200_CUDA_TRY(cudaMemcpyAsync(nd, e->loops_dev, sizeof(LoopDev) * e->loops_dev_cap, cudaMemcpyDeviceToDevice, e->stream));
            B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
            if (e->loops_dev) B200_CUDA_TRY(cudaFree(e->loops_dev));
            if (e->loop_items_dev) B200_CUDA_TRY(cudaFree(e->loop_items_dev));
            if (e->loop_taps_dev) B200_CUDA_TRY(cudaFree(e->loop_taps_dev));
            if (e->loop_nrec_dev) B200_CUDA_TRY(cudaFree(e->loop_nrec_dev));
            e->loops_dev = nd;
            e->loops_dev_cap = cap;
            B200_CUDA_TRY(cudaMalloc(&e->loop_items_dev, sizeof(b200_trk_item) * cap));
            B200_CUDA_TRY(cudaMalloc(&e->loop_taps_dev, sizeof(float2) * kLoopTapStride * cap));
            B200_CUDA_TRY(cudaMalloc(&e->loop_nrec_dev, sizeof(int) * cap));
            B200_CUDA_TRY(cudaMemsetAsync(e->loop_items_dev, 0, sizeof(b200_trk_item) * cap, e->stream));
            B200_CUDA_TRY(cudaMemsetAsync(e->loop_taps_dev, 0, sizeof(float2) * kLoopTapStride * cap, e->stream));
        }
    const size_t words = static_cast<size_t>(n) * static_cast<size_t>(rec_capacity > 0 ? rec_capacity : 1) * kLoopRecordWords;
    if (words > e->loop_rec_cap)
        {
            if (e->loop_rec_dev) B200_CUDA_TRY(cudaFree(e->loop_rec_dev));
            e->loop_rec_cap = words + words / 4;
            B200_CUDA_TRY(cudaMalloc(&e->loop_rec_dev, sizeof(unsigned int) * e->loop_rec_cap));
        }
    return B200_OK;
}

LoopAvail snapshot_avail(const b200_engine* e)
{
    LoopAvail a{};
    for (int i = 0; i < kMaxBands; i++)
        {
            const Band& b = e->bands[i];
            if (!b.in_use) continue;
            if (b.attached)
                {
                    a.lo[i] = b.first_index;
                    a.hi[i] = b.first_index + b.capacity;
                }
            else
                {
                    a.hi[i] = b.write_index;
                    a.lo[i] = (b.write_index - b.first_index > b.capacity) ? b.write_index - b.capacity : b.first_index;
                }
        }
    return a;
}

int batch_slices(int n_items)
{
    int slices = 1;
    if (n_items < 592) slices = (592 + n_items - 1) / n_items;
    if (slices > 64) slices = 64;
    return slices;
}

int correlate_loop_items(b200_engine* e, int n, int slices)
{
    int rc = launch_trk_batch(e->loop_items_dev, n, e->chans_dev, e->bands_dev, e->loop_taps_dev, kLoopTapStride, slices, e->partial,
        e->counters, e->max_code_len, e->taps_uniform, e->stream, e->taps_uniform == 0 ? e->taps_mask : 0u);
    if (rc == B200_OK) e->launches++;
    return rc;
}
}  // namespace

namespace b200
{
void loops_free(b200_engine* e)
{
    if (e->loops_dev) cudaFree(e->loops_dev);
    if (e->loop_items_dev) cudaFree(e->loop_items_dev);
    if (e->loop_taps_dev) cudaFree(e->loop_taps_dev);
    if (e->loop_nrec_dev) cudaFree(e->loop_nrec_dev);
    if (e->loop_rec_dev) cudaFree(e->loop_rec_dev);
}
}  // namespace b200

extern "C"
{
    int b200_trk_loop_create(b200_engine* e, int channel, const b200_trk_loop_conf* conf, int* loop_id)
    {
        if (!e || !conf || !loop_id) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        if (channel < 0 || channel >= static_cast<int>(e->chans.size()) || e->chans[channel].desc.code == nullptr)
            {
                set_error("loop_create: channel %d has no code table", channel);
                return B200_ERR_STATE;
            }
        const int need_taps = conf->veml ? 5 : 3;
        if (e->chans[channel].desc.taps != need_taps)
            {
                set_error("loop_create: channel %d has %d correlators, the loop needs %d", channel, e->chans[channel].desc.taps, need_taps);
                return B200_ERR_ARG;
            }
        if (e->chans[channel].desc.high_dyn)
            {
                set_error("loop_create: high-dynamics channels are not supported by the device loop");
                return B200_ERR_ARG;
            }
        if (conf->cn0_samples < 1 || conf->cn0_samples > kLoopMaxCn0Samples || conf->fs_in < 1.0 || conf->vector_length < 1 ||
            conf->code_length_chips < 1 || conf->code_chip_rate <= 0.0 || conf->pll_filter_order < 2 || conf->pll_filter_order > 3 ||
            conf->dll_filter_order < 1 || conf->dll_filter_order > 3 || conf->code_samples_per_chip < 1 || conf->slope == 0.0F)
            {
                set_error("loop_create: configuration out of range");
                return B200_ERR_RANGE;
            }
        LoopDev L;
        std::memset(&L, 0, sizeof(L));
        L.c = *conf;
        L.channel = channel;
        L.taps = need_taps;
        L.band = e->chans[channel].desc.band;
        L.carrier_lock_threshold = conf->carrier_lock_th;
        L.code_freq_chips = conf->code_chip_rate;
        code_filter_coefficients(L, static_cast<float>(conf->code_period), conf->dll_bw_hz, conf->dll_filter_order);
        carrier_filter_params(L, conf->fll_bw_hz, conf->pll_bw_hz, conf->pll_filter_order);
        int cn0_init = 200;
        if (conf->code_period > 0.0)
            {
                const int ms = static_cast<int>(conf->code_period * 1000.0);
                if (ms < 1)
                    {
                        set_error("loop_create: code_period below 1 ms");
                        return B200_ERR_RANGE;
                    }
                cn0_init = conf->cn0_smoother_samples / ms;
            }
        smoother_init(L.cn0_smoother, conf->cn0_smoother_alpha, 25.0F, 12.0F, cn0_init);
        smoother_init(L.carrier_lock_test_smoother, conf->carrier_lock_test_smoother_alpha, -1.0F, 0.0F, conf->carrier_lock_test_smoother_samples);
        L.spc = conf->early_late_space_chips;
        L.pull_in_transitory = 1;
        L.cloop = 1;
        L.state = 0;
        e->loops.push_back(L);
        *loop_id = static_cast<int>(e->loops.size()) - 1;
        B200_ENTER_DEVICE(e->device);
        int rc = ensure_loop_buffers(e, 1);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(e->loops_dev + *loop_id, &e->loops[*loop_id], sizeof(LoopDev), cudaMemcpyHostToDevice, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200_OK;
    }

    int b200_trk_loop_start(b200_engine* e, int loop_id, double acq_delay_samples, double acq_doppler_hz, uint64_t acq_samplestamp_samples,
        uint64_t nitems_read)
    {
        if (!e) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        if (loop_id < 0 || loop_id >= static_cast<int>(e->loops.size())) return B200_ERR_ARG;
        LoopDev& L = e->loops[loop_id];
        const b200_trk_loop_conf& c = L.c;
        L.acq_code_phase_samples = acq_delay_samples;
        L.acq_carrier_doppler_hz = acq_doppler_hz;
        L.acq_sample_stamp = acq_samplestamp_samples;
        L.nitems_read = nitems_read;
        L.carrier_doppler_hz = L.acq_carrier_doppler_hz;
        L.carrier_phase_step_rad = kTwoPi * L.carrier_doppler_hz / c.fs_in;
        L.carrier_phase_rate_step_rad = 0.0;
        L.carrier_lock_fail_counter = 0;
        L.code_lock_fail_counter = 0;
        L.rem_code_phase_samples = 0.0;
        L.rem_carr_phase_rad = 0.0F;
        L.rem_code_phase_chips = 0.0;
        L.acc_carrier_phase_rad = 0.0;
        L.cn0_estimation_counter = 0;
        L.carrier_lock_test = 1.0;
        L.CN0_SNV_dB_Hz = 0.0;
        L.current_correlation_time_s = c.code_period;
        carrier_filter_params(L, c.fll_bw_hz, c.pll_bw_hz, c.pll_filter_order);
        code_filter_coefficients(L, static_cast<float>(c.code_period), c.dll_bw_hz, c.dll_filter_order);
        // d_carrier_loop_filter.initialize(acq doppler); d_code_loop_filter.initialize()
        if (L.pll_order == 3)
            {
                L.pll_x = 2.0F * static_cast<float>(L.acq_carrier_doppler_hz);
                L.pll_w = 0;
            }
        else
            {
                L.pll_w = static_cast<float>(L.acq_carrier_doppler_hz);
                L.pll_x = 0;
            }
        for (int i = 0; i < 4; i++)
            {
                L.dll_inputs[i] = 0.0F;
                L.dll_outputs[i] = 0.0F;
            }
        L.dll_index = 3;
        L.state = 1;
        L.cloop = c.cloop ? 1 : 0;
        L.pull_in_transitory = 1;
        L.loss_of_lock = 0;
        L.epochs = 0;
        L.pending = 0;
        // the device copy is replaced by the start-of-tracking state (what start_tracking leaves alone -
        // d_P_accu_old, the error terms - was zeroed by clear_tracking_vars when the previous run ended)
        B200_ENTER_DEVICE(e->device);
        int rc = ensure_loop_buffers(e, 1);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(e->loops_dev + loop_id, &L, sizeof(LoopDev), cudaMemcpyHostToDevice, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200_OK;
    }

    int b200_trk_loop_run(b200_engine* e, int max_epochs, b200_trk_dump_record* records_host, int* n_records_host)
    {
        if (!e || max_epochs < 0) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        const int n = static_cast<int>(e->loops.size());
        if (n == 0 || max_epochs == 0) return B200_OK;
        B200_ENTER_DEVICE(e->device);
        int rc = upload_tables(e);
        if (rc) return rc;
        rc = ensure_loop_buffers(e, max_epochs);
        if (rc) return rc;
        if (e->loop_mode < 0)
            {
                const char* env = std::getenv("B200_LOOP_MODE");
                e->loop_mode = env ? std::atoi(env) : 0;
            }
        const LoopAvail avail = snapshot_avail(e);
        B200_CUDA_TRY(cudaMemsetAsync(e->loop_nrec_dev, 0, sizeof(int) * n, e->stream));
        if (e->loop_mode == 0)
            {
                rc = launch_loop_persistent(e->loops_dev, n, max_epochs, avail, e->chans_dev, e->bands_dev, e->loop_rec_dev, max_epochs,
                    e->loop_nrec_dev, e->max_code_len, e->stream);
                if (rc) return rc;
                e->launches++;
            }
        else
            {
                const int slices = e->loop_mode == 1 ? 1 : batch_slices(n);
                rc = ensure_partials(e, n, slices);
                if (rc) return rc;
                rc = launch_loop_cycle(e->loops_dev, n, kLoopPrepare | kLoopCheckAvail, avail, e->loop_items_dev, e->loop_taps_dev, nullptr, 0,
                    nullptr, e->stream);
                if (rc) return rc;
                e->launches++;
                for (int k = 0; k < max_epochs; k++)
                    {
                        rc = correlate_loop_items(e, n, slices);
                        if (rc) return rc;
                        rc = launch_loop_cycle(e->loops_dev, n, kLoopUpdate | kLoopPrepare | kLoopCheckAvail, avail, e->loop_items_dev,
                            e->loop_taps_dev, e->loop_rec_dev, max_epochs, e->loop_nrec_dev, e->stream);
                        if (rc) return rc;
                        e->launches++;
                    }
            }
        std::vector<int> nrec(n);
        B200_CUDA_TRY(cudaMemcpyAsync(nrec.data(), e->loop_nrec_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
        if (records_host)
            B200_CUDA_TRY(cudaMemcpyAsync(records_host, e->loop_rec_dev, sizeof(b200_trk_dump_record) * static_cast<size_t>(n) * max_epochs,
                cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        if (n_records_host)
            for (int i = 0; i < n; i++) n_records_host[i] = nrec[i] < max_epochs ? nrec[i] : max_epochs;
        return B200_OK;
    }

    int b200_trk_loop_set_mode(b200_engine* e, int mode)
    {
        if (!e || mode < 0 || mode > 2) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        e->loop_mode = mode;
        return B200_OK;
    }

    int b200_trk_loop_peek_items(b200_engine* e, b200_trk_item* items_host)
    {
        if (!e || !items_host) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        const int n = static_cast<int>(e->loops.size());
        if (n == 0) return B200_OK;
        B200_ENTER_DEVICE(e->device);
        int rc = ensure_loop_buffers(e, 1);
        if (rc) return rc;
        const LoopAvail avail{};
        rc = launch_loop_cycle(e->loops_dev, n, kLoopPrepare, avail, e->loop_items_dev, e->loop_taps_dev, nullptr, 0, nullptr, e->stream);
        if (rc) return rc;
        e->launches++;
        B200_CUDA_TRY(cudaMemcpyAsync(items_host, e->loop_items_dev, sizeof(b200_trk_item) * n, cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200_OK;
    }

    int b200_trk_loop_step_taps(b200_engine* e, const b200_cf32* taps_host, b200_trk_dump_record* records_host, int* logged_host)
    {
        if (!e || !taps_host) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        const int n = static_cast<int>(e->loops.size());
        if (n == 0) return B200_OK;
        B200_ENTER_DEVICE(e->device);
        int rc = ensure_loop_buffers(e, 1);
        if (rc) return rc;
        const LoopAvail avail{};
        B200_CUDA_TRY(cudaMemsetAsync(e->loop_nrec_dev, 0, sizeof(int) * n, e->stream));
        rc = launch_loop_cycle(e->loops_dev, n, kLoopPrepare, avail, e->loop_items_dev, e->loop_taps_dev, nullptr, 0, nullptr, e->stream);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(e->loop_taps_dev, taps_host, sizeof(float2) * kLoopTapStride * n, cudaMemcpyHostToDevice, e->stream));
        rc = launch_loop_cycle(e->loops_dev, n, kLoopUpdate, avail, e->loop_items_dev, e->loop_taps_dev, e->loop_rec_dev, 1, e->loop_nrec_dev,
            e->stream);
        if (rc) return rc;
        e->launches += 2;
        if (records_host)
            B200_CUDA_TRY(cudaMemcpyAsync(records_host, e->loop_rec_dev, sizeof(b200_trk_dump_record) * n, cudaMemcpyDeviceToHost, e->stream));
        if (logged_host) B200_CUDA_TRY(cudaMemcpyAsync(logged_host, e->loop_nrec_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200_OK;
    }

    int b200_trk_loop_status_get(b200_engine* e, int loop_id, b200_trk_loop_status* out)
    {
        if (!e || !out) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        if (loop_id < 0 || loop_id >= static_cast<int>(e->loops.size())) return B200_ERR_ARG;
        B200_ENTER_DEVICE(e->device);
        LoopDev L;
        B200_CUDA_TRY(cudaMemcpyAsync(&L, e->loops_dev + loop_id, sizeof(LoopDev), cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        out->state = L.state;
        out->loss_of_lock = L.loss_of_lock;
        out->sample_counter = L.nitems_read;
        out->epochs = L.epochs;
        out->carrier_doppler_hz = L.carrier_doppler_hz;
        out->code_freq_chips = L.code_freq_chips;
        out->rem_code_phase_samples = L.rem_code_phase_samples;
        out->acc_carrier_phase_rad = L.acc_carrier_phase_rad;
        out->CN0_SNV_dB_Hz = L.CN0_SNV_dB_Hz;
        out->carrier_lock_test = L.carrier_lock_test;
        return B200_OK;
    }

    int b200_trk_dump_write(const char* filename, const b200_trk_dump_record* records, int n_records, int append)
    {
        if (!filename || n_records < 0 || (n_records && !records)) return B200_ERR_ARG;
        FILE* f = std::fopen(filename, append ? "ab" : "wb");
        if (!f)
            {
                set_error("dump_write: cannot open %s", filename);
                return B200_ERR_STATE;
            }
        const size_t w = std::fwrite(records, sizeof(b200_trk_dump_record), static_cast<size_t>(n_records), f);
        const int bad = std::fclose(f);
        if (w != static_cast<size_t>(n_records) || bad)
            {
                set_error("dump_write: short write to %s", filename);
                return B200_ERR_STATE;
            }
        return B200_OK;
    }
}
