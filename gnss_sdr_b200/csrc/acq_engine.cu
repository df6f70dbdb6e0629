// Host side of the acquisition C ABI (include/b200gnss.h, b200_acq_*): owns what one
// pcps_acquisition block owns (src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.h:
// FFT plans, d_grid_doppler_wipeoffs, d_fft_codes, d_magnitude_grid) as device buffers.

#include "acq_fft.cuh"
#include "engine.cuh"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace b200
{
struct AcqRowStat
{
    float max;
    unsigned int argmax;
    float sum;
    float pad;
};
}  // namespace b200

using namespace b200;

struct b200_acq
{
    b200_engine* e{nullptr};
    b200_acq_conf c{};
    FftPlan plan{};
    cudaStream_t stream{nullptr};
    bool own_stream{false};
    int doppler_center{0};
    int doppler_bias{0};
    float2* tw{nullptr};
    float2* wipe{nullptr};     // bins x n
    float2* X{nullptr};        // bins x n   forward spectra (digit-reversed)
    float2* codes{nullptr};    // slots x n  conj(FFT(code)) (digit-reversed)
    float2* in_dev{nullptr};   // consumed
    short* raw_dev{nullptr};   // 2 x consumed int16 (cshort input, converted on the device)
    float2* code_stage{nullptr};
    float* grid{nullptr};      // slots x bins x ne (optional)
    float2* wipe2{nullptr};    // step-two wipe-offs (bins2 x n)
    float2* Z{nullptr};        // two-level FFT only: rows x fft_size inverse-transformed blocks
    AcqRowStat* partial{nullptr};
    float center2{0.f}, step2{0.f};
    uint32_t bins2{0};
    AcqRowStat* rowstat{nullptr};
    int* slot_list{nullptr};
    void* best{nullptr};
    float* second_peak{nullptr};
    b200_acq_result* results_dev{nullptr};
    b200_acq_result* results_pin{nullptr};
    int* slot_pin{nullptr};
    std::vector<char> slot_set;
    // asynchronous search (b200_acq_search_submit / _wait): one sweep in flight per object
    float2* in_pin{nullptr};
    cudaEvent_t done{nullptr};
    uint32_t pending_slots{0};
    bool pending{false};
    std::vector<int> slots_on_device;  // slot list last uploaded (sweeps usually repeat it)
    // Bluestein (chirp-z) path for transform sizes the mixed-radix planner cannot factor (e.g. 16 368 = 2^4 * 3 * 11 * 31):
    // every N-point DFT is a circular convolution with a chirp, done with M-point transforms (M >= 2N-1, two-level plan).
    struct Bluestein
    {
        int M{0};
        FftPlan plM{};
        float2* twM{nullptr};
        float2* chirp{nullptr};        // w[k] = exp(+j pi k^2 / N), k < N
        float2* chirp_conj_M{nullptr}; // conj(w[k]) / M
        float2* chirp_M{nullptr};      // w[k] / M
        float2* chirp_conj{nullptr};   // conj(w[k])
        float2* filt{nullptr};         // 2 x M: FFT_M(h), FFT_M(conj h), h = w extended symmetrically
        float2* mult{nullptr};         // bins x N: wipe-off . conj(w)
        float2* A{nullptr};            // bins x M spectra of the padded, chirped rows
        float2* Zw{nullptr};           // bins x M inverse workspace
        float2* Xs{nullptr};           // bins x N: DFT_N(signal . wipe-off)
        float2* CW{nullptr};           // slots x N: conj(DFT_N(code)) . w / M
        int* slot_ids{nullptr};        // device {0, 1}
        AcqRowStat* partial{nullptr};
    };
    Bluestein* bl{nullptr};
};

namespace
{
int bluestein_search(b200_acq* a, const float2* in_dev, const uint32_t* slots, uint32_t n_slots, uint32_t dwell_counter, b200_acq_result* results_dev);

int search_impl(b200_acq* a, const float2* in_dev, const uint32_t* slots, uint32_t n_slots, uint32_t dwell_counter,
    b200_acq_result* results_dev, int step_two = 0, float prev_input_power = 0.f)
{
    const b200_acq_conf& c = a->c;
    if (n_slots == 0) return B200_OK;
    if (n_slots > c.n_code_slots || !slots) return B200_ERR_ARG;
    if (dwell_counter < 1) dwell_counter = 1;
    if (dwell_counter > 1 && !a->grid)
        {
            set_error("dwell_counter %u needs the magnitude grid (create with max_dwells > 1 or keep_grid)", dwell_counter);
            return B200_ERR_STATE;
        }
    for (uint32_t i = 0; i < n_slots; i++)
        {
            if (slots[i] >= c.n_code_slots || !a->slot_set[slots[i]])
                {
                    set_error("slot %u has no local code", slots[i]);
                    return B200_ERR_STATE;
                }
        }
    if (a->bl)
        {
            if (step_two)
                {
                    set_error("two-step acquisition is not available for chirp-z transform sizes");
                    return B200_ERR_STATE;
                }
            return bluestein_search(a, in_dev, slots, n_slots, dwell_counter, results_dev);
        }
    cudaStream_t st = a->stream;
    bool same = a->slots_on_device.size() == n_slots;
    for (uint32_t i = 0; same && i < n_slots; i++) same = a->slots_on_device[i] == static_cast<int>(slots[i]);
    if (!same)
        {
            // slot_pin is read by the asynchronous copy: wait for a previous upload's reader before rewriting it
            B200_CUDA_TRY(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < n_slots; i++) a->slot_pin[i] = static_cast<int>(slots[i]);
            B200_CUDA_TRY(cudaMemcpyAsync(a->slot_list, a->slot_pin, sizeof(int) * n_slots, cudaMemcpyHostToDevice, st));
            a->slots_on_device.assign(a->slot_pin, a->slot_pin + n_slots);
        }
    const int n = static_cast<int>(c.fft_size);
    if (step_two && (a->bins2 == 0 || !a->wipe2 || n_slots != 1))
        {
            set_error("step two needs b200_acq_set_step_two first and exactly one slot");
            return B200_ERR_STATE;
        }
    const int bins = step_two ? static_cast<int>(a->bins2) : static_cast<int>(c.num_doppler_bins);
    const float2* wipe = step_two ? a->wipe2 : a->wipe;
    const int ne = static_cast<int>(c.effective_fft_size);
    const int off = c.bit_transition_flag ? ne : 0;
    int rc = acq_launch_fwd(in_dev, static_cast<int>(c.consumed_samples), wipe, a->X, bins, a->plan, a->tw, st);
    if (rc) return rc;
    rc = acq_launch_corr(a->X, a->codes, a->slot_list, static_cast<int>(n_slots), bins, a->plan, a->tw, off, ne, a->rowstat, a->grid,
        dwell_counter > 1 ? 1 : 0, 0, nullptr, 0, nullptr, a->Z, a->partial, st);
    if (rc) return rc;
    rc = acq_launch_stats(a->rowstat, static_cast<int>(n_slots), bins, ne, c.doppler_max, a->doppler_center, c.doppler_step,
        dwell_counter, c.use_cfar, a->best, results_dev, step_two, a->center2, a->step2, prev_input_power, st);
    if (rc) return rc;
    uint64_t launched = 3;
    if (!c.use_cfar)
        {
            rc = acq_launch_corr(a->X, a->codes, a->slot_list, static_cast<int>(n_slots), bins, a->plan, a->tw, off, ne, a->rowstat,
                a->grid, 0, 1, a->best, static_cast<int>(c.samples_per_chip), a->second_peak, a->Z, a->partial, st);
            if (rc) return rc;
            rc = acq_launch_finish_second_peak(a->second_peak, static_cast<int>(n_slots), results_dev, st);
            if (rc) return rc;
            launched += 2;
        }
    {
        std::lock_guard<std::mutex> lk(a->e->mu);
        a->e->launches += launched;
    }
    (void)n;
    return B200_OK;
}
}  // namespace

namespace
{
// smallest M >= 2N - 1 that the two-level planner accepts
int bluestein_size(int n, FftPlan* pl)
{
    for (long long m = 2LL * n - 1; m <= 10LL * kAcqMaxSmemPoints; m++)
        {
            long long r = m;
            for (int p : {2, 3, 5, 7})
                while (r % p == 0) r /= p;
            if (r != 1) continue;
            if (acq_plan_make_two_level(static_cast<int>(m), pl) == B200_OK) return static_cast<int>(m);
        }
    return 0;
}

int bluestein_refresh_mult(b200_acq* a)
{
    // mult[d][i] = wipe[d][i] * conj(w[i]): the wipe-off grid keeps the reference's float32 sincos values, the chirp is extra
    b200_acq::Bluestein* b = a->bl;
    const int n = static_cast<int>(a->c.fft_size);
    return acq_launch_rows_times_vector(a->wipe, n, b->chirp_conj, n, 0, b->mult, n, static_cast<int>(a->c.num_doppler_bins), a->stream);
}

int bluestein_setup_queue(b200_acq* a, std::vector<float2> (&host)[6]);

int bluestein_setup(b200_acq* a)
{
    // the host tables must outlive the asynchronous copies queued from them, on the error paths too
    std::vector<float2> host[6];
    const int rc = bluestein_setup_queue(a, host);
    if (a->stream) cudaStreamSynchronize(a->stream);
    return rc;
}

int bluestein_setup_queue(b200_acq* a, std::vector<float2> (&host)[6])
{
    const b200_acq_conf& c = a->c;
    const int n = static_cast<int>(c.fft_size);
    auto* b = new (std::nothrow) b200_acq::Bluestein();
    if (!b) return B200_ERR_NOMEM;
    a->bl = b;
    b->M = bluestein_size(n, &b->plM);
    if (b->M == 0)
        {
            set_error("fft_size %d: no chirp-z size M >= 2N-1 within %d points", n, 10 * kAcqMaxSmemPoints);
            return B200_ERR_RANGE;
        }
    const size_t M = static_cast<size_t>(b->M), bins = c.num_doppler_bins, slots = c.n_code_slots;
    B200_CUDA_TRY(cudaMalloc(&b->twM, sizeof(float2) * M));
    B200_CUDA_TRY(cudaMalloc(&b->chirp, sizeof(float2) * n));
    B200_CUDA_TRY(cudaMalloc(&b->chirp_conj, sizeof(float2) * n));
    B200_CUDA_TRY(cudaMalloc(&b->chirp_conj_M, sizeof(float2) * n));
    B200_CUDA_TRY(cudaMalloc(&b->chirp_M, sizeof(float2) * n));
    B200_CUDA_TRY(cudaMalloc(&b->filt, sizeof(float2) * 2 * M));
    B200_CUDA_TRY(cudaMalloc(&b->mult, sizeof(float2) * bins * n));
    B200_CUDA_TRY(cudaMalloc(&b->A, sizeof(float2) * bins * M));
    B200_CUDA_TRY(cudaMalloc(&b->Zw, sizeof(float2) * bins * M));
    B200_CUDA_TRY(cudaMalloc(&b->Xs, sizeof(float2) * bins * n));
    B200_CUDA_TRY(cudaMalloc(&b->CW, sizeof(float2) * slots * n));
    B200_CUDA_TRY(cudaMalloc(&b->slot_ids, sizeof(int) * 2));
    B200_CUDA_TRY(cudaMalloc(&b->partial, sizeof(AcqRowStat) * bins * acq_final_chunks(b->plM)));
    // every copy below goes through a->stream: a synchronous cudaMemcpy from pageable memory may return before its DMA has
    // landed and is ordered only against the legacy default stream, which a->stream (non-blocking) does not wait for - the
    // filter kernels launched right after it read a half-written chirp (found on sizes with M > 32768)
    static const int ids[2] = {0, 1};
    B200_CUDA_TRY(cudaMemcpyAsync(b->slot_ids, ids, sizeof(ids), cudaMemcpyHostToDevice, a->stream));
    int rc = acq_launch_twiddles(b->twM, b->plM, a->stream);
    if (rc) return rc;
    // chirp tables in double: w[k] = exp(j pi k^2 / N) with k^2 reduced mod 2N
    std::vector<float2>&w = host[0], &wc = host[1], &wcm = host[2], &wm = host[3], &h = host[4], &hc = host[5];
    w.resize(n);
    wc.resize(n);
    wcm.resize(n);
    wm.resize(n);
    h.resize(M);
    hc.resize(M);
    const double inv_m = 1.0 / static_cast<double>(M);
    for (size_t i = 0; i < M; i++) h[i] = hc[i] = make_float2(0.f, 0.f);
    for (int k = 0; k < n; k++)
        {
            const long long k2 = (static_cast<long long>(k) * k) % (2LL * n);
            const double ang = 3.14159265358979323846 * static_cast<double>(k2) / static_cast<double>(n);
            const double cr = std::cos(ang), ci = std::sin(ang);
            w[k] = make_float2(static_cast<float>(cr), static_cast<float>(ci));
            wc[k] = make_float2(static_cast<float>(cr), static_cast<float>(-ci));
            wcm[k] = make_float2(static_cast<float>(cr * inv_m), static_cast<float>(-ci * inv_m));
            wm[k] = make_float2(static_cast<float>(cr * inv_m), static_cast<float>(ci * inv_m));
            h[k] = w[k];
            hc[k] = wc[k];
            if (k > 0)
                {
                    h[M - k] = w[k];
                    hc[M - k] = wc[k];
                }
        }
    B200_CUDA_TRY(cudaMemcpyAsync(b->chirp, w.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, a->stream));
    B200_CUDA_TRY(cudaMemcpyAsync(b->chirp_conj, wc.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, a->stream));
    B200_CUDA_TRY(cudaMemcpyAsync(b->chirp_conj_M, wcm.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, a->stream));
    B200_CUDA_TRY(cudaMemcpyAsync(b->chirp_M, wm.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, a->stream));
    // filter spectra through the code-spectrum kernel, which stores conj(FFT(.)): for a symmetric h, conj(FFT(conj h)) = FFT(h)
    B200_CUDA_TRY(cudaMemcpyAsync(b->A, hc.data(), sizeof(float2) * M, cudaMemcpyHostToDevice, a->stream));
    rc = acq_launch_code_fft(b->A, b->M, 0, b->filt, b->plM, b->twM, a->stream);
    if (rc) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(b->Zw, h.data(), sizeof(float2) * M, cudaMemcpyHostToDevice, a->stream));
    rc = acq_launch_code_fft(b->Zw, b->M, 0, b->filt + M, b->plM, b->twM, a->stream);
    return rc;
}

void bluestein_free(b200_acq* a)
{
    b200_acq::Bluestein* b = a->bl;
    if (!b) return;
    cudaFree(b->twM);
    cudaFree(b->chirp);
    cudaFree(b->chirp_conj);
    cudaFree(b->chirp_conj_M);
    cudaFree(b->chirp_M);
    cudaFree(b->filt);
    cudaFree(b->mult);
    cudaFree(b->A);
    cudaFree(b->Zw);
    cudaFree(b->Xs);
    cudaFree(b->CW);
    cudaFree(b->slot_ids);
    cudaFree(b->partial);
    delete b;
    a->bl = nullptr;
}

// rows x N: out[r] = DFT_N(in[r * in_stride .. ] . mult[r * mult_stride ..])   (forward chirp-z, natural order)
int bluestein_forward_rows(b200_acq* a, const float2* in, size_t in_stride, const float2* mult, size_t mult_stride, int rows, float2* out)
{
    b200_acq::Bluestein* b = a->bl;
    const int n = static_cast<int>(a->c.fft_size);
    int rc = acq_launch_fwd_rows(in, in_stride, n, mult, mult_stride, b->A, rows, b->plM, b->twM, a->stream);
    if (rc) return rc;
    return acq_launch_inverse_store_rows(b->A, b->filt, b->slot_ids, rows, b->plM, b->twM, b->Zw, b->chirp_conj_M, n, out, n, a->stream);
}

int bluestein_search(b200_acq* a, const float2* in_dev, const uint32_t* slots, uint32_t n_slots, uint32_t dwell_counter, b200_acq_result* results_dev)
{
    b200_acq::Bluestein* b = a->bl;
    const b200_acq_conf& c = a->c;
    const int n = static_cast<int>(c.fft_size);
    const int bins = static_cast<int>(c.num_doppler_bins);
    // X_d = DFT_N(x . wipe_d) for every Doppler bin, once per sweep
    int rc = bluestein_forward_rows(a, in_dev, 0, b->mult, n, bins, b->Xs);
    if (rc) return rc;
    uint64_t launched = 4;
    for (uint32_t i = 0; i < n_slots; i++)
        {
            // rows of this PRN: y = IDFT_N(X_d . C) = w . IFFT_M(FFT_M(pad(X_d . C . w)) . FFT_M(conj h)) / M ; |w| = 1
            rc = acq_launch_fwd_rows(b->Xs, n, n, b->CW + static_cast<size_t>(slots[i]) * n, 0, b->A, bins, b->plM, b->twM, a->stream);
            if (rc) return rc;
            // the correlation kernel addresses the grid by code-slot id, and the "code" here is filter slot 1
            float* grid = a->grid ? a->grid + (static_cast<ptrdiff_t>(slots[i]) - 1) * static_cast<ptrdiff_t>(bins) * n : nullptr;
            rc = acq_launch_corr(b->A, b->filt, b->slot_ids + 1, 1, bins, b->plM, b->twM, 0, n, a->rowstat + static_cast<size_t>(i) * bins, grid,
                dwell_counter > 1 ? 1 : 0, 0, nullptr, 0, nullptr, b->Zw, b->partial, a->stream);
            if (rc) return rc;
            launched += 5;
        }
    rc = acq_launch_stats(a->rowstat, static_cast<int>(n_slots), bins, n, c.doppler_max, a->doppler_center, c.doppler_step, dwell_counter, c.use_cfar,
        a->best, results_dev, 0, 0.f, 0.f, 0.f, a->stream);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(a->e->mu);
        a->e->launches += launched + 1;
    }
    return B200_OK;
}
}  // namespace

extern "C"
{
    int b200_acq_create(b200_engine* e, const b200_acq_conf* conf, b200_acq** out)
    {
        if (!e || !conf || !out) return B200_ERR_ARG;
        *out = nullptr;
        const b200_acq_conf& c = *conf;
        if (c.fft_size < 2 || c.effective_fft_size < 1 || c.effective_fft_size > c.fft_size || c.consumed_samples < 1 ||
            c.consumed_samples > c.fft_size || c.num_doppler_bins < 1 || c.n_code_slots < 1 || c.code_layout > 2 || c.fs_in <= 0)
            {
                set_error("b200_acq_create: inconsistent configuration");
                return B200_ERR_ARG;
            }
        if (c.bit_transition_flag && 2 * c.effective_fft_size > c.fft_size) return B200_ERR_ARG;
        FftPlan pl{};
        int rc = acq_plan_make(static_cast<int>(c.fft_size), &pl);
        bool use_bluestein = false;
        if (rc)
            {
                // sizes with prime factors > 7 (16 368 = 16.368 Msps x 1 ms, a standard front-end rate): chirp-z through M-point
                // transforms; plain layout and the CFAR statistic only
                FftPlan probe{};
                if (c.code_layout == 0 && !c.bit_transition_flag && c.effective_fft_size == c.fft_size && c.consumed_samples == c.fft_size &&
                    c.use_cfar && bluestein_size(static_cast<int>(c.fft_size), &probe) != 0)
                    {
                        use_bluestein = true;
                        pl = FftPlan{};
                        pl.n = static_cast<int>(c.fft_size);
                        pl.n_total = pl.n;
                        pl.n1 = 1;
                    }
                else
                    {
                        set_error("fft_size %u unsupported: prime factors must be 2,3,5,7 and fft_size <= 10 x %d (sizes with larger prime "
                                  "factors are supported up to %d points for the plain layout with the CFAR statistic)",
                            c.fft_size, kAcqMaxSmemPoints, 5 * kAcqMaxSmemPoints);
                        return rc;
                    }
            }
        B200_ENTER_DEVICE(e->device);
        b200_acq* a = new (std::nothrow) b200_acq();
        if (!a) return B200_ERR_NOMEM;
        a->e = e;
        // every early return of the body below destroys the half-built object
        rc = [&]() -> int {
        a->c = c;
        a->plan = pl;
        a->slot_set.assign(c.n_code_slots, 0);
        const size_t n = c.fft_size, bins = c.num_doppler_bins, slots = c.n_code_slots, ne = c.effective_fft_size;
        if (e->own_stream)
            {
                // one private stream per acquisition object: blocks of different channels overlap
                B200_CUDA_TRY(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
                a->own_stream = true;
            }
        else
            {
                a->stream = e->stream;  // caller-owned stream (e.g. for event timing on that stream)
            }
        B200_CUDA_TRY(cudaMalloc(&a->tw, sizeof(float2) * n));
        B200_CUDA_TRY(cudaMalloc(&a->wipe, sizeof(float2) * n * bins));
        B200_CUDA_TRY(cudaMalloc(&a->X, sizeof(float2) * n * bins));
        B200_CUDA_TRY(cudaMalloc(&a->codes, sizeof(float2) * n * slots));
        B200_CUDA_TRY(cudaMalloc(&a->in_dev, sizeof(float2) * n));
        B200_CUDA_TRY(cudaMalloc(&a->code_stage, sizeof(float2) * n));
        B200_CUDA_TRY(cudaMalloc(&a->rowstat, sizeof(AcqRowStat) * bins * slots));
        B200_CUDA_TRY(cudaMalloc(&a->slot_list, sizeof(int) * slots));
        B200_CUDA_TRY(cudaMalloc(&a->best, 8 * slots));
        B200_CUDA_TRY(cudaMalloc(&a->second_peak, sizeof(float) * slots));
        B200_CUDA_TRY(cudaMalloc(&a->results_dev, sizeof(b200_acq_result) * slots));
        B200_CUDA_TRY(cudaMallocHost(&a->results_pin, sizeof(b200_acq_result) * slots));
        B200_CUDA_TRY(cudaMallocHost(&a->slot_pin, sizeof(int) * slots));
        if (pl.n1 > 1)
            {
                const size_t rows = bins * slots;
                cudaError_t zerr = cudaMalloc(&a->Z, sizeof(float2) * n * rows);
                if (zerr != cudaSuccess)
                    {
                        set_error("two-level FFT workspace (%zu bytes): %s", sizeof(float2) * n * rows, cudaGetErrorString(zerr));
                        return B200_ERR_NOMEM;
                    }
                B200_CUDA_TRY(cudaMalloc(&a->partial, sizeof(AcqRowStat) * rows * acq_final_chunks(pl)));
            }
        if (c.max_dwells > 1 || c.keep_grid)
            {
                cudaError_t err = cudaMalloc(&a->grid, sizeof(float) * ne * bins * slots);
                if (err != cudaSuccess)
                    {
                        set_error("magnitude grid (%zu bytes): %s", sizeof(float) * ne * bins * slots, cudaGetErrorString(err));
                        return B200_ERR_NOMEM;
                    }
                B200_CUDA_TRY(cudaMemsetAsync(a->grid, 0, sizeof(float) * ne * bins * slots, a->stream));
            }
        int r2 = B200_OK;
        if (use_bluestein)
            {
                r2 = acq_launch_wipeoff(a->wipe, static_cast<int>(n), static_cast<int>(bins), c.doppler_max, 0, c.doppler_step, 0, c.fs_in, 0, 0.f, 0.f, a->stream);
                if (r2) return r2;
                r2 = bluestein_setup(a);
                if (r2) return r2;
                r2 = bluestein_refresh_mult(a);
                if (r2) return r2;
                B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
                return B200_OK;
            }
        r2 = acq_launch_twiddles(a->tw, a->plan, a->stream);
        if (r2) return r2;
        r2 = acq_launch_wipeoff(a->wipe, static_cast<int>(n), static_cast<int>(bins), c.doppler_max, 0, c.doppler_step, 0, c.fs_in, 0, 0.f, 0.f, a->stream);
        if (r2) return r2;
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
        }();
        if (rc != B200_OK)
            {
                b200_acq_destroy(a);
                return rc;
            }
        *out = a;
        return B200_OK;
    }


    int b200_acq_set_local_code(b200_acq* a, uint32_t slot, const b200_cf32* code_host)
    {
        if (!a || !code_host || slot >= a->c.n_code_slots) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        const b200_acq_conf& c = a->c;
        const size_t need = (c.code_layout == 1) ? c.fft_size / 2 : c.consumed_samples;
        B200_CUDA_TRY(cudaMemcpyAsync(a->code_stage, code_host, sizeof(float2) * need, cudaMemcpyHostToDevice, a->stream));
        if (a->bl)
            {
                // CW = conj(DFT_N(code)) . w / M  (volk_32fc_conjugate_32fc of set_local_code :250, then the inverse chirp-z pre-multiplier)
                b200_acq::Bluestein* b = a->bl;
                const int n = static_cast<int>(c.fft_size);
                int rcb = bluestein_forward_rows(a, a->code_stage, 0, b->chirp_conj, 0, 1, b->Xs);
                if (rcb) return rcb;
                rcb = acq_launch_rows_times_vector(b->Xs, n, b->chirp_M, n, 1, b->CW + static_cast<size_t>(slot) * n, n, 1, a->stream);
                if (rcb) return rcb;
                B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
                a->slot_set[slot] = 1;
                return B200_OK;
            }
        int rc = acq_launch_code_fft(a->code_stage, static_cast<int>(c.consumed_samples), static_cast<int>(c.code_layout),
            a->codes + static_cast<size_t>(slot) * c.fft_size, a->plan, a->tw, a->stream);
        if (rc) return rc;
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        a->slot_set[slot] = 1;
        return B200_OK;
    }

    int b200_acq_set_doppler_center(b200_acq* a, int32_t doppler_center, int32_t doppler_bias)
    {
        if (!a) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        a->doppler_center = doppler_center;
        a->doppler_bias = doppler_bias;
        const b200_acq_conf& c = a->c;
        int rc = acq_launch_wipeoff(a->wipe, static_cast<int>(c.fft_size), static_cast<int>(c.num_doppler_bins), c.doppler_max,
            doppler_center, c.doppler_step, doppler_bias, c.fs_in, 0, 0.f, 0.f, a->stream);
        if (rc) return rc;
        if (a->bl)
            {
                rc = bluestein_refresh_mult(a);
                if (rc) return rc;
            }
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_search_submit(b200_acq* a, const b200_cf32* in_host, const uint32_t* slots, uint32_t n_slots, uint32_t dwell_counter)
    {
        if (!a || !in_host) return B200_ERR_ARG;
        if (a->pending)
            {
                set_error("a search is already in flight on this acquisition object: call b200_acq_search_wait");
                return B200_ERR_STATE;
            }
        B200_ENTER_DEVICE(a->e->device);
        if (!a->in_pin) B200_CUDA_TRY(cudaMallocHost(&a->in_pin, sizeof(float2) * a->c.fft_size));
        if (!a->done) B200_CUDA_TRY(cudaEventCreateWithFlags(&a->done, cudaEventDisableTiming));
        // the caller's buffer is free as soon as this returns: stage it in pinned memory so that the copy is truly asynchronous
        std::memcpy(a->in_pin, in_host, sizeof(float2) * a->c.consumed_samples);
        B200_CUDA_TRY(cudaMemcpyAsync(a->in_dev, a->in_pin, sizeof(float2) * a->c.consumed_samples, cudaMemcpyHostToDevice, a->stream));
        int rc = search_impl(a, a->in_dev, slots, n_slots, dwell_counter, a->results_dev);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(a->results_pin, a->results_dev, sizeof(b200_acq_result) * n_slots, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaEventRecord(a->done, a->stream));
        a->pending = true;
        a->pending_slots = n_slots;
        return B200_OK;
    }

    int b200_acq_search_wait(b200_acq* a, b200_acq_result* results_host)
    {
        if (!a || !results_host) return B200_ERR_ARG;
        if (!a->pending)
            {
                set_error("no search in flight on this acquisition object");
                return B200_ERR_STATE;
            }
        B200_ENTER_DEVICE(a->e->device);
        B200_CUDA_TRY(cudaEventSynchronize(a->done));
        std::memcpy(results_host, a->results_pin, sizeof(b200_acq_result) * a->pending_slots);
        a->pending = false;
        return B200_OK;
    }

    int b200_acq_search(b200_acq* a, const b200_cf32* in_host, const uint32_t* slots, uint32_t n_slots,
        uint32_t dwell_counter, b200_acq_result* results_host)
    {
        if (!a || !in_host || !results_host) return B200_ERR_ARG;
        const int rc = b200_acq_search_submit(a, in_host, slots, n_slots, dwell_counter);
        if (rc) return rc;
        return b200_acq_search_wait(a, results_host);
    }

    // cshort input (pcps_acquisition.cc:653-656 converts with volk_gnsssdr_16ic_convert_32fc on the host): the raw 16-bit
    // pairs cross PCIe (half the bytes) and become float2 on the device; int -> float is exact, so results are those
    // of the float path on the converted samples.
    static int upload_i16(b200_acq* a, const int16_t* in_host_iq)
    {
        const size_t n = a->c.consumed_samples;
        if (!a->in_pin) B200_CUDA_TRY(cudaMallocHost(&a->in_pin, sizeof(float2) * a->c.fft_size));
        if (!a->raw_dev) B200_CUDA_TRY(cudaMalloc(&a->raw_dev, sizeof(short) * 2 * a->c.fft_size));
        std::memcpy(a->in_pin, in_host_iq, sizeof(int16_t) * 2 * n);
        B200_CUDA_TRY(cudaMemcpyAsync(a->raw_dev, a->in_pin, sizeof(int16_t) * 2 * n, cudaMemcpyHostToDevice, a->stream));
        const int rc = launch_convert_i16(a->raw_dev, a->in_dev, ~0ULL, 0ULL, n, a->stream);
        if (rc == B200_OK)
            {
                std::lock_guard<std::mutex> lk(a->e->mu);
                a->e->launches++;
            }
        return rc;
    }

    int b200_acq_search_i16(b200_acq* a, const int16_t* in_host_iq, const uint32_t* slots, uint32_t n_slots, uint32_t dwell_counter,
        b200_acq_result* results_host)
    {
        if (!a || !in_host_iq || !results_host) return B200_ERR_ARG;
        if (a->pending)
            {
                set_error("a search is already in flight on this acquisition object: call b200_acq_search_wait");
                return B200_ERR_STATE;
            }
        B200_ENTER_DEVICE(a->e->device);
        int rc = upload_i16(a, in_host_iq);
        if (rc) return rc;
        rc = search_impl(a, a->in_dev, slots, n_slots, dwell_counter, a->results_dev);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(a->results_pin, a->results_dev, sizeof(b200_acq_result) * n_slots, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        std::memcpy(results_host, a->results_pin, sizeof(b200_acq_result) * n_slots);
        return B200_OK;
    }

    int b200_acq_search_step_two_i16(b200_acq* a, const int16_t* in_host_iq, uint32_t slot, uint32_t dwell_counter, float prev_input_power,
        b200_acq_result* result_host)
    {
        if (!a || !in_host_iq || !result_host) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        int rc = upload_i16(a, in_host_iq);
        if (rc) return rc;
        rc = search_impl(a, a->in_dev, &slot, 1, dwell_counter, a->results_dev, 1, prev_input_power);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(a->results_pin, a->results_dev, sizeof(b200_acq_result), cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        *result_host = a->results_pin[0];
        return B200_OK;
    }

    int b200_acq_set_step_two(b200_acq* a, float doppler_center_step_two, float doppler_step2, uint32_t num_doppler_bins_step2)
    {
        if (!a || num_doppler_bins_step2 < 1) return B200_ERR_ARG;
        if (num_doppler_bins_step2 > a->c.num_doppler_bins)
            {
                set_error("step-two bins %u exceed the grid's %u rows", num_doppler_bins_step2, a->c.num_doppler_bins);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(a->e->device);
        if (!a->wipe2 || a->bins2 < num_doppler_bins_step2)
            {
                B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
                if (a->wipe2) B200_CUDA_TRY(cudaFree(a->wipe2));
                B200_CUDA_TRY(cudaMalloc(&a->wipe2, sizeof(float2) * a->c.fft_size * num_doppler_bins_step2));
            }
        a->bins2 = num_doppler_bins_step2;
        a->center2 = doppler_center_step_two;
        a->step2 = doppler_step2;
        int rc = acq_launch_wipeoff(a->wipe2, static_cast<int>(a->c.fft_size), static_cast<int>(a->bins2), 0, 0, 0, a->doppler_bias, a->c.fs_in, 1,
            doppler_center_step_two, doppler_step2, a->stream);
        if (rc) return rc;
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_search_step_two(b200_acq* a, const b200_cf32* in_host, uint32_t slot, uint32_t dwell_counter, float prev_input_power,
        b200_acq_result* result_host)
    {
        if (!a || !in_host || !result_host) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        B200_CUDA_TRY(cudaMemcpyAsync(a->in_dev, in_host, sizeof(float2) * a->c.consumed_samples, cudaMemcpyHostToDevice, a->stream));
        int rc = search_impl(a, a->in_dev, &slot, 1, dwell_counter, a->results_dev, 1, prev_input_power);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(a->results_pin, a->results_dev, sizeof(b200_acq_result), cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        *result_host = a->results_pin[0];
        return B200_OK;
    }

    int b200_acq_search_dev(b200_acq* a, const b200_cf32* in_dev, const uint32_t* slots_host, uint32_t n_slots,
        uint32_t dwell_counter, b200_acq_result* results_dev)
    {
        if (!a || !in_dev || !results_dev) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        // (search_impl synchronises by itself when - and only when - the slot list changed and slot_pin must be rewritten; a
        // repeated sweep over the same slots issues no host synchronisation and can be captured in a CUDA graph)
        return search_impl(a, reinterpret_cast<const float2*>(in_dev), slots_host, n_slots, dwell_counter, results_dev);
    }

    int b200_acq_sweep_best_dev(b200_acq* a, const b200_acq_result* results_dev, const uint32_t* prn_of_result_dev, uint32_t n_results,
        b200_acq_peak* peak_dev)
    {
        if (!a || !results_dev || !prn_of_result_dev || !peak_dev) return B200_ERR_ARG;
        const int rc = acq_launch_sweep_best(results_dev, prn_of_result_dev, static_cast<int>(n_results), peak_dev, a->stream);
        if (rc == B200_OK)
            {
                std::lock_guard<std::mutex> lk(a->e->mu);
                a->e->launches++;
            }
        return rc;
    }

    int b200_acq_selftest_dft(b200_acq* a, const b200_cf32* in_host, b200_cf32* out_host)
    {
        if (!a || !in_host || !out_host) return B200_ERR_ARG;
        if (!a->bl)
            {
                set_error("selftest_dft: not a chirp-z object");
                return B200_ERR_STATE;
            }
        B200_ENTER_DEVICE(a->e->device);
        const size_t n = a->c.fft_size;
        B200_CUDA_TRY(cudaMemcpyAsync(a->code_stage, in_host, sizeof(float2) * n, cudaMemcpyHostToDevice, a->stream));
        const int rc = bluestein_forward_rows(a, a->code_stage, 0, a->bl->chirp_conj, 0, 1, a->bl->Xs);
        if (rc) return rc;
        B200_CUDA_TRY(cudaMemcpyAsync(out_host, a->bl->Xs, sizeof(float2) * n, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_selftest_read(b200_acq* a, int what, b200_cf32* out_host)
    {
        if (!a || !out_host || what < 0 || what > 1) return B200_ERR_ARG;
        if (!a->bl) return B200_ERR_STATE;
        B200_ENTER_DEVICE(a->e->device);
        const size_t n = a->c.fft_size;
        const size_t rows = what == 0 ? a->c.num_doppler_bins : a->c.n_code_slots;
        B200_CUDA_TRY(cudaMemcpyAsync(out_host, what == 0 ? a->bl->Xs : a->bl->CW, sizeof(float2) * n * rows, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_read_grid(b200_acq* a, uint32_t slot, float* grid_host)
    {
        if (!a || !grid_host || slot >= a->c.n_code_slots) return B200_ERR_ARG;
        if (!a->grid)
            {
                set_error("no magnitude grid: create with keep_grid or max_dwells > 1");
                return B200_ERR_STATE;
            }
        B200_ENTER_DEVICE(a->e->device);
        const size_t row = static_cast<size_t>(a->c.num_doppler_bins) * a->c.effective_fft_size;
        B200_CUDA_TRY(cudaMemcpyAsync(grid_host, a->grid + row * slot, sizeof(float) * row, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_read_wipeoffs(b200_acq* a, b200_cf32* wipe_host)
    {
        if (!a || !wipe_host) return B200_ERR_ARG;
        B200_ENTER_DEVICE(a->e->device);
        B200_CUDA_TRY(cudaMemcpyAsync(wipe_host, a->wipe, sizeof(float2) * a->c.fft_size * a->c.num_doppler_bins, cudaMemcpyDeviceToHost, a->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(a->stream));
        return B200_OK;
    }

    int b200_acq_destroy(b200_acq* a)
    {
        if (!a) return B200_ERR_ARG;
        cudaSetDevice(a->e->device);
        if (a->stream) cudaStreamSynchronize(a->stream);
        bluestein_free(a);
        cudaFree(a->tw);
        cudaFree(a->wipe);
        cudaFree(a->wipe2);
        cudaFree(a->Z);
        cudaFree(a->partial);
        cudaFree(a->X);
        cudaFree(a->codes);
        if (a->in_pin) cudaFreeHost(a->in_pin);
        if (a->done) cudaEventDestroy(a->done);
        cudaFree(a->in_dev);
        cudaFree(a->raw_dev);
        cudaFree(a->code_stage);
        cudaFree(a->grid);
        cudaFree(a->rowstat);
        cudaFree(a->slot_list);
        cudaFree(a->best);
        cudaFree(a->second_peak);
        cudaFree(a->results_dev);
        if (a->results_pin) cudaFreeHost(a->results_pin);
        if (a->slot_pin) cudaFreeHost(a->slot_pin);
        if (a->stream && a->own_stream) cudaStreamDestroy(a->stream);
        delete a;
        return B200_OK;
    }
}
