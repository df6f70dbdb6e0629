// Host side of libb200gnss.so: engine, IQ band stores, channel registry and the tracking C ABI.
// The shapes mirror the reference's correlator class
// (src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.{h,cc}); see include/b200gnss.h
// for the per-function citations.

#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace b200
{
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

}  // namespace b200

#include "engine.cuh"

using namespace b200;

namespace b200
{
int upload_tables(b200_engine* e)
{
    if (!e->tables_dirty) return B200_OK;
    BandDesc bd[kMaxBands];
    for (int i = 0; i < kMaxBands; i++)
        {
            bd[i].base = e->bands[i].base;
            bd[i].mask = e->bands[i].mask;
            bd[i].first_index = e->bands[i].first_index;
            bd[i].limit = e->bands[i].capacity;
        }
    B200_CUDA_TRY(cudaMemcpyAsync(e->bands_dev, bd, sizeof(bd), cudaMemcpyHostToDevice, e->stream));
    const int n = static_cast<int>(e->chans.size());
    if (n > e->chans_dev_cap)
        {
            if (e->chans_dev) B200_CUDA_TRY(cudaFree(e->chans_dev));
            e->chans_dev_cap = n + 64;
            B200_CUDA_TRY(cudaMalloc(&e->chans_dev, sizeof(ChanDesc) * e->chans_dev_cap));
        }
    e->max_code_len = 0;
    e->taps_uniform = -1;
    e->taps_mask = 0u;
    e->any_high_dyn = false;
    if (n > 0)
        {
            std::vector<ChanDesc> cd(n);
            for (int i = 0; i < n; i++)
                {
                    cd[i] = e->chans[i].desc;
                    if (cd[i].code_len > e->max_code_len) e->max_code_len = cd[i].code_len;
                    if (cd[i].code == nullptr) continue;
                    if (cd[i].high_dyn) e->any_high_dyn = true;
                    e->taps_mask |= 1u << cd[i].taps;
                    if (e->taps_uniform == -1)
                        e->taps_uniform = cd[i].taps;
                    else if (e->taps_uniform != cd[i].taps)
                        e->taps_uniform = 0;
                }
            // synchronous-safe: source is a temporary, so block until the copy has been staged
            B200_CUDA_TRY(cudaMemcpyAsync(e->chans_dev, cd.data(), sizeof(ChanDesc) * n, cudaMemcpyHostToDevice, e->stream));
            B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        }
    if (e->taps_uniform < 0) e->taps_uniform = 0;
    e->tables_dirty = false;
    return B200_OK;
}

int ensure_partials(b200_engine* e, int n_items, int slices)
{
    if (slices <= 1) return B200_OK;
    const size_t need = trk_partial_elems(n_items, slices);
    if (need > e->partial_cap)
        {
            if (e->partial) B200_CUDA_TRY(cudaFree(e->partial));
            e->partial_cap = need + need / 4;
            B200_CUDA_TRY(cudaMalloc(&e->partial, sizeof(float2) * e->partial_cap));
        }
    if (n_items > e->counters_cap)
        {
            if (e->counters) B200_CUDA_TRY(cudaFree(e->counters));
            e->counters_cap = n_items + n_items / 4 + 16;
            B200_CUDA_TRY(cudaMalloc(&e->counters, sizeof(unsigned int) * e->counters_cap));
            B200_CUDA_TRY(cudaMemsetAsync(e->counters, 0, sizeof(unsigned int) * e->counters_cap, e->stream));
        }
    return B200_OK;
}

unsigned long long next_pow2(unsigned long long v)
{
    unsigned long long p = 2;
    while (p < v) p <<= 1;
    return p;
}
}  // namespace

extern "C"
{
    int b200_version(void) { return 100; }
    const char* b200_last_error(void) { return get_error(); }

    int b200_device_count(int* count)
    {
        if (!count) return B200_ERR_ARG;
        int n = 0;
        cudaError_t err = cudaGetDeviceCount(&n);
        if (err != cudaSuccess)
            {
                set_error("cudaGetDeviceCount: %s", cudaGetErrorString(err));
                *count = 0;
                return B200_ERR_NODEV;
            }
        *count = n;
        return B200_OK;
    }

    int b200_engine_create(b200_engine** out, int device, void* stream)
    {
        if (!out) return B200_ERR_ARG;
        *out = nullptr;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
            {
                set_error("no CUDA device visible");
                return B200_ERR_NODEV;
            }
        if (device < 0 || device >= n)
            {
                set_error("device %d out of range (have %d)", device, n);
                return B200_ERR_ARG;
            }
        cudaDeviceProp prop{};
        B200_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            {
                set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
                return B200_ERR_NODEV;
            }
        B200_ENTER_DEVICE(device);
        b200_engine* e = new (std::nothrow) b200_engine();
        if (!e) return B200_ERR_NOMEM;
        e->device = device;
        const int rc = [&]() -> int {
            if (stream)
                {
                    e->stream = static_cast<cudaStream_t>(stream);
                }
            else
                {
                    B200_CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
                    e->own_stream = true;
                }
            B200_CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
            B200_CUDA_TRY(cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming));
            B200_CUDA_TRY(cudaEventCreate(&e->t0));
            B200_CUDA_TRY(cudaEventCreate(&e->t1));
            B200_CUDA_TRY(cudaMalloc(&e->bands_dev, sizeof(BandDesc) * kMaxBands));
            return B200_OK;
        }();
        if (rc != B200_OK)
            {
                b200_engine_destroy(e);  // tolerates a half-built engine
                return rc;
            }
        *out = e;
        return B200_OK;
    }

    int b200_engine_destroy(b200_engine* e)
    {
        if (!e) return B200_ERR_ARG;
        cudaSetDevice(e->device);
        if (e->stream) cudaStreamSynchronize(e->stream);
        if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
        for (auto& b : e->bands)
            {
                if (b.dev) cudaFree(b.dev);
                for (int k = 0; k < 2; k++)
                    {
                        if (b.raw_stage[k]) cudaFree(b.raw_stage[k]);
                        if (b.raw_free[k]) cudaEventDestroy(b.raw_free[k]);
                    }
            }
        for (auto& c : e->chans)
            if (c.code_dev) cudaFree(c.code_dev);
        if (e->bands_dev) cudaFree(e->bands_dev);
        if (e->chans_dev) cudaFree(e->chans_dev);
        if (e->items_dev) cudaFree(e->items_dev);
        if (e->items_pin) cudaFreeHost(e->items_pin);
        if (e->out_dev) cudaFree(e->out_dev);
        if (e->out_pin) cudaFreeHost(e->out_pin);
        if (e->partial) cudaFree(e->partial);
        if (e->counters) cudaFree(e->counters);
        loops_free(e);
        for (auto& sl : e->slots)
            {
                if (sl.items_dev) cudaFree(sl.items_dev);
                if (sl.items_pin) cudaFreeHost(sl.items_pin);
                if (sl.out_dev) cudaFree(sl.out_dev);
                if (sl.out_pin) cudaFreeHost(sl.out_pin);
                if (sl.done) cudaEventDestroy(sl.done);
                if (sl.items_ready) cudaEventDestroy(sl.items_ready);
            }
        if (e->copy_done) cudaEventDestroy(e->copy_done);
        if (e->t0) cudaEventDestroy(e->t0);
        if (e->t1) cudaEventDestroy(e->t1);
        if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
        if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
        delete e;
        return B200_OK;
    }

    int b200_engine_sync(b200_engine* e)
    {
        if (!e) return B200_ERR_ARG;
        B200_CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200_OK;
    }

    int b200_engine_timer_start(b200_engine* e)
    {
        if (!e) return B200_ERR_ARG;
        B200_CUDA_TRY(cudaEventRecord(e->t0, e->stream));
        return B200_OK;
    }

    int b200_engine_timer_stop_ms(b200_engine* e, float* ms)
    {
        if (!e || !ms) return B200_ERR_ARG;
        B200_CUDA_TRY(cudaEventRecord(e->t1, e->stream));
        B200_CUDA_TRY(cudaEventSynchronize(e->t1));
        B200_CUDA_TRY(cudaEventElapsedTime(ms, e->t0, e->t1));
        return B200_OK;
    }

    int b200_engine_launch_count(b200_engine* e, uint64_t* n)
    {
        if (!e || !n) return B200_ERR_ARG;
        *n = e->launches;
        return B200_OK;
    }

    // ---- bands -----------------------------------------------------------------------------
    int b200_iq_create(b200_engine* e, int band, uint64_t capacity_samples)
    {
        if (!e || band < 0 || band >= kMaxBands || capacity_samples == 0) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        B200_ENTER_DEVICE(e->device);
        Band& b = e->bands[band];
        if (b.dev)
            {
                B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
                B200_CUDA_TRY(cudaFree(b.dev));
                b.dev = nullptr;
            }
        const unsigned long long cap = next_pow2(capacity_samples);
        cudaError_t err = cudaMalloc(&b.dev, cap * sizeof(float2));
        if (err != cudaSuccess)
            {
                set_error("band %d: cudaMalloc(%llu samples): %s", band, cap, cudaGetErrorString(err));
                return B200_ERR_NOMEM;
            }
        b.base = b.dev;
        b.capacity = cap;
        b.mask = cap - 1;
        b.first_index = 0;
        b.write_index = 0;
        b.valid_from = 0;
        b.attached = false;
        b.in_use = true;
        e->tables_dirty = true;
        return B200_OK;
    }

    int b200_iq_push(b200_engine* e, int band, const b200_cf32* host, uint64_t n, uint64_t* first_index)
    {
        if (!e || band < 0 || band >= kMaxBands || (!host && n)) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        Band& b = e->bands[band];
        if (!b.in_use || b.attached || !b.dev)
            {
                set_error("band %d is not an owned ring (call b200_iq_create)", band);
                return B200_ERR_STATE;
            }
        if (n > b.capacity)
            {
                set_error("push of %llu samples exceeds ring capacity %llu", (unsigned long long)n, b.capacity);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(e->device);
        if (first_index) *first_index = b.write_index;
        // copies run on the copy stream; later launches on the compute stream wait on copy_done.
        // (Overwriting samples that an in-flight launch still reads is the caller's ring-sizing
        // responsibility, exactly as with any circular sample buffer.)
        const unsigned long long off = b.write_index & b.mask;
        const unsigned long long first = (n < b.capacity - off) ? n : (b.capacity - off);
        if (first) B200_CUDA_TRY(cudaMemcpyAsync(b.dev + off, host, first * sizeof(float2), cudaMemcpyHostToDevice, e->copy_stream));
        if (n > first) B200_CUDA_TRY(cudaMemcpyAsync(b.dev, host + first, (n - first) * sizeof(float2), cudaMemcpyHostToDevice, e->copy_stream));
        B200_CUDA_TRY(cudaEventRecord(e->copy_done, e->copy_stream));
        B200_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->copy_done, 0));
        b.write_index += n;
        return B200_OK;
    }

    // shared implementation of the integer-sample pushes: raw bytes H2D, then conversion into the ring
    static int push_raw(b200_engine* e, int band, const void* host, uint64_t n, uint64_t* first_index, int bytes_per_component)
    {
        if (!e || band < 0 || band >= kMaxBands || (!host && n)) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        Band& b = e->bands[band];
        if (!b.in_use || b.attached || !b.dev)
            {
                set_error("band %d is not an owned ring (call b200_iq_create)", band);
                return B200_ERR_STATE;
            }
        if (n > b.capacity)
            {
                set_error("push of %llu samples exceeds ring capacity %llu", (unsigned long long)n, b.capacity);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(e->device);
        if (first_index) *first_index = b.write_index;
        const unsigned long long bytes = n * 2ULL * static_cast<unsigned long long>(bytes_per_component);
        const int k = b.raw_next;
        b.raw_next ^= 1;
        if (!b.raw_free[k]) B200_CUDA_TRY(cudaEventCreateWithFlags(&b.raw_free[k], cudaEventDisableTiming));
        if (bytes > b.raw_cap[k])
            {
                B200_CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
                B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
                if (b.raw_stage[k]) B200_CUDA_TRY(cudaFree(b.raw_stage[k]));
                b.raw_cap[k] = bytes + bytes / 4 + 256;
                B200_CUDA_TRY(cudaMalloc(&b.raw_stage[k], b.raw_cap[k]));
                B200_CUDA_TRY(cudaEventRecord(b.raw_free[k], e->stream));
            }
        if (n)
            {
                // raw integers cross PCIe on the copy stream; the conversion into the float ring runs on the compute
                // stream (ordered before every later correlator launch), so the copy engine goes straight on to the
                // next block, which lands in the other staging buffer
                B200_CUDA_TRY(cudaStreamWaitEvent(e->copy_stream, b.raw_free[k], 0));
                B200_CUDA_TRY(cudaMemcpyAsync(b.raw_stage[k], host, bytes, cudaMemcpyHostToDevice, e->copy_stream));
                B200_CUDA_TRY(cudaEventRecord(e->copy_done, e->copy_stream));
                B200_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->copy_done, 0));
                int rc = (bytes_per_component == 2)
                             ? launch_convert_i16(static_cast<const short*>(b.raw_stage[k]), b.dev, b.mask, b.write_index, n, e->stream)
                             : launch_convert_i8(static_cast<const signed char*>(b.raw_stage[k]), b.dev, b.mask, b.write_index, n, e->stream);
                if (rc) return rc;
                e->launches++;
                B200_CUDA_TRY(cudaEventRecord(b.raw_free[k], e->stream));
            }
        b.write_index += n;
        return B200_OK;
    }

    int b200_iq_push_at(b200_engine* e, int band, uint64_t abs_index, const b200_cf32* host, uint64_t n, uint64_t* n_new)
    {
        if (!e || band < 0 || band >= kMaxBands || (!host && n)) return B200_ERR_ARG;
        if (n_new) *n_new = 0;
        // competing producers of one band (every tracking block of a flowgraph sees the same stream and offers the same
        // samples) are serialised here; e->mu itself is only held for the bookkeeping and inside b200_iq_push
        std::lock_guard<std::mutex> plk(e->push_mu[band]);
        uint64_t skip = 0;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            Band& b = e->bands[band];
            if (!b.in_use || b.attached || !b.dev)
                {
                    set_error("band %d is not an owned ring (call b200_iq_create)", band);
                    return B200_ERR_STATE;
                }
            // what the ring holds right now
            const unsigned long long oldest = b.write_index > b.capacity ? b.write_index - b.capacity : 0ULL;
            const unsigned long long lo = b.valid_from > oldest ? b.valid_from : oldest;
            if (abs_index >= lo && abs_index + n <= b.write_index) return B200_OK;  // every sample of this block is already in the band
            if (abs_index > b.write_index || abs_index < lo)
                {
                    // a gap (the first block to track starts long after sample 0) or a stream that starts over at an older
                    // index (a new capture, a test): what the ring held is history, the band restarts at abs_index
                    b.valid_from = abs_index;
                    b.write_index = abs_index;
                }
            skip = b.write_index - abs_index;
        }
        uint64_t first = 0;
        const int rc = b200_iq_push(e, band, host + skip, n - skip, &first);
        if (rc == B200_OK && n_new) *n_new = n - skip;
        return rc;
    }

    int b200_iq_forget(b200_engine* e, int band)
    {
        if (!e || band < 0 || band >= kMaxBands) return B200_ERR_ARG;
        std::lock_guard<std::mutex> plk(e->push_mu[band]);
        std::lock_guard<std::mutex> lk(e->mu);
        Band& b = e->bands[band];
        if (!b.in_use || b.attached) return B200_ERR_STATE;
        b.valid_from = b.write_index;
        return B200_OK;
    }

    int b200_iq_window(b200_engine* e, int band, uint64_t* valid_from, uint64_t* write_index)
    {
        if (!e || band < 0 || band >= kMaxBands) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        const Band& b = e->bands[band];
        if (!b.in_use) return B200_ERR_STATE;
        const unsigned long long oldest = (!b.attached && b.write_index > b.capacity) ? b.write_index - b.capacity : b.first_index;
        if (valid_from) *valid_from = b.valid_from > oldest ? b.valid_from : oldest;
        if (write_index) *write_index = b.write_index;
        return B200_OK;
    }

    int b200_iq_push_i16(b200_engine* e, int band, const int16_t* host_iq, uint64_t n, uint64_t* first_index)
    {
        return push_raw(e, band, host_iq, n, first_index, 2);
    }

    int b200_iq_push_i8(b200_engine* e, int band, const int8_t* host_iq, uint64_t n, uint64_t* first_index)
    {
        return push_raw(e, band, host_iq, n, first_index, 1);
    }

    int b200_iq_push_file(b200_engine* e, int band, const char* path, const char* item_type, uint64_t header_bytes, uint64_t skip_samples,
        uint64_t max_samples, uint64_t chunk_samples, uint64_t* first_index, uint64_t* samples_pushed)
    {
        if (!e || !path || !item_type || band < 0 || band >= kMaxBands) return B200_ERR_ARG;
        // FileSourceBase::itemTypeToSize (file_source_base.cc:340-378): complex samples as interleaved (I, Q) pairs
        int bytes_per_component;
        if (std::strcmp(item_type, "gr_complex") == 0)
            bytes_per_component = 4;
        else if (std::strcmp(item_type, "ishort") == 0)
            bytes_per_component = 2;
        else if (std::strcmp(item_type, "ibyte") == 0)
            bytes_per_component = 1;
        else
            {
                set_error("iq_push_file: item type %s is not an interleaved complex type (gr_complex, ishort, ibyte)", item_type);
                return B200_ERR_ARG;
            }
        unsigned long long capacity;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            const Band& b = e->bands[band];
            if (!b.in_use || b.attached || !b.dev)
                {
                    set_error("band %d is not an owned ring (call b200_iq_create)", band);
                    return B200_ERR_STATE;
                }
            capacity = b.capacity;
        }
        if (chunk_samples == 0) chunk_samples = 1ULL << 20;
        if (chunk_samples > capacity / 2) chunk_samples = capacity / 2 ? capacity / 2 : 1;
        B200_ENTER_DEVICE(e->device);
        FILE* f = std::fopen(path, "rb");
        if (!f)
            {
                set_error("iq_push_file: cannot open %s", path);
                return B200_ERR_STATE;
            }
        const size_t sample_bytes = 2u * static_cast<size_t>(bytes_per_component);
        // samplesToSkip (:385-414): header, then whole samples
        if (std::fseek(f, static_cast<long>(header_bytes + skip_samples * sample_bytes), SEEK_SET) != 0)
            {
                std::fclose(f);
                set_error("iq_push_file: cannot seek in %s", path);
                return B200_ERR_RANGE;
            }
        // two pinned staging buffers: the disk read of block k+1 overlaps the host->device copy of block k
        void* stage[2] = {nullptr, nullptr};
        cudaEvent_t freed[2] = {nullptr, nullptr};
        int rc = B200_OK;
        for (int k = 0; k < 2 && rc == B200_OK; k++)
            {
                if (cudaMallocHost(&stage[k], chunk_samples * sample_bytes) != cudaSuccess ||
                    cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming) != cudaSuccess)
                    {
                        set_error("iq_push_file: pinned staging allocation failed");
                        rc = B200_ERR_NOMEM;
                    }
            }
        uint64_t total = 0;
        bool have_first = false;
        for (int k = 0; rc == B200_OK && (max_samples == 0 || total < max_samples); k ^= 1)
            {
                if (cudaEventSynchronize(freed[k]) != cudaSuccess)  // the copy that last read this buffer has finished
                    {
                        rc = B200_ERR_CUDA;
                        break;
                    }
                uint64_t want = chunk_samples;
                if (max_samples != 0 && max_samples - total < want) want = max_samples - total;
                const size_t got = std::fread(stage[k], sample_bytes, static_cast<size_t>(want), f);
                if (got == 0) break;
                uint64_t first = 0;
                if (bytes_per_component == 4)
                    rc = b200_iq_push(e, band, static_cast<const b200_cf32*>(stage[k]), got, &first);
                else if (bytes_per_component == 2)
                    rc = b200_iq_push_i16(e, band, static_cast<const int16_t*>(stage[k]), got, &first);
                else
                    rc = b200_iq_push_i8(e, band, static_cast<const int8_t*>(stage[k]), got, &first);
                if (rc != B200_OK) break;
                if (cudaEventRecord(freed[k], e->copy_stream) != cudaSuccess) rc = B200_ERR_CUDA;
                if (!have_first)
                    {
                        have_first = true;
                        if (first_index) *first_index = first;
                    }
                total += got;
                if (got < want) break;  // end of file
            }
        std::fclose(f);
        cudaStreamSynchronize(e->copy_stream);
        for (int k = 0; k < 2; k++)
            {
                if (stage[k]) cudaFreeHost(stage[k]);
                if (freed[k]) cudaEventDestroy(freed[k]);
            }
        if (samples_pushed) *samples_pushed = total;
        return rc;
    }

    int b200_iq_attach_dev(b200_engine* e, int band, const b200_cf32* dev, uint64_t n_samples, uint64_t first_index)
    {
        if (!e || band < 0 || band >= kMaxBands || !dev) return B200_ERR_ARG;
        if (reinterpret_cast<uintptr_t>(dev) & 15)
            {
                set_error("attached band must be 16-byte aligned");
                return B200_ERR_ARG;
            }
        std::lock_guard<std::mutex> lk(e->mu);
        Band& b = e->bands[band];
        if (b.dev)
            {
                B200_ENTER_DEVICE(e->device);
                B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
                B200_CUDA_TRY(cudaFree(b.dev));
                b.dev = nullptr;
            }
        b.base = reinterpret_cast<const float2*>(dev);
        b.capacity = n_samples;
        b.mask = ~0ULL;
        b.first_index = first_index;
        b.write_index = first_index + n_samples;
        b.valid_from = first_index;
        b.attached = true;
        b.in_use = true;
        e->tables_dirty = true;
        return B200_OK;
    }

    int b200_iq_refill(b200_engine* e, int band, const b200_cf32* host, uint64_t n, uint64_t first_index)
    {
        if (!e || band < 0 || band >= kMaxBands || (!host && n)) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        Band& b = e->bands[band];
        if (!b.in_use || !b.attached)
            {
                set_error("band %d is not an attached device buffer (call b200_iq_attach_dev)", band);
                return B200_ERR_STATE;
            }
        if (n > b.capacity)
            {
                set_error("refill of %llu samples exceeds the attached buffer (%llu)", (unsigned long long)n, b.capacity);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(e->device);
        // same ordering as b200_iq_push: the copy runs on the copy stream, later launches on the compute stream wait for it;
        // overwriting samples an in-flight launch still reads is the caller's double-buffering responsibility
        if (n) B200_CUDA_TRY(cudaMemcpyAsync(const_cast<float2*>(b.base), host, n * sizeof(float2), cudaMemcpyHostToDevice, e->copy_stream));
        B200_CUDA_TRY(cudaEventRecord(e->copy_done, e->copy_stream));
        B200_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->copy_done, 0));
        if (b.first_index != first_index)
            {
                b.first_index = first_index;
                e->tables_dirty = true;
            }
        b.valid_from = first_index;
        b.write_index = first_index + n;
        return B200_OK;
    }

    // ---- channels ------------------------------------------------------------------------------
    int b200_trk_channel_create(b200_engine* e, int band, int n_correlators, int* channel_id)
    {
        if (!e || !channel_id || band < 0 || band >= kMaxBands) return B200_ERR_ARG;
        if (n_correlators < 1 || n_correlators > B200_MAX_TAPS)
            {
                set_error("n_correlators %d outside 1..%d", n_correlators, B200_MAX_TAPS);
                return B200_ERR_RANGE;
            }
        std::lock_guard<std::mutex> lk(e->mu);
        Channel c;
        c.desc.band = band;
        c.desc.taps = n_correlators;
        c.desc.code = nullptr;
        c.desc.code_len = 0;
        c.desc.high_dyn = 0;
        for (float& s : c.desc.shifts) s = 0.f;
        e->chans.push_back(c);
        *channel_id = static_cast<int>(e->chans.size()) - 1;
        e->tables_dirty = true;
        return B200_OK;
    }

    int b200_trk_channel_set_code(b200_engine* e, int channel_id, int code_length_chips, const float* local_code_in, const float* shifts_chips, int high_dynamics)
    {
        if (!e || !local_code_in || !shifts_chips || code_length_chips < 1) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        if (channel_id < 0 || channel_id >= static_cast<int>(e->chans.size())) return B200_ERR_ARG;
        B200_ENTER_DEVICE(e->device);
        Channel& c = e->chans[channel_id];
        // a new table may be read by launches already queued: allocate fresh if it must grow
        if (code_length_chips > c.code_cap)
            {
                B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
                if (c.code_dev) B200_CUDA_TRY(cudaFree(c.code_dev));
                c.code_cap = code_length_chips;
                B200_CUDA_TRY(cudaMalloc(&c.code_dev, sizeof(float) * c.code_cap));
            }
        B200_CUDA_TRY(cudaMemcpyAsync(c.code_dev, local_code_in, sizeof(float) * code_length_chips, cudaMemcpyHostToDevice, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));  // caller's buffer may go away
        c.desc.code = c.code_dev;
        c.desc.code_len = code_length_chips;
        c.desc.high_dyn = high_dynamics ? 1 : 0;
        for (int t = 0; t < c.desc.taps; t++) c.desc.shifts[t] = shifts_chips[t];
        e->tables_dirty = true;
        return B200_OK;
    }

    int b200_trk_channel_set_taps(b200_engine* e, int channel_id, const float* shifts_chips)
    {
        if (!e || !shifts_chips) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        if (channel_id < 0 || channel_id >= static_cast<int>(e->chans.size())) return B200_ERR_ARG;
        Channel& c = e->chans[channel_id];
        bool changed = false;
        for (int t = 0; t < c.desc.taps; t++)
            {
                if (c.desc.shifts[t] != shifts_chips[t]) changed = true;
                c.desc.shifts[t] = shifts_chips[t];
            }
        // the descriptor table is re-uploaded in stream order before the next launch (earlier launches keep the old taps)
        if (changed) e->tables_dirty = true;
        return B200_OK;
    }

    // caller holds e->mu
    // layout_hint: 1 = consecutive items overlap in the band (a receiver's channels on one stream), 0 = they do not (every
    // item reads its own samples: nothing to share, the per-item kernel streams them faster), -1 = unknown (device items)
    static int batch_dev_impl(b200_engine* e, const b200_trk_item* items_dev, int n_items, b200_cf32* out_dev, int out_stride, int slices,
        int layout_hint = -1)
    {
        B200_ENTER_DEVICE(e->device);
        int rc = upload_tables(e);
        if (rc) return rc;
        if (slices < 1) slices = 1;
        rc = ensure_partials(e, n_items, slices);
        if (rc) return rc;
        if (e->shared_mode < 0)
            {
                const char* env = std::getenv("B200_TRK_SHARED");
                e->shared_mode = env ? std::atoi(env) : 2;  // 2 = automatic
            }
        const bool shared_legal = (e->taps_uniform == 1 || e->taps_uniform == 3 || e->taps_uniform == 5) && !e->any_high_dyn && slices == 1;
        const bool use_shared = shared_legal && (e->shared_mode == 1 || (e->shared_mode == 2 && n_items >= 1024 && layout_hint != 0 &&
                                                                            e->max_code_len <= trk_shared_max_code_len()));
        if (use_shared)
            rc = launch_trk_shared(items_dev, n_items, e->chans_dev, e->bands_dev, reinterpret_cast<float2*>(out_dev), out_stride,
                e->taps_uniform, e->stream);
        else
            rc = launch_trk_batch(items_dev, n_items, e->chans_dev, e->bands_dev, reinterpret_cast<float2*>(out_dev), out_stride,
                slices, e->partial, e->counters, e->max_code_len, e->taps_uniform, e->stream, e->taps_uniform == 0 ? e->taps_mask : 0u);
        if (rc == B200_OK)
            {
                // a batch with several tap counts on the per-item path is one launch per specialised kernel (trk_kernels.cu)
                int n_launch = 1;
                if (!use_shared && e->taps_uniform == 0 && e->taps_mask != 0u)
                    {
                        const unsigned int special = e->taps_mask & ((1u << 1) | (1u << 3) | (1u << 5));
                        n_launch = __builtin_popcount(special) + ((e->taps_mask & ~special) ? 1 : 0);
                    }
                e->launches += static_cast<uint64_t>(n_launch);
            }
        return rc;
    }

    int b200_trk_kernel_choice(b200_engine* e, int mode)
    {
        if (!e || mode < -1 || mode > 2) return B200_ERR_ARG;
        std::lock_guard<std::mutex> lk(e->mu);
        e->shared_mode = mode;
        return B200_OK;
    }

    int b200_trk_batch_dev(b200_engine* e, const b200_trk_item* items_dev, int n_items, b200_cf32* out_dev, int out_stride, int slices)
    {
        if (!e || n_items < 0 || (n_items && (!items_dev || !out_dev)) || out_stride < 1) return B200_ERR_ARG;
        if (n_items == 0) return B200_OK;
        std::lock_guard<std::mutex> lk(e->mu);
        return batch_dev_impl(e, items_dev, n_items, out_dev, out_stride, slices);
    }

    int b200_trk_submit(b200_engine* e, const b200_trk_item* items_host, int n_items, int out_stride, uint64_t* ticket)
    {
        if (!e || !ticket || n_items < 1 || !items_host || out_stride < 1) return B200_ERR_ARG;
        // one lock for the whole submission: concurrent submitters (the reference runs one thread per tracking block)
        // queue whole batches, never interleave the item copy of one with the launch of another
        std::lock_guard<std::mutex> lk(e->mu);
        B200_ENTER_DEVICE(e->device);
        for (int i = 0; i < n_items; i++)
            {
                const int ch = items_host[i].channel;
                if (ch < 0 || ch >= static_cast<int>(e->chans.size()) || e->chans[ch].desc.code == nullptr)
                    {
                        set_error("item %d: channel %d has no code table", i, ch);
                        return B200_ERR_STATE;
                    }
                if (e->chans[ch].desc.taps > out_stride)
                    {
                        set_error("item %d: out_stride %d < taps %d", i, out_stride, e->chans[ch].desc.taps);
                        return B200_ERR_ARG;
                    }
                // ring bands: the epoch must lie inside what has been pushed and not yet overwritten
                const Band& b = e->bands[e->chans[ch].desc.band];
                if (b.in_use && !b.attached && items_host[i].n > 0)
                    {
                        const unsigned long long s0 = items_host[i].sample_index, s1 = s0 + static_cast<unsigned long long>(items_host[i].n);
                        const unsigned long long oldest = b.write_index > b.capacity ? b.write_index - b.capacity : 0ULL;
                        if (s0 < b.valid_from || s0 < oldest || s1 > b.write_index)
                            {
                                set_error("item %d: samples [%llu, %llu) are outside the band's window [%llu, %llu)", i, s0, s1,
                                    b.valid_from > oldest ? b.valid_from : oldest, b.write_index);
                                return B200_ERR_RANGE;
                            }
                    }
            }
        b200_engine::Slot* sl = nullptr;
        for (auto& s : e->slots)
            if (!s.busy)
                {
                    sl = &s;
                    break;
                }
        if (!sl)
            {
                set_error("more than %d batches in flight: call b200_trk_wait", b200_engine::kSlots);
                return B200_ERR_STATE;
            }
        // the slot is released again on every early return below (a transient CUDA error must not leak it)
        struct BusyGuard
        {
            b200_engine::Slot* s;
            bool keep{false};
            ~BusyGuard()
            {
                if (!keep) s->busy = false;
            }
        } guard{sl};
        sl->busy = true;
        if (!sl->done) B200_CUDA_TRY(cudaEventCreateWithFlags(&sl->done, cudaEventDisableTiming));
        if (!sl->items_ready) B200_CUDA_TRY(cudaEventCreateWithFlags(&sl->items_ready, cudaEventDisableTiming));
        if (n_items > sl->items_cap)
            {
                if (sl->items_dev) B200_CUDA_TRY(cudaFree(sl->items_dev));
                sl->items_dev = nullptr;
                if (sl->items_pin) B200_CUDA_TRY(cudaFreeHost(sl->items_pin));
                sl->items_pin = nullptr;
                sl->items_cap = 0;
                const int cap = n_items + n_items / 2 + 64;
                B200_CUDA_TRY(cudaMalloc(&sl->items_dev, sizeof(b200_trk_item) * cap));
                B200_CUDA_TRY(cudaMallocHost(&sl->items_pin, sizeof(b200_trk_item) * cap));
                sl->items_cap = cap;
            }
        const int out_elems = n_items * out_stride;
        if (out_elems > sl->out_cap)
            {
                if (sl->out_dev) B200_CUDA_TRY(cudaFree(sl->out_dev));
                sl->out_dev = nullptr;
                if (sl->out_pin) B200_CUDA_TRY(cudaFreeHost(sl->out_pin));
                sl->out_pin = nullptr;
                sl->out_cap = 0;
                const int cap = out_elems + out_elems / 2 + 64;
                B200_CUDA_TRY(cudaMalloc(&sl->out_dev, sizeof(float2) * cap));
                B200_CUDA_TRY(cudaMallocHost(&sl->out_pin, sizeof(float2) * cap));
                sl->out_cap = cap;
            }
        sl->n_items = n_items;
        sl->out_stride = out_stride;
        // The shared-window kernel serves items 8g..8g+7 from one copy of the samples, so items
        // should be ordered by start sample.  Already-ordered input (the usual epoch-major layout)
        // is copied as is; otherwise sort a permutation and undo it in b200_trk_wait.
        bool sorted = true;
        for (int i = 1; i < n_items && sorted; i++) sorted = items_host[i - 1].sample_index <= items_host[i].sample_index;
        sl->perm.clear();
        if (sorted)
            {
                std::memcpy(sl->items_pin, items_host, sizeof(b200_trk_item) * n_items);
            }
        else
            {
                sl->perm.resize(n_items);
                for (int i = 0; i < n_items; i++) sl->perm[i] = i;
                std::stable_sort(sl->perm.begin(), sl->perm.end(),
                    [&](int a, int b) { return items_host[a].sample_index < items_host[b].sample_index; });
                for (int i = 0; i < n_items; i++) sl->items_pin[i] = items_host[sl->perm[i]];
            }
        // few items: split epochs into slices so the whole chip works on them
        int slices = 1;
        if (n_items < 592) slices = (592 + n_items - 1) / n_items;
        if (slices > 64) slices = 64;
        // Latency path (a handful of channels waiting for their taps, e.g. the block threads behind the coalescer): the
        // kernel reads the work items straight from the pinned host buffer and writes the taps straight into pinned host
        // memory (unified addressing: cudaMallocHost memory is device-accessible at the same address), so the submission
        // is one launch + one event instead of two copies through the copy engine, two events and a stream hop.
        sl->zero_copy = n_items <= 1024;
        if (sl->zero_copy)
            {
                const int rc = batch_dev_impl(e, sl->items_pin, n_items, reinterpret_cast<b200_cf32*>(sl->out_pin), out_stride, slices);
                if (rc) return rc;
                B200_CUDA_TRY(cudaEventRecord(sl->done, e->stream));
            }
        else
            {
                // The work items travel on the COPY stream, queued behind the sample pushes made so far, and the compute
                // stream waits on an event.  Issued on the compute stream instead, this small host->device copy sits in
                // the copy engine's queue until the previous batch's kernel has finished and holds up every sample
                // push queued behind it (measured: 41.7 instead of 54.6 GB/s of sustained host->device traffic).
                B200_CUDA_TRY(cudaMemcpyAsync(sl->items_dev, sl->items_pin, sizeof(b200_trk_item) * n_items, cudaMemcpyHostToDevice, e->copy_stream));
                B200_CUDA_TRY(cudaEventRecord(sl->items_ready, e->copy_stream));
                B200_CUDA_TRY(cudaStreamWaitEvent(e->stream, sl->items_ready, 0));
                // do the groups of consecutive (sorted) items the shared-window kernel would form really overlap?
                unsigned long long sum_n = 0, sum_hull = 0;
                for (int g = 0; g < n_items; g += 8)
                    {
                        unsigned long long lo = ~0ULL, hi = 0;
                        for (int i = g; i < n_items && i < g + 8; i++)
                            {
                                const b200_trk_item& it = sl->items_pin[i];
                                const unsigned long long a = it.sample_index + (static_cast<unsigned long long>(e->chans[it.channel].desc.band) << 48);
                                lo = a < lo ? a : lo;
                                hi = a + it.n > hi ? a + it.n : hi;
                                sum_n += it.n > 0 ? it.n : 0;
                            }
                        sum_hull += hi - lo;
                    }
                const int layout_hint = (sum_n > 2 * sum_hull) ? 1 : 0;
                const int rc = batch_dev_impl(e, sl->items_dev, n_items, reinterpret_cast<b200_cf32*>(sl->out_dev), out_stride, slices, layout_hint);
                if (rc) return rc;
                B200_CUDA_TRY(cudaMemcpyAsync(sl->out_pin, sl->out_dev, sizeof(float2) * n_items * out_stride, cudaMemcpyDeviceToHost, e->stream));
                B200_CUDA_TRY(cudaEventRecord(sl->done, e->stream));
            }
        sl->ticket = e->next_ticket++;
        *ticket = sl->ticket;
        guard.keep = true;
        return B200_OK;
    }

    int b200_trk_wait(b200_engine* e, uint64_t ticket, b200_cf32* out_host)
    {
        if (!e || !out_host) return B200_ERR_ARG;
        b200_engine::Slot* sl = nullptr;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            for (auto& s : e->slots)
                if (s.busy && s.ticket == ticket) sl = &s;
        }
        if (!sl)
            {
                set_error("unknown ticket %llu", static_cast<unsigned long long>(ticket));
                return B200_ERR_ARG;
            }
        B200_ENTER_DEVICE(e->device);
        B200_CUDA_TRY(cudaEventSynchronize(sl->done));
        if (sl->perm.empty())
            {
                std::memcpy(out_host, sl->out_pin, sizeof(float2) * sl->n_items * sl->out_stride);
            }
        else
            {
                for (int k = 0; k < sl->n_items; k++)
                    std::memcpy(out_host + static_cast<size_t>(sl->perm[k]) * sl->out_stride, sl->out_pin + static_cast<size_t>(k) * sl->out_stride,
                        sizeof(float2) * sl->out_stride);
            }
        std::lock_guard<std::mutex> lk(e->mu);
        sl->busy = false;
        return B200_OK;
    }

    int b200_trk_batch(b200_engine* e, const b200_trk_item* items_host, int n_items, b200_cf32* out_host, int out_stride)
    {
        if (!e || n_items < 0 || (n_items && (!items_host || !out_host)) || out_stride < 1) return B200_ERR_ARG;
        if (n_items == 0) return B200_OK;
        uint64_t ticket = 0;
        int rc = b200_trk_submit(e, items_host, n_items, out_stride, &ticket);
        if (rc) return rc;
        return b200_trk_wait(e, ticket, out_host);
    }
}

// ---- single correlator object (Cpu_Multicorrelator_Real_Codes shape) ----------------------------
struct b200_trk
{
    b200_engine* e{nullptr};
    int max_len{0};
    int taps{0};
    int high_dyn{0};
    cudaStream_t stream{nullptr};
    float2* sig_dev{nullptr};
    float* code_dev{nullptr};
    int code_cap{0};
    // host-mapped control block: the kernel reads the item and descriptors and writes the taps
    // straight into pinned host memory (no D2H memcpy on the latency path)
    struct Ctl
    {
        b200_trk_item item;
        ChanDesc chan;
        BandDesc band;
        float2 out[B200_MAX_TAPS];
    };
    Ctl* ctl{nullptr};
    float2* partial{nullptr};
    unsigned int* counter{nullptr};
    int slices_cap{0};
    bool have_code{false};
    // legacy variants (trk_variants.cu): complex local code / 16-bit samples and code
    float2* cplx_code_dev{nullptr};
    int cplx_code_len{0};
    short* code16_dev{nullptr};
    int code16_len{0};
    short* sig16_dev{nullptr};
    short* out16_dev{nullptr};
    float2* out_cplx_dev{nullptr};
    float var_shifts[B200_MAX_TAPS] = {0};
};

extern "C"
{
    int b200_trk_create(b200_engine* e, b200_trk** out, int max_signal_length_samples, int n_correlators)
    {
        if (!e || !out || max_signal_length_samples < 1) return B200_ERR_ARG;
        if (n_correlators < 1 || n_correlators > B200_MAX_TAPS)
            {
                set_error("n_correlators %d outside 1..%d", n_correlators, B200_MAX_TAPS);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(e->device);
        b200_trk* t = new (std::nothrow) b200_trk();
        if (!t) return B200_ERR_NOMEM;
        t->e = e;
        t->max_len = max_signal_length_samples;
        t->taps = n_correlators;
        const int rc = [&]() -> int {
            B200_CUDA_TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
            B200_CUDA_TRY(cudaMalloc(&t->sig_dev, sizeof(float2) * (static_cast<size_t>(t->max_len) + 2)));
            B200_CUDA_TRY(cudaHostAlloc(&t->ctl, sizeof(b200_trk::Ctl), cudaHostAllocMapped));
            std::memset(t->ctl, 0, sizeof(b200_trk::Ctl));
            t->slices_cap = 64;
            B200_CUDA_TRY(cudaMalloc(&t->partial, sizeof(float2) * trk_partial_elems(1, t->slices_cap)));
            B200_CUDA_TRY(cudaMalloc(&t->counter, sizeof(unsigned int)));
            B200_CUDA_TRY(cudaMemsetAsync(t->counter, 0, sizeof(unsigned int), t->stream));
            return B200_OK;
        }();
        if (rc != B200_OK)
            {
                b200_trk_destroy(t);  // tolerates a half-built correlator
                return rc;
            }
        *out = t;
        return B200_OK;
    }

    int b200_trk_set_high_dynamics_resampler(b200_trk* t, int use_high_dynamics_resampler)
    {
        if (!t) return B200_ERR_ARG;
        t->high_dyn = use_high_dynamics_resampler ? 1 : 0;
        return B200_OK;
    }

    int b200_trk_set_local_code_and_taps(b200_trk* t, int code_length_chips, const float* local_code_in, const float* shifts_chips)
    {
        if (!t || !local_code_in || !shifts_chips || code_length_chips < 1) return B200_ERR_ARG;
        B200_ENTER_DEVICE(t->e->device);
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        if (code_length_chips > t->code_cap)
            {
                if (t->code_dev) B200_CUDA_TRY(cudaFree(t->code_dev));
                t->code_cap = code_length_chips;
                B200_CUDA_TRY(cudaMalloc(&t->code_dev, sizeof(float) * t->code_cap));
            }
        B200_CUDA_TRY(cudaMemcpyAsync(t->code_dev, local_code_in, sizeof(float) * code_length_chips, cudaMemcpyHostToDevice, t->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        ChanDesc& c = t->ctl->chan;
        c.code = t->code_dev;
        c.code_len = code_length_chips;
        c.taps = t->taps;
        c.band = 0;
        for (int k = 0; k < t->taps; k++) c.shifts[k] = shifts_chips[k];
        t->have_code = true;
        return B200_OK;
    }

    int b200_trk_set_taps(b200_trk* t, const float* shifts_chips)
    {
        if (!t || !shifts_chips) return B200_ERR_ARG;
        // the descriptor lives in the host-mapped control block the kernel reads at launch; calls on one handle are
        // serialised by the caller and b200_trk_correlate is synchronous, so no launch is in flight here
        for (int k = 0; k < t->taps; k++) t->ctl->chan.shifts[k] = shifts_chips[k];
        return B200_OK;
    }

    int b200_trk_correlate(b200_trk* t, const b200_cf32* sig_in_host,
        float rem_carrier_phase_in_rad, float phase_step_rad, float phase_rate_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips,
        int signal_length_samples, b200_cf32* corr_out_host)
    {
        if (!t || !sig_in_host || !corr_out_host) return B200_ERR_ARG;
        if (!t->have_code)
            {
                set_error("correlate before set_local_code_and_taps");
                return B200_ERR_STATE;
            }
        if (signal_length_samples < 0 || signal_length_samples > t->max_len)
            {
                set_error("signal_length_samples %d outside 0..%d", signal_length_samples, t->max_len);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(t->e->device);
        const int n = signal_length_samples;
        if (n > 0) B200_CUDA_TRY(cudaMemcpyAsync(t->sig_dev, sig_in_host, sizeof(float2) * n, cudaMemcpyHostToDevice, t->stream));
        b200_trk::Ctl* c = t->ctl;
        c->chan.high_dyn = t->high_dyn;
        c->band.base = t->sig_dev;
        c->band.mask = ~0ULL;
        c->band.first_index = 0;
        c->band.limit = static_cast<unsigned long long>(t->max_len) + 2ULL;
        c->item.channel = 0;
        c->item.n = n;
        c->item.sample_index = 0;
        c->item.rem_carrier_phase_rad = rem_carrier_phase_in_rad;
        c->item.phase_step_rad = phase_step_rad;
        c->item.phase_rate_step_rad = phase_rate_step_rad;
        c->item.rem_code_phase_chips = rem_code_phase_chips;
        c->item.code_phase_step_chips = code_phase_step_chips;
        c->item.code_phase_rate_step_chips = code_phase_rate_step_chips;
        int slices = (n + 4 * kTrkTile - 1) / (4 * kTrkTile);
        if (slices < 1) slices = 1;
        if (slices > t->slices_cap) slices = t->slices_cap;
        int rc = launch_trk_batch(&c->item, 1, &c->chan, &c->band, c->out, B200_MAX_TAPS, slices, t->partial, t->counter,
            c->chan.code_len, t->taps == 1 || t->taps == 3 || t->taps == 5 ? t->taps : 0, t->stream);
        if (rc) return rc;
        {
            std::lock_guard<std::mutex> lk(t->e->mu);
            t->e->launches++;
        }
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        for (int k = 0; k < t->taps; k++)
            {
                corr_out_host[k].re = c->out[k].x;
                corr_out_host[k].im = c->out[k].y;
            }
        return B200_OK;
    }

    // ---- Cpu_Multicorrelator (complex local code), cpu_multicorrelator.cc:53-100 --------------------------------------
    int b200_trk_set_local_code_and_taps_cplx(b200_trk* t, int code_length_chips, const b200_cf32* local_code_in, const float* shifts_chips)
    {
        if (!t || !local_code_in || !shifts_chips || code_length_chips < 1) return B200_ERR_ARG;
        B200_ENTER_DEVICE(t->e->device);
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        if (code_length_chips > t->cplx_code_len)
            {
                if (t->cplx_code_dev) B200_CUDA_TRY(cudaFree(t->cplx_code_dev));
                t->cplx_code_dev = nullptr;
                B200_CUDA_TRY(cudaMalloc(&t->cplx_code_dev, sizeof(float2) * code_length_chips));
            }
        if (!t->out_cplx_dev) B200_CUDA_TRY(cudaMalloc(&t->out_cplx_dev, sizeof(float2) * B200_MAX_TAPS));
        t->cplx_code_len = code_length_chips;
        B200_CUDA_TRY(cudaMemcpyAsync(t->cplx_code_dev, local_code_in, sizeof(float2) * code_length_chips, cudaMemcpyHostToDevice, t->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        for (int k = 0; k < t->taps; k++) t->var_shifts[k] = shifts_chips[k];
        return B200_OK;
    }

    int b200_trk_correlate_cplx(b200_trk* t, const b200_cf32* sig_in_host, float rem_carrier_phase_in_rad, float phase_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, int signal_length_samples, b200_cf32* corr_out_host)
    {
        if (!t || !sig_in_host || !corr_out_host) return B200_ERR_ARG;
        if (!t->cplx_code_dev)
            {
                set_error("correlate before set_local_code_and_taps_cplx");
                return B200_ERR_STATE;
            }
        const int n = signal_length_samples;
        if (n < 0 || n > t->max_len)
            {
                set_error("signal_length_samples %d outside 0..%d", n, t->max_len);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(t->e->device);
        if (n > 0) B200_CUDA_TRY(cudaMemcpyAsync(t->sig_dev, sig_in_host, sizeof(float2) * n, cudaMemcpyHostToDevice, t->stream));
        const int rc = launch_trk_cplx_code(t->sig_dev, t->cplx_code_dev, n, t->cplx_code_len, t->taps, t->var_shifts, rem_carrier_phase_in_rad,
            phase_step_rad, rem_code_phase_chips, code_phase_step_chips, t->out_cplx_dev, t->stream);
        if (rc) return rc;
        {
            std::lock_guard<std::mutex> lk(t->e->mu);
            t->e->launches++;
        }
        B200_CUDA_TRY(cudaMemcpyAsync(corr_out_host, t->out_cplx_dev, sizeof(float2) * t->taps, cudaMemcpyDeviceToHost, t->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        return B200_OK;
    }

    // ---- Cpu_Multicorrelator_16sc (16-bit samples and code), cpu_multicorrelator_16sc.cc:47-91 -------------------------
    int b200_trk_set_local_code_and_taps_16sc(b200_trk* t, int code_length_chips, const int16_t* local_code_iq, const float* shifts_chips)
    {
        if (!t || !local_code_iq || !shifts_chips || code_length_chips < 1) return B200_ERR_ARG;
        B200_ENTER_DEVICE(t->e->device);
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        if (code_length_chips > t->code16_len)
            {
                if (t->code16_dev) B200_CUDA_TRY(cudaFree(t->code16_dev));
                t->code16_dev = nullptr;
                B200_CUDA_TRY(cudaMalloc(&t->code16_dev, sizeof(short) * 2 * code_length_chips));
            }
        if (!t->sig16_dev) B200_CUDA_TRY(cudaMalloc(&t->sig16_dev, sizeof(short) * 2 * (static_cast<size_t>(t->max_len) + 2)));
        if (!t->out16_dev) B200_CUDA_TRY(cudaMalloc(&t->out16_dev, sizeof(short) * 2 * B200_MAX_TAPS));
        t->code16_len = code_length_chips;
        B200_CUDA_TRY(cudaMemcpyAsync(t->code16_dev, local_code_iq, sizeof(short) * 2 * code_length_chips, cudaMemcpyHostToDevice, t->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        for (int k = 0; k < t->taps; k++) t->var_shifts[k] = shifts_chips[k];
        return B200_OK;
    }

    int b200_trk_correlate_16sc(b200_trk* t, const int16_t* sig_in_iq_host, float rem_carrier_phase_in_rad, float phase_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, int signal_length_samples, int16_t* corr_out_iq_host)
    {
        if (!t || !sig_in_iq_host || !corr_out_iq_host) return B200_ERR_ARG;
        if (!t->code16_dev)
            {
                set_error("correlate before set_local_code_and_taps_16sc");
                return B200_ERR_STATE;
            }
        const int n = signal_length_samples;
        if (n < 0 || n > t->max_len)
            {
                set_error("signal_length_samples %d outside 0..%d", n, t->max_len);
                return B200_ERR_RANGE;
            }
        B200_ENTER_DEVICE(t->e->device);
        if (n > 0) B200_CUDA_TRY(cudaMemcpyAsync(t->sig16_dev, sig_in_iq_host, sizeof(short) * 2 * n, cudaMemcpyHostToDevice, t->stream));
        const int rc = launch_trk_16sc(t->sig16_dev, t->code16_dev, n, t->code16_len, t->taps, t->var_shifts, rem_carrier_phase_in_rad, phase_step_rad,
            rem_code_phase_chips, code_phase_step_chips, t->out16_dev, t->stream);
        if (rc) return rc;
        {
            std::lock_guard<std::mutex> lk(t->e->mu);
            t->e->launches++;
        }
        B200_CUDA_TRY(cudaMemcpyAsync(corr_out_iq_host, t->out16_dev, sizeof(short) * 2 * t->taps, cudaMemcpyDeviceToHost, t->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(t->stream));
        return B200_OK;
    }

    int b200_trk_destroy(b200_trk* t)
    {
        if (!t) return B200_ERR_ARG;
        cudaSetDevice(t->e->device);
        if (t->stream) cudaStreamSynchronize(t->stream);
        if (t->sig_dev) cudaFree(t->sig_dev);
        if (t->code_dev) cudaFree(t->code_dev);
        if (t->ctl) cudaFreeHost(t->ctl);
        if (t->partial) cudaFree(t->partial);
        if (t->counter) cudaFree(t->counter);
        if (t->cplx_code_dev) cudaFree(t->cplx_code_dev);
        if (t->code16_dev) cudaFree(t->code16_dev);
        if (t->sig16_dev) cudaFree(t->sig16_dev);
        if (t->out16_dev) cudaFree(t->out16_dev);
        if (t->out_cplx_dev) cudaFree(t->out_cplx_dev);
        if (t->stream) cudaStreamDestroy(t->stream);
        delete t;
        return B200_OK;
    }
}
