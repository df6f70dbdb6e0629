// Acquisition dump writer: the variables pcps_acquisition::dump_results stores per acquisition
// (src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.cc:354-406), written as a MATLAB Level-5 MAT-file.
// The reference goes through matio with MAT_FT_MAT73 (HDF5 container); neither matio nor HDF5 exists in this
// image, so the same variable names, classes and shapes are written in the Level-5 format, which matio's
// Mat_Open (the reference's tests/unit-tests/signal-processing-blocks/libs/acquisition_dump_reader.cc:26-131),
// MATLAB, Octave and scipy.io.loadmat read transparently.  Host-only code: no CUDA calls.
#include "common.cuh"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace b200;

namespace
{
// MAT-file Level 5 constants (MathWorks "MAT-File Format", tables 1-1 and 1-3)
enum : uint32_t
{
    miINT8 = 1,
    miINT32 = 5,
    miUINT32 = 6,
    miSINGLE = 7,
    miUINT64 = 13,
    miMATRIX = 14
};
enum : uint32_t
{
    mxSINGLE_CLASS = 7,
    mxINT32_CLASS = 12,
    mxUINT32_CLASS = 13,
    mxUINT64_CLASS = 15
};

struct Buf
{
    std::vector<unsigned char> b;
    void put(const void* p, size_t n)
    {
        const auto* c = static_cast<const unsigned char*>(p);
        b.insert(b.end(), c, c + n);
    }
    void u32(uint32_t v) { put(&v, 4); }
    void pad8()
    {
        while (b.size() % 8) b.push_back(0);
    }
    // data element: tag (type, bytes) + payload padded to 8 bytes
    void element(uint32_t type, const void* p, uint32_t nbytes)
    {
        u32(type);
        u32(nbytes);
        put(p, nbytes);
        pad8();
    }
};

void matrix(Buf& out, const char* name, uint32_t mx_class, uint32_t mi_type, const void* data, uint32_t elem_size, uint32_t rows, uint32_t cols)
{
    Buf m;
    const uint32_t flags[2] = {mx_class, 0u};
    m.element(miUINT32, flags, 8);
    const int32_t dims[2] = {static_cast<int32_t>(rows), static_cast<int32_t>(cols)};
    m.element(miINT32, dims, 8);
    m.element(miINT8, name, static_cast<uint32_t>(std::strlen(name)));
    m.element(mi_type, data, elem_size * rows * cols);
    out.u32(miMATRIX);
    out.u32(static_cast<uint32_t>(m.b.size()));
    out.put(m.b.data(), m.b.size());
}

template <typename T>
void scalar(Buf& out, const char* name, uint32_t mx_class, uint32_t mi_type, T v)
{
    matrix(out, name, mx_class, mi_type, &v, sizeof(T), 1, 1);
}
}  // namespace

extern "C"
{
    int b200_acq_dump_write(const char* filename, const b200_acq_dump* d)
    {
        if (!filename || !d || !d->acq_grid || d->effective_fft_size == 0 || d->num_doppler_bins == 0) return B200_ERR_ARG;
        if (d->acq_grid_narrow && d->num_doppler_bins_step2 == 0) return B200_ERR_ARG;
        Buf out;
        // 128-byte header: 116 bytes of text, 8 bytes subsystem offset, version 0x0100, endian indicator "IM"
        char text[116];
        std::memset(text, ' ', sizeof(text));
        const char* banner = "MATLAB 5.0 MAT-file, Platform: b200gnss, acquisition dump (pcps_acquisition::dump_results)";
        std::memcpy(text, banner, std::strlen(banner));
        out.put(text, sizeof(text));
        const unsigned char zeros[8] = {0};
        out.put(zeros, 8);
        const uint16_t version = 0x0100;
        out.put(&version, 2);
        const char endian[2] = {'I', 'M'};
        out.put(endian, 2);
        // d_grid is arma::fmat(effective_fft_size, num_doppler_bins): column d = Doppler bin d, which is exactly
        // the engine's bins x effective_fft_size row-major grid
        matrix(out, "acq_grid", mxSINGLE_CLASS, miSINGLE, d->acq_grid, 4, d->effective_fft_size, d->num_doppler_bins);
        scalar<int32_t>(out, "doppler_max", mxINT32_CLASS, miINT32, d->doppler_max);
        scalar<int32_t>(out, "doppler_step", mxINT32_CLASS, miINT32, d->doppler_step);
        scalar<int32_t>(out, "positive_acq", mxINT32_CLASS, miINT32, d->positive_acq ? 1 : 0);
        scalar<float>(out, "acq_doppler_hz", mxSINGLE_CLASS, miSINGLE, d->acq_doppler_hz);
        scalar<float>(out, "acq_delay_samples", mxSINGLE_CLASS, miSINGLE, d->acq_delay_samples);
        scalar<float>(out, "test_statistic", mxSINGLE_CLASS, miSINGLE, d->test_statistic);
        scalar<float>(out, "threshold", mxSINGLE_CLASS, miSINGLE, d->threshold);
        scalar<float>(out, "input_power", mxSINGLE_CLASS, miSINGLE, d->input_power);
        scalar<uint64_t>(out, "sample_counter", mxUINT64_CLASS, miUINT64, d->sample_counter);
        scalar<uint32_t>(out, "PRN", mxUINT32_CLASS, miUINT32, d->prn);
        scalar<int32_t>(out, "num_dwells", mxINT32_CLASS, miINT32, d->num_dwells);
        if (d->acq_grid_narrow)
            {
                matrix(out, "acq_grid_narrow", mxSINGLE_CLASS, miSINGLE, d->acq_grid_narrow, 4, d->effective_fft_size, d->num_doppler_bins_step2);
                scalar<float>(out, "doppler_step_narrow", mxSINGLE_CLASS, miSINGLE, d->doppler_step_narrow);
                scalar<float>(out, "doppler_grid_narrow_min", mxSINGLE_CLASS, miSINGLE, d->doppler_grid_narrow_min);
            }
        FILE* f = std::fopen(filename, "wb");
        if (!f)
            {
                set_error("acq_dump_write: cannot open %s", filename);
                return B200_ERR_STATE;
            }
        const size_t w = std::fwrite(out.b.data(), 1, out.b.size(), f);
        const int bad = std::fclose(f);
        if (w != out.b.size() || bad)
            {
                set_error("acq_dump_write: short write to %s", filename);
                return B200_ERR_STATE;
            }
        return B200_OK;
    }

    // filename of dump_results (:357-370): <base>_<System>_<Sig0><Sig1>_ch_<channel>_<dump_number>_sat_<PRN>.mat
    int b200_acq_dump_filename(const char* base, char system, const char* signal2, uint32_t channel, uint32_t dump_number, uint32_t prn,
        char* out, size_t out_size)
    {
        if (!base || !signal2 || !out || out_size == 0) return B200_ERR_ARG;
        std::string s(base);
        s += "_";
        s += system;
        s += "_";
        s += signal2[0];
        s += signal2[1];
        s += "_ch_" + std::to_string(channel) + "_" + std::to_string(dump_number) + "_sat_" + std::to_string(prn) + ".mat";
        if (s.size() + 1 > out_size) return B200_ERR_RANGE;
        std::memcpy(out, s.c_str(), s.size() + 1);
        return B200_OK;
    }
}
