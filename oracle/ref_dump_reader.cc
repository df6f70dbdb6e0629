// TEST INFRASTRUCTURE ONLY - exposes the reference's own Tracking_Dump_Reader
// (tests/unit-tests/signal-processing-blocks/libs/tracking_dump_reader.{h,cc}, compiled where it lies) so that
// tests/test_dump_format.py can read files written by b200_trk_dump_write with the reference's reader.
#include "tracking_dump_reader.h"

#include <cstdint>
#include <string>

extern "C"
{
    // fields per epoch, as doubles: the 24 members in the order the reader declares them
    int64_t ref_trk_dump_read(const char* filename, double* out, int64_t max_epochs)
    {
        Tracking_Dump_Reader rd;
        if (!rd.open_obs_file(std::string(filename))) return -1;
        const int64_t n = rd.num_epochs();
        int64_t k = 0;
        for (; k < n && k < max_epochs; k++)
            {
                if (!rd.read_binary_obs()) break;
                double* o = out + k * 24;
                o[0] = rd.abs_VE; o[1] = rd.abs_E; o[2] = rd.abs_P; o[3] = rd.abs_L; o[4] = rd.abs_VL;
                o[5] = rd.prompt_I; o[6] = rd.prompt_Q; o[7] = static_cast<double>(rd.PRN_start_sample_count);
                o[8] = rd.acc_carrier_phase_rad; o[9] = rd.carrier_doppler_hz; o[10] = rd.carrier_doppler_rate_hz_s;
                o[11] = rd.code_freq_chips; o[12] = rd.code_freq_rate_chips; o[13] = rd.carr_error_hz;
                o[14] = rd.carr_error_filt_hz; o[15] = rd.code_error_chips; o[16] = rd.code_error_filt_chips;
                o[17] = rd.CN0_SNV_dB_Hz; o[18] = rd.carrier_lock_test; o[19] = rd.aux1; o[20] = rd.aux2;
                o[21] = rd.PRN; o[22] = static_cast<double>(rd.TOW_ms); o[23] = rd.WN;
            }
        return k;
    }
}
