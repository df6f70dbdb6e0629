/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- builds into oracle/liboracle_port.so.
 *
 * Restatement of the reference's PRN replica generators, used to make synthetic
 * inputs for tests and bench.py.  The product library takes code tables as
 * arguments (exactly like Cpu_Multicorrelator_Real_Codes::set_local_code_and_taps
 * and pcps_acquisition::set_local_code) and never generates codes itself.
 *
 * Pin: the reference has no known-answer test for the C/A generator
 * (tests/unit-tests/arithmetic/code_generation_test.cc:28-49 is timing only), so
 * tests/test_codes.py pins it EXTERNALLY against IS-GPS-200 Table 3-Ia
 * "first 10 chips octal" (PRN 1 = 1440, ...), labelled as such.
 */
#include <math.h>
#include <stdint.h>

/* src/algorithms/libs/gps_sdr_signal_replica.cc:24-100 (gps_l1_ca_code_gen_int):
 * G1 = 1 + x^3 + x^10, G2 = 1 + x^2 + x^3 + x^6 + x^8 + x^9 + x^10, both seeded all-ones;
 * PRN selected by a G2 delay (table :41-44); chip = G1 xor G2_delayed, mapped to +1/-1. */
static const int32_t g2_delays[51] = {5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
    473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862,
    145, 175, 52, 21, 237, 235, 886, 657, 634, 762,
    355, 1012, 176, 603, 130, 359, 595, 68, 386};

int port_gps_l1_ca_code_gen_int(int32_t* dest, int32_t prn, uint32_t chip_shift)
{
    enum
    {
        CL = 1023
    };
    uint8_t G1[CL], G2[CL];
    uint8_t r1[10], r2[10];
    int32_t prn_idx = (prn >= 120 && prn <= 138) ? prn - 88 : prn - 1;
    if (prn_idx < 0 || prn_idx > 50) return -1;
    for (int i = 0; i < 10; i++) r1[i] = r2[i] = 1;
    for (int k = 0; k < CL; k++)
        {
            G1[k] = r1[0];
            G2[k] = r2[0];
            const uint8_t f1 = r1[7] ^ r1[0];
            const uint8_t f2 = r2[8] ^ r2[7] ^ r2[4] ^ r2[2] ^ r2[1] ^ r2[0];
            for (int j = 0; j < 9; j++)
                {
                    r1[j] = r1[j + 1];
                    r2[j] = r2[j + 1];
                }
            r1[9] = f1;
            r2[9] = f2;
        }
    uint32_t delay = (uint32_t)(CL - g2_delays[prn_idx]);
    delay += chip_shift;
    delay %= CL;
    for (uint32_t k = 0; k < CL; k++)
        {
            const uint8_t aux = G1[(k + chip_shift) % CL] ^ G2[delay];
            dest[k] = aux ? 1 : -1;
            delay++;
            delay %= CL;
        }
    return 0;
}

/* gps_sdr_signal_replica.cc:103-114 */
int port_gps_l1_ca_code_gen_float(float* dest, int32_t prn, uint32_t chip_shift)
{
    int32_t tmp[1023];
    if (port_gps_l1_ca_code_gen_int(tmp, prn, chip_shift)) return -1;
    for (int i = 0; i < 1023; i++) dest[i] = (float)tmp[i];
    return 0;
}

/* gps_sdr_signal_replica.cc:135-173 (gps_l1_ca_code_gen_complex_sampled): nearest-chip
 * upsampling with float index floor(ts*i/tc), last sample forced to the last chip; the
 * complex C/A replica is IMAGINARY, (0, +-1) (:117-129).  dest: interleaved cf32,
 * samplesPerCode = (int)(fs / (1.023e6/1023)) entries.  Returns samplesPerCode. */
int port_gps_l1_ca_code_gen_complex_sampled(float* dest_iq, uint32_t prn, int32_t sampling_freq, uint32_t chip_shift)
{
    const int32_t codeFreqBasis = 1023000;
    const int32_t codeLength = 1023;
    const float tc = 1.0F / (float)codeFreqBasis;
    const int32_t samplesPerCode = (int32_t)((double)sampling_freq / ((double)codeFreqBasis / (double)codeLength));
    const float ts = 1.0F / (float)sampling_freq;
    int32_t code[1023];
    if (port_gps_l1_ca_code_gen_int(code, (int32_t)prn, chip_shift)) return -1;
    for (int32_t i = 0; i < samplesPerCode; i++)
        {
            const int32_t codeValueIndex = (int32_t)floorf(ts * (float)i / tc);
            const int32_t v = (i == samplesPerCode - 1) ? code[codeLength - 1] : code[codeValueIndex];
            dest_iq[2 * i] = 0.0F;
            dest_iq[2 * i + 1] = (float)v;
        }
    return samplesPerCode;
}

/* src/algorithms/libs/galileo_e1_signal_replica.cc:98-108 (galileo_e1_code_gen_sinboc11_float)
 * applied to a caller-supplied +-1 primary code: each chip c -> (c, -c), 2 samples/chip.
 * (The E1 primary codes themselves are ICD memory-code tables, Galileo_E1.h; they are data,
 * not arithmetic, and are not restated here - tests use seeded random +-1 primaries.) */
int port_sinboc11_from_primary(float* dest, const int32_t* primary, uint32_t n_chips)
{
    for (uint32_t i = 0; i < n_chips; i++)
        {
            dest[2 * i] = (float)primary[i];
            dest[2 * i + 1] = -(float)primary[i];
        }
    return 0;
}
