/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- builds into oracle/_ref/liboracle_ref_blocks.so.
 *
 * No algorithm here.  Defines the volk_gnsssdr dispatcher pointers (upstream: Mako-generated
 * tmpl/volk_gnsssdr.tmpl.c:146-184) that the reference's BLOCKS and its 16-bit / complex-code correlators
 * call, pointing at the reference's own kernel implementations included where they lie.
 */
#include <volk_gnsssdr/volk_gnsssdr.h>

#include "volk_gnsssdr_s32f_sincos_32fc.h"
#include "volk_gnsssdr_32f_index_max_32u.h"
#include "volk_gnsssdr_16ic_convert_32fc.h"
#include "volk_gnsssdr_16ic_xn_resampler_16ic_xn.h"
#include "volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn.h"
#include "volk_gnsssdr_32fc_xn_resampler_32fc_xn.h"
#include "volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn.h"

#include <string.h>

/* the variants volk_gnsssdr_profile would pick on an AVX2 x86-64 host (widest SIMD implementation present) */
p_s32f_sincos_32fc volk_gnsssdr_s32f_sincos_32fc = volk_gnsssdr_s32f_sincos_32fc_u_avx2;
p_32f_index_max_32u volk_gnsssdr_32f_index_max_32u = volk_gnsssdr_32f_index_max_32u_u_avx;
p_16ic_convert_32fc volk_gnsssdr_16ic_convert_32fc = volk_gnsssdr_16ic_convert_32fc_u_avx2;
p_16ic_xn_resampler_16ic_xn volk_gnsssdr_16ic_xn_resampler_16ic_xn = volk_gnsssdr_16ic_xn_resampler_16ic_xn_u_avx;
p_16ic_x2_rotator_dot_prod_16ic_xn volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn = volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn_generic;
p_32fc_xn_resampler_32fc_xn volk_gnsssdr_32fc_xn_resampler_32fc_xn = volk_gnsssdr_32fc_xn_resampler_32fc_xn_u_avx;
p_32fc_x2_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn_u_avx;

/* "generic" or "simd": which implementations the block-level dispatchers use. */
int ref_blocks_select_arch(const char* arch)
{
    if (strcmp(arch, "generic") == 0)
        {
            volk_gnsssdr_s32f_sincos_32fc = volk_gnsssdr_s32f_sincos_32fc_generic;
            volk_gnsssdr_32f_index_max_32u = volk_gnsssdr_32f_index_max_32u_generic;
            volk_gnsssdr_16ic_convert_32fc = volk_gnsssdr_16ic_convert_32fc_generic;
            volk_gnsssdr_16ic_xn_resampler_16ic_xn = volk_gnsssdr_16ic_xn_resampler_16ic_xn_generic;
            volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn = volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn_generic;
            volk_gnsssdr_32fc_xn_resampler_32fc_xn = volk_gnsssdr_32fc_xn_resampler_32fc_xn_generic;
            volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn_generic;
            return 0;
        }
    if (strcmp(arch, "simd") == 0)
        {
            volk_gnsssdr_s32f_sincos_32fc = volk_gnsssdr_s32f_sincos_32fc_u_avx2;
            volk_gnsssdr_32f_index_max_32u = volk_gnsssdr_32f_index_max_32u_u_avx;
            volk_gnsssdr_16ic_convert_32fc = volk_gnsssdr_16ic_convert_32fc_u_avx2;
            volk_gnsssdr_16ic_xn_resampler_16ic_xn = volk_gnsssdr_16ic_xn_resampler_16ic_xn_u_avx;
            volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn = volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn_a_avx2;
            volk_gnsssdr_32fc_xn_resampler_32fc_xn = volk_gnsssdr_32fc_xn_resampler_32fc_xn_u_avx;
            volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn = volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn_u_avx;
            return 0;
        }
    return -1;
}
