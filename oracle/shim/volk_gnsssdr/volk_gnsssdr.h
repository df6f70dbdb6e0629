/*
 * TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Stand-in for the Mako-generated dispatcher header <volk_gnsssdr/volk_gnsssdr.h>
 * (generated upstream from tmpl/volk_gnsssdr.tmpl.h:64-82, which needs Mako and
 * the reference's build system; neither is run here).
 *
 *  - C translation units (oracle/ref_kernels.c) include the reference kernel
 *    headers *in place* under /root/reference (via -I); those only need the
 *    common/complex/malloc headers and the intrinsics, which this shim forwards.
 *  - C++ translation units (the reference's own cpu_multicorrelator_real_codes.cc,
 *    compiled where it lies) see the dispatcher as what it really is upstream:
 *    mutable global C function pointers (tmpl/volk_gnsssdr.tmpl.c:146-184).
 *    They are defined in oracle/ref_kernels.c and re-pointed by
 *    ref_select_arch().
 *
 * No reference source is copied.
 */
#ifndef B200_ORACLE_VOLK_GNSSSDR_SHIM_H
#define B200_ORACLE_VOLK_GNSSSDR_SHIM_H

#include <immintrin.h>
#include <stdint.h>
#include <volk_gnsssdr/volk_gnsssdr_common.h>
#include <volk_gnsssdr/volk_gnsssdr_complex.h>
#include <volk_gnsssdr/volk_gnsssdr_malloc.h>

__VOLK_DECL_BEGIN

typedef void (*p_32f_xn_resampler_32f_xn)(float** result, const float* local_code, float rem_code_phase_chips, float code_phase_step_chips, float* shifts_chips, unsigned int code_length_chips, int num_out_vectors, unsigned int num_points);
typedef void (*p_32f_xn_high_dynamics_resampler_32f_xn)(float** result, const float* local_code, float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips, float* shifts_chips, unsigned int code_length_chips, int num_out_vectors, unsigned int num_points);
typedef void (*p_32fc_32f_rotator_dot_prod_32fc_xn)(lv_32fc_t* result, const lv_32fc_t* in_common, const lv_32fc_t phase_inc, lv_32fc_t* phase, const float** in_a, int num_a_vectors, unsigned int num_points);
typedef void (*p_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn)(lv_32fc_t* result, const lv_32fc_t* in_common, const lv_32fc_t phase_inc, const lv_32fc_t phase_inc_rate, lv_32fc_t* phase, const float** in_a, int num_a_vectors, unsigned int num_points);

extern p_32f_xn_resampler_32f_xn volk_gnsssdr_32f_xn_resampler_32f_xn;
extern p_32f_xn_high_dynamics_resampler_32f_xn volk_gnsssdr_32f_xn_high_dynamics_resampler_32f_xn;
extern p_32fc_32f_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn;
extern p_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_32f_high_dynamic_rotator_dot_prod_32fc_xn;


/* dispatchers used by the reference's BLOCKS and the 16-bit / complex-code correlators (oracle/ref_kernels_blocks.c) */
typedef void (*p_s32f_sincos_32fc)(lv_32fc_t* out, const float phase_inc, float* phase, unsigned int num_points);
typedef void (*p_32f_index_max_32u)(uint32_t* target, const float* src0, uint32_t num_points);
typedef void (*p_16ic_convert_32fc)(lv_32fc_t* outputVector, const lv_16sc_t* inputVector, unsigned int num_points);
typedef void (*p_16ic_xn_resampler_16ic_xn)(lv_16sc_t** result, const lv_16sc_t* local_code, float rem_code_phase_chips, float code_phase_step_chips, float* shifts_chips, unsigned int code_length_chips, int num_out_vectors, unsigned int num_points);
typedef void (*p_16ic_x2_rotator_dot_prod_16ic_xn)(lv_16sc_t* result, const lv_16sc_t* in_common, const lv_32fc_t phase_inc, lv_32fc_t* phase, const lv_16sc_t** in_a, int num_a_vectors, unsigned int num_points);
typedef void (*p_32fc_xn_resampler_32fc_xn)(lv_32fc_t** result, const lv_32fc_t* local_code, float rem_code_phase_chips, float code_phase_step_chips, float* shifts_chips, unsigned int code_length_chips, int num_out_vectors, unsigned int num_points);
typedef void (*p_32fc_x2_rotator_dot_prod_32fc_xn)(lv_32fc_t* result, const lv_32fc_t* in_common, const lv_32fc_t phase_inc, lv_32fc_t* phase, const lv_32fc_t** in_a, int num_a_vectors, unsigned int num_points);
extern p_s32f_sincos_32fc volk_gnsssdr_s32f_sincos_32fc;
extern p_32f_index_max_32u volk_gnsssdr_32f_index_max_32u;
extern p_16ic_convert_32fc volk_gnsssdr_16ic_convert_32fc;
extern p_16ic_xn_resampler_16ic_xn volk_gnsssdr_16ic_xn_resampler_16ic_xn;
extern p_16ic_x2_rotator_dot_prod_16ic_xn volk_gnsssdr_16ic_x2_rotator_dot_prod_16ic_xn;
extern p_32fc_xn_resampler_32fc_xn volk_gnsssdr_32fc_xn_resampler_32fc_xn;
extern p_32fc_x2_rotator_dot_prod_32fc_xn volk_gnsssdr_32fc_x2_rotator_dot_prod_32fc_xn;

size_t volk_gnsssdr_get_alignment(void);

__VOLK_DECL_END

#endif
