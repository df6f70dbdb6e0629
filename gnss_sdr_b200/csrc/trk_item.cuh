// Per-item correlation of the multi-tap tracking correlator (device templates shared by the batch kernel in
// trk_kernels.cu and the persistent free-running kernel in loop_kernels_nofma.cu).  See trk_kernels.cu for the
// mapping to the reference's volk_gnsssdr kernels.  Every floating-point contraction in here is written
// explicitly (fmaf / __fmul_rn / __fadd_rn / packed intrinsics), so the arithmetic does not depend on the
// translation unit's --fmad setting; tests/test_loop_gpu.py checks the two instantiations bit for bit.
#pragma once

#include "common.cuh"
#include "trk_device.cuh"

// tiles of the packed main loop unrolled per thread (one 16-byte load each): memory-level parallelism vs registers.  Measured
// on B200 (tools/ab_item.sh): with five taps (C3, Galileo E1) 4.66 ms at 4, 3.92 at 8, 3.89 at 16 (= a whole re-seed group:
// the kernel is waiting on L2, more loads in flight is what it wants); with three taps (distinct-IQ C2) 0.612 ms at 4, 0.633 at
// 8, 0.908 at 2; the one-tap data correlator of a tracked pilot (C3 + pilot) 6.50 ms at 4, 6.18 at 8, 6.14 at 16.
#ifndef TRK_TILE_UNROLL_3
#define TRK_TILE_UNROLL_3 4
#endif
#ifndef TRK_TILE_UNROLL_5
#define TRK_TILE_UNROLL_5 16
#endif
#ifndef TRK_TILE_UNROLL_1
#define TRK_TILE_UNROLL_1 16
#endif

namespace b200
{
namespace
{
constexpr int kTileUnroll3 = TRK_TILE_UNROLL_3;   // (#pragma unroll takes constant expressions, not macros)
constexpr int kTileUnroll5 = TRK_TILE_UNROLL_5;
constexpr int kTileUnroll1 = TRK_TILE_UNROLL_1;   // one tap: the data prompt of a tracked pilot

struct ItemCtx
{
    const float2* base;
    unsigned long long mask;
    unsigned long long s0;  // offset of the epoch's first sample in the band
    unsigned long long T0;  // -rem_carrier in turns
    unsigned long long DT;  // -phase_step in turns
    float step, rem;
    int N, body;            // body = 8*(N/8): samples using the AVX association
};

// Code lookup policies -----------------------------------------------------------------------
// FAST: extended table in smem covering [tbl_base, tbl_base+span): no modulo in the loop.
struct LookupExt
{
    const float* tb;  // smem_table - tbl_base
    __device__ __forceinline__ float operator()(int idx) const { return tb[idx]; }
};
// GENERAL: any index range (multi-period epochs, pathological parameters): integer modulo,
// table in smem when it fits, else global.
struct LookupMod
{
    const float* tbl;
    int L;
    __device__ __forceinline__ float operator()(int idx) const { return tbl[mod_pos(idx, L)]; }
};

template <int TAPS, class Lookup>
__device__ __forceinline__ void accumulate_sample(float2 x, float2 z, float m, const float (&aux2)[TAPS],
    const Lookup& lut, float2 (&acc)[TAPS])
{
    const float wr = fmaf(x.x, z.x, -x.y * z.y);
    const float wi = fmaf(x.x, z.y, x.y * z.x);
#pragma unroll
    for (int t = 0; t < TAPS; t++)
        {
            const int idx = __float2int_rd(__fadd_rn(m, aux2[t]));
            const float c = lut(idx);
            acc[t].x = fmaf(wr, c, acc[t].x);
            acc[t].y = fmaf(wi, c, acc[t].y);
        }
}

template <int TAPS, class Lookup, int THREADS = kTrkThreads>
__device__ __forceinline__ void correlate_range(const ItemCtx& cx, const float (&shifts)[TAPS], const Lookup& lut,
    int tile_begin, int tile_end, int head, bool do_remainder, int n_main_end, float2 (&acc)[TAPS])
{
    constexpr int kTile = 2 * THREADS;  // samples per CTA iteration (one LDG.128 per thread)
    const int tid = threadIdx.x;
    float aux2[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) aux2[t] = __fsub_rn(shifts[t], cx.rem);

    // ---- main tiles: all samples < body, 16-byte aligned pairs -------------------------------
    const float2 D = phasor_from_turns(cx.DT * static_cast<unsigned long long>(kTile));
    for (int tg = tile_begin; tg < tile_end; tg += kTrkReseed)
        {
            const int tg_end = min(tg + kTrkReseed, tile_end);
            int n0 = head + tg * kTile + 2 * tid;
            float2 za = phasor_from_turns(cx.T0 + cx.DT * static_cast<unsigned long long>(n0));
            float2 zb = phasor_from_turns(cx.T0 + cx.DT * static_cast<unsigned long long>(n0 + 1));
            float nf = static_cast<float>(n0);
#pragma unroll 4
            for (int tile = tg; tile < tg_end; tile++)
                {
                    const float4 v = ldg_stream16(cx.base + ((cx.s0 + static_cast<unsigned long long>(n0)) & cx.mask));
                    accumulate_sample<TAPS>(make_float2(v.x, v.y), za, __fmul_rn(cx.step, nf), aux2, lut, acc);
                    accumulate_sample<TAPS>(make_float2(v.z, v.w), zb, __fmul_rn(cx.step, nf + 1.0f), aux2, lut, acc);
                    za = cmulf(za, D);
                    zb = cmulf(zb, D);
                    n0 += kTile;
                    nf += static_cast<float>(kTile);
                }
        }

    // ---- remainder: optional head sample 0 and everything from n_main_end to N ------------------
    if (do_remainder)
        {
            const int count = head + (cx.N - n_main_end);
            for (int r = tid; r < count; r += THREADS)
                {
                    const int n = (r < head) ? 0 : n_main_end + (r - head);
                    const float2 x = ldg_stream8(cx.base + ((cx.s0 + static_cast<unsigned long long>(n)) & cx.mask));
                    const float2 z = phasor_from_turns(cx.T0 + cx.DT * static_cast<unsigned long long>(n));
                    const float wr = fmaf(x.x, z.x, -x.y * z.y);
                    const float wi = fmaf(x.x, z.y, x.y * z.x);
                    const float nf = static_cast<float>(n);
#pragma unroll
                    for (int t = 0; t < TAPS; t++)
                        {
                            const int idx = (n < cx.body) ? chip_index_avx(cx.step, nf, aux2[t])
                                                          : chip_index_generic(cx.step, nf, shifts[t], cx.rem);
                            const float c = lut(idx);
                            acc[t].x = fmaf(wr, c, acc[t].x);
                            acc[t].y = fmaf(wi, c, acc[t].y);
                        }
                }
        }
}


// ---- main tiles, FAST path: packed arithmetic on sample pairs, no modulo, no F2I -------------------
// Lane .x of every packed value belongs to sample n0, lane .y to sample n0+1 (one LDG.128).
// floor() uses the 1.5*2^23 trick: fl_rm(aux + 12582912.f) has floor(aux) in its low mantissa bits
// (exact for |aux| < 2^22, guaranteed by the caller's range check), so the table address is one LEA:
//   addr = (bits << 2) + tbl_off,  tbl_off = smem(table) - 4*(tbl_base + 0x4B400000)  (mod 2^32).
// NOTE: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit rounding
// modifiers and --fmad=false, which would change chip indices; the products step*n therefore use
// the scalar __fmul_rn (never contracted) and only the additions are packed.
template <int TAPS, bool WRAPS, int THREADS = kTrkThreads>
__device__ __forceinline__ void correlate_tiles_fast(const ItemCtx& cx, const float (&shifts)[TAPS], unsigned int tbl_off,
    int tile_begin, int tile_end, int head, float2 (&acc)[TAPS])
{
    constexpr int kTile = 2 * THREADS;  // samples per CTA iteration (one LDG.128 per thread)
    const int tid = threadIdx.x;
    float2 aux2[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++)
        {
            const float a = __fsub_rn(shifts[t], cx.rem);
            aux2[t] = make_float2(a, a);
        }
    float2 are[TAPS], aim[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) are[t] = aim[t] = make_float2(0.f, 0.f);

    const float2 magic2 = make_float2(12582912.0f, 12582912.0f);
    const float2 D = phasor_from_turns(cx.DT * static_cast<unsigned long long>(kTile));
    const float2 Dr2 = make_float2(D.x, D.x), Di2 = make_float2(D.y, D.y);
    // group-to-group phasor step: kTrkReseed tiles
    const float2 G = phasor_from_turns(cx.DT * static_cast<unsigned long long>(kTile * kTrkReseed));
    const float2 Gr2 = make_float2(G.x, G.x), Gi2 = make_float2(G.y, G.y);

    int n0 = head + tile_begin * kTile + 2 * tid;
    float2 zr2, zi2;   // phasors of the two samples at the start of the current group
    {
        const float2 za = phasor_from_turns(cx.T0 + cx.DT * static_cast<unsigned long long>(n0));
        const float2 zb = phasor_from_turns(cx.T0 + cx.DT * static_cast<unsigned long long>(n0 + 1));
        zr2 = make_float2(za.x, zb.x);
        zi2 = make_float2(za.y, zb.y);
    }
    float nfa = static_cast<float>(n0), nfb = static_cast<float>(n0 + 1);
    const float2* ptr = cx.base + ((cx.s0 + static_cast<unsigned long long>(n0)) & cx.mask);

    for (int tg = tile_begin; tg < tile_end; tg += kTrkReseed)
        {
            const int tg_end = min(tg + kTrkReseed, tile_end);
            float2 zr = zr2, zi = zi2;   // running phasors inside the group
#pragma unroll (TAPS >= 4 ? kTileUnroll5 : (TAPS == 1 ? kTileUnroll1 : kTileUnroll3))
            for (int tile = tg; tile < tg_end; tile++)
                {
                    float4 v;
                    if (WRAPS)
                        v = ldg_stream16(cx.base + ((cx.s0 + static_cast<unsigned long long>(n0)) & cx.mask));
                    else
                        v = ldg_stream16(ptr);
                    // w = x * z, scalar: the LDG.128 delivers (re,im,re,im), so packed operands would
                    // need 8 register moves per tile; scalar results land directly in register pairs
                    float2 wr2, wi2;
                    wr2.x = fmaf(v.x, zr.x, -v.y * zi.x);
                    wi2.x = fmaf(v.x, zi.x, v.y * zr.x);
                    wr2.y = fmaf(v.z, zr.y, -v.w * zi.y);
                    wi2.y = fmaf(v.z, zi.y, v.w * zr.y);
                    const float2 m2 = make_float2(__fmul_rn(cx.step, nfa), __fmul_rn(cx.step, nfb));
#pragma unroll
                    for (int t = 0; t < TAPS; t++)
                        {
                            const float2 tt = __fadd2_rd(__fadd2_rn(m2, aux2[t]), magic2);
                            const float ca = lds_f32((__float_as_uint(tt.x) << 2) + tbl_off);
                            const float cb = lds_f32((__float_as_uint(tt.y) << 2) + tbl_off);
                            const float2 c2 = make_float2(ca, cb);
                            are[t] = __ffma2_rn(wr2, c2, are[t]);
                            aim[t] = __ffma2_rn(wi2, c2, aim[t]);
                        }
                    // z *= D
                    const float2 t1 = __fmul2_rn(zi, Di2);
                    const float2 nzr = __ffma2_rn(zr, Dr2, make_float2(-t1.x, -t1.y));
                    zi = __ffma2_rn(zr, Di2, __fmul2_rn(zi, Dr2));
                    zr = nzr;
                    nfa += static_cast<float>(kTile);
                    nfb += static_cast<float>(kTile);
                    n0 += kTile;
                    ptr += kTile;
                }
            // group seed advances by G (few steps per epoch: error stays ~1e-7 per step)
            const float2 t2 = __fmul2_rn(zi2, Gi2);
            const float2 ngr = __ffma2_rn(zr2, Gr2, make_float2(-t2.x, -t2.y));
            zi2 = __ffma2_rn(zr2, Gi2, __fmul2_rn(zi2, Gr2));
            zr2 = ngr;
        }
#pragma unroll
    for (int t = 0; t < TAPS; t++)
        {
            acc[t].x += are[t].x + are[t].y;
            acc[t].y += aim[t].x + aim[t].y;
        }
}

// High-dynamics variant (a4): quadratic code phase on tap 0, other taps are circular
// integer-sample shifts of tap 0's resampled sequence; carrier has a phase-rate term that lags
// one sample (..._high_dynamic_rotator_dot_prod_32fc_xn.h:92-103).  Not the throughput path:
// one sample per thread per step, exact phasor per sample.
template <int TAPS, class Lookup, int THREADS = kTrkThreads>
__device__ __forceinline__ void correlate_range_hd(const ItemCtx& cx, float rate, unsigned long long RT,
    const float (&shifts)[TAPS], const Lookup& lut, int n_begin, int n_end, float2 (&acc)[TAPS])
{
    int shift_samples[TAPS];
    shift_samples[0] = 0;
    unsigned int ss = 0;
#pragma unroll
    for (int t = 1; t < TAPS; t++)
        {
            // (int)round((shifts[t]-shifts[t-1])/step) in double like C's round() on a float expr
            ss += static_cast<unsigned int>(static_cast<int>(round(static_cast<double>(__fdiv_rn(__fsub_rn(shifts[t], shifts[t - 1]), cx.step)))));
            shift_samples[t] = static_cast<int>(ss);
        }
    for (int n = n_begin + threadIdx.x; n < n_end; n += THREADS)
        {
            const float2 x = ldg_stream8(cx.base + ((cx.s0 + static_cast<unsigned long long>(n)) & cx.mask));
            // rate exponent: (n-1)^2 for n>=1 with the reference's uint32 wrap of k*k, as float
            unsigned long long T = cx.T0 + cx.DT * static_cast<unsigned long long>(n);
            if (n >= 1)
                {
                    const unsigned int k = static_cast<unsigned int>(n - 1);
                    const float e = static_cast<float>(k * k);
                    // RT is the per-unit rate in turns; e is an integer-valued float < 2^32
                    T += RT * static_cast<unsigned long long>(e);
                }
            const float2 z = phasor_from_turns(T);
            const float wr = fmaf(x.x, z.x, -x.y * z.y);
            const float wi = fmaf(x.x, z.y, x.y * z.x);
#pragma unroll
            for (int t = 0; t < TAPS; t++)
                {
                    unsigned int m = static_cast<unsigned int>(n) + static_cast<unsigned int>(shift_samples[t]);
                    if (m >= static_cast<unsigned int>(cx.N)) m -= static_cast<unsigned int>(cx.N);
                    // samples below 8*(N/8) follow the AVX kernel's association, the tail the generic one
                    const int idx = (m < static_cast<unsigned int>(cx.body))
                                        ? chip_index_hd_avx(cx.step, rate, static_cast<float>(m), __fsub_rn(shifts[0], cx.rem))
                                        : chip_index_hd(cx.step, rate, m, shifts[0], cx.rem);
                    const float c = lut(idx);
                    acc[t].x = fmaf(wr, c, acc[t].x);
                    acc[t].y = fmaf(wi, c, acc[t].y);
                }
        }
}

// REUSE (persistent tracker: one CTA keeps serving the same channel): tbl_cache[0..1] = [begin, end) of the chip-index
// window already staged in smem_tbl; an epoch whose window lies inside it skips the staging pass.
template <int TAPS, bool REUSE = false, int THREADS = kTrkThreads>
__device__ void process_item(const b200_trk_item& it, const ChanDesc& ch, const BandDesc& bd, float* smem_tbl,
    int tbl_cap, float2* smem_red, int slice, int slices, float2 (&result)[TAPS], int* tbl_cache = nullptr)
{
    constexpr int kTile = 2 * THREADS;  // samples per CTA iteration (one LDG.128 per thread)
    const int tid = threadIdx.x;
    ItemCtx cx;
    cx.base = bd.base;
    cx.mask = bd.mask;
    cx.s0 = it.sample_index - bd.first_index;
    cx.N = it.n;
    cx.body = (it.n / 8) * 8;
    cx.step = it.code_phase_step_chips;
    cx.rem = it.rem_code_phase_chips;
    cx.T0 = turns_from_rad(-static_cast<double>(it.rem_carrier_phase_rad));
    cx.DT = turns_from_rad(-static_cast<double>(it.phase_step_rad));
    const int L = ch.code_len;

    float shifts[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) shifts[t] = ch.shifts[t];

    float2 acc[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) acc[t] = make_float2(0.f, 0.f);

    if (ch.high_dyn)
        {
            // table of exactly L entries (smem if it fits), modulo lookup
            const bool in_smem = L <= tbl_cap;
            if (in_smem)
                {
                    for (int j = tid; j < L; j += THREADS) smem_tbl[j] = ch.code[j];
                }
            __syncthreads();
            LookupMod lut{in_smem ? smem_tbl : ch.code, L};
            const unsigned long long RT = turns_from_rad(-static_cast<double>(it.phase_rate_step_rad));
            const int per = (cx.N + slices - 1) / slices;
            const int nb = min(cx.N, slice * per), ne = min(cx.N, nb + per);
            correlate_range_hd<TAPS, LookupMod, THREADS>(cx, it.code_phase_rate_step_chips, RT, shifts, lut, nb, ne, acc);
        }
    else
        {
            // index range over the epoch (monotone in n within each association)
            long long lo = 0x7fffffff, hi = -0x7fffffff - 1LL;
            const float nl_avx = static_cast<float>(max(cx.body - 1, 0));
            const float n_last = static_cast<float>(max(cx.N - 1, 0));
            const float n_body = static_cast<float>(cx.body);
#pragma unroll
            for (int t = 0; t < TAPS; t++)
                {
                    const float a2 = __fsub_rn(shifts[t], cx.rem);
                    int v[4];
                    v[0] = chip_index_avx(cx.step, 0.f, a2);
                    v[1] = chip_index_avx(cx.step, nl_avx, a2);
                    v[2] = chip_index_generic(cx.step, n_body, shifts[t], cx.rem);
                    v[3] = chip_index_generic(cx.step, n_last, shifts[t], cx.rem);
                    // association 0 also evaluated at n = 0 (epochs shorter than 8 samples)
                    const int v4 = chip_index_generic(cx.step, 0.f, shifts[t], cx.rem);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        {
                            lo = min(lo, static_cast<long long>(v[q]));
                            hi = max(hi, static_cast<long long>(v[q]));
                        }
                    lo = min(lo, static_cast<long long>(v4));
                    hi = max(hi, static_cast<long long>(v4));
                }
            const long long tbl_base = lo - 2;
            const long long span = hi - lo + 5;

            const int head = static_cast<int>(cx.s0 & 1ULL);
            const int ntiles = (cx.body > head) ? (cx.body - head) / kTile : 0;
            const int n_main_end = head + ntiles * kTile;
            const int tb = static_cast<int>((static_cast<long long>(ntiles) * slice) / slices);
            const int te = static_cast<int>((static_cast<long long>(ntiles) * (slice + 1)) / slices);
            const bool rem_here = (slice == slices - 1);

            if (span <= static_cast<long long>(tbl_cap))
                {
                    int base_i = static_cast<int>(tbl_base);
                    int span_st = static_cast<int>(span);
                    bool stage = true;
                    if (REUSE)
                        {
                            const int cb = tbl_cache[0], ce = tbl_cache[1];
                            if (ce > cb && base_i >= cb && base_i + span_st <= ce)
                                {
                                    stage = false;
                                    base_i = cb;
                                }
                            else
                                {
                                    // a little wider than needed: the window drifts by a chip or two per epoch
                                    int margin = (tbl_cap - span_st) / 2;
                                    margin = margin > 8 ? 8 : (margin < 0 ? 0 : margin);
                                    base_i -= margin;
                                    span_st += 2 * margin;
                                }
                        }
                    if (stage)
                        {
                            int r = mod_pos(base_i + tid, L);
                            const int stride = THREADS % L;
                            for (int j = tid; j < span_st; j += THREADS)
                                {
                                    smem_tbl[j] = ch.code[r];
                                    r += stride;
                                    if (r >= L) r -= L;
                                }
                        }
                    __syncthreads();
                    if (REUSE && stage && tid == 0)
                        {
                            tbl_cache[0] = base_i;
                            tbl_cache[1] = base_i + span_st;
                        }
                    LookupExt lut{smem_tbl - base_i};
                    if (lo > -4000000LL && hi < 4000000LL)
                        {
                            // computed inside an asm so the optimiser cannot split the constant back out of
                            // the per-lookup LEA
                            unsigned int tbl_off;
                            asm("sub.u32 %0, %1, %2;"
                                : "=r"(tbl_off)
                                : "r"(static_cast<unsigned int>(__cvta_generic_to_shared(smem_tbl))),
                                  "r"(4u * (static_cast<unsigned int>(base_i) + 0x4B400000u)));
                            // ring wrap inside the epoch? (uniform per item)
                            const bool wraps = ((cx.s0 & cx.mask) + static_cast<unsigned long long>(cx.N)) > cx.mask;
                            if (wraps)
                                correlate_tiles_fast<TAPS, true, THREADS>(cx, shifts, tbl_off, tb, te, head, acc);
                            else
                                correlate_tiles_fast<TAPS, false, THREADS>(cx, shifts, tbl_off, tb, te, head, acc);
                            // remainder samples only (no main tiles) through the scalar path
                            correlate_range<TAPS, LookupExt, THREADS>(cx, shifts, lut, 0, 0, head, rem_here, n_main_end, acc);
                        }
                    else
                        {
                            correlate_range<TAPS, LookupExt, THREADS>(cx, shifts, lut, tb, te, head, rem_here, n_main_end, acc);
                        }
                }
            else
                {
                    const bool in_smem = L <= tbl_cap;
                    if (REUSE && tid == 0) tbl_cache[1] = tbl_cache[0];  // smem_tbl no longer holds an extended window
                    if (in_smem)
                        {
                            for (int j = tid; j < L; j += THREADS) smem_tbl[j] = ch.code[j];
                        }
                    __syncthreads();
                    LookupMod lut{in_smem ? smem_tbl : ch.code, L};
                    correlate_range<TAPS, LookupMod, THREADS>(cx, shifts, lut, tb, te, head, rem_here, n_main_end, acc);
                }
        }

    // ---- CTA reduction: shuffles, then 8 warp partials through shared memory -------------------
#pragma unroll
    for (int t = 0; t < TAPS; t++)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                {
                    acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, o);
                    acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, o);
                }
        }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0)
        {
#pragma unroll
            for (int t = 0; t < TAPS; t++) smem_red[warp * B200_MAX_TAPS + t] = acc[t];
        }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TAPS; t++)
        {
            float2 s = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < THREADS / 32; w++)
                {
                    const float2 p = smem_red[w * B200_MAX_TAPS + t];
                    s.x += p.x;
                    s.y += p.y;
                }
            result[t] = s;
        }
}

}  // namespace
}  // namespace b200
