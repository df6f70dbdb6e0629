// Tracking correlator, "shared window" kernel for sm_100a.
//
// Why: the per-item kernel (trk_kernels.cu) re-reads every IQ sample from L2 once per channel.  With
// 32 channels on one band that is 6.4 GB through the L2->SM crossbar per second of signal, and the
// crossbar (~6.7 TB/s measured, profiles/roofline_traffic.json) becomes the bound.  Channels that are
// in lock on the same band correlate the SAME samples, so here one CTA serves a GROUP of up to 8 work
// items (one consumer warp each) from one copy of the samples:
//
//   producer warp   cp.async.bulk (TMA bulk copy, UBLKCP) of 512-sample tiles of the band store into a
//                   4-stage shared-memory ring, completion signalled on mbarriers (expect_tx bytes);
//   8 consumer warps each owns one (channel, epoch) item of the group: waits for the tile, reads its
//                   samples with conflict-free LDS.128, rotates, looks its three (five, ...) chips up
//                   in a warp-private shared-memory code table, accumulates E/P/L in packed f32x2
//                   registers, releases the tile.
//
// The arithmetic per sample is exactly that of trk_kernels.cu (same chip-index float sequence, same
// 64-bit fixed-point carrier phase), so results follow the same contract.  Items of a group need
// not start at the same sample: the tile range is the hull of the group and a warp only works on
// tiles that intersect its own epoch (masked at the ends).  Group membership is the caller's item
// ORDER (items 8g .. 8g+7); the host API sorts by start sample, bench.py's epoch-major order already
// has that shape.
#include "trk_device.cuh"

namespace b200
{
constexpr int kShK = 8;          // items (consumer warps) per CTA
constexpr int kShTile = 512;     // samples per tile (4 KB)
#ifndef SH_STAGES
#define SH_STAGES 4
#endif
#ifndef SH_MINB
#define SH_MINB 3
#endif
constexpr int kShStages = SH_STAGES;
// SH_STRESS (test builds only, tests/test_ring_stress_gpu.py): pseudo-random delays in the producer and in every consumer
// warp, before the tile is read and before the slot is released, so that warps drift apart by several tiles and every
// full/empty hand-over of the TMA ring is exercised with 2, 4 and 8 stages.
#ifndef SH_STRESS
#define SH_STRESS 0
#endif
#if SH_STRESS
__device__ __forceinline__ void stress_delay(unsigned int salt)
{
    unsigned int h = salt * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    if ((h & 3u) == 0u) __nanosleep(100u + ((h >> 4) & 1023u));
}
#endif
constexpr int kShTblCap = 1152;  // floats of warp-private code storage: whole-epoch table when it fits ...
constexpr int kShWin = 256;      // ... else two 256-entry windows (double-buffered, refilled per tile with cp.async)
constexpr int kShThreads = (kShK + 1) * 32;
constexpr int kShReseed = 64;    // 64-sample steps between exact phasor re-seeds (4096 samples)

namespace
{
struct __align__(128) ShSmem
{
    float2 tiles[kShStages][kShTile];
    float tbl[kShK][kShTblCap];
    unsigned long long full[kShStages];
    unsigned long long empty[kShStages];
    unsigned long long item_start[kShK];  // offset of the item's first sample in the band (band-relative)
    int item_n[kShK];
    int item_band[kShK];
    unsigned long long hull_start;
    int n_tiles;
    int share;
};

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return static_cast<unsigned int>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)  // suspend-time hint: the warp sleeps in hardware instead of spinning
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned int bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Correct-for-anything fallback for one warp: direct global loads, exact phasor per sample, integer
// modulo lookup in the global code table.  Used when an item cannot take the table path.
template <int TAPS>
__device__ void warp_correlate_general(const b200_trk_item& it, const ChanDesc& ch, const BandDesc& bd, float2 (&acc)[TAPS])
{
    const int lane = threadIdx.x & 31;
    const unsigned long long s0 = it.sample_index - bd.first_index;
    const unsigned long long T0 = turns_from_rad(-static_cast<double>(it.rem_carrier_phase_rad));
    const unsigned long long DT = turns_from_rad(-static_cast<double>(it.phase_step_rad));
    const int body = (it.n / 8) * 8;
    for (int n = lane; n < it.n; n += 32)
        {
            const float2 x = ldg_stream8(bd.base + ((s0 + static_cast<unsigned long long>(n)) & bd.mask));
            const float2 z = phasor_from_turns(T0 + DT * static_cast<unsigned long long>(n));
            const float wr = fmaf(x.x, z.x, -x.y * z.y);
            const float wi = fmaf(x.x, z.y, x.y * z.x);
            const float nf = static_cast<float>(n);
#pragma unroll
            for (int t = 0; t < TAPS; t++)
                {
                    const int idx = (n < body) ? chip_index_avx(it.code_phase_step_chips, nf, __fsub_rn(ch.shifts[t], it.rem_code_phase_chips))
                                               : chip_index_generic(it.code_phase_step_chips, nf, ch.shifts[t], it.rem_code_phase_chips);
                    const float c = ch.code[mod_pos(idx, ch.code_len)];
                    acc[t].x = fmaf(wr, c, acc[t].x);
                    acc[t].y = fmaf(wi, c, acc[t].y);
                }
        }
}

// One tile (512 samples in shared memory) for one warp: 8 steps of 64 samples, lane l takes the pair
// (2l, 2l+1) of every step.  fa/fb (sample index within the epoch, as floats) and the running phasors
// are lane state that continues seamlessly from tile to tile (tiles are contiguous).
// MASKED = the tile sticks out of [0, body): samples outside contribute zero and evaluate the chip
// index of the nearest sample inside (always within the code window).
// where a warp gets the tile's samples from: the CTA's shared ring (normal case) ...
struct SmemTileLoader
{
    const float4* p;  // tile + lane
    __device__ __forceinline__ float4 operator()(int k) const { return p[32 * k]; }
};
// ... or straight from the band store, when the group's items cannot share a window (different bands,
// epochs far apart): same arithmetic, the samples just come through L2 per item as in trk_kernels.cu.
struct GlobalTileLoader
{
    const float2* base;
    unsigned long long mask, off, limit;  // off = band offset of the tile's first sample + 2*lane (even)
    __device__ __forceinline__ float4 operator()(int k) const
    {
        const unsigned long long o = off + 64ULL * static_cast<unsigned long long>(k);
        if (mask == ~0ULL && o + 2ULL > limit) return make_float4(0.f, 0.f, 0.f, 0.f);  // linear band: stay inside
        return ldg_stream16(base + (o & mask));
    }
};

template <int TAPS, bool MASKED, class Loader>
__device__ __forceinline__ void warp_tile(const Loader& load, float vlo, float vhi, float step, const float2 (&aux2)[TAPS],
    unsigned int tbl_off, float2 Dr2, float2 Di2, float& fa, float& fb, float2& zr, float2& zi, float2 (&are)[TAPS], float2 (&aim)[TAPS])
{
    const float2 magic2 = make_float2(12582912.0f, 12582912.0f);
#pragma unroll
    for (int k = 0; k < kShTile / 64; k++)
        {
            const float4 v = load(k);
            float ua = fa, ub = fb;
            float2 xa = make_float2(v.x, v.y), xb = make_float2(v.z, v.w);
            if (MASKED)
                {
                    // [vlo, vhi] = the tile's samples that belong to the epoch (indices are exact in float);
                    // the others contribute zero and look the chip of the nearest valid sample up
                    if (fa < vlo || fa > vhi) xa = make_float2(0.f, 0.f);
                    if (fb < vlo || fb > vhi) xb = make_float2(0.f, 0.f);
                    ua = fminf(fmaxf(fa, vlo), vhi);
                    ub = fminf(fmaxf(fb, vlo), vhi);
                }
            float2 wr2, wi2;
            wr2.x = fmaf(xa.x, zr.x, -xa.y * zi.x);
            wi2.x = fmaf(xa.x, zi.x, xa.y * zr.x);
            wr2.y = fmaf(xb.x, zr.y, -xb.y * zi.y);
            wi2.y = fmaf(xb.x, zi.y, xb.y * zr.y);
            const float2 m2 = make_float2(__fmul_rn(step, ua), __fmul_rn(step, ub));
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
            const float2 t1 = __fmul2_rn(zi, Di2);
            const float2 nzr = __ffma2_rn(zr, Dr2, make_float2(-t1.x, -t1.y));
            zi = __ffma2_rn(zr, Di2, __fmul2_rn(zi, Dr2));
            zr = nzr;
            fa += 64.0f;
            fb += 64.0f;
        }
}

template <int TAPS>
__global__ void __launch_bounds__(kShThreads, SH_MINB) trk_shared_kernel(const b200_trk_item* __restrict__ items, int n_items,
    const ChanDesc* __restrict__ chans, const BandDesc* __restrict__ bands, float2* __restrict__ out, int out_stride)
{
    extern __shared__ __align__(128) unsigned char sh_raw[];
    ShSmem& sm = *reinterpret_cast<ShSmem*>(sh_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x;
    const int item_id = group * kShK + warp;
    const bool consumer = warp < kShK;
    const bool have_item = consumer && item_id < n_items;

    b200_trk_item it;
    ChanDesc const* ch = nullptr;
    BandDesc bd;
    if (have_item)
        {
            it = items[item_id];
            ch = &chans[it.channel];
            bd = bands[ch->band];
        }
    if (threadIdx.x == 0)
        {
            for (int s = 0; s < kShStages; s++)
                {
                    mbar_init(&sm.full[s], 1);
                    mbar_init(&sm.empty[s], kShK);
                }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    if (consumer && lane == 0)
        {
            sm.item_n[warp] = have_item ? it.n : 0;
            sm.item_band[warp] = have_item ? ch->band : -1;
            sm.item_start[warp] = have_item ? (it.sample_index - bd.first_index) : 0ULL;
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
            // hull of the group's sample ranges (same band only, no ring wrap inside an item's range
            // relative to the hull start, bounded sparsity) -- otherwise the warps load for themselves
            unsigned long long lo = ~0ULL, hi = 0ULL;
            int band = -1, share = 1, nmax = 0;
            for (int w = 0; w < kShK; w++)
                {
                    if (sm.item_n[w] <= 0) continue;
                    if (band < 0) band = sm.item_band[w];
                    if (sm.item_band[w] != band) share = 0;
                    lo = min(lo, sm.item_start[w]);
                    hi = max(hi, sm.item_start[w] + static_cast<unsigned long long>(sm.item_n[w]));
                    nmax = max(nmax, sm.item_n[w]);
                }
            if (band < 0) share = 0;
            lo &= ~1ULL;  // 16-byte aligned tile starts
            if (share && (hi - lo) > 3ULL * static_cast<unsigned long long>(nmax) + 2ULL * kShTile) share = 0;
            sm.hull_start = lo;
            sm.n_tiles = share ? static_cast<int>((hi - lo + kShTile - 1) / kShTile) : 0;
            sm.share = share;
        }
    __syncthreads();
    const int n_tiles = sm.n_tiles;
    const unsigned long long hull_start = sm.hull_start;

    if (!consumer)
        {
            // ---- producer warp: stream the hull through the ring ---------------------------------------
            if (lane == 0 && n_tiles > 0)
                {
                    BandDesc pb;
                    {
                        int w0 = 0;
                        while (sm.item_n[w0] <= 0) w0++;
                        pb = bands[sm.item_band[w0]];
                    }
                    const bool ring = (pb.mask != ~0ULL);
                    const unsigned long long cap = ring ? pb.mask + 1ULL : pb.limit;
                    for (int t = 0; t < n_tiles; t++)
                        {
                            const int s = t % kShStages;
                            const unsigned int par = static_cast<unsigned int>((t / kShStages) & 1);
                            mbar_wait(&sm.empty[s], par ^ 1u);
#if SH_STRESS
                            stress_delay(static_cast<unsigned int>(t) * 31u + blockIdx.x * 7u + 3u);
#endif
                            unsigned long long src = hull_start + static_cast<unsigned long long>(t) * kShTile;
                            unsigned int n1 = kShTile, n2 = 0;
                            if (ring)
                                {
                                    src &= pb.mask;
                                    if (src + kShTile > cap)
                                        {
                                            n1 = static_cast<unsigned int>(cap - src);
                                            n2 = kShTile - n1;
                                        }
                                }
                            else
                                {
                                    // linear band: never read past its end (the clipped samples are outside every item)
                                    if (src >= cap)
                                        n1 = 0;
                                    else if (src + kShTile > cap)
                                        n1 = static_cast<unsigned int>(cap - src) & ~1u;
                                }
                            mbar_arrive_expect_tx(&sm.full[s], (n1 + n2) * 8u);
                            if (n1) bulk_g2s(&sm.tiles[s][0], pb.base + src, n1 * 8u, &sm.full[s]);
                            if (n2) bulk_g2s(&sm.tiles[s][n1], pb.base, n2 * 8u, &sm.full[s]);
                        }
                }
            return;
        }

    // ---- consumer warp: one item ---------------------------------------------------------------------
    float2 acc[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) acc[t] = make_float2(0.f, 0.f);

    bool table_path = false, whole_table = false;
    unsigned long long my_hull = 0;
    int body = 0, a_i = 0, t_first = 0, t_last = 0;
    float step = 0.f;
    unsigned int tbl_off = 0;
    float2 aux2[TAPS];
    float2 Dr2, Di2, Gr2, Gi2, zr2, zi2, zr, zi;
    unsigned long long T0 = 0, DT = 0;
    if (have_item && it.n > 0)
        {
            body = (it.n / 8) * 8;
            step = it.code_phase_step_chips;
            const float rem = it.rem_code_phase_chips;
            T0 = turns_from_rad(-static_cast<double>(it.rem_carrier_phase_rad));
            DT = turns_from_rad(-static_cast<double>(it.phase_step_rad));
            float smin = ch->shifts[0], smax = ch->shifts[0];
#pragma unroll
            for (int t = 0; t < TAPS; t++)
                {
                    const float sh = ch->shifts[t];
                    const float a2 = __fsub_rn(sh, rem);
                    aux2[t] = make_float2(a2, a2);
                    smin = fminf(smin, sh);
                    smax = fmaxf(smax, sh);
                }
            // The code replica is staged per TILE: a tile spans 512*step chips (+ the tap spread), so a
            // 256-entry window per warp serves any code length (1023-chip C/A, 8184-value E1, 10230-chip L5).
            const float span_bound = ceilf(fabsf(step) * static_cast<float>(kShTile)) + ceilf(smax - smin) + 6.0f;
            table_path = !ch->high_dyn && span_bound <= static_cast<float>(kShWin) && fabsf(step) * static_cast<float>(it.n) < 3.0e6f &&
                         fabsf(smax) < 1.0e5f && fabsf(smin) < 1.0e5f && fabsf(rem) < 1.0e6f;
            // (3.0e6 + 1.0e6 + 1.0e5 < 2^22: every chip index of the epoch stays inside the range where the
            //  1.5*2^23 floor trick and the table-offset LEA of warp_tile are exact)
            if (table_path)
                {
                    // whole-epoch table when the epoch's chip range fits the warp's storage (C/A at any rate):
                    // filled once, no per-tile work.  Range from the epoch ends (index monotone in n).
                    {
                        int elo = 0x7fffffff, ehi = -0x7fffffff - 1;
                        const float nl = static_cast<float>(max(body - 1, 0));
#pragma unroll
                        for (int q = 0; q < TAPS; q++)
                            {
                                const int i0 = chip_index_avx(step, 0.f, aux2[q].x);
                                const int i1 = chip_index_avx(step, nl, aux2[q].x);
                                elo = min(elo, min(i0, i1));
                                ehi = max(ehi, max(i0, i1));
                            }
                        whole_table = (static_cast<long long>(ehi) - elo + 3) <= kShTblCap;
                        if (whole_table)
                            {
                                const int wb = elo - 1;
                                const int wspan = ehi - elo + 3;
                                const int L = ch->code_len;
                                int r = mod_pos(wb + lane, L);
                                const int stride = 32 % L;
                                for (int j = lane; j < wspan; j += 32)
                                    {
                                        sm.tbl[warp][j] = __ldg(ch->code + r);
                                        r += stride;
                                        if (r >= L) r -= L;
                                    }
                                asm("sub.u32 %0, %1, %2;" : "=r"(tbl_off) : "r"(smem_u32(&sm.tbl[warp][0])), "r"(4u * (static_cast<unsigned int>(wb) + 0x4B400000u)));
                                __syncwarp();
                            }
                    }
                    // tiles are counted from the group's hull when the window is shared, else from this item's
                    // own (even-aligned) start
                    my_hull = sm.share ? hull_start : (sm.item_start[warp] & ~1ULL);
                    a_i = static_cast<int>(sm.item_start[warp] - my_hull);
                    t_first = a_i / kShTile;
                    t_last = (a_i + body + kShTile - 1) / kShTile;
                    const float2 D = phasor_from_turns(DT * 64ULL);
                    const float2 G = phasor_from_turns(DT * static_cast<unsigned long long>(64 * kShReseed));
                    Dr2 = make_float2(D.x, D.x);
                    Di2 = make_float2(D.y, D.y);
                    Gr2 = make_float2(G.x, G.x);
                    Gi2 = make_float2(G.y, G.y);
                    // phasors of the lane's two samples at the first tile (n may be negative: modular phase)
                    const long long n0 = static_cast<long long>(t_first) * kShTile + 2 * lane - a_i;
                    const float2 za = phasor_from_turns(T0 + DT * static_cast<unsigned long long>(n0));
                    const float2 zb = phasor_from_turns(T0 + DT * static_cast<unsigned long long>(n0 + 1));
                    zr2 = make_float2(za.x, zb.x);
                    zi2 = make_float2(za.y, zb.y);
                    zr = zr2;
                    zi = zi2;
                }
        }

    float2 are[TAPS], aim[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; t++) are[t] = aim[t] = make_float2(0.f, 0.f);
    int tiles_in_group = 0;
    int next_wb = 0;
    float fa = static_cast<float>(t_first * kShTile + 2 * lane - a_i), fb = fa + 1.0f;
    // one tile of this warp's epoch; `load` says where the samples come from
    auto process_tile = [&](int t, const auto& load) {
        const int n_tile0 = t * kShTile - a_i;
        const bool interior = (n_tile0 >= 0) && (n_tile0 + kShTile <= body);
        // samples of this tile that belong to the epoch
        const int n_lo = max(n_tile0, 0), n_hi = min(n_tile0 + kShTile, body) - 1;
        const float vlo = static_cast<float>(n_lo), vhi = static_cast<float>(n_hi);
        if (!whole_table)
            {
                // sliding code window: the replica values this tile touches, staged with cp.async
                // (LDGSTS, no registers) one tile AHEAD so the L2 latency hides behind the
                // correlation of the current tile.  window(t) lives in half (t & 1) of tbl[warp].
                auto stage_window = [&](int tt) {
                    const int m0 = tt * kShTile - a_i;
                    const float lo_f = static_cast<float>(max(m0, 0)), hi_f = static_cast<float>(min(m0 + kShTile, body) - 1);
                    int wlo = 0x7fffffff, whi = -0x7fffffff - 1;
#pragma unroll
                    for (int q = 0; q < TAPS; q++)
                        {
                            const int i0 = chip_index_avx(step, lo_f, aux2[q].x);
                            const int i1 = chip_index_avx(step, hi_f, aux2[q].x);
                            wlo = min(wlo, min(i0, i1));
                            whi = max(whi, max(i0, i1));
                        }
                    const int wb = wlo - 1;
                    const int wspan = min(whi - wlo + 3, kShWin);  // <= kShWin by the span_bound test
                    const int L = ch->code_len;
                    int r = mod_pos(wb + lane, L);
                    const int stride = 32 % L;
                    float* dst = &sm.tbl[warp][(tt & 1) * kShWin];
                    for (int j = lane; j < wspan; j += 32)
                        {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + j)), "l"(ch->code + r) : "memory");
                            r += stride;
                            if (r >= L) r -= L;
                        }
                    return wb;
                };
                if (t == t_first) next_wb = stage_window(t);
                const int wb = next_wb;
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                if (t + 1 < t_last) next_wb = stage_window(t + 1);
                asm("sub.u32 %0, %1, %2;" : "=r"(tbl_off) : "r"(smem_u32(&sm.tbl[warp][(t & 1) * kShWin])), "r"(4u * (static_cast<unsigned int>(wb) + 0x4B400000u)));
            }
        if (interior)
            warp_tile<TAPS, false>(load, vlo, vhi, step, aux2, tbl_off, Dr2, Di2, fa, fb, zr, zi, are, aim);
        else
            warp_tile<TAPS, true>(load, vlo, vhi, step, aux2, tbl_off, Dr2, Di2, fa, fb, zr, zi, are, aim);
        if (++tiles_in_group == kShReseed / 8)
            {
                // re-seed: the group seed advances by G = exp(j DT 64*kShReseed) and replaces the
                // running phasor (bounds the drift of the 64 recurrence steps in between)
                const float2 t2 = __fmul2_rn(zi2, Gi2);
                const float2 ngr = __ffma2_rn(zr2, Gr2, make_float2(-t2.x, -t2.y));
                zi2 = __ffma2_rn(zr2, Gi2, __fmul2_rn(zi2, Gr2));
                zr2 = ngr;
                zr = zr2;
                zi = zi2;
                tiles_in_group = 0;
            }
        __syncwarp();
    };

    for (int t = 0; t < n_tiles; t++)
        {
            const int s = t % kShStages;
            const unsigned int par = static_cast<unsigned int>((t / kShStages) & 1);
            // Every consumer waits for every tile, also the ones it skips: a warp that ran ahead could
            // otherwise arrive twice on empty[s] within one phase and let the producer overwrite a slot
            // that a slower warp is still reading.
            mbar_wait(&sm.full[s], par);
#if SH_STRESS
            stress_delay(static_cast<unsigned int>(t) * 17u + static_cast<unsigned int>(warp) * 101u + blockIdx.x * 13u);
#endif
            if (table_path && t >= t_first && t < t_last)
                {
                    const SmemTileLoader ld{reinterpret_cast<const float4*>(&sm.tiles[s][0]) + lane};
                    process_tile(t, ld);
                }
#if SH_STRESS
            stress_delay(static_cast<unsigned int>(t) * 29u + static_cast<unsigned int>(warp) * 53u + blockIdx.x * 5u + 1u);
            __syncwarp();
#endif
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
    if (!sm.share && table_path)
        {
            // no shared window for this group: stream this item's own tiles from the band store
            for (int t = t_first; t < t_last; t++)
                {
                    const GlobalTileLoader ld{bd.base, bd.mask, my_hull + static_cast<unsigned long long>(t) * kShTile + 2ULL * lane, bd.limit};
                    process_tile(t, ld);
                }
        }
    if (have_item && it.n > 0)
        {
            if (table_path)
                {
#pragma unroll
                    for (int t = 0; t < TAPS; t++)
                        {
                            acc[t].x = are[t].x + are[t].y;
                            acc[t].y = aim[t].x + aim[t].y;
                        }
                    // tail n in [body, N): generic association, a handful of samples
                    const int n = body + lane;
                    if (n < it.n)
                        {
                            const unsigned long long s0 = it.sample_index - bd.first_index;
                            const float2 x = ldg_stream8(bd.base + ((s0 + static_cast<unsigned long long>(n)) & bd.mask));
                            const float2 z = phasor_from_turns(T0 + DT * static_cast<unsigned long long>(n));
                            const float wr = fmaf(x.x, z.x, -x.y * z.y);
                            const float wi = fmaf(x.x, z.y, x.y * z.x);
#pragma unroll
                            for (int t = 0; t < TAPS; t++)
                                {
                                    const int idx = chip_index_generic(step, static_cast<float>(n), ch->shifts[t], it.rem_code_phase_chips);
                                    const float c = __ldg(ch->code + mod_pos(idx, ch->code_len));
                                    acc[t].x = fmaf(wr, c, acc[t].x);
                                    acc[t].y = fmaf(wi, c, acc[t].y);
                                }
                        }
                }
            else
                {
                    warp_correlate_general<TAPS>(it, *ch, bd, acc);
                }
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
            if (lane < TAPS)
                {
                    float2 r = acc[0];
#pragma unroll
                    for (int t = 1; t < TAPS; t++)
                        if (lane == t) r = acc[t];
                    out[static_cast<size_t>(item_id) * out_stride + lane] = r;
                }
        }
    else if (have_item && lane < ch->taps)
        {
            out[static_cast<size_t>(item_id) * out_stride + lane] = make_float2(0.f, 0.f);
        }
}
}  // namespace

int launch_trk_shared(const b200_trk_item* items, int n_items, const ChanDesc* chans, const BandDesc* bands, float2* out,
    int out_stride, int taps_uniform, cudaStream_t stream)
{
    if (n_items <= 0) return B200_OK;
    const int groups = (n_items + kShK - 1) / kShK;
    static DeviceOnce once;  // per device: a second engine on another GPU needs its own opt-in
    const int once_dev = once.begin();
    if (once_dev >= 0)
        {
            B200_CUDA_TRY(cudaFuncSetAttribute(trk_shared_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ShSmem))));
            B200_CUDA_TRY(cudaFuncSetAttribute(trk_shared_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ShSmem))));
            B200_CUDA_TRY(cudaFuncSetAttribute(trk_shared_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ShSmem))));
            once.done(once_dev);
        }
    const size_t smem = sizeof(ShSmem);
    switch (taps_uniform)
        {
        case 1: trk_shared_kernel<1><<<groups, kShThreads, smem, stream>>>(items, n_items, chans, bands, out, out_stride); break;
        case 3: trk_shared_kernel<3><<<groups, kShThreads, smem, stream>>>(items, n_items, chans, bands, out, out_stride); break;
        case 5: trk_shared_kernel<5><<<groups, kShThreads, smem, stream>>>(items, n_items, chans, bands, out, out_stride); break;
        default: return B200_ERR_ARG;
        }
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

// Longest code table for which every epoch takes the whole-table mode (measured faster than the per-item
// kernel: 0.85 vs 0.92 ms on C2).  Longer tables (E1 8184, L5 10230) work through the sliding window but
// measure slower than the per-item kernel (C3: 5.4 vs 4.7 ms), so the engine only picks them when forced.
int trk_shared_max_code_len() { return kShTblCap - 64; }

}  // namespace b200
