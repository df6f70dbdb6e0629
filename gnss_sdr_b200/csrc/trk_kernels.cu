// Multi-tap tracking correlator for sm_100a.
//
// One CTA per (work item, slice).  A work item is one (channel, epoch): N complex samples of the
// band store, one code table, up to 8 taps.  The kernel fuses what the reference does in three
// passes over memory (VG = src/algorithms/libs/volk_gnsssdr_module/volk_gnsssdr):
//   a1  VG kernels/volk_gnsssdr/volk_gnsssdr_32f_xn_resampler_32f_xn.h:362-435  (code resampling,
//       never materialised here: the chip index is computed per sample in registers with the
//       SAME float32 operation order as the a_avx/u_avx kernel, so indices are bit-identical),
//   a2  VG kernels/volk_gnsssdr/volk_gnsssdr_32fc_32f_rotator_dot_prod_32fc_xn.h:155-314 (carrier
//       rotation + E/P/L dot products),
//   a3  src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.cc:103-127 (phase set-up).
//
// Data movement: each sample is read exactly once per channel with coalesced 16-byte loads
// (LDG.128 = 2 samples per lane); the code table is staged once per CTA in shared memory as an
// "extended" table that already contains the wrap-around, so the inner loop has no modulo.
// Reductions: registers -> warp shuffles -> shared memory -> (optionally) a deterministic
// cross-CTA combine when an epoch is split into slices for latency.
//
// Carrier: phase(n) = -(rem + n*step) is kept as a 64-bit fixed-point fraction of a turn, so
// every re-seed is exact to 2^-32 turn regardless of n; between re-seeds (kTrkReseed tiles) the
// phasor advances by a complex multiply.  No tensor cores: there is no dense contraction here.

#include "trk_item.cuh"

namespace b200
{
namespace
{
template <int TAPS>
__device__ __forceinline__ void run_item(const b200_trk_item& it, const ChanDesc& ch, const BandDesc& bd,
    float* smem_tbl, int tbl_cap, float2* smem_red, int* smem_flag, int item_id, int slice, int slices,
    float2* out, int out_stride, float2* partial, unsigned int* counters)
{
    float2 result[TAPS];
    process_item<TAPS>(it, ch, bd, smem_tbl, tbl_cap, smem_red, slice, slices, result);
    const int tid = threadIdx.x;
    if (slices == 1)
        {
            if (tid < TAPS) out[static_cast<size_t>(item_id) * out_stride + tid] = result[tid];
            return;
        }
    // deterministic cross-CTA combine: every slice publishes its partial; the last one to arrive
    // adds them in slice order.
    float2* my = partial + (static_cast<size_t>(item_id) * slices + slice) * B200_MAX_TAPS;
    if (tid < TAPS) my[tid] = result[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0)
        {
            const unsigned int prev = atomicAdd(&counters[item_id], 1u);
            *smem_flag = (prev == static_cast<unsigned int>(slices - 1));
        }
    __syncthreads();
    if (*smem_flag)
        {
            __threadfence();
            if (tid < TAPS)
                {
                    float2 s = make_float2(0.f, 0.f);
                    const volatile float2* p = partial + static_cast<size_t>(item_id) * slices * B200_MAX_TAPS;
                    for (int k = 0; k < slices; k++)
                        {
                            s.x += p[k * B200_MAX_TAPS + tid].x;
                            s.y += p[k * B200_MAX_TAPS + tid].y;
                        }
                    out[static_cast<size_t>(item_id) * out_stride + tid] = s;
                }
            if (tid == 0) counters[item_id] = 0u;  // ready for the next launch
        }
}

// TAPS_T > 0: every channel in the launch has exactly TAPS_T taps (specialised registers).
// TAPS_T == 0: taps read per item.
// No minimum-blocks hint by default: ptxas then settles on 64 registers = 4 CTAs per SM, which measures best by far
// (tools/ab_item.sh, C3: 3.89 ms; TRK_MIN_BLOCKS=2 -> 128 registers 4.11 ms, =3 -> 80 registers 7.37 ms).
#ifdef TRK_MIN_BLOCKS
#define TRK_LAUNCH_BOUNDS __launch_bounds__(kTrkThreads, TRK_MIN_BLOCKS)
#else
#define TRK_LAUNCH_BOUNDS __launch_bounds__(kTrkThreads)
#endif
template <int TAPS_T>
__global__ void TRK_LAUNCH_BOUNDS trk_correlate_kernel(const b200_trk_item* __restrict__ items, int n_items,
    const ChanDesc* __restrict__ chans, const BandDesc* __restrict__ bands, float2* __restrict__ out, int out_stride,
    int slices, float2* partial, unsigned int* counters, int tbl_cap, unsigned int only_mask)
{
    // only_mask != 0: a batch with several tap counts runs as one launch per specialisation; this launch serves the items whose
    // channel has (1 << taps) in the mask and leaves the others to their own launch
    extern __shared__ __align__(16) float smem[];
    float* smem_tbl = smem;
    float2* smem_red = reinterpret_cast<float2*>(smem + tbl_cap);
    int* smem_flag = reinterpret_cast<int*>(smem_red + (kTrkThreads / 32) * B200_MAX_TAPS);

    for (int w = blockIdx.x; w < n_items * slices; w += gridDim.x)
        {
            const int item_id = w / slices;
            const int slice = w - item_id * slices;
            const b200_trk_item it = items[item_id];
            const ChanDesc& ch = chans[it.channel];
            if (only_mask != 0u && ((1u << ch.taps) & only_mask) == 0u) continue;   // uniform per CTA iteration
            const BandDesc bd = bands[ch.band];
            if (w != static_cast<int>(blockIdx.x)) __syncthreads();  // smem reuse across items
            if (it.n <= 0)
                {
                    if (slice == 0 && threadIdx.x < ch.taps) out[static_cast<size_t>(item_id) * out_stride + threadIdx.x] = make_float2(0.f, 0.f);
                    continue;
                }
            if (TAPS_T > 0)
                {
                    run_item<(TAPS_T > 0 ? TAPS_T : 1)>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters);
                }
            else
                {
                    switch (ch.taps)
                        {
                        case 1: run_item<1>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 2: run_item<2>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 3: run_item<3>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 4: run_item<4>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 5: run_item<5>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 6: run_item<6>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        case 7: run_item<7>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        default: run_item<8>(it, ch, bd, smem_tbl, tbl_cap, smem_red, smem_flag, item_id, slice, slices, out, out_stride, partial, counters); break;
                        }
                }
        }
}

template <int T>
int launch_one(const b200_trk_item* items, int n_items, const ChanDesc* chans, const BandDesc* bands, float2* out,
    int out_stride, int slices, float2* partial, unsigned int* counters, int tbl_cap, size_t smem_bytes, cudaStream_t stream, unsigned int only_mask = 0u)
{
    static DeviceOnce once;
    const int once_dev = once.begin();
    if (once_dev >= 0)
        {
            B200_CUDA_TRY(cudaFuncSetAttribute(trk_correlate_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.done(once_dev);
        }
    const long long work = static_cast<long long>(n_items) * slices;
    // plain grid for moderate sizes, grid-stride beyond (keeps blockIdx math in int)
    const int grid = static_cast<int>(work < (1LL << 20) ? work : (1LL << 20));
    trk_correlate_kernel<T><<<grid, kTrkThreads, smem_bytes, stream>>>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap,
        only_mask);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
}  // namespace

size_t trk_partial_elems(int n_items, int slices)
{
    return static_cast<size_t>(n_items) * static_cast<size_t>(slices) * B200_MAX_TAPS;
}

int launch_trk_batch(const b200_trk_item* items, int n_items, const ChanDesc* chans, const BandDesc* bands,
    float2* out, int out_stride, int slices, float2* partial, unsigned int* counters,
    int max_code_len, int taps_uniform, cudaStream_t stream, unsigned int taps_mask)
{
    if (n_items <= 0) return B200_OK;
    if (slices < 1) slices = 1;
    int tbl_cap = max_code_len + kTrkTablePad;
    const int cap_limit = (200 * 1024 - 1024) / 4;
    if (tbl_cap > cap_limit) tbl_cap = cap_limit;
    tbl_cap = (tbl_cap + 3) & ~3;
    const size_t smem_bytes = static_cast<size_t>(tbl_cap) * 4 + (kTrkThreads / 32) * B200_MAX_TAPS * sizeof(float2) + 16;
    switch (taps_uniform)
        {
        case 1: return launch_one<1>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream);
        case 3: return launch_one<3>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream);
        case 5: return launch_one<5>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream);
        default: break;
        }
    // Several tap counts in one batch (a tracked pilot's 5 + 1 taps, a multi-signal receiver): one launch per specialised
    // kernel, each serving its own items - the generic kernel (tap count read per item) needs 128 registers and halves the
    // occupancy (C5 share and C3 + pilot ran on it).  Tap counts without a specialisation still take the generic kernel.
    if (taps_mask == 0u) return launch_one<0>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream);
    int rc = B200_OK;
    if (taps_mask & (1u << 1)) rc = launch_one<1>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream, 1u << 1);
    if (rc == B200_OK && (taps_mask & (1u << 3)))
        rc = launch_one<3>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream, 1u << 3);
    if (rc == B200_OK && (taps_mask & (1u << 5)))
        rc = launch_one<5>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream, 1u << 5);
    const unsigned int others = taps_mask & ~((1u << 1) | (1u << 3) | (1u << 5));
    if (rc == B200_OK && others) rc = launch_one<0>(items, n_items, chans, bands, out, out_stride, slices, partial, counters, tbl_cap, smem_bytes, stream, others);
    return rc;
}

}  // namespace b200
