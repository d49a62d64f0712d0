// Stable LSD radix sort with a device-side item count (see sort.cuh).
//
// Per 8-bit digit ONE kernel ranks, resolves the global offsets by decoupled look-back and scatters:
//   1. every 256-thread block takes a dynamic tile id (atomic ticket) and loads 4096 keys, warp-striped, so that
//      (warp, item, lane) order == memory order;
//   2. ranking inside a warp with MATCH.ANY: the lanes holding the same digit elect a leader that bumps the warp's
//      private shared-memory counter once; rank = old count + number of equal-digit lanes below;
//   3. thread d owns digit d: exclusive prefix of the 8 warp counters, block aggregate published to the look-back
//      array, predecessors summed until an inclusive prefix is met (flag and 30-bit value share one 32-bit word, so no
//      fence is needed between them);
//   4. keys and values are reordered through shared memory so that each digit's run leaves as contiguous stores.
// Keys move 2 x passes times through L2 (126 MB holds the c2 front list entirely), nothing else touches DRAM.
#include "sort.cuh"

namespace dqo {

#define RS_WARPS (RS_THREADS / 32)
#define RS_FLAG_PARTIAL 0x40000000u
#define RS_FLAG_INCLUSIVE 0x80000000u
#define RS_VALUE_MASK 0x3FFFFFFFu

__device__ __forceinline__ int64_t sort_count(const int *count, const int *skip, int64_t capacity) {
    int64_t n = capacity;
    if (count) {
        const int64_t c = (int64_t)(*reinterpret_cast<const volatile int *>(count));
        n = c < capacity ? c : capacity;
    }
    if (n < 0) n = 0;
    if (skip && *reinterpret_cast<const volatile int *>(skip)) n = 0;
    return n;
}

// all digit histograms of the input in one read of the keys
template <typename KeyT>
__global__ void __launch_bounds__(256) radix_hist_kernel(const KeyT *__restrict__ keys, const int *count, const int *skip,
                                                         int64_t capacity, int nbits, uint32_t *hist) {
    __shared__ uint32_t s_h[RS_MAX_PASSES][256];
    const int passes = (nbits + 7) / 8;
    for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const int64_t n = sort_count(count, skip, capacity);
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += stride) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < n;
        const uint32_t k = valid ? (uint32_t)keys[i] : 0u;
        for (int p = 0; p < passes; p++) {
            const int bits = min(8, nbits - 8 * p);
            const uint32_t d = (k >> (8 * p)) & ((1u << bits) - 1u);
            // warp-aggregated: sorted-ish inputs (depth keys share their top byte) would serialise 32 ways otherwise
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0x100u);
            if (valid && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_h[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += 256) {
        const uint32_t c = (&s_h[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS, 3)
    radix_onesweep_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out, const uint32_t *__restrict__ vals_in,
                          uint32_t *__restrict__ vals_out, const int *count, const int *skip, int64_t capacity, int shift,
                          int bits, const uint32_t *__restrict__ hist, uint32_t *tile_status, uint32_t *ticket) {
    __shared__ uint32_t s_warp_hist[RS_WARPS][256];
    __shared__ uint32_t s_excl[256];      // block-local position of the first key of each digit
    __shared__ uint32_t s_out_base[256];  // global position of that key minus s_excl: out = s_out_base[d] + local position
    __shared__ uint32_t s_scan[2][RS_WARPS];
    __shared__ KeyT s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t n = sort_count(count, skip, capacity);
    const int64_t tile_start = (int64_t)tile * RS_TILE;
    if (tile_start >= n) return;
    const int valid_count = (int)((n - tile_start) < RS_TILE ? (n - tile_start) : RS_TILE);
    const uint32_t mask = (1u << bits) - 1u;
    const unsigned lanes_below = (1u << lane) - 1u;

    // 1. load, warp-striped (32-bit indices relative to the tile)
    keys_in += tile_start;
    if (vals_in) vals_in += tile_start;
    const int wbase = warp * (32 * RS_ITEMS) + lane;
    uint32_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int idx = wbase + i * 32;
        k[i] = (idx < valid_count) ? (uint32_t)keys_in[idx] : 0u;
    }
    // 2. rank inside the warp, in (item, lane) order
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool valid = (wbase + i * 32) < valid_count;
        const uint32_t d = (k[i] >> shift) & mask;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, valid ? d : (0x100u | lane));
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + (uint32_t)__popc(peers);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rank[i] = (uint16_t)(old + (uint32_t)__popc(peers & lanes_below));
        __syncwarp();
    }
    // the values are not needed before the scatter: their loads overlap the look-back
    const uint32_t implicit_base = (uint32_t)tile_start;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int idx = wbase + i * 32;
        v[i] = (idx < valid_count) ? (vals_in ? vals_in[idx] : implicit_base + (uint32_t)idx) : 0u;
    }
    __syncthreads();
    // 3. thread d owns digit d
    uint32_t block_count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        const uint32_t c = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = block_count;
        block_count += c;
    }
    // publish as early as possible, then the two 256-wide exclusive scans (block-local digit offsets, global bin bases)
    uint32_t *my_status = tile_status + (size_t)tile * 256 + tid;
    if (tile == 0)
        *reinterpret_cast<volatile uint32_t *>(my_status) = RS_FLAG_INCLUSIVE | block_count;
    else
        *reinterpret_cast<volatile uint32_t *>(my_status) = RS_FLAG_PARTIAL | block_count;
    const uint32_t total_d = hist[tid];
    uint32_t a = block_count, b = total_d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xFFFFFFFFu, a, o), tb = __shfl_up_sync(0xFFFFFFFFu, b, o);
        if (lane >= o) {
            a += ta;
            b += tb;
        }
    }
    if (lane == 31) {
        s_scan[0][warp] = a;
        s_scan[1][warp] = b;
    }
    __syncthreads();
    uint32_t wa = 0, wb = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        if (w < warp) {
            wa += s_scan[0][w];
            wb += s_scan[1][w];
        }
    }
    const uint32_t excl_local = wa + a - block_count;  // keys of smaller digits in this block
    const uint32_t bin_base = wb + b - total_d;        // keys of smaller digits in the whole input
    uint32_t before = 0;                               // keys of this digit in earlier tiles
    if (tile > 0) {
        int j = tile - 1;
        while (true) {
            const uint32_t s = *reinterpret_cast<const volatile uint32_t *>(tile_status + (size_t)j * 256 + tid);
            if ((s & (RS_FLAG_PARTIAL | RS_FLAG_INCLUSIVE)) == 0) continue;
            before += s & RS_VALUE_MASK;
            if (s & RS_FLAG_INCLUSIVE) break;
            j--;
        }
        *reinterpret_cast<volatile uint32_t *>(my_status) = RS_FLAG_INCLUSIVE | (before + block_count);
    }
    s_excl[tid] = excl_local;
    s_out_base[tid] = bin_base + before - excl_local;
    __syncthreads();
    // 4. reorder through shared memory, then contiguous runs per digit
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        if ((wbase + i * 32) < valid_count) {
            const uint32_t d = (k[i] >> shift) & mask;
            const uint32_t pos = s_excl[d] + s_warp_hist[warp][d] + rank[i];
            s_keys[pos] = (KeyT)k[i];
            s_vals[pos] = v[i];
        }
    }
    __syncthreads();
    for (int j = tid; j < valid_count; j += RS_THREADS) {
        const KeyT key = s_keys[j];
        const uint32_t d = ((uint32_t)key >> shift) & mask;
        const uint32_t out = s_out_base[d] + (uint32_t)j;
        keys_out[out] = key;
        vals_out[out] = s_vals[j];
    }
}

template <typename KeyT>
int radix_sort_pairs(KeyT *keys_a, KeyT *keys_b, uint32_t *vals_a, uint32_t *vals_b, bool implicit_vals, const int *count,
                     const int *skip, int64_t capacity, int nbits, void *temp, cudaStream_t stream) {
    if (capacity <= 0) return DQO_OK;
    if (capacity >= (1ll << 30)) {
        set_error("radix_sort_pairs: at most 2^30 - 1 items");
        return DQO_ERR_INVALID_ARG;
    }
    if (nbits < 1) nbits = 1;
    if (nbits > (int)sizeof(KeyT) * 8) nbits = (int)sizeof(KeyT) * 8;
    SortTemp T;
    make_sort_temp(capacity, nbits, &T);
    const int passes = radix_passes(nbits);
    char *tp = (char *)temp;
    uint32_t *hist = (uint32_t *)(tp + T.hist), *ticket = (uint32_t *)(tp + T.ticket), *status = (uint32_t *)(tp + T.status);
    DQO_CUDA_CHECK(cudaMemsetAsync(temp, 0, T.total, stream));
    int hb = (int)((capacity + 256 * 8 - 1) / (256 * 8));
    if (hb > 148 * 4) hb = 148 * 4;
    radix_hist_kernel<KeyT><<<hb, 256, 0, stream>>>(keys_a, count, skip, capacity, nbits, hist);
    DQO_LAUNCH_CHECK("radix histogram", 0, stream);
    KeyT *kin = keys_a, *kout = keys_b;
    const uint32_t *vin = implicit_vals ? nullptr : vals_a;
    uint32_t *vout = vals_b;
    for (int p = 0; p < passes; p++) {
        const int bits = nbits - 8 * p < 8 ? nbits - 8 * p : 8;
        radix_onesweep_kernel<KeyT><<<T.tiles, RS_THREADS, 0, stream>>>(kin, kout, vin, vout, count, skip, capacity, 8 * p, bits,
                                                                        hist + 256 * p, status + (size_t)p * T.tiles * 256,
                                                                        ticket + p);
        DQO_LAUNCH_CHECK("radix onesweep", 0, stream);
        KeyT *tk = kin;
        kin = kout;
        kout = tk;
        vin = vout;
        vout = (vout == vals_b) ? vals_a : vals_b;
    }
    return DQO_OK;
}

template int radix_sort_pairs<uint16_t>(uint16_t *, uint16_t *, uint32_t *, uint32_t *, bool, const int *, const int *,
                                        int64_t, int, void *, cudaStream_t);
template int radix_sort_pairs<uint32_t>(uint32_t *, uint32_t *, uint32_t *, uint32_t *, bool, const int *, const int *,
                                        int64_t, int, void *, cudaStream_t);

} // namespace dqo

using namespace dqo;

// C-ABI entry (tests, and callers that want the sort on its own): see include/dqo_b200.h
extern "C" size_t dqo_sort_pairs_temp_bytes(int64_t capacity, int32_t key_bits) {
    SortTemp T;
    make_sort_temp(capacity, key_bits, &T);
    return T.total;
}
extern "C" int dqo_sort_pairs_u32(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, int32_t implicit_vals,
                                  const int32_t *count, const int32_t *skip, int64_t capacity, int32_t key_bits,
                                  void *temp, void *stream) {
    if (capacity < 0 || (capacity > 0 && (!keys_a || !keys_b || !vals_b || !temp)) || key_bits < 1 || key_bits > 32 ||
        (capacity > 0 && !vals_a && (!implicit_vals || radix_passes(key_bits) > 1))) {
        set_error("dqo_sort_pairs_u32: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    return radix_sort_pairs<uint32_t>(keys_a, keys_b, vals_a, vals_b, implicit_vals != 0, count, skip, capacity, key_bits, temp,
                                      (cudaStream_t)stream);
}
extern "C" int dqo_sort_pairs_u16(uint16_t *keys_a, uint16_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, int32_t implicit_vals,
                                  const int32_t *count, const int32_t *skip, int64_t capacity, int32_t key_bits,
                                  void *temp, void *stream) {
    if (capacity < 0 || (capacity > 0 && (!keys_a || !keys_b || !vals_b || !temp)) || key_bits < 1 || key_bits > 16 ||
        (capacity > 0 && !vals_a && (!implicit_vals || radix_passes(key_bits) > 1))) {
        set_error("dqo_sort_pairs_u16: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    return radix_sort_pairs<uint16_t>(keys_a, keys_b, vals_a, vals_b, implicit_vals != 0, count, skip, capacity, key_bits, temp,
                                      (cudaStream_t)stream);
}
