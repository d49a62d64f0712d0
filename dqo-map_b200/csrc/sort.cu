// Stable LSD radix sort with a device-side item count (see sort.cuh).
//
// Per 8-bit digit three short kernels without any inter-block dependency chain:
//   count    every 256-thread block histograms the digit of its 4096 keys (one VOTE-based peer mask per key, the group
//            leader adds the group size to a shared-memory counter) and writes its column of the digit-major count
//            matrix counts[digit][tile];
//   scan     exclusive prefix sum over the flattened matrix (digit-major order = output order) with a single-pass
//            decoupled look-back over 4096-element blocks -- the only look-back left is over a few dozen blocks with one
//            word each, instead of 256 digit chains over hundreds of tiles;
//   scatter  every block reloads its keys (warp-striped, so that (warp, item, lane) order == memory order), ranks them
//            inside the warp (lanes with the same digit read the warp's counter, the leader bumps it), reorders keys and
//            values through shared memory so that each digit's run leaves as contiguous stores, and adds the scanned
//            offset of (digit, tile).
// A first version used the "onesweep" scheme (rank + look-back + scatter in one kernel per digit); at the sizes of this
// workload (2.5 M keys = 600 tiles, all resident at once) every tile spent most of its time summing its predecessors'
// partial counts -- 45 us per digit against 22 us for CUB -- so the look-back was taken out of the per-tile path.
// Keys move through L2 (126 MB holds the c2 front list entirely); only the first read and the last write touch DRAM.
#include "sort.cuh"

namespace dqo {

#define RS_WARPS (RS_THREADS / 32)
#define RS_SCAN_BLOCK 4096 // elements per block of the count-matrix scan (1024 threads x 4)

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

// Lanes of the warp holding the same digit as this lane (invalid lanes: only themselves).  MATCH.ANY is a slow,
// partially serialised instruction; one VOTE + LOP3 per digit bit is several times faster for <= 8 bits.  Digits are
// masked to `bits` bits, so the ballots of the unused upper bits would be no-ops: the upper four are skipped as a group.
__device__ __forceinline__ unsigned digit_peers(uint32_t d, int bits, bool valid, int lane) {
    unsigned peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const unsigned set = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
        peers &= ~(set ^ (0u - ((d >> b) & 1u)));
    }
    if (bits > 4) {
#pragma unroll
        for (int b = 4; b < 8; b++) {
            const unsigned set = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
            peers &= ~(set ^ (0u - ((d >> b) & 1u)));
        }
    }
    return valid ? peers : (1u << lane);
}

// counts[d * tiles_used + tile] = number of keys of tile `tile` whose digit is d (tiles_used = ceil(n / keys per tile))
template <typename KeyT, int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS) radix_count_kernel(const KeyT *__restrict__ keys, const int *count, const int *skip,
                                                                 int64_t capacity, int shift, int bits, int tiles,
                                                                 uint32_t *__restrict__ counts) {
    pdl_enter();
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    const int64_t tile_start = (int64_t)tile * RS_TILE;
    // the loads only depend on the capacity (the buffers are capacity-sized): they are in flight while the count arrives
    // (reading the count first costs the front tile sort 11 us at c2 and saves the oversized back region less)
    const int cap_count = (int)((capacity - tile_start) < RS_TILE ? (capacity - tile_start) : RS_TILE);
    keys += tile_start;
    const int wbase = warp * (32 * RS_ITEMS) + lane;
    uint32_t k[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int idx = wbase + i * 32;
        k[i] = (idx < cap_count) ? (uint32_t)keys[idx] : 0u;
    }
    const int64_t n = sort_count(count, skip, capacity);
    if (tile_start < n) {
        const int valid_count = (int)((n - tile_start) < RS_TILE ? (n - tile_start) : RS_TILE);
        const uint32_t mask = (1u << bits) - 1u;
        // Runs of equal digits in consecutive lanes (the high digit of tile ids, sorted inputs) are counted once by
        // their first lane; isolated repeats simply add separately.
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const bool valid = (wbase + i * 32) < valid_count;
            const uint32_t d = valid ? ((k[i] >> shift) & mask) : 0x100u;
            const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, d, 1);
            const bool head = (lane == 0) || (prev != d);
            const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
            if (head && valid) {
                const unsigned above = heads & ~((2u << lane) - 1u); // heads in higher lanes
                const int end = above ? (__ffs(above) - 1) : 32;
                atomicAdd(&s_hist[d], (uint32_t)(end - lane));
            }
        }
    }
    __syncthreads();
    // the count matrix is laid out for the tiles that hold keys (digit-major, row stride = used tiles), not for the
    // capacity: a generously sized region costs empty blocks, not scan work
    const int tiles_used = (int)((n + RS_TILE - 1) / RS_TILE);
    if (tile < tiles_used) counts[(size_t)tid * tiles_used + tile] = s_hist[tid];
}

// in-place exclusive prefix sum of data[0, n): single pass, decoupled look-back (sort.cuh) over blocks of 4096 elements
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t *data, const int *count, const int *skip, int64_t capacity,
                                                          int keys_per_tile, unsigned long long *lb, uint32_t *ticket) {
    pdl_enter();
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_excl, s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t items = sort_count(count, skip, capacity);
    const int64_t n = 256 * ((items + keys_per_tile - 1) / keys_per_tile); // the used part of the count matrix
    if ((int64_t)tile * RS_SCAN_BLOCK >= n) return; // (every later ticket is out of range too: nobody waits for this block)
    const int64_t base = (int64_t)tile * RS_SCAN_BLOCK + tid * 4;
    uint32_t v[4] = {0, 0, 0, 0};
    if (base + 3 < n) {
        const uint4 t = *reinterpret_cast<const uint4 *>(data + base);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (base + q < n) v[q] = data[base + q];
    }
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w; // inclusive over warps
        const uint32_t block_total = __shfl_sync(0xFFFFFFFFu, w, 31);
        if (lane == 0) lb_store(&lb[tile], tile == 0 ? LB_INCLUSIVE : LB_PARTIAL, block_total);
        uint32_t excl = 0;
        if (tile > 0) {
            excl = lb_exclusive_prefix(lb, tile, lane);
            if (lane == 0) lb_store(&lb[tile], LB_INCLUSIVE, excl + block_total);
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    uint32_t run = s_excl + (warp > 0 ? s_warp[warp - 1] : 0u) + incl - mine;
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        o[q] = run;
        run += v[q];
    }
    if (base + 3 < n) {
        *reinterpret_cast<uint4 *>(data + base) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (base + q < n) data[base + q] = o[q];
    }
}

template <typename KeyT, int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, 3)
    radix_scatter_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out, const uint32_t *__restrict__ vals_in,
                         uint32_t *__restrict__ vals_out, const int *count, const int *skip, int64_t capacity, int shift,
                         int bits, int tiles, const uint32_t *__restrict__ offsets) {
    pdl_enter();
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    __shared__ uint32_t s_warp_hist[RS_WARPS][256];
    __shared__ uint32_t s_excl[256];      // block-local position of the first key of each digit
    __shared__ uint32_t s_out_base[256];  // global position of that key minus s_excl: out = s_out_base[d] + local position
    __shared__ uint32_t s_scan[RS_WARPS];
    __shared__ KeyT s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t tile_start = (int64_t)tile * RS_TILE;
    // 1. load, warp-striped (32-bit indices relative to the tile).  The key loads only depend on the capacity (the
    // buffers are capacity-sized), so they are in flight while the item count arrives.
    const int cap_count = (int)((capacity - tile_start) < RS_TILE ? (capacity - tile_start) : RS_TILE);
    keys_in += tile_start;
    if (vals_in) vals_in += tile_start;
    const int wbase = warp * (32 * RS_ITEMS) + lane;
    uint32_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    uint32_t info[RS_ITEMS]; // bits 0-7: equal-digit lanes below, 8-15: equal-digit lanes in the warp, 16: group leader;
                             // after the serial part: the rank of the key among its warp's keys of the same digit
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int idx = wbase + i * 32;
        k[i] = (idx < cap_count) ? (uint32_t)keys_in[idx] : 0u;
    }
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
    const int64_t n = sort_count(count, skip, capacity);
    if (tile_start >= n) return;
    const int tiles_used = (int)((n + RS_TILE - 1) / RS_TILE);
    const uint32_t my_offset = offsets[(size_t)tid * tiles_used + tile]; // global position of this tile's first key of digit tid
    __syncthreads();
    const int valid_count = (int)((n - tile_start) < RS_TILE ? (n - tile_start) : RS_TILE);
    const uint32_t mask = (1u << bits) - 1u;
    const unsigned lanes_below = (1u << lane) - 1u;
    // 2. rank inside the warp, in (item, lane) order: the peer masks first (independent of the counters), then the
    // serial part, one LDS + one STS per item
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool valid = (wbase + i * 32) < valid_count;
        const uint32_t d = (k[i] >> shift) & mask;
        const unsigned peers = digit_peers(d, bits, valid, lane);
        const bool leader = valid && ((peers & lanes_below) == 0);
        info[i] = (uint32_t)__popc(peers & lanes_below) | ((uint32_t)__popc(peers) << 8) | (leader ? 0x10000u : 0u);
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const bool valid = (wbase + i * 32) < valid_count;
        const uint32_t d = (k[i] >> shift) & mask;
        const uint32_t old = valid ? s_warp_hist[warp][d] : 0u; // every lane of the group reads the same counter
        __syncwarp();
        if (info[i] & 0x10000u) s_warp_hist[warp][d] = old + ((info[i] >> 8) & 0xFFu);
        __syncwarp();
        info[i] = old + (info[i] & 0xFFu);
    }
    // the values are not needed before the reorder: their loads overlap the scans below
    const uint32_t implicit_base = (uint32_t)tile_start;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int idx = wbase + i * 32;
        v[i] = (idx < valid_count) ? (vals_in ? vals_in[idx] : implicit_base + (uint32_t)idx) : 0u;
    }
    __syncthreads();
    // 3. thread d owns digit d: exclusive prefix over the warps, then over the digits
    uint32_t block_count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        const uint32_t c = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = block_count;
        block_count += c;
    }
    uint32_t a = block_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xFFFFFFFFu, a, o);
        if (lane >= o) a += ta;
    }
    if (lane == 31) s_scan[warp] = a;
    __syncthreads();
    uint32_t wa = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++)
        if (w < warp) wa += s_scan[w];
    const uint32_t excl_local = wa + a - block_count; // keys of smaller digits in this block
    s_excl[tid] = excl_local;
    s_out_base[tid] = my_offset - excl_local;
    __syncthreads();
    // 4. reorder through shared memory, then contiguous runs per digit
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        if ((wbase + i * 32) < valid_count) {
            const uint32_t d = (k[i] >> shift) & mask;
            const uint32_t pos = s_excl[d] + s_warp_hist[warp][d] + info[i];
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
                     const int *skip, int64_t capacity, int nbits, void *temp, cudaStream_t stream, int pass_slot0, bool clear,
                     int64_t temp_capacity) {
    if (capacity <= 0) return DQO_OK;
    if (capacity >= (1ll << 30)) {
        set_error("radix_sort_pairs: at most 2^30 - 1 items");
        return DQO_ERR_INVALID_ARG;
    }
    if (nbits < 1) nbits = 1;
    if (nbits > (int)sizeof(KeyT) * 8) nbits = (int)sizeof(KeyT) * 8;
    if (temp_capacity < capacity) temp_capacity = capacity;
    SortTemp T; // offsets (and the look-back stride) from the capacity the scratch was laid out for, grids from this sort's
    make_sort_temp(temp_capacity, nbits, &T);
    const int lb_stride = T.scan_blocks;
    T.tiles = (int)((capacity + radix_tile(capacity) - 1) / radix_tile(capacity));
    if (T.tiles < 1) T.tiles = 1;
    T.scan_blocks = (int)(((int64_t)256 * T.tiles + 4095) / 4096);
    const int passes = radix_passes(nbits);
    char *tp = (char *)temp;
    uint32_t *counts = (uint32_t *)(tp + T.counts), *ticket = (uint32_t *)(tp + T.ticket);
    unsigned long long *lb = (unsigned long long *)(tp + T.lb);
    if (pass_slot0 < 0 || pass_slot0 + passes > RS_MAX_PASSES) {
        set_error("radix_sort_pairs: pass slots out of range");
        return DQO_ERR_INVALID_ARG;
    }
    if (clear) DQO_CUDA_CHECK(cudaMemsetAsync(tp + T.ticket, 0, T.clear_bytes, stream)); // tickets + look-back words of every pass
    KeyT *kin = keys_a, *kout = keys_b;
    const uint32_t *vin = implicit_vals ? nullptr : vals_a;
    uint32_t *vout = vals_b;
    for (int p = 0; p < passes; p++) {
        const int bits = nbits - 8 * p < 8 ? nbits - 8 * p : 8;
        if (radix_items(capacity) == 8) {
            launch_pdl(radix_count_kernel<KeyT, 8>, dim3(T.tiles), dim3(RS_THREADS), 0, stream, kin, count, skip, capacity, 8 * p,
                       bits, T.tiles, counts);
            launch_pdl(radix_scan_kernel, dim3(T.scan_blocks), dim3(1024), 0, stream, counts, count, skip, capacity,
                       RS_THREADS * 8, lb + (size_t)(pass_slot0 + p) * lb_stride, ticket + pass_slot0 + p);
            launch_pdl(radix_scatter_kernel<KeyT, 8>, dim3(T.tiles), dim3(RS_THREADS), 0, stream, kin, kout, vin, vout, count,
                       skip, capacity, 8 * p, bits, T.tiles, counts);
        } else {
            launch_pdl(radix_count_kernel<KeyT, 16>, dim3(T.tiles), dim3(RS_THREADS), 0, stream, kin, count, skip, capacity, 8 * p,
                       bits, T.tiles, counts);
            launch_pdl(radix_scan_kernel, dim3(T.scan_blocks), dim3(1024), 0, stream, counts, count, skip, capacity,
                       RS_THREADS * 16, lb + (size_t)(pass_slot0 + p) * lb_stride, ticket + pass_slot0 + p);
            launch_pdl(radix_scatter_kernel<KeyT, 16>, dim3(T.tiles), dim3(RS_THREADS), 0, stream, kin, kout, vin, vout, count,
                       skip, capacity, 8 * p, bits, T.tiles, counts);
        }
        DQO_LAUNCH_CHECK("radix pass", 0, stream);
        note_launch(2);
        KeyT *tk = kin;
        kin = kout;
        kout = tk;
        vin = vout;
        vout = (vout == vals_b) ? vals_a : vals_b;
    }
    return DQO_OK;
}

template int radix_sort_pairs<uint16_t>(uint16_t *, uint16_t *, uint32_t *, uint32_t *, bool, const int *, const int *,
                                        int64_t, int, void *, cudaStream_t, int, bool, int64_t);
template int radix_sort_pairs<uint32_t>(uint32_t *, uint32_t *, uint32_t *, uint32_t *, bool, const int *, const int *,
                                        int64_t, int, void *, cudaStream_t, int, bool, int64_t);

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
