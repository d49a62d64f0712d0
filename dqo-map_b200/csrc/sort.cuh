// Hand-written stable LSD radix sort (per 8-bit digit: count -> scan -> scatter, see sort.cu) and a single-pass
// decoupled look-back prefix sum for sm_100a.
//
// Replaces the library sorts / scans of the reference's binning (cub::DeviceRadixSort::SortPairs,
// RAST/cuda_rasterizer/rasterizer_impl.cu:327-336; cub::DeviceScan::InclusiveSum, :303; simple_knn.cu:237) on the hot
// path.  What the library calls cannot do and these can:
//   * the item count is read from DEVICE memory (a status word written by the kernel that produced the keys), so
//     the host never has to know R: no sentinel padding of the unused tail, no sort over the whole capacity;
//   * a "skip" word (the overflow flag) turns the whole sort into a no-op on the device;
//   * values may be implicit (value = index), which saves the iota array of the depth sort.
// Stability (equal keys keep their input order) is what makes "depth-rank emission + sort by tile id" reproduce the
// reference's (tile, depth, index) order bit for bit.
#pragma once
#include "common.cuh"

namespace dqo {

#define RS_THREADS 256
#define RS_MAX_PASSES 4
// Keys per thread: small sorts (the 1 M depth keys, the back-phase lists) are latency-bound -- half-size tiles give twice
// the blocks and half the serial ranking loop per block (-18 us on the 4-pass depth sort); large sorts amortise the
// per-block scans better with 16 (measured at 2.7 M keys: 79 vs 87 us).
#define RS_SMALL_SORT 1500000
inline int radix_items(int64_t capacity) { return capacity <= RS_SMALL_SORT ? 8 : 16; }
inline int radix_tile(int64_t capacity) { return RS_THREADS * radix_items(capacity); }

struct SortTemp {
    size_t ticket;  // u32[RS_MAX_PASSES] dynamic block ids of the scan kernel (+ padding)           } fixed position: sorts of
    size_t lb;      // u64[RS_MAX_PASSES][scan_blocks] look-back words of the scan kernel            } different sizes can share
    size_t clear_bytes; // ticket + lb: cleared once before the first sort that uses the scratch    } one scratch buffer
    size_t counts;  // u32[256][tiles] digit-major count matrix of the current pass, scanned in place
    size_t total;
    int tiles, scan_blocks;
};

inline int radix_passes(int nbits) { return nbits <= 0 ? 1 : (nbits + 7) / 8; }

// Layout of the scratch for sorts of up to `capacity` items (the worst case over both tile sizes).
inline void make_sort_temp(int64_t capacity, int nbits, SortTemp *T) {
    (void)nbits;
    const int64_t tiles8 = (capacity + RS_THREADS * 8 - 1) / (RS_THREADS * 8);
    T->tiles = (int)((capacity + radix_tile(capacity) - 1) / radix_tile(capacity));
    if (T->tiles < 1) T->tiles = 1;
    const int64_t max_tiles = tiles8 < 1 ? 1 : tiles8; // a smaller sort sharing this scratch may use 8-key tiles
    T->scan_blocks = (int)(((int64_t)256 * max_tiles + 4095) / 4096);
    size_t cur = 0;
    T->ticket = cur;
    cur += 256;
    T->lb = cur;
    cur += align_up((size_t)RS_MAX_PASSES * T->scan_blocks * 8, 256);
    T->clear_bytes = cur - T->ticket;
    T->counts = cur;
    cur += align_up((size_t)256 * max_tiles * 4, 256);
    T->total = align_up(cur, 256);
}

// Sorts (key, value) pairs by key bits [0, nbits), stable.  n = min(*count, capacity) (count == nullptr: capacity), 0 when
// *skip != 0.  The data ping-pongs between the two buffer pairs: with an even number of passes the result ends up in
// (keys_a, vals_a), with an odd number in (keys_b, vals_b) -- see radix_result_in_a().  vals_a == nullptr on input means
// value = index (vals_b and, for an even number of passes, a scratch vals_a are still needed: pass it as vals_scratch).
// `temp` must hold make_sort_temp(capacity, nbits).total bytes.  Enqueues 1 memset + 3 kernels per pass on `stream`.
// `clear` = false: the caller has zeroed the scratch words itself (sort_clear_region) before the first sort that uses
// `temp`; several sorts may then share one `temp` as long as their passes use disjoint slots [pass_slot0, pass_slot0 +
// passes) of the RS_MAX_PASSES available -- no memset node between the kernels that produce the keys and the sort.
// `temp_capacity` (0: = capacity) is the capacity the scratch was laid out for (make_sort_temp); it fixes the offsets.
template <typename KeyT>
int radix_sort_pairs(KeyT *keys_a, KeyT *keys_b, uint32_t *vals_a, uint32_t *vals_b, bool implicit_vals,
                     const int *count, const int *skip, int64_t capacity, int nbits, void *temp, cudaStream_t stream,
                     int pass_slot0 = 0, bool clear = true, int64_t temp_capacity = 0);
inline void sort_clear_region(void *temp, int64_t capacity, int nbits, void **ptr, size_t *bytes) {
    SortTemp T;
    make_sort_temp(capacity, nbits, &T);
    *ptr = (char *)temp + T.ticket;
    *bytes = T.clear_bytes;
}

inline bool radix_result_in_a(int nbits) { return radix_passes(nbits) % 2 == 0; }

// ---- single-pass prefix sum building blocks (used inside the emission kernels) -------------------------------------
// look-back word: flag in the high 32 bits (0 = empty, 1 = block aggregate, 2 = inclusive prefix), value in the low 32
#define LB_PARTIAL 1ull
#define LB_INCLUSIVE 2ull
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long flag, uint32_t value) {
    *reinterpret_cast<volatile unsigned long long *>(p) = (flag << 32) | value;
}
// Exclusive prefix of block `tile` given every earlier block's aggregate; executed by one full warp.  Block j publishes
// lb_store(&status[j], LB_PARTIAL, aggregate) before calling this, and LB_INCLUSIVE afterwards.
__device__ __forceinline__ uint32_t lb_exclusive_prefix(const unsigned long long *status, int tile, int lane) {
    uint32_t excl = 0;
    int j = tile - 1;
    while (j >= 0) {
        const int idx = j - lane;
        unsigned long long s = (idx >= 0) ? lb_load(&status[idx]) : (LB_INCLUSIVE << 32);
        while (__any_sync(0xFFFFFFFFu, (s >> 32) == 0)) {
            if ((s >> 32) == 0) s = lb_load(&status[idx]);
        }
        const unsigned incl = __ballot_sync(0xFFFFFFFFu, (s >> 32) == LB_INCLUSIVE);
        const int first = incl ? (__ffs(incl) - 1) : 31;
        uint32_t v = (lane <= first) ? (uint32_t)s : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        excl += v;
        if (incl) break;
        j -= 32;
    }
    return excl;
}

} // namespace dqo
