// Multinomial resampling with SORTED uniforms (throughput mode of SMCSamples.resample,
// /root/reference/src/aspire/samples.py:1251-1287: idx = searchsorted(cumsum(w), rng.random(M), 'right')).
//
// The reference draws M iid uniforms and binary-searches each one: M dependent chains of ~27 random reads over a
// CDF far larger than L2 (4.2 ms + 1.1 ms for a guide table at M = N = 1e8).  Only the MULTISET of ancestors
// matters - the particles are exchangeable - and the order statistics of M iid uniforms are the normalised
// partial sums of M + 1 iid exponentials.  So in throughput mode the uniforms are generated already sorted:
//
//   q_j   = floor(-ln(1 - u_j) 2^26), u_j the Philox uniform stream          (spacing_math.cuh; j = 0 .. M)
//   S_j   = q_0 + ... + q_j  (integers: exact, any scan order, any sharding),   total = S_M
//   T_k   = rint(cdf_k * total)                                               (cdf: the exact fp64 CDF, as before)
//   idx_j = #{k : T_k <= S_j}  (clamped to N - 1)                             == searchsorted(T, S, 'right')
//
// i.e. the same searchsorted-right of the same CDF, at the sorted uniforms S_j / total.  Two sorted sequences are
// merged by streaming: the slot tile of a block and the CDF tile(s) its slots fall into meet in shared memory, every
// HBM access is sequential, and the ancestors come out sorted - the gather of the following sweep reads ascending
// rows.  Slot j is keyed by its rank among the uniforms, so a GPU that owns the CDF slice [k0, k1) owns the
// contiguous slot range {j : T_{k0-1} <= S_j < T_{k1-1}} and can regenerate those spacings itself: no uniform and
// no index crosses NVLink.  Supplied uniforms / PCG64 parity mode keep the guided search (search_gather.cu).
// Spec: oracle/sorted_resample.py.
#include "common.cuh"
#include "philox.cuh"
#include "spacing_math.cuh"

namespace asp {

constexpr int kSpBlock = 256;
constexpr int kSpItems = kSpacingTile / kSpBlock;  // 8 consecutive slots per thread

__device__ __forceinline__ uint64_t warp_scan_incl_u64(uint64_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (kSpBlock threads); total returned through `total`
__device__ __forceinline__ uint64_t block_scan_excl_u64(uint64_t v, uint64_t* s_warp /*[8]*/, uint64_t& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t incl = warp_scan_incl_u64(v);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kSpBlock / 32; ++w) {
        const uint64_t t = s_warp[w];
        if (w < wid) base += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return base + incl - v;
}

// q for elements [j0, j0 + tiles * 2048) (j0 a multiple of 2048; elements > m_total get 0) and their tile sums
__global__ void __launch_bounds__(kSpBlock) spacings_kernel(PhiloxKeys K, uint32_t stream, int64_t j0, int64_t n_tiles,
                                                            int64_t m_total, uint32_t* __restrict__ q,
                                                            unsigned long long* __restrict__ tile_sums) {
    __shared__ uint64_t s_warp[kSpBlock / 32];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = j0 + tile * kSpacingTile + (int64_t)threadIdx.x * kSpItems;  // global element index, even
        uint32_t v[kSpItems];
        uint64_t sum = 0;
#pragma unroll
        for (int b = 0; b < kSpItems / 2; ++b) {
            const uint64_t blk = (uint64_t)(e0 >> 1) + b;  // Philox block of elements (2 blk, 2 blk + 1): asp_philox_uniform's layout
            const uint4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u, K);
            const uint64_t ka = ((((uint64_t)r.x) << 32) | r.y) >> 11, kb = ((((uint64_t)r.z) << 32) | r.w) >> 11;
            v[2 * b] = e0 + 2 * b <= m_total ? spacing_q(ka) : 0u;
            v[2 * b + 1] = e0 + 2 * b + 1 <= m_total ? spacing_q(kb) : 0u;
            sum += (uint64_t)v[2 * b] + v[2 * b + 1];
        }
        uint4* dst = reinterpret_cast<uint4*>(q + tile * kSpacingTile + threadIdx.x * kSpItems);
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
        uint64_t tot;
        (void)block_scan_excl_u64(sum, s_warp, tot);
        if (threadIdx.x == 0) tile_sums[tile] = tot;
    }
}

// exclusive prefix of the tile sums and the grand total (one block; integer sums: the order is immaterial).
// Thread t owns the contiguous chunk [t per, (t + 1) per); loads are issued eight at a time so that the chunk's
// latency is paid once, not per element.
__global__ void __launch_bounds__(1024) tile_scan_kernel(const unsigned long long* __restrict__ sums, int64_t n_tiles,
                                                         unsigned long long* __restrict__ prefix,
                                                         unsigned long long* __restrict__ total) {
    __shared__ uint64_t s_warp[32];
    const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const int64_t a = (int64_t)threadIdx.x * per, b = a + per < n_tiles ? a + per : n_tiles;
    uint64_t s = 0;
    for (int64_t t = a; t < b; t += 8) {
        uint64_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = t + i < b ? __ldg(sums + t + i) : 0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
    }
    // block exclusive scan of the chunk sums
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t incl = warp_scan_incl_u64(s);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint64_t w = s_warp[lane];
        const uint64_t wi = warp_scan_incl_u64(w);
        s_warp[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    uint64_t run = s_warp[wid] + incl - s;
    for (int64_t t = a; t < b; t += 8) {
        uint64_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = t + i < b ? __ldg(sums + t + i) : 0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (t + i < b) prefix[t + i] = run;
            run += v[i];
        }
    }
}

// largest CDF tile ct in [0, n_ct) whose first threshold is <= s (0 if none)
__device__ __forceinline__ int64_t coarse_tile(const double* __restrict__ cdf, int64_t n_ct, double total, uint64_t s) {
    int64_t lo = 0, hi = n_ct;  // answer in [lo, hi)
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (spacing_threshold(__ldg(cdf + mid * kSpacingTile), total) <= s) lo = mid;
        else hi = mid;
    }
    return lo;
}

// tile_ct[b] = CDF tile holding the ancestor of the first slot of slot tile j_lo / 2048 + b (one thread per slot
// tile: the dependent binary-search chain of every tile runs in parallel instead of at the head of its block)
__global__ void __launch_bounds__(kSpBlock) slot_tile_map_kernel(const double* __restrict__ cdf, int64_t n_cdf,
                                                                 const uint32_t* __restrict__ q, int64_t q_j0,
                                                                 const unsigned long long* __restrict__ tile_prefix,
                                                                 const unsigned long long* __restrict__ total_p,
                                                                 int64_t j_lo, int64_t n_slot_tiles,
                                                                 int32_t* __restrict__ tile_ct) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_slot_tiles) return;
    const double total = (double)*total_p;
    const int64_t n_ct = (n_cdf + kSpacingTile - 1) / kSpacingTile;
    const int64_t gt = j_lo / kSpacingTile + b;
    // partial sum of the tile's first slot that is >= j_lo: a lower bound (the tile's first slot) is enough for a
    // starting tile - the search loop moves forward from it
    const uint64_t s = tile_prefix[gt] + q[gt * kSpacingTile - q_j0];
    tile_ct[b] = (int32_t)coarse_tile(cdf, n_ct, total, s);
}

// idx[j - j_lo] = min(#{k < n_cdf : T_k <= S_j}, n_cdf - 1) for the slots j_lo <= j < j_hi; q[0] is element q_j0.
// A block owns one slot tile (2048 slots, 8 consecutive per thread, partial sums in registers) and walks the CDF
// forward from tile_ct in windows of two CDF tiles staged in shared memory as integer thresholds.
constexpr int kSpWin = 2 * kSpacingTile;
// Thread t looks at thresholds around position 8 t (its slots are consecutive, the thresholds nearly evenly spread),
// so a plain array puts a half-warp on two bank pairs.  One padding word per 8 entries makes lane t land on
// word 9 t + d: 16 distinct bank pairs.
__device__ __forceinline__ int sp_pad(int e) { return e + (e >> 3); }
__global__ void __launch_bounds__(kSpBlock) search_sorted_kernel(const double* __restrict__ cdf, int64_t n_cdf,
                                                                 const uint32_t* __restrict__ q, int64_t q_j0,
                                                                 const unsigned long long* __restrict__ tile_prefix,
                                                                 const unsigned long long* __restrict__ total_p,
                                                                 int64_t j_lo, int64_t j_hi,
                                                                 const int32_t* __restrict__ tile_ct,
                                                                 int64_t* __restrict__ idx) {
    __shared__ uint64_t s_T[kSpWin + kSpWin / 8];
    __shared__ uint64_t s_warp[kSpBlock / 32];
    __shared__ uint64_t s_sc;
    __shared__ int s_cnt;
    const double total = (double)*total_p;
    const int64_t n_ct = (n_cdf + kSpacingTile - 1) / kSpacingTile;
    const int64_t gt = j_lo / kSpacingTile + blockIdx.x;  // global slot tile
    const int64_t tile0 = gt * kSpacingTile;
    int64_t ct = tile_ct[blockIdx.x];
    // ---- partial sums of this slot tile
    uint64_t S[kSpItems];
    {
        const uint4* src = reinterpret_cast<const uint4*>(q + (tile0 - q_j0) + threadIdx.x * kSpItems);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        const uint32_t v[kSpItems] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint64_t run = 0;
#pragma unroll
        for (int i = 0; i < kSpItems; ++i) {
            run += v[i];
            S[i] = run;
        }
        uint64_t tot;
        const uint64_t base = block_scan_excl_u64(run, s_warp, tot) + tile_prefix[gt];
#pragma unroll
        for (int i = 0; i < kSpItems; ++i) S[i] += base;
    }
    const int first = (int)(j_lo > tile0 ? j_lo - tile0 : 0);
    const int end = (int)(j_hi - tile0 < kSpacingTile ? j_hi - tile0 : kSpacingTile);
    int64_t anc[kSpItems];
#pragma unroll
    for (int i = 0; i < kSpItems; ++i) anc[i] = n_cdf - 1;
    int cur = first;  // slots [first, cur) are resolved
    while (cur < end) {
        // window: CDF entries [k0, k0 + cnt), cnt <= 4096, as integer thresholds; t_next = the entry after it
        const int64_t k0 = ct * kSpacingTile;
        const int cnt = (int)(n_cdf - k0 < kSpWin ? n_cdf - k0 : kSpWin);
        const bool last_win = k0 + cnt >= n_cdf;
#pragma unroll
        for (int i = 0; i < kSpWin / kSpBlock; ++i) {
            const int e = i * kSpBlock + threadIdx.x;
            s_T[sp_pad(e)] = e < cnt ? spacing_threshold(ld_stream(cdf + k0 + e), total) : ~0ull;
        }
        const uint64_t t_next = last_win ? ~0ull : spacing_threshold(__ldg(cdf + k0 + cnt), total);
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        // slots with S < t_next have their ancestor in [k0, k0 + cnt] (k0 + cnt: the first entry after the window)
        int hint = 0, done = 0;
#pragma unroll
        for (int i = 0; i < kSpItems; ++i) {
            const int ls = threadIdx.x * kSpItems + i;
            if (ls >= cur && ls < end && (S[i] < t_next || last_win)) {
                const uint64_t s = S[i];
                int lo = hint, hi = cnt, step = 1;  // count of thresholds <= s, in [lo, hi]
                while (lo + step <= hi && s_T[sp_pad(lo + step - 1)] <= s) {
                    lo += step;
                    step <<= 1;
                }
                if (lo + step - 1 < hi) hi = lo + step - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_T[sp_pad(mid - 1)] <= s) lo = mid;
                    else hi = mid - 1;
                }
                hint = lo;
                const int64_t a = k0 + lo;
                anc[i] = a < n_cdf - 1 ? a : n_cdf - 1;
                ++done;
            }
        }
        // the slots are sorted, so the resolved ones are the next `sum(done)` slots
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
        if ((threadIdx.x & 31) == 0 && done) atomicAdd(&s_cnt, done);
        __syncthreads();
        cur += s_cnt;
        if (cur >= end) break;
        // first unresolved slot: S >= t_next, so its ancestor lies after the window - the following window if that
        // still covers it, otherwise wherever the coarse search says (a long stretch of negligible weights)
        {
            const int owner = cur / kSpItems, which = cur % kSpItems;
            if (threadIdx.x == owner) {
                uint64_t v = S[0];
#pragma unroll
                for (int i = 1; i < kSpItems; ++i) v = which == i ? S[i] : v;
                s_sc = v;
            }
        }
        __syncthreads();
        const uint64_t sc = s_sc;
        const int64_t ct_follow = ct + kSpWin / kSpacingTile;
        const int64_t ct_after = ct_follow + kSpWin / kSpacingTile;
        if (ct_after >= n_ct || sc < spacing_threshold(__ldg(cdf + ct_after * kSpacingTile), total)) ct = ct_follow;
        else ct = coarse_tile(cdf, n_ct, total, sc);  // every thread runs the same search: uniform, no extra barrier
        __syncthreads();  // s_T, s_cnt and s_sc are rewritten by the next round
    }
    // the ancestors leave through shared memory: consecutive lanes store consecutive slots
    __syncthreads();
    int64_t* s_out = reinterpret_cast<int64_t*>(s_T);
#pragma unroll
    for (int i = 0; i < kSpItems; ++i) s_out[sp_pad(threadIdx.x * kSpItems + i)] = anc[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSpItems; ++i) {
        const int ls = i * kSpBlock + threadIdx.x;
        if (ls >= first && ls < end) idx[tile0 + ls - j_lo] = s_out[sp_pad(ls)];
    }
}

// out[i] = #{j < m : S_j < t_i}: the first slot whose partial sum reaches t_i (shard boundaries).  One block per
// threshold: binary search of the global tile table, then the tile's spacings are regenerated in place.
__global__ void __launch_bounds__(kSpBlock) lower_bound_kernel(PhiloxKeys K, uint32_t stream,
                                                               const unsigned long long* __restrict__ tile_prefix,
                                                               int64_t n_tiles, int64_t m,
                                                               const unsigned long long* __restrict__ t_p,
                                                               long long* __restrict__ out) {
    __shared__ uint64_t s_warp[kSpBlock / 32];
    __shared__ int s_count;
    const uint64_t t = t_p[blockIdx.x];
    if (threadIdx.x == 0) s_count = 0;
    int64_t lo = 0, hi = n_tiles;  // the last tile whose exclusive prefix is < t (prefix[0] = 0; t == 0 below)
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (tile_prefix[mid] < t) lo = mid;
        else hi = mid;
    }
    const int64_t e0 = lo * kSpacingTile + (int64_t)threadIdx.x * kSpItems;
    uint64_t run = 0, S[kSpItems];
#pragma unroll
    for (int b = 0; b < kSpItems / 2; ++b) {
        const uint64_t blk = (uint64_t)(e0 >> 1) + b;
        const uint4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u, K);
        const uint64_t ka = ((((uint64_t)r.x) << 32) | r.y) >> 11, kb = ((((uint64_t)r.z) << 32) | r.w) >> 11;
        run += e0 + 2 * b <= m ? spacing_q(ka) : 0u;
        S[2 * b] = run;
        run += e0 + 2 * b + 1 <= m ? spacing_q(kb) : 0u;
        S[2 * b + 1] = run;
    }
    uint64_t tot;
    const uint64_t base = block_scan_excl_u64(run, s_warp, tot) + tile_prefix[lo];
    int c = 0;
#pragma unroll
    for (int i = 0; i < kSpItems; ++i) c += (S[i] + base < t) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t res = t == 0 ? 0 : lo * kSpacingTile + s_count;
        out[blockIdx.x] = res < m ? res : m;
    }
}

}  // namespace asp

using namespace asp;

extern "C" {

int64_t asp_sorted_tiles(int64_t m_total) { return (m_total + 1 + kSpacingTile - 1) / kSpacingTile; }

int asp_sorted_spacings(uint64_t seed, uint32_t stream_id, int64_t j0, int64_t n_tiles, int64_t m_total, uint32_t* d_q,
                        unsigned long long* d_tile_sums, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_q && d_tile_sums, "null pointer");
    ASP_REQUIRE(m_total > 0 && n_tiles > 0, "sizes must be positive");
    ASP_REQUIRE(j0 >= 0 && j0 % kSpacingTile == 0, "j0 must be a multiple of 2048");
    const int64_t cap = (int64_t)kNumSMs * 8;
    spacings_kernel<<<(int)(n_tiles < cap ? n_tiles : cap), kSpBlock, 0, (cudaStream_t)stream>>>(
        philox_keys(seed), stream_id, j0, n_tiles, m_total, d_q, d_tile_sums);
    return check_launch("asp_sorted_spacings");
}

int asp_sorted_tile_scan(const unsigned long long* d_tile_sums, int64_t n_tiles, unsigned long long* d_tile_prefix,
                         unsigned long long* d_total, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_tile_sums && d_tile_prefix && d_total, "null pointer");
    ASP_REQUIRE(n_tiles > 0, "n_tiles must be positive");
    tile_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_tile_sums, n_tiles, d_tile_prefix, d_total);
    return check_launch("asp_sorted_tile_scan");
}

size_t asp_search_sorted_workspace_bytes(int64_t n_slots) {
    return sizeof(int32_t) * (size_t)(n_slots / kSpacingTile + 2) + 16;
}

int asp_search_sorted(const double* d_cdf, int64_t n_cdf, const uint32_t* d_q, int64_t q_j0,
                      const unsigned long long* d_tile_prefix, const unsigned long long* d_total, int64_t j_lo,
                      int64_t j_hi, int64_t* d_idx, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_cdf && d_q && d_tile_prefix && d_total && d_idx && ws, "null pointer");
    ASP_REQUIRE(n_cdf > 0, "n_cdf must be positive");
    ASP_REQUIRE(q_j0 >= 0 && q_j0 % kSpacingTile == 0, "q_j0 must be a multiple of 2048");
    ASP_REQUIRE(j_lo >= q_j0 && j_hi >= j_lo, "bad slot range");
    if (j_hi == j_lo) return ASP_OK;
    const int64_t tiles = (j_hi - 1) / kSpacingTile - j_lo / kSpacingTile + 1;
    ASP_REQUIRE(tiles < (1ll << 31), "too many slot tiles");
    ASP_REQUIRE((n_cdf + kSpacingTile - 1) / kSpacingTile < (1ll << 31), "CDF too long");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* tile_ct = (int32_t*)ws;
    slot_tile_map_kernel<<<(int)((tiles + kSpBlock - 1) / kSpBlock), kSpBlock, 0, s>>>(d_cdf, n_cdf, d_q, q_j0,
                                                                                       d_tile_prefix, d_total, j_lo,
                                                                                       tiles, tile_ct);
    search_sorted_kernel<<<(int)tiles, kSpBlock, 0, s>>>(d_cdf, n_cdf, d_q, q_j0, d_tile_prefix, d_total, j_lo, j_hi,
                                                        tile_ct, d_idx);
    return check_launch("asp_search_sorted");
}

int asp_sorted_lower_bound(uint64_t seed, uint32_t stream_id, const unsigned long long* d_tile_prefix, int64_t n_tiles,
                           int64_t m_total, const unsigned long long* d_thresholds, int32_t n_thresholds,
                           long long* d_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_tile_prefix && d_thresholds && d_out, "null pointer");
    ASP_REQUIRE(n_tiles > 0 && n_thresholds > 0 && m_total > 0, "sizes must be positive");
    lower_bound_kernel<<<n_thresholds, kSpBlock, 0, (cudaStream_t)stream>>>(philox_keys(seed), stream_id, d_tile_prefix,
                                                                           n_tiles, m_total, d_thresholds, d_out);
    return check_launch("asp_sorted_lower_bound");
}

}  // extern "C"
