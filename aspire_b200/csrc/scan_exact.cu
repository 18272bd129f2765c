// K5: sequential-order-exact fp64 prefix sum (the CDF of numpy's Generator.choice,
// called at /root/reference/src/aspire/samples.py:1278), tile-parallel.
//
// Result is bit-identical to numpy.cumsum(p) (strictly sequential fp64 adds) and
// then to `cdf /= cdf[-1]` (IEEE division).  Algorithm: oracle/exact_scan.py and
// SURVEY App. E.  While the running sum stays in one binade it is m*u with integer
// m, and adding w is m += round_half_even(w/u): integer, hence associative; ties
// depend on the parity of m, so an element is the pair (increment if m even,
// increment if m odd) and pairs compose associatively.  Binade crossings are
// resolved by one genuine fp64 add each, inside "dirty" tiles that one block
// walks in order; everything else is embarrassingly parallel.
//
// HBM traffic per weight: tile stats 8 B (free when asp_smc_weights produced
// them), aggregate 8 B, write 8 B + 8 B.  All integer/fp64 ALU, no tensor work.
#include <math.h>

#include "common.cuh"
#include "scan_math.cuh"

namespace asp {

constexpr int kTile = ASP_CDF_TILE;       // 2048 weights per tile
constexpr int kItems = kTile / kBlock;    // 8 per thread (blocked)
constexpr int kResolveThreads = 1024;
constexpr int kDirty = -100000;           // tile_e flag

enum { ERR_PLAN = 1, ERR_CROSSED = 2, ERR_BAD_WEIGHT = 4 };

// Multi-GPU carry chain over NVLink peer memory (asp_cdf_resolve_p2p / asp_cdf_write_p2p): the chain region of
// a rank's mailbox (include/aspire_b200.h, asp_peer_t) holds, per epoch parity, {running sum handed over by rank
// r-1, its epoch flag, exact total written by the last rank, its epoch flag}; the epoch counter follows.
constexpr int kChainBase = 2 * ASP_MAX_PEERS * 16 + 8;
constexpr int kChainCounter = kChainBase + 8;

// waits (bounded: 20 s of %globaltimer) until *flag == epoch; false = the writer never arrived
__device__ __forceinline__ bool chain_wait(volatile double* flag, double epoch) {
    unsigned long long t0 = 0, t1;
    while (*flag != epoch) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (!t0) t0 = t1;
        if (t1 - t0 > 20000000000ull) return false;
    }
    __threadfence_system();
    return true;
}

struct CdfLayout {
    int64_t n_tiles;
    size_t off_stats, off_e, off_agg, off_start, off_dirty, total;
};

struct CdfHeader {
    ReduceWs red;
    long long info[4];  // n_tiles, n_dirty, n_phases, error flags
    double total[2];    // exact outgoing sum, running max
    int n_dirty;
    int pad[3];
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static CdfLayout cdf_layout(int64_t n) {
    CdfLayout L;
    L.n_tiles = (n + kTile - 1) / kTile;
    size_t o = align256(sizeof(CdfHeader));
    L.off_stats = o; o = align256(o + sizeof(double) * 2 * L.n_tiles);
    L.off_e = o;     o = align256(o + sizeof(int) * L.n_tiles);
    L.off_agg = o;   o = align256(o + sizeof(Pair) * L.n_tiles);
    L.off_start = o; o = align256(o + sizeof(double) * L.n_tiles);
    L.off_dirty = o; o = align256(o + sizeof(int) * (L.n_tiles + 1));
    L.total = o;
    return L;
}

// the exact-arithmetic helpers (binade_of, scale_pow2, element_pair, compose, apply) live in scan_math.cuh,
// host + device, so that the CPU tests run the shipped arithmetic against numpy.cumsum
__device__ __forceinline__ Pair shfl_up_pair(Pair a, int o) {
    Pair r;
    r.d0 = __shfl_up_sync(0xffffffffu, a.d0, o);
    r.d1 = __shfl_up_sync(0xffffffffu, a.d1, o);
    return r;
}
__device__ __forceinline__ Pair shfl_pair(Pair a, int src) {
    Pair r;
    r.d0 = __shfl_sync(0xffffffffu, a.d0, src);
    r.d1 = __shfl_sync(0xffffffffu, a.d1, src);
    return r;
}

// Ordered block scan of pairs: returns the EXCLUSIVE prefix of this thread and,
// through *total, the composition of the whole block.  smem: Pair[32].
__device__ __forceinline__ Pair block_exclusive_scan(Pair mine, Pair* smem, Pair* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    Pair incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Pair t = shfl_up_pair(incl, o);
        if (lane >= o) incl = compose(t, incl);
    }
    __syncthreads();
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        Pair w = lane < nw ? smem[lane] : Pair{0, 0};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Pair t = shfl_up_pair(w, o);
            if (lane >= o) w = compose(t, w);
        }
        smem[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    Pair excl = shfl_up_pair(incl, 1);
    if (lane == 0) excl = Pair{0, 0};
    if (wid > 0) excl = compose(smem[wid - 1], excl);
    *total = smem[nw - 1];
    return excl;
}

// ---------------------------------------------------------------------------
// stage 1: per-tile sum and max
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) cdf_tile_stats_kernel(const double* __restrict__ p, int64_t n,
                                                                double* __restrict__ stats, CdfHeader* hdr) {
    __shared__ double smem[3 * 32];
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTile;
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            int64_t i = base + k * kBlock + threadIdx.x;
            double w = i < n ? ld_stream(p + i) : 0.0;
            t[0] += w;
            t[1] = fmax(t[1], w);
            t[2] += (w >= 0.0 && w < identity(OP_MIN)) ? 0.0 : 1.0;  // negative, NaN or inf
        }
        const int ops[3] = {OP_SUM, OP_MAX, OP_SUM};
        block_reduce<3>(t, ops, smem);
        if (threadIdx.x == 0) {
            stats[2 * tile] = t[0];
            stats[2 * tile + 1] = t[1];
            if (t[2] > 0.0) atomicOr((unsigned long long*)&hdr->info[3], (unsigned long long)ERR_BAD_WEIGHT);
        }
    }
}

// ---------------------------------------------------------------------------
// stage 2: approximate prefixes -> per-tile binade or dirty flag; dirty list
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kResolveThreads) cdf_plan_kernel(int64_t n_tiles, const double* __restrict__ carry,
                                                                   const double* __restrict__ stats,
                                                                   int* __restrict__ tile_e, int* __restrict__ dirty,
                                                                   CdfHeader* hdr) {
    __shared__ double s_sum[kResolveThreads];
    __shared__ double s_max[kResolveThreads];
    __shared__ int s_cnt[kResolveThreads];
    const int t = threadIdx.x;
    const double offset = carry ? carry[0] : 0.0;
    const double max_in = carry ? carry[1] : 0.0;
    const double n_total = (carry && carry[3] > 0.0) ? carry[3] : (double)n_tiles * kTile;
    const double tol = (n_total + kTile) * 2.220446049250313e-16;  // 2^-52
    const int64_t chunk = (n_tiles + kResolveThreads - 1) / kResolveThreads;
    const int64_t lo_i = (int64_t)t * chunk;
    const int64_t hi_i = lo_i + chunk < n_tiles ? lo_i + chunk : n_tiles;
    double ls = 0.0, lm = 0.0;
    for (int64_t i = lo_i; i < hi_i; ++i) {
        ls += stats[2 * i];
        lm = fmax(lm, stats[2 * i + 1]);
    }
    s_sum[t] = ls;
    s_max[t] = lm;
    __syncthreads();
    // Hillis-Steele inclusive scans over the 1024 thread aggregates
    for (int o = 1; o < kResolveThreads; o <<= 1) {
        double a = 0.0, b = 0.0;
        if (t >= o) {
            a = s_sum[t - o];
            b = s_max[t - o];
        }
        __syncthreads();
        if (t >= o) {
            s_sum[t] = a + s_sum[t];
            s_max[t] = fmax(b, s_max[t]);
        }
        __syncthreads();
    }
    double P = offset + (t > 0 ? s_sum[t - 1] : 0.0);
    double runmax = fmax(max_in, t > 0 ? s_max[t - 1] : 0.0);
    int cnt = 0;
    for (int64_t i = lo_i; i < hi_i; ++i) {
        double T = stats[2 * i];
        double lo = fmax(P * (1.0 - tol), runmax);
        double hi = (P + T) * (1.0 + tol);
        int e_lo = binade_of(lo), e_hi = binade_of(hi);
        bool clean = (e_lo == e_hi) && (hi < identity(OP_MIN));
        tile_e[i] = clean ? e_lo : kDirty;
        cnt += clean ? 0 : 1;
        P += T;
        runmax = fmax(runmax, stats[2 * i + 1]);
    }
    s_cnt[t] = cnt;
    __syncthreads();
    for (int o = 1; o < kResolveThreads; o <<= 1) {
        int a = 0;
        if (t >= o) a = s_cnt[t - o];
        __syncthreads();
        if (t >= o) s_cnt[t] += a;
        __syncthreads();
    }
    int pos = t > 0 ? s_cnt[t - 1] : 0;
    for (int64_t i = lo_i; i < hi_i; ++i)
        if (tile_e[i] == kDirty) dirty[pos++] = (int)i;
    if (t == kResolveThreads - 1) {
        hdr->n_dirty = s_cnt[t];
        hdr->info[0] = n_tiles;
        hdr->info[1] = s_cnt[t];
        hdr->info[2] = 0;
        hdr->total[1] = fmax(max_in, s_max[t]);
    }
}

// ---------------------------------------------------------------------------
// stage 3: integer increment pair of every clean tile
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_blocked(const double* __restrict__ p, int64_t n, int64_t base, double (&w)[kItems]) {
    const int64_t first = base + (int64_t)threadIdx.x * kItems;
    if (first + kItems <= n && ((reinterpret_cast<uintptr_t>(p + first) & 15) == 0)) {
#pragma unroll
        for (int k = 0; k < kItems; k += 2) {
            double2 v = *reinterpret_cast<const double2*>(p + first + k);
            w[k] = v.x;
            w[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k) w[k] = first + k < n ? p[first + k] : 0.0;
    }
}

__global__ void __launch_bounds__(kBlock) cdf_aggregate_kernel(const double* __restrict__ p, int64_t n,
                                                               const int* __restrict__ tile_e,
                                                               Pair* __restrict__ agg) {
    __shared__ Pair smem[32];
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int e = tile_e[tile];
        if (e == kDirty) continue;  // uniform per block
        double w[kItems];
        load_blocked(p, n, tile * kTile, w);
        Pair h = element_pair(w[0], e);
#pragma unroll
        for (int k = 1; k < kItems; ++k) h = compose(h, element_pair(w[k], e));
        Pair total;
        block_exclusive_scan(h, smem, &total);
        if (threadIdx.x == 0) agg[tile] = total;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// stage 4: one block walks the tile sequence: clean runs by an ordered block
// scan of their pairs, dirty tiles by the phase loop (integer scan to the first
// crossing, one genuine add, repeat)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kResolveThreads) cdf_resolve_kernel(
    const double* __restrict__ p, int64_t n, const double* __restrict__ carry, const int* __restrict__ tile_e,
    const Pair* __restrict__ agg, const int* __restrict__ dirty, double* __restrict__ tile_start,
    double* __restrict__ cdf, double* __restrict__ carry_out, CdfHeader* hdr, const asp_peer_t P,
    double* __restrict__ chain_out) {
    __shared__ Pair smem[32];
    __shared__ double s_s;
    __shared__ double s_epoch;
    __shared__ int s_j;
    __shared__ long long s_phases;
    __shared__ unsigned int s_err;
    const int t = threadIdx.x;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int n_dirty = hdr->n_dirty;
    if (t == 0) {
        s_s = carry ? carry[2] : 0.0;
        s_phases = 0;
        s_err = 0;
        if (P.world > 1) {  // the running sum at the end of rank r-1's slice arrives in this rank's mailbox
            double* counter = P.box[P.rank] + kChainCounter;
            s_epoch = *counter + 1.0;
            *counter = s_epoch;
            s_s = 0.0;
            if (P.rank > 0) {
                volatile double* m = P.box[P.rank] + kChainBase + ((long long)s_epoch & 1) * 4;
                s_s = chain_wait(m + 1, s_epoch) ? m[0] : nan("");  // NaN poisons the CDF: the host raises
            }
            if (chain_out) chain_out[0] = s_s;
        }
    }
    __syncthreads();
    int64_t prev = 0;
    for (int k = 0; k <= n_dirty; ++k) {
        const int64_t run_end = k < n_dirty ? (int64_t)dirty[k] : n_tiles;
        // ---- clean run [prev, run_end) ----
        if (run_end > prev) {
            const double s = s_s;
            const int e = tile_e[prev];
            const long long m0 = __double2ll_rn(scale_pow2(s, 52 - e));
            if (t == 0 && binade_of(s) != e) s_err |= ERR_PLAN;
            const int64_t len = run_end - prev;
            const int64_t chunk = (len + kResolveThreads - 1) / kResolveThreads;
            const int64_t a = prev + (int64_t)t * chunk;
            const int64_t b = a + chunk < run_end ? a + chunk : run_end;
            Pair h{0, 0};
            for (int64_t i = a; i < b; ++i) h = compose(h, agg[i]);
            Pair total;
            Pair g = block_exclusive_scan(h, smem, &total);
            for (int64_t i = a; i < b; ++i) {
                tile_start[i] = scale_pow2((double)apply(g, m0), e - 52);
                g = compose(g, agg[i]);
            }
            __syncthreads();
            if (t == 0) {
                long long m_end = apply(total, m0);
                if (m_end >= kTwo53) s_err |= ERR_CROSSED;
                s_s = scale_pow2((double)m_end, e - 52);
            }
            __syncthreads();
        }
        // ---- dirty tile run_end ----
        if (k < n_dirty) {
            const int64_t base = run_end * kTile;
            const int64_t i0 = base + 2 * t, i1 = i0 + 1;
            const double w0 = i0 < n ? p[i0] : 0.0;
            const double w1 = i1 < n ? p[i1] : 0.0;
            if (t == 0) tile_start[run_end] = s_s;
            int pos = 0;
            while (pos < kTile) {
                const double s = s_s;
                const int e = binade_of(s);
                const long long m0 = __double2ll_rn(scale_pow2(s, 52 - e));
                Pair f0 = (2 * t >= pos) ? element_pair(w0, e) : Pair{0, 0};
                Pair f1 = (2 * t + 1 >= pos) ? element_pair(w1, e) : Pair{0, 0};
                Pair total;
                Pair g = block_exclusive_scan(compose(f0, f1), smem, &total);
                const long long m_prev = apply(g, m0);
                const long long m_a = apply(compose(g, f0), m0);
                const long long m_b = apply(compose(compose(g, f0), f1), m0);
                if (t == 0) s_j = kTile;
                __syncthreads();
                int cand = kTile;
                if (2 * t >= pos && m_a >= kTwo53) cand = 2 * t;
                else if (2 * t + 1 >= pos && m_b >= kTwo53) cand = 2 * t + 1;
                if (cand < kTile) atomicMin(&s_j, cand);
                __syncthreads();
                const int j = s_j;
                if (2 * t >= pos && 2 * t < j && i0 < n) cdf[i0] = scale_pow2((double)m_a, e - 52);
                if (2 * t + 1 >= pos && 2 * t + 1 < j && i1 < n) cdf[i1] = scale_pow2((double)m_b, e - 52);
                if (j == kTile) {
                    if (t == kResolveThreads - 1) s_s = scale_pow2((double)m_b, e - 52);
                } else if (j == 2 * t) {
                    double sn = __dadd_rn(scale_pow2((double)m_prev, e - 52), w0);
                    if (i0 < n) cdf[i0] = sn;
                    s_s = sn;
                } else if (j == 2 * t + 1) {
                    double sn = __dadd_rn(scale_pow2((double)m_a, e - 52), w1);
                    if (i1 < n) cdf[i1] = sn;
                    s_s = sn;
                }
                if (t == 0) s_phases += 1;
                pos = j + 1;
                __syncthreads();
            }
        }
        prev = run_end + 1;
    }
    if (t == 0) {
        hdr->total[0] = s_s;
        hdr->info[2] = s_phases;
        if (s_err) atomicOr((unsigned long long*)&hdr->info[3], (unsigned long long)s_err);
        if (carry_out) {
            carry_out[0] = s_s;
            carry_out[1] = hdr->total[1];
        }
        if (P.world > 1) {
            const int par = (int)((long long)s_epoch & 1) * 4;
            if (P.rank < P.world - 1) {  // hand the running sum to the next rank
                volatile double* m = P.box[P.rank + 1] + kChainBase + par;
                m[0] = s_s;
                __threadfence_system();
                m[1] = s_epoch;
            } else {  // the last rank holds the exact total: every rank normalises by it
                for (int q = 0; q < P.world; ++q) (P.box[q] + kChainBase + par)[2] = s_s;
                __threadfence_system();
                for (int q = 0; q < P.world; ++q) ((volatile double*)(P.box[q] + kChainBase + par))[3] = s_epoch;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// stage 5: write the (normalised) CDF
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) cdf_write_kernel(const double* __restrict__ p, int64_t n,
                                                           const double* __restrict__ total_ptr, int normalise,
                                                           const int* __restrict__ tile_e,
                                                           const double* __restrict__ tile_start,
                                                           double* __restrict__ cdf, const asp_peer_t P,
                                                           double* __restrict__ chain_out) {
    __shared__ Pair smem[32];
    __shared__ double s_total;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    if (P.world > 1) {  // the total comes from the last rank's resolve stage, through this rank's mailbox
        if (threadIdx.x == 0) {
            const double epoch = P.box[P.rank][kChainCounter];  // advanced by this rank's resolve stage
            volatile double* m = P.box[P.rank] + kChainBase + ((long long)epoch & 1) * 4;
            s_total = chain_wait(m + 3, epoch) ? m[2] : nan("");
            if (blockIdx.x == 0 && chain_out) chain_out[1] = s_total;
        }
        __syncthreads();
    }
    const double total = P.world > 1 ? (normalise ? s_total : 1.0) : (normalise ? total_ptr[0] : 1.0);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int e = tile_e[tile];
        const int64_t base = tile * kTile;
        if (e == kDirty) {  // values already written by the resolve stage
            if (normalise) {
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    int64_t i = base + k * kBlock + threadIdx.x;
                    if (i < n) cdf[i] = __ddiv_rn(cdf[i], total);
                }
            }
            continue;
        }
        double w[kItems];
        load_blocked(p, n, base, w);
        Pair f[kItems];
        f[0] = element_pair(w[0], e);
        Pair h = f[0];
#pragma unroll
        for (int k = 1; k < kItems; ++k) {
            f[k] = element_pair(w[k], e);
            h = compose(h, f[k]);
        }
        Pair tot;
        Pair g = block_exclusive_scan(h, smem, &tot);
        const long long m0 = __double2ll_rn(scale_pow2(tile_start[tile], 52 - e));
        long long m = apply(g, m0);
        double out[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            m += (m & 1) ? f[k].d1 : f[k].d0;
            double v = scale_pow2((double)m, e - 52);
            out[k] = normalise ? __ddiv_rn(v, total) : v;
        }
        const int64_t first = base + (int64_t)threadIdx.x * kItems;
        if (first + kItems <= n && ((reinterpret_cast<uintptr_t>(cdf + first) & 15) == 0)) {
#pragma unroll
            for (int k = 0; k < kItems; k += 2) *reinterpret_cast<double2*>(cdf + first + k) = make_double2(out[k], out[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < kItems; ++k)
                if (first + k < n) cdf[first + k] = out[k];
        }
        __syncthreads();
    }
}

__global__ void cdf_reset_kernel(CdfHeader* hdr) {
    hdr->info[0] = hdr->info[1] = hdr->info[2] = hdr->info[3] = 0;
}

__global__ void cdf_copy_stats_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

static inline int tile_grid(int64_t n_tiles) {
    int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(n_tiles < cap ? n_tiles : cap);
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_cdf_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return cdf_layout(n).total;
}

int asp_cdf_tile_stats(const double* d_p, int64_t n, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cudaStream_t s = (cudaStream_t)stream;
    cdf_reset_kernel<<<1, 1, 0, s>>>((CdfHeader*)b);
    cdf_tile_stats_kernel<<<tile_grid(L.n_tiles), kBlock, 0, s>>>(d_p, n, (double*)(b + L.off_stats), (CdfHeader*)b);
    return check_launch("asp_cdf_tile_stats");
}

int asp_cdf_set_tile_stats(const double* d_tile_stats, int64_t n, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_tile_stats && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cudaStream_t s = (cudaStream_t)stream;
    cdf_reset_kernel<<<1, 1, 0, s>>>((CdfHeader*)b);
    cdf_copy_stats_kernel<<<grid_for(2 * L.n_tiles, 1), kBlock, 0, s>>>(d_tile_stats, (double*)(b + L.off_stats),
                                                                         2 * L.n_tiles);
    return check_launch("asp_cdf_set_tile_stats");
}

int asp_cdf_plan(int64_t n, const double* d_carry, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cdf_plan_kernel<<<1, kResolveThreads, 0, (cudaStream_t)stream>>>(L.n_tiles, d_carry, (double*)(b + L.off_stats),
                                                                      (int*)(b + L.off_e), (int*)(b + L.off_dirty),
                                                                      (CdfHeader*)b);
    return check_launch("asp_cdf_plan");
}

int asp_cdf_aggregate(const double* d_p, int64_t n, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cdf_aggregate_kernel<<<tile_grid(L.n_tiles), kBlock, 0, (cudaStream_t)stream>>>(d_p, n, (int*)(b + L.off_e),
                                                                                    (Pair*)(b + L.off_agg));
    return check_launch("asp_cdf_aggregate");
}

int asp_cdf_resolve(const double* d_p, int64_t n, const double* d_carry, double* d_cdf, double* d_carry_out, void* ws,
                    void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && d_cdf && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cdf_resolve_kernel<<<1, kResolveThreads, 0, (cudaStream_t)stream>>>(
        d_p, n, d_carry, (int*)(b + L.off_e), (Pair*)(b + L.off_agg), (int*)(b + L.off_dirty),
        (double*)(b + L.off_start), d_cdf, d_carry_out, (CdfHeader*)b, asp_peer_t{}, nullptr);
    return check_launch("asp_cdf_resolve");
}

static int check_chain_peer(const asp_peer_t* peer) {
    ASP_REQUIRE(peer && peer->world >= 2 && peer->world <= ASP_MAX_PEERS && peer->rank >= 0 && peer->rank < peer->world,
                "bad peer descriptor");
    for (int r = 0; r < peer->world; ++r) ASP_REQUIRE(peer->box[r], "null mailbox pointer");
    return ASP_OK;
}

int asp_cdf_resolve_p2p(const double* d_p, int64_t n, double* d_cdf, double* d_chain_out, const asp_peer_t* peer,
                        void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && d_cdf && ws && d_chain_out, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    int rc = check_chain_peer(peer);
    if (rc) return rc;
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cdf_resolve_kernel<<<1, kResolveThreads, 0, (cudaStream_t)stream>>>(
        d_p, n, nullptr, (int*)(b + L.off_e), (Pair*)(b + L.off_agg), (int*)(b + L.off_dirty),
        (double*)(b + L.off_start), d_cdf, nullptr, (CdfHeader*)b, *peer, d_chain_out);
    return check_launch("asp_cdf_resolve_p2p");
}

int asp_cdf_write_p2p(const double* d_p, int64_t n, int normalise, double* d_cdf, double* d_chain_out,
                      const asp_peer_t* peer, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && d_cdf && ws && d_chain_out, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    int rc = check_chain_peer(peer);
    if (rc) return rc;
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    cdf_write_kernel<<<tile_grid(L.n_tiles), kBlock, 0, (cudaStream_t)stream>>>(
        d_p, n, nullptr, normalise, (int*)(b + L.off_e), (double*)(b + L.off_start), d_cdf, *peer, d_chain_out);
    return check_launch("asp_cdf_write_p2p");
}

int asp_cdf_write(const double* d_p, int64_t n, const double* d_total, int normalise, double* d_cdf, void* ws,
                  void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_p && d_cdf && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    CdfLayout L = cdf_layout(n);
    char* b = (char*)ws;
    const double* total = d_total ? d_total : ((CdfHeader*)b)->total;
    cdf_write_kernel<<<tile_grid(L.n_tiles), kBlock, 0, (cudaStream_t)stream>>>(
        d_p, n, total, normalise, (int*)(b + L.off_e), (double*)(b + L.off_start), d_cdf, asp_peer_t{}, nullptr);
    return check_launch("asp_cdf_write");
}

int asp_cdf_exact(const double* d_p, int64_t n, int normalise, double* d_cdf, void* ws, void* stream) {
    int rc = asp_cdf_tile_stats(d_p, n, ws, stream);
    if (rc) return rc;
    if ((rc = asp_cdf_plan(n, nullptr, ws, stream))) return rc;
    if ((rc = asp_cdf_aggregate(d_p, n, ws, stream))) return rc;
    if ((rc = asp_cdf_resolve(d_p, n, nullptr, d_cdf, nullptr, ws, stream))) return rc;
    return asp_cdf_write(d_p, n, nullptr, normalise, d_cdf, ws, stream);
}

int asp_cdf_info(void* ws, int64_t* h_info, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(ws && h_info, "null pointer");
    cudaError_t e = cudaMemcpyAsync(h_info, ((CdfHeader*)ws)->info, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("asp_cdf_info: %s", cudaGetErrorString(e));
        return ASP_ERR_CUDA;
    }
    return ASP_OK;
}

}  // extern "C"
