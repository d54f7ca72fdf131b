// kb_watch.cu -- watch fan-out: one revision-ordered event slab x every registered watcher's
// (key-prefix, min-revision) predicate -> per-watcher ordered delivery lists.
//
// Replaces (reference file:line):
//   WatcherHub.Stream broadcast           pkg/backend/watcherhub.go:78-92   (every batch handed to every watcher)
//   processEvents / filterByRevision /    pkg/backend/watch.go:119-159      (per watcher: strip leading events below
//   filterByPrefix                                                           min_rev, keep bytes.HasPrefix matches)
//
// Brute force is W x E prefix tests (1e9 for 10k watchers x 100k events).  Here watchers are grouped by
// distinct prefix; an event probes a device hash table once per DISTINCT PREFIX LENGTH with the FNV-1a
// hash of its own leading bytes (verified byte-exactly), so the work is O(E * #lengths + deliveries):
//   k_batch_pm       per collector batch: running max of Event.Revision (the "leading strip" predicate becomes
//                    pm[i] >= min_rev) + a flag "revisions globally non-decreasing"
//   k_match_count    per event and prefix length: matched group (kept in ematch) + warp-aggregated group counts
//   (scan)           group segment offsets
//   k_classify       groups by match count: small (<=32), medium (<= BIG_T), large (global bitmap)
//   k_scatter        small/medium: warp-aggregated slot claim into the group's segment (unordered across warps);
//                    large: the warp's ballot IS the 32-event bitmap word of the group (plain store, no atomics)
//   k_sort_small / k_sort_medium / k_expand_large   every segment ascending (warp rank-sort / shared-memory bitmap /
//                    ordered expansion of the global bitmap)
//   k_expand_count   per watcher: deliveries = its group's segment filtered by min_rev
//   (scan)           per-watcher output offsets
//   k_expand_write   ordered event indices, one thread per delivery (suffix copy when revisions are monotone)
#include <algorithm>
#include <map>
#include <unordered_map>

#include "kb_internal.cuh"

struct kb_events_dev {
    DBuf keys;       // n x stride bytes: the first `stride` bytes of every event key (zero padded)
    DBuf klen;       // n x u32 true key lengths
    DBuf rev;        // n x u64
    DBuf batch_off;  // (nb+1) x u64
    uint32_t n = 0, nb = 0, stride = 0;
};

struct WatchTablesDev {
    uint32_t n_ids = 0, n_groups = 0, n_lens = 0, table_size = 0, max_len = 0;
    uint64_t d_hint = 0;  // deliveries of the previous match (sizes the next output buffer)
    DBuf gprefix, goff16, glen, ghash, gstart, gmember, wgroup, wminrev, lens, table;
    // per-call scratch
    DBuf zeros, gbase, gclass, ematch, seg, seg_sorted, bitmaps, pm, wcnt, wlo, wstart, total;
};

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint64_t FNV_OFFSET = 14695981039346656037ull;
constexpr uint64_t FNV_PRIME = 1099511628211ull;

__host__ __device__ __forceinline__ uint32_t slot_of(uint64_t h, uint32_t len, uint32_t mask)
{
    uint64_t m = h ^ ((uint64_t)len * 0x9E3779B97F4A7C15ull);
    m ^= m >> 29;
    return (uint32_t)(m ^ (m >> 32)) & mask;
}

struct EvDev {
    const uint4 *keys;
    const uint32_t *klen;
    const uint64_t *rev;
    const uint64_t *batch_off;
    uint32_t n, nb, stride16;
};

struct TabDev {
    const uint4 *gprefix;
    const uint32_t *goff16, *glen;
    const uint64_t *ghash;
    const uint32_t *gstart, *gmember, *wgroup;
    const uint64_t *wminrev;
    const uint32_t *lens;
    const uint32_t *table;
    uint32_t n_ids, n_groups, n_lens, mask, max_len;
};

// ---- running max of the revisions inside each collector batch (filterByRevision strips only the LEADING
//      events below min_rev, watch.go:153-159, so event i survives iff max(rev[batch start..i]) >= min_rev)
__device__ __forceinline__ void d_batch_pm(const EvDev &ev, uint64_t *__restrict__ pm, uint32_t *__restrict__ nonmono,
                                           uint32_t vblock)
{
    const uint32_t w = (vblock * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ev.nb) return;
    const uint64_t lo = ev.batch_off[w], hi = ev.batch_off[w + 1];
    uint64_t carry = 0;
    bool bad = false;
    for (uint64_t c = lo; c < hi; c += 32) {
        const uint64_t i = c + lane;
        uint64_t v = i < hi ? ev.rev[i] : 0;
        if (i < hi && i > 0 && ev.rev[i - 1] > v) bad = true;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(FULL, v, d);
            if (lane >= (unsigned)d) v = max(v, o);
        }
        v = max(v, carry);
        if (i < hi) pm[i] = v;
        carry = __shfl_sync(FULL, v, 31);
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(nonmono, 1u);
}

// ---- per event and distinct prefix length: the group whose prefix the key starts with (or NONE).
//      ematch[li * E + i]; group counts are aggregated inside the warp before touching memory.
__device__ __forceinline__ void d_match_count(const EvDev &ev, const TabDev &tb, uint32_t *__restrict__ ematch,
                                              uint32_t *__restrict__ gcnt, uint32_t vblock)
{
    const uint32_t i = vblock * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool valid = i < ev.n;
    const uint32_t klen = valid ? ev.klen[i] : 0;
    const uint4 *kp = ev.keys + (uint64_t)(valid ? i : 0) * ev.stride16;
    uint64_t h = FNV_OFFSET;
    uint32_t pos = 0;  // bytes hashed so far
    uint4 chunk = make_uint4(0, 0, 0, 0);
    for (uint32_t li = 0; li < tb.n_lens; li++) {  // uniform trip count: the warp collectives below need all lanes
        const uint32_t L = tb.lens[li];
        uint32_t g = KB_NONE;
        if (valid && L <= klen) {
            while (pos < L) {
                if ((pos & 15) == 0) chunk = kp[pos >> 4];
                h = (h ^ (uint64_t)byte_of(chunk, pos & 15)) * FNV_PRIME;
                pos++;
            }
            // probe (hash, length); verify the bytes so the result is exact
            uint32_t s = slot_of(h, L, tb.mask);
            for (;;) {
                const uint32_t c = tb.table[s];
                if (c == KB_NONE) break;
                if (tb.ghash[c] == h && tb.glen[c] == L) {
                    const uint4 *gp = tb.gprefix + tb.goff16[c];
                    bool eq = true;
                    for (uint32_t k = 0; k * 16 < L && eq; k++) {
                        uint4 a = kp[k], b = gp[k];
                        int p = first_diff16(a, b);
                        if (p < 16 && k * 16 + p < L) eq = false;
                    }
                    if (eq) {
                        g = c;
                        break;  // prefixes are unique per group
                    }
                }
                s = (s + 1) & tb.mask;
            }
        }
        if (valid) ematch[(uint64_t)li * ev.n + i] = g;
        const unsigned peers = __match_any_sync(FULL, g);
        if (g != KB_NONE && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&gcnt[g], (uint32_t)__popc(peers));
    }
}

// one launch: blocks [0, pm_blocks) compute the per-batch running max, the rest match events against the groups
__global__ void __launch_bounds__(256)
k_match_count(EvDev ev, TabDev tb, uint32_t pm_blocks, uint64_t *__restrict__ pm, uint32_t *__restrict__ nonmono,
              uint32_t *__restrict__ ematch, uint32_t *__restrict__ gcnt)
{
    if (blockIdx.x < pm_blocks)
        d_batch_pm(ev, pm, nonmono, blockIdx.x);
    else if (tb.n_groups)
        d_match_count(ev, tb, ematch, gcnt, blockIdx.x - pm_blocks);
}

// single CTA: exclusive scan of the group counts (segment offsets) + classification by count.
// lists layout: [0]=n_medium [1]=n_large [2..2+G) medium groups [2+G..2+2G) large groups
__global__ void __launch_bounds__(1024)
k_group_scan_classify(uint32_t n_groups, const uint32_t *__restrict__ gcnt, uint32_t big_t, uint32_t max_large,
                      uint32_t *__restrict__ gbase, uint32_t *__restrict__ gclass, uint32_t *__restrict__ lists)
{
    __shared__ uint32_t wsum[33];
    __shared__ uint32_t carry_s;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_groups; c0 += 1024) {
        const uint32_t g = c0 + threadIdx.x;
        const uint32_t n = g < n_groups ? gcnt[g] : 0;
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t v = wsum[lane], iv = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(FULL, iv, d);
                if (lane >= (unsigned)d) iv += o;
            }
            wsum[lane] = iv - v;
            if (lane == 31) wsum[32] = iv;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        if (g < n_groups) {
            gbase[g] = carry + wsum[wid] + inc - n;
            uint32_t cls = KB_NONE;  // NONE: small or medium (segment scatter); otherwise the bitmap slot
            if (n > big_t) {
                const uint32_t slot = atomicAdd(&lists[1], 1u);
                if (slot < max_large) {  // cannot overflow: sum(gcnt) <= E * n_lens
                    cls = slot;
                    lists[2 + n_groups + slot] = g;
                }
            } else if (n > 32) {
                lists[2 + atomicAdd(&lists[0], 1u)] = g;
            }
            gclass[g] = cls;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + wsum[32];
        __syncthreads();
    }
}

// single CTA: exclusive scan of the per-watcher delivery counts; wstart[W] = total
__global__ void __launch_bounds__(1024)
k_watcher_scan(uint32_t n, const uint64_t *__restrict__ wcnt, uint64_t *__restrict__ wstart, uint64_t *__restrict__ total)
{
    __shared__ uint64_t wsum[33];
    __shared__ uint64_t carry_s;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n; c0 += 1024) {
        const uint32_t w = c0 + threadIdx.x;
        const uint64_t v0 = w < n ? wcnt[w] : 0;
        uint64_t inc = v0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint64_t v = wsum[lane], iv = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t o = __shfl_up_sync(FULL, iv, d);
                if (lane >= (unsigned)d) iv += o;
            }
            wsum[lane] = iv - v;
            if (lane == 31) wsum[32] = iv;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        if (w < n) wstart[w] = carry + wsum[wid] + inc - v0;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + wsum[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        wstart[n] = carry_s;
        *total = carry_s;
    }
}

__global__ void __launch_bounds__(256)
k_scatter(uint32_t n_events, uint32_t n_lens, const uint32_t *__restrict__ ematch,
          const uint32_t *__restrict__ gclass, const uint32_t *__restrict__ gbase, uint32_t *__restrict__ gfill,
          uint32_t *__restrict__ seg, uint32_t *__restrict__ bitmaps, uint32_t bm_words)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool valid = i < n_events;
    for (uint32_t li = 0; li < n_lens; li++) {
        const uint32_t g = valid ? ematch[(uint64_t)li * n_events + i] : KB_NONE;
        const unsigned peers = __match_any_sync(FULL, g);
        if (g == KB_NONE) continue;
        const uint32_t leader = __ffs(peers) - 1;
        const uint32_t cls = gclass[g];
        if (cls != KB_NONE) {
            // the 32 events of this warp are exactly one bitmap word of the group
            if (lane == leader) bitmaps[(uint64_t)cls * bm_words + (i >> 5)] = peers;
        } else {
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&gfill[g], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            seg[gbase[g] + base + __popc(peers & ((1u << lane) - 1))] = i;
        }
    }
}

// ---- segment sort: ascending event index per group
__device__ __forceinline__ void d_sort_small(uint32_t n_groups, const uint32_t *__restrict__ gcnt,
                                             const uint32_t *__restrict__ gbase, const uint32_t *__restrict__ seg,
                                             uint32_t *__restrict__ sorted, uint32_t vblock)
{
    const uint32_t g = (vblock * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const uint32_t n = gcnt[g];
    if (n == 0 || n > 32) return;
    const uint32_t base = gbase[g];
    const uint32_t v = lane < n ? seg[base + lane] : 0xFFFFFFFFu;
    uint32_t rank = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t o = __shfl_sync(FULL, v, j);
        rank += (o < v) ? 1u : 0u;  // event indices inside one group are distinct
    }
    if (lane < n) sorted[base + rank] = v;
}

constexpr uint32_t BM_WORDS = 8192;  // 262144 event indices per window (32 KiB of shared memory)

__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *wsum /* 9 */, uint32_t &total)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(FULL, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int k = 0; k < 8; k++) {
            uint32_t x = wsum[k];
            wsum[k] = run;
            run += x;
        }
        wsum[8] = run;
    }
    __syncthreads();
    const uint32_t ex = wsum[wid] + inc - v;
    total = wsum[8];
    __syncthreads();
    return ex;
}

// one CTA per medium group (33..big_t entries): shared-memory bitmap over [min, max] of the segment
__device__ __forceinline__ void d_sort_medium(const uint32_t *__restrict__ lists, const uint32_t *__restrict__ gcnt,
                                              const uint32_t *__restrict__ gbase, const uint32_t *__restrict__ seg,
                                              uint32_t *__restrict__ sorted, uint32_t *bm, uint32_t *wsum, uint32_t *red,
                                              uint32_t vblock, uint32_t vgrid)
{
    const uint32_t nmed = lists[0];
    for (uint32_t bi = vblock; bi < nmed; bi += vgrid) {
        const uint32_t g = lists[2 + bi];
        const uint32_t n = gcnt[g], base = gbase[g];
        if (threadIdx.x == 0) {
            red[0] = 0xFFFFFFFFu;
            red[1] = 0;
        }
        __syncthreads();
        uint32_t mn = 0xFFFFFFFFu, mx = 0;
        for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t v = seg[base + j];
            mn = min(mn, v);
            mx = max(mx, v);
        }
        atomicMin(&red[0], mn);
        atomicMax(&red[1], mx);
        __syncthreads();
        mn = red[0];
        mx = red[1];
        uint32_t outpos = 0;
        for (uint64_t w0 = mn & ~31u; w0 <= mx; w0 += (uint64_t)BM_WORDS * 32) {
            const uint64_t need_words = ((uint64_t)mx - w0) / 32 + 1;
            const uint32_t nwords = need_words < BM_WORDS ? (uint32_t)need_words : BM_WORDS;
            for (uint32_t j = threadIdx.x; j < nwords; j += blockDim.x) bm[j] = 0;
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
                const uint64_t v = seg[base + j];
                if (v >= w0 && v < w0 + (uint64_t)nwords * 32) {
                    const uint32_t d = (uint32_t)(v - w0);
                    atomicOr(&bm[d >> 5], 1u << (d & 31));
                }
            }
            __syncthreads();
            // ordered expansion: each thread owns a contiguous run of words
            const uint32_t per = (nwords + blockDim.x - 1) / blockDim.x;
            const uint32_t wlo = min(nwords, threadIdx.x * per), whi = min(nwords, wlo + per);
            uint32_t cnt = 0;
            for (uint32_t j = wlo; j < whi; j++) cnt += __popc(bm[j]);
            uint32_t total;
            uint32_t at = outpos + block_excl_scan_u32(cnt, wsum, total);
            for (uint32_t j = wlo; j < whi; j++) {
                uint32_t bits = bm[j];
                while (bits) {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    sorted[base + at++] = (uint32_t)(w0 + (uint64_t)j * 32 + b);
                }
            }
            outpos += total;
            __syncthreads();
        }
        __syncthreads();
    }
}

// large groups: ordered expansion of the global bitmap; CTA = (large group, chunk of 256 words)
__device__ __forceinline__ void d_expand_large(const uint32_t *__restrict__ lists, uint32_t n_groups,
                                               const uint32_t *__restrict__ gbase, const uint32_t *__restrict__ bitmaps,
                                               uint32_t bm_words, uint32_t chunks_per_group,
                                               uint32_t *__restrict__ sorted, uint32_t *wsum, uint32_t *pre_sp,
                                               uint32_t vblock, uint32_t vgrid)
{
    uint32_t &pre_s = *pre_sp;
    const uint32_t nlarge = lists[1];
    for (uint32_t job = vblock; job < nlarge * chunks_per_group; job += vgrid) {
        const uint32_t slot = job / chunks_per_group, chunk = job % chunks_per_group;
        const uint32_t g = lists[2 + n_groups + slot];
        const uint32_t *bm = bitmaps + (uint64_t)slot * bm_words;
        // matches in the words before this chunk
        uint32_t part = 0;
        for (uint32_t j = threadIdx.x; j < chunk * 256; j += blockDim.x) part += __popc(bm[j]);
        uint32_t tot;
        block_excl_scan_u32(part, wsum, tot);
        if (threadIdx.x == 0) pre_s = tot;
        __syncthreads();
        const uint32_t wi = chunk * 256 + threadIdx.x;
        uint32_t bits = wi < bm_words ? bm[wi] : 0;
        uint32_t total;
        uint32_t at = gbase[g] + pre_s + block_excl_scan_u32(__popc(bits), wsum, total);
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            sorted[at++] = wi * 32 + b;
        }
        __syncthreads();
    }
}

// one launch for the three segment-ordering paths: blocks [0,nb_small) warp rank-sort, [nb_small, nb_small+nb_med)
// shared-memory bitmap sort of medium groups, the rest expand the large groups' global bitmaps
__global__ void __launch_bounds__(256)
k_sort(uint32_t n_groups, const uint32_t *__restrict__ gcnt, const uint32_t *__restrict__ gbase,
       const uint32_t *__restrict__ seg, uint32_t *__restrict__ sorted, const uint32_t *__restrict__ lists,
       const uint32_t *__restrict__ bitmaps, uint32_t bm_words, uint32_t chunks_per_group, uint32_t nb_small,
       uint32_t nb_med)
{
    __shared__ uint32_t bm[BM_WORDS];
    __shared__ uint32_t wsum[9];
    __shared__ uint32_t red[2];
    if (blockIdx.x < nb_small)
        d_sort_small(n_groups, gcnt, gbase, seg, sorted, blockIdx.x);
    else if (blockIdx.x < nb_small + nb_med)
        d_sort_medium(lists, gcnt, gbase, seg, sorted, bm, wsum, red, blockIdx.x - nb_small, nb_med);
    else
        d_expand_large(lists, n_groups, gbase, bitmaps, bm_words, chunks_per_group, sorted, wsum, red,
                       blockIdx.x - nb_small - nb_med, gridDim.x - nb_small - nb_med);
}

// ---- per watcher: deliveries = its group's sorted segment filtered by pm[e] >= min_rev
__global__ void __launch_bounds__(256)
k_expand_count(TabDev tb, const uint32_t *__restrict__ gcnt, const uint32_t *__restrict__ gbase,
               const uint32_t *__restrict__ sorted, const uint64_t *__restrict__ pm,
               const uint32_t *__restrict__ nonmono, uint64_t *__restrict__ wcnt, uint32_t *__restrict__ wlo)
{
    if (*nonmono == 0) {
        // revisions non-decreasing over the whole slab: the survivors are a suffix of the segment -> one thread
        // per watcher, binary search for the first event at or above min_rev
        const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
        if (w >= tb.n_ids) return;
        const uint32_t g = tb.wgroup[w];
        uint32_t n = 0, lo = 0;
        if (g != KB_NONE) {
            n = gcnt[g];
            const uint32_t *M = sorted + gbase[g];
            const uint64_t mr = tb.wminrev[w];
            uint32_t hi = n;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (pm[M[mid]] >= mr) hi = mid; else lo = mid + 1;
            }
        }
        wcnt[w] = n - lo;
        wlo[w] = lo;
        return;
    }
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= tb.n_ids) return;
    const uint32_t g = tb.wgroup[w];
    if (g == KB_NONE) {
        if (lane == 0) {
            wcnt[w] = 0;
            wlo[w] = 0;
        }
        return;
    }
    const uint32_t n = gcnt[g];
    const uint32_t *M = sorted + gbase[g];
    const uint64_t mr = tb.wminrev[w];
    uint64_t total = 0;
    for (uint32_t c = 0; c < n; c += 32) {
        const uint32_t j = c + lane;
        const bool keep = j < n && pm[M[j]] >= mr;
        total += __popc(__ballot_sync(FULL, keep));
    }
    if (lane == 0) {
        wcnt[w] = total;
        wlo[w] = 0;
    }
}

// monotone revisions: one thread per delivery (suffix copy)
__device__ __forceinline__ void d_expand_write(const TabDev &tb, const uint32_t *__restrict__ gbase,
                                               const uint32_t *__restrict__ sorted, const uint64_t *__restrict__ wstart,
                                               const uint32_t *__restrict__ wlo, uint64_t n_deliveries,
                                               uint32_t *__restrict__ out)
{
    for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_deliveries;
         d += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = tb.n_ids;  // last watcher with wstart[w] <= d
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (wstart[mid] <= d) lo = mid; else hi = mid;
    }
    const uint32_t w = lo;
    const uint32_t g = tb.wgroup[w];
    out[d] = sorted[gbase[g] + wlo[w] + (uint32_t)(d - wstart[w])];
    }
}

// non-monotone revisions (general case): warp per watcher, ordered filtered copy
__device__ __forceinline__ void d_expand_write_general(const TabDev &tb, const uint32_t *__restrict__ gcnt,
                                                       const uint32_t *__restrict__ gbase,
                                                       const uint32_t *__restrict__ sorted,
                                                       const uint64_t *__restrict__ pm,
                                                       const uint64_t *__restrict__ wstart, uint32_t *__restrict__ out)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= tb.n_ids) return;
    const uint32_t g = tb.wgroup[w];
    if (g == KB_NONE) return;
    const uint32_t n = gcnt[g];
    const uint32_t *M = sorted + gbase[g];
    const uint64_t mr = tb.wminrev[w];
    const uint64_t o = wstart[w];
    uint64_t total = 0;
    for (uint32_t c = 0; c < n; c += 32) {
        const uint32_t j = c + lane;
        const uint32_t e = j < n ? M[j] : 0;
        const bool keep = j < n && pm[e] >= mr;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) out[o + total + __popc(m & ((1u << lane) - 1))] = e;
        total += __popc(m);
    }
}

// one launch for both output paths: revisions monotone -> one thread per delivery; otherwise warp per watcher
__global__ void __launch_bounds__(256)
k_expand_write(TabDev tb, const uint32_t *__restrict__ gcnt, const uint32_t *__restrict__ gbase,
               const uint32_t *__restrict__ sorted, const uint64_t *__restrict__ pm,
               const uint32_t *__restrict__ nonmono, const uint64_t *__restrict__ wstart,
               const uint32_t *__restrict__ wlo, uint64_t capacity, uint32_t *__restrict__ out)
{
    const uint64_t n_deliveries = wstart[tb.n_ids];
    if (n_deliveries > capacity) return;  // the host sees the total, grows the buffer and launches again
    if (*nonmono == 0)
        d_expand_write(tb, gbase, sorted, wstart, wlo, n_deliveries, out);
    else
        d_expand_write_general(tb, gcnt, gbase, sorted, pm, wstart, out);
}

uint64_t fnv1a(const std::string &s)
{
    uint64_t h = FNV_OFFSET;
    for (unsigned char c : s) h = (h ^ c) * FNV_PRIME;
    return h;
}

int upload(kb_ctx *ctx, DBuf &b, const void *src, size_t bytes)
{
    KB_TRY(dbuf_ensure(ctx, b, std::max<size_t>(bytes, 16)));
    if (bytes) KB_CUDA(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KB_OK;
}

// rebuild the device tables after kb_watch_add / kb_watch_del
int rebuild_tables(kb_ctx *ctx)
{
    if (!ctx->wt) ctx->wt = new WatchTablesDev();
    WatchTablesDev &T = *ctx->wt;
    const uint32_t n_ids = (uint32_t)ctx->watchers.size();
    std::map<std::string, std::vector<uint32_t>> groups;  // ordered: deterministic group ids
    for (uint32_t w = 0; w < n_ids; w++)
        if (ctx->watchers[w].live) groups[ctx->watchers[w].prefix].push_back(w);
    const uint32_t G = (uint32_t)groups.size();
    std::vector<uint8_t> gprefix;
    std::vector<uint32_t> goff16(G + 1), glen(G), gstart(G + 1), gmember, wgroup(std::max(n_ids, 1u), KB_NONE), lens;
    std::vector<uint64_t> ghash(std::max(G, 1u)), wminrev(std::max(n_ids, 1u), 0);
    uint32_t gi = 0, max_len = 0;
    for (auto &kv : groups) {
        const std::string &p = kv.first;
        goff16[gi] = (uint32_t)(gprefix.size() / 16);
        glen[gi] = (uint32_t)p.size();
        ghash[gi] = fnv1a(p);
        gprefix.insert(gprefix.end(), p.begin(), p.end());
        gprefix.resize((gprefix.size() + 15) / 16 * 16 + 16, 0);
        gstart[gi] = (uint32_t)gmember.size();
        for (uint32_t w : kv.second) {
            gmember.push_back(w);
            wgroup[w] = gi;
        }
        lens.push_back((uint32_t)p.size());
        max_len = std::max<uint32_t>(max_len, (uint32_t)p.size());
        gi++;
    }
    goff16[G] = (uint32_t)(gprefix.size() / 16);
    gstart[G] = (uint32_t)gmember.size();
    for (uint32_t w = 0; w < n_ids; w++) wminrev[w] = ctx->watchers[w].min_rev;
    std::sort(lens.begin(), lens.end());
    lens.erase(std::unique(lens.begin(), lens.end()), lens.end());
    uint32_t tsize = 16;
    while (tsize < 2 * G + 2) tsize <<= 1;
    std::vector<uint32_t> table(tsize, KB_NONE);
    for (uint32_t g = 0; g < G; g++) {
        uint32_t s = slot_of(ghash[g], glen[g], tsize - 1);
        while (table[s] != KB_NONE) s = (s + 1) & (tsize - 1);
        table[s] = g;
    }
    if (gprefix.empty()) gprefix.resize(16, 0);
    KB_TRY(upload(ctx, T.gprefix, gprefix.data(), gprefix.size()));
    KB_TRY(upload(ctx, T.goff16, goff16.data(), goff16.size() * 4));
    KB_TRY(upload(ctx, T.glen, glen.data(), glen.size() * 4));
    KB_TRY(upload(ctx, T.ghash, ghash.data(), ghash.size() * 8));
    KB_TRY(upload(ctx, T.gstart, gstart.data(), gstart.size() * 4));
    KB_TRY(upload(ctx, T.gmember, gmember.data(), gmember.size() * 4));
    KB_TRY(upload(ctx, T.wgroup, wgroup.data(), wgroup.size() * 4));
    KB_TRY(upload(ctx, T.wminrev, wminrev.data(), wminrev.size() * 8));
    KB_TRY(upload(ctx, T.lens, lens.data(), lens.size() * 4));
    KB_TRY(upload(ctx, T.table, table.data(), table.size() * 4));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors die here
    T.n_ids = n_ids;
    T.n_groups = G;
    T.n_lens = (uint32_t)lens.size();
    T.table_size = tsize;
    T.max_len = max_len;
    ctx->watch_dirty = false;
    return KB_OK;
}

uint32_t needed_stride(kb_ctx *ctx)
{
    uint32_t max_len = 0;
    for (auto &w : ctx->watchers)
        if (w.live) max_len = std::max<uint32_t>(max_len, (uint32_t)w.prefix.size());
    return std::max<uint32_t>(16, (max_len + 15) / 16 * 16);
}

int events_upload_locked(kb_ctx *ctx, const kb_events *ev, kb_events_dev *d)
{
    if (ev->n >= 0xFFFFFFF0ull) return kb_fail(ctx, KB_ELIMIT, "too many events");
    const uint32_t n = (uint32_t)ev->n;
    const uint32_t stride = needed_stride(ctx);
    const uint32_t nb = ev->batch_off && ev->n_batches ? (uint32_t)ev->n_batches : 1;
    // pinned staging: [keys n*stride][klen n*4][rev n*8][batch_off (nb+1)*8]
    const size_t kbytes = (size_t)n * stride, total = kbytes + (size_t)n * 12 + (size_t)(nb + 1) * 8 + 64;
    KB_TRY(hbuf_ensure(ctx, ctx->h_stage, total));
    uint8_t *h = (uint8_t *)ctx->h_stage.p;
    uint32_t *hl = (uint32_t *)(h + kbytes);
    uint64_t *hr = (uint64_t *)(h + kbytes + (size_t)n * 4);
    // keep 8-byte alignment for the u64 arrays
    size_t rev_off = (kbytes + (size_t)n * 4 + 7) & ~(size_t)7;
    hr = (uint64_t *)(h + rev_off);
    uint64_t *hb = hr + n;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t o = ev->key_off[i], l = ev->key_off[i + 1] - o;
        const uint32_t c = (uint32_t)std::min<uint64_t>(l, stride);
        uint8_t *dst = h + (size_t)i * stride;
        memcpy(dst, ev->keys + o, c);
        if (c < stride) memset(dst + c, 0, stride - c);
        hl[i] = (uint32_t)std::min<uint64_t>(l, 0xFFFFFFFFull);
    }
    if (n) memcpy(hr, ev->rev, (size_t)n * 8);
    if (ev->batch_off && ev->n_batches) {
        memcpy(hb, ev->batch_off, (size_t)(nb + 1) * 8);
    } else {
        hb[0] = 0;
        hb[1] = n;
    }
    KB_TRY(dbuf_ensure(ctx, d->keys, std::max<size_t>(kbytes, 16)));
    KB_TRY(dbuf_ensure(ctx, d->klen, std::max<size_t>((size_t)n * 4, 16)));
    KB_TRY(dbuf_ensure(ctx, d->rev, std::max<size_t>((size_t)n * 8, 16)));
    KB_TRY(dbuf_ensure(ctx, d->batch_off, (size_t)(nb + 1) * 8));
    if (n) {
        KB_CUDA(ctx, cudaMemcpyAsync(d->keys.p, h, kbytes, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(ctx, cudaMemcpyAsync(d->klen.p, hl, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(ctx, cudaMemcpyAsync(d->rev.p, hr, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    KB_CUDA(ctx, cudaMemcpyAsync(d->batch_off.p, hb, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    d->n = n;
    d->nb = nb;
    d->stride = stride;
    return KB_OK;
}

void events_release(kb_events_dev *d)
{
    if (d->keys.p) cudaFree(d->keys.p);
    if (d->klen.p) cudaFree(d->klen.p);
    if (d->rev.p) cudaFree(d->rev.p);
    if (d->batch_off.p) cudaFree(d->batch_off.p);
    d->keys = d->klen = d->rev = d->batch_off = DBuf();
}

}  // namespace

void watch_tables_free(kb_ctx *ctx)
{
    if (!ctx->wt) return;
    WatchTablesDev &T = *ctx->wt;
    DBuf *all[] = {&T.gprefix, &T.goff16, &T.glen, &T.ghash, &T.gstart, &T.gmember, &T.wgroup, &T.wminrev,
                   &T.lens, &T.table, &T.zeros, &T.gbase, &T.gclass, &T.ematch, &T.seg, &T.seg_sorted, &T.bitmaps, &T.wlo, &T.pm,
                   &T.wcnt, &T.wstart, &T.total};
    for (DBuf *b : all)
        if (b->p) cudaFree(b->p);
    delete ctx->wt;
    ctx->wt = nullptr;
    if (ctx->ev_scratch) {
        events_release(ctx->ev_scratch);
        delete ctx->ev_scratch;
        ctx->ev_scratch = nullptr;
    }
}

extern "C" int kb_watch_add(kb_ctx *ctx, const uint8_t *prefix, uint64_t prefix_len, uint64_t min_rev, uint32_t *id)
{
    if (!ctx || !id || (prefix_len && !prefix)) return KB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (prefix_len > 65535) return kb_fail(ctx, KB_ELIMIT, "watch prefix longer than 65535 bytes");
    Watcher w;
    w.prefix.assign((const char *)prefix, (size_t)prefix_len);
    w.min_rev = min_rev;
    w.live = true;
    if (!ctx->free_watch_ids.empty()) {  // ids of cancelled watches are reused so the id space stays dense
        *id = ctx->free_watch_ids.back();
        ctx->free_watch_ids.pop_back();
        ctx->watchers[*id] = w;
    } else {
        ctx->watchers.push_back(w);
        *id = (uint32_t)ctx->watchers.size() - 1;
    }
    ctx->watch_dirty = true;
    return KB_OK;
}

extern "C" int kb_watch_del(kb_ctx *ctx, uint32_t id)
{
    if (!ctx) return KB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (id >= ctx->watchers.size() || !ctx->watchers[id].live) return kb_fail(ctx, KB_EINVAL, "unknown watcher %u", id);
    ctx->watchers[id].live = false;
    ctx->free_watch_ids.push_back(id);
    ctx->watch_dirty = true;
    return KB_OK;
}

extern "C" int kb_watch_count(kb_ctx *ctx, uint64_t *n)
{
    if (!ctx || !n) return KB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    uint64_t c = 0;
    for (auto &w : ctx->watchers) c += w.live ? 1 : 0;
    *n = c;
    return KB_OK;
}

extern "C" int kb_events_upload(kb_ctx *ctx, const kb_events *ev, kb_events_dev **out)
{
    if (!ctx || !ev || !out || (ev->n && (!ev->keys || !ev->key_off || !ev->rev))) return KB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    kb_events_dev *d = new kb_events_dev();
    int rc = events_upload_locked(ctx, ev, d);
    if (rc == KB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = kb_fail(ctx, KB_ECUDA, "event upload");
    if (rc != KB_OK) {
        events_release(d);
        delete d;
        return rc;
    }
    *out = d;
    return KB_OK;
}

extern "C" void kb_events_free(kb_ctx *ctx, kb_events_dev *ev)
{
    if (!ev) return;
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->mu);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        events_release(ev);
    }
    delete ev;
}

// the delivery total, handed to the host through mapped pinned memory as soon as it is known (behind the offsets copy,
// in front of the write kernel): a device-resident match returns while its delivery lists are still being written
__global__ void k_publish_total(const uint64_t *__restrict__ total, uint64_t *host, uint64_t epoch)
{
    if (threadIdx.x == 0) {
        host[1] = *total;
        __threadfence_system();
        *(volatile uint64_t *)host = epoch;
    }
}

static int wpub_wait(kb_ctx *ctx, uint64_t epoch)
{
    volatile uint64_t *flag = ctx->h_wpub;
    for (uint64_t spins = 1;; spins++) {
        if (*flag == epoch) return KB_OK;
        kb_cpu_relax();
        if ((spins & 0xFFFF) == 0) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) return *flag == epoch ? KB_OK : kb_fail(ctx, KB_ECUDA, "watch match: total was not published");
            if (q != cudaErrorNotReady) return kb_cuda_fail(ctx, q, "watch match");
        }
    }
}

static int match_locked(kb_ctx *ctx, const kb_events_dev *d, int out_mode, kb_result **out)
{
    kb_tp tseg = kb_now();
    if (ctx->watch_dirty || !ctx->wt) KB_TRY(rebuild_tables(ctx));
    WatchTablesDev &T = *ctx->wt;
    if (d->stride < needed_stride(ctx))
        return kb_fail(ctx, KB_ESTATE, "event slab was uploaded for shorter watcher prefixes (stride %u < %u); upload again",
                       d->stride, needed_stride(ctx));
    const uint32_t E = d->n, G = T.n_groups, W = T.n_ids, NL = std::max(T.n_lens, 1u);
    const uint64_t seg_cap = std::max<uint64_t>((uint64_t)E * NL, 1);
    if (seg_cap >= 0xFFFFFFF0ull) return kb_fail(ctx, KB_ELIMIT, "events x distinct prefix lengths exceeds 2^32");
    // groups matching more than big_t events get a global bitmap; at most E*NL/big_t of them can exist
    const uint32_t big_t = std::max<uint32_t>(1024, E / 64);
    const uint32_t max_large = (uint32_t)(seg_cap / big_t) + 1;
    const uint32_t bm_words = (E + 31) / 32;
    const uint32_t chunks_per_group = (bm_words + 255) / 256;
    KB_TRY(dbuf_ensure(ctx, T.zeros, (size_t)(4 + 2 * (G + 1) + 2 * G + 8) * 4));
    KB_TRY(dbuf_ensure(ctx, T.gbase, (size_t)(G + 1) * 4));
    KB_TRY(dbuf_ensure(ctx, T.gclass, (size_t)(G + 1) * 4));
    KB_TRY(dbuf_ensure(ctx, T.ematch, seg_cap * 4));
    KB_TRY(dbuf_ensure(ctx, T.seg, seg_cap * 4));
    KB_TRY(dbuf_ensure(ctx, T.seg_sorted, seg_cap * 4));
    KB_TRY(dbuf_ensure(ctx, T.bitmaps, std::max<size_t>((size_t)max_large * bm_words * 4, 16)));
    KB_TRY(dbuf_ensure(ctx, T.pm, std::max<size_t>((size_t)E * 8, 16)));
    KB_TRY(dbuf_ensure(ctx, T.wcnt, (size_t)(W + 1) * 8));
    KB_TRY(dbuf_ensure(ctx, T.wlo, (size_t)(W + 1) * 4));
    KB_TRY(dbuf_ensure(ctx, T.wstart, (size_t)(W + 2) * 8));
    KB_TRY(dbuf_ensure(ctx, T.total, 16));

    EvDev ev;
    ev.keys = (const uint4 *)d->keys.p;
    ev.klen = (const uint32_t *)d->klen.p;
    ev.rev = (const uint64_t *)d->rev.p;
    ev.batch_off = (const uint64_t *)d->batch_off.p;
    ev.n = E;
    ev.nb = d->nb;
    ev.stride16 = d->stride / 16;
    TabDev tb;
    tb.gprefix = (const uint4 *)T.gprefix.p;
    tb.goff16 = (const uint32_t *)T.goff16.p;
    tb.glen = (const uint32_t *)T.glen.p;
    tb.ghash = (const uint64_t *)T.ghash.p;
    tb.gstart = (const uint32_t *)T.gstart.p;
    tb.gmember = (const uint32_t *)T.gmember.p;
    tb.wgroup = (const uint32_t *)T.wgroup.p;
    tb.wminrev = (const uint64_t *)T.wminrev.p;
    tb.lens = (const uint32_t *)T.lens.p;
    tb.table = (const uint32_t *)T.table.p;
    tb.n_ids = W;
    tb.n_groups = G;
    tb.n_lens = T.n_lens;
    tb.mask = T.table_size - 1;
    tb.max_len = T.max_len;

    // zeroed region: [flag x4][gcnt G+1][gfill G+1][lists 2G+4]
    const size_t zero_bytes = (size_t)(4 + 2 * (G + 1) + 2) * 4;
    uint32_t *zr = (uint32_t *)T.zeros.p;
    uint32_t *gcnt = zr + 4, *gfill = zr + 4 + (G + 1), *lists = zr + 4 + 2 * (G + 1);
    uint32_t *gbase = (uint32_t *)T.gbase.p;
    uint32_t *gclass = (uint32_t *)T.gclass.p, *ematch = (uint32_t *)T.ematch.p;
    uint32_t *seg = (uint32_t *)T.seg.p, *sorted = (uint32_t *)T.seg_sorted.p, *bitmaps = (uint32_t *)T.bitmaps.p;
    uint64_t *pm = (uint64_t *)T.pm.p;
    uint32_t *flag = zr, *wlo = (uint32_t *)T.wlo.p;
    uint64_t *wcnt = (uint64_t *)T.wcnt.p, *wstart = (uint64_t *)T.wstart.p, *total = (uint64_t *)T.total.p;

    // gcnt | gfill | lists header | flag live in one zeroed region (one memset per call)
    KB_CUDA(ctx, cudaMemsetAsync(T.zeros.p, 0, zero_bytes, ctx->stream));
    const uint64_t ev_bytes = (uint64_t)E * (d->stride + 4);
    const bool work = E && W && G;
    if (E && W) {
        const uint32_t pm_blocks = (d->nb * 32 + 255) / 256;
        KB_LAUNCH(ctx, "k_match_count", ev_bytes + (uint64_t)E * NL * 4 + (uint64_t)E * 16,
                  (k_match_count<<<pm_blocks + (G ? (E + 255) / 256 : 0), 256, 0, ctx->stream>>>(ev, tb, pm_blocks, pm,
                                                                                            flag, ematch, gcnt)));
    }
    if (work) {
        KB_CUDA(ctx, cudaMemsetAsync(bitmaps, 0, (size_t)max_large * bm_words * 4, ctx->stream));
        KB_LAUNCH(ctx, "k_group_scan_classify", (uint64_t)G * 12,
                  (k_group_scan_classify<<<1, 1024, 0, ctx->stream>>>(G, gcnt, big_t, max_large, gbase, gclass, lists)));
        KB_LAUNCH(ctx, "k_scatter", (uint64_t)E * NL * 8,
                  (k_scatter<<<(E + 255) / 256, 256, 0, ctx->stream>>>(E, T.n_lens, ematch, gclass, gbase, gfill, seg,
                                                                     bitmaps, bm_words)));
        const uint32_t nb_small = (G * 32 + 255) / 256, nb_med = 148 * 2;
        const uint32_t nb_large = std::min<uint32_t>(148 * 4, max_large * chunks_per_group);
        KB_LAUNCH(ctx, "k_sort", (uint64_t)E * NL * 8,
                  (k_sort<<<nb_small + nb_med + nb_large, 256, 0, ctx->stream>>>(G, gcnt, gbase, seg, sorted, lists, bitmaps,
                                                                               bm_words, chunks_per_group, nb_small,
                                                                               nb_med)));
    }
    if (W) {
        KB_LAUNCH(ctx, "k_expand_count", (uint64_t)W * 24,
                  (k_expand_count<<<(W * 32 + 255) / 256, 256, 0, ctx->stream>>>(tb, gcnt, gbase, sorted, pm, flag, wcnt,
                                                                              wlo)));
    }
    KB_LAUNCH(ctx, "k_watcher_scan", (uint64_t)W * 16,
              (k_watcher_scan<<<1, 1024, 0, ctx->stream>>>(W, wcnt, wstart, total)));
    // Output [start (W+1) x u64][event_idx D x u32].  D is only known on the device; the buffer is sized from the
    // previous call's D (+25 %) and the write kernel refuses to run when it would not fit, so the steady state needs
    // no round trip before the write.
    KB_TRY(hbuf_ensure(ctx, ctx->h_stage2, 64));
    if (!ctx->h_wpub) {
        KB_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_wpub, 64, cudaHostAllocMapped));
        memset(ctx->h_wpub, 0, 64);
    }
    uint64_t cap = std::max<uint64_t>(T.d_hint + T.d_hint / 4 + 4096, 1 << 16);
    uint64_t D = 0;
    DBuf d_out;
    HBuf h_out;
    int rc = KB_OK;
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t out_bytes = (size_t)(W + 1) * 8 + cap * 4 + 16;
        KB_TRY(pool_get_dev(ctx, out_bytes, &d_out));
        uint64_t *o_start = (uint64_t *)d_out.p;
        uint32_t *o_idx = (uint32_t *)(o_start + W + 1);
        cudaMemcpyAsync(o_start, wstart, (size_t)(W + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        uint64_t wepoch = 0;
        if (out_mode != KB_OUT_HOST) {
            if (!h_out.p) KB_TRY(pool_get_host(ctx, (size_t)(W + 1) * 8 + 16, &h_out));
            cudaMemcpyAsync(h_out.p, wstart, (size_t)(W + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
            wepoch = ++ctx->wpub_epoch;
            k_publish_total<<<1, 32, 0, ctx->stream>>>(total, ctx->h_wpub, wepoch);
        }
        if (W) {
            const unsigned wgrid = (unsigned)std::max<uint64_t>(std::min<uint64_t>((cap + 255) / 256, 148 * 16),
                                                                ((uint64_t)W * 32 + 255) / 256);
            KB_LAUNCH(ctx, "k_expand_write", cap * 8,
                      (k_expand_write<<<wgrid, 256, 0, ctx->stream>>>(tb, gcnt, gbase, sorted, pm, flag, wstart, wlo, cap,
                                                                     o_idx)));
        }
        cudaError_t e0 = cudaSuccess;
        if (out_mode != KB_OUT_HOST) {
            // device-resident result: the offsets have travelled and the total is published in front of the write
            // kernel; the host returns on the flag while the delivery lists are still being written (stream order)
            kb_seg(ctx, "host:match_launch", tseg);
            rc = wpub_wait(ctx, wepoch);
            kb_seg(ctx, "host:match_sync", tseg);
            if (rc != KB_OK) {
                pool_put_dev(ctx, d_out);
                pool_put_host(ctx, h_out);
                return rc;
            }
            D = ctx->h_wpub[1];
        } else {
            KB_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage2.p, total, 8, cudaMemcpyDeviceToHost, ctx->stream));
            kb_seg(ctx, "host:match_launch", tseg);
            e0 = cudaStreamSynchronize(ctx->stream);
            kb_seg(ctx, "host:match_sync", tseg);
            if (e0 != cudaSuccess) {
                pool_put_dev(ctx, d_out);
                return kb_cuda_fail(ctx, e0, "watch match");
            }
            D = *(uint64_t *)ctx->h_stage2.p;
        }
        T.d_hint = D;
        if (D <= cap) break;
        pool_put_dev(ctx, d_out);  // first call or a burst larger than the hint: grow and write again
        d_out = DBuf();
        cap = D;
    }
    if (out_mode == KB_OUT_HOST) {
        rc = pool_get_host(ctx, (size_t)(W + 1) * 8 + D * 4 + 16, &h_out);
        if (rc == KB_OK)
            cudaMemcpyAsync(h_out.p, d_out.p, (size_t)(W + 1) * 8 + D * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        kb_seg(ctx, "host:match_d2h", tseg);
        if (rc == KB_OK && e != cudaSuccess) rc = kb_cuda_fail(ctx, e, "watch match");
    }
    if (rc != KB_OK) {
        pool_put_dev(ctx, d_out);
        pool_put_host(ctx, h_out);
        return rc;
    }
    kb_result *res = kb_result_new(3, out_mode);
    if (out_mode == KB_OUT_HOST) {
        pool_put_dev(ctx, d_out);
        d_out = DBuf();
    }
    res->n_watchers = W;
    res->n_deliveries = D;
    res->h_match = h_out;
    res->d_match = d_out;
    *out = res;
    return KB_OK;
}

extern "C" int kb_watch_match_dev(kb_ctx *ctx, const kb_events_dev *ev, int out_mode, kb_result **out)
{
    if (!ctx || !ev || !out || (out_mode != KB_OUT_HOST && out_mode != KB_OUT_DEVICE)) return KB_EINVAL;
    *out = nullptr;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    return match_locked(ctx, ev, out_mode, out);
}

extern "C" int kb_watch_match(kb_ctx *ctx, const kb_events *ev, int out_mode, kb_result **out)
{
    if (!ctx || !ev || !out || (out_mode != KB_OUT_HOST && out_mode != KB_OUT_DEVICE)) return KB_EINVAL;
    if (ev->n && (!ev->keys || !ev->key_off || !ev->rev)) return KB_EINVAL;
    *out = nullptr;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!ctx->ev_scratch) ctx->ev_scratch = new kb_events_dev();
    KB_TRY(events_upload_locked(ctx, ev, ctx->ev_scratch));
    return match_locked(ctx, ctx->ev_scratch, out_mode, out);
}

extern "C" int kb_match_view_get(const kb_result *res, kb_match_view *v)
{
    if (!res || !v || res->type != 3) return KB_EINVAL;
    memset(v, 0, sizeof(*v));
    v->n_watchers = res->n_watchers;
    v->n_deliveries = res->n_deliveries;
    v->on_device = res->out_mode == KB_OUT_DEVICE;
    v->start = (const uint64_t *)res->h_match.p;  // offsets are always host readable
    if (v->on_device)
        v->event_idx = (const uint32_t *)((const uint64_t *)res->d_match.p + res->n_watchers + 1);
    else
        v->event_idx = (const uint32_t *)((const uint64_t *)res->h_match.p + res->n_watchers + 1);
    return KB_OK;
}
