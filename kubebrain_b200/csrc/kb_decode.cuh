// kb_decode.cuh -- k_decode_lcp: the HBM-bound pass of the scan (included by kb_scan.cu).
//
// Streams the raw internal keys of every examined record once (one bulk-TMA copy per 32-record sub-tile into a
// per-warp shared-memory ring, two stages deep so the next sub-tile is in flight while the current one is decoded),
// and reduces each record to one 32-bit meta word:
//   bits 0..15  LCP with the preceding key (common-prefix length, the input of the "same user key" test)
//   bits 16..23 decode / visibility / tombstone / compaction-class flags (KB_M_*)
// plus one (last PREVOK slot, min LCP after it) aggregate per 32-record sub-tile for the cross-tile carry.
//
// Replaces coder.Decode (pkg/backend/coder/normal.go:58-70) and the per-record front half of worker.run
// (pkg/backend/scanner/scanner.go:430-453, 471-491, 566-591): decode, TTL expiry, revision visibility,
// tombstone test, deleted-flag revision-record test.  Warps are persistent and fully independent (no CTA barrier).
//
// Per-warp software pipeline (every stage one iteration apart, so no load is waited for in the iteration that
// issues it):   tile descriptor -> record directory (koff16/klen/vlen/voff16) -> key bytes (cp.async.bulk, completion
// on an mbarrier) + 16-byte value probe of 9-byte values -> decode.
#pragma once

#include "kb_internal.cuh"

namespace {

constexpr uint32_t MAGIC_LE = 0x8b80fb57u;  // bytes 57 fb 80 8b (coder/normal.go:26)
constexpr int DECODE_WARPS = 12;
constexpr int DECODE_STAGES = 2;

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ---- bulk (TMA) copy + mbarrier helpers: one instruction moves a whole sub-tile's key bytes into shared memory
__device__ __forceinline__ uint32_t dsmem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmbar_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dsmem_u32(bar)));
}
__device__ __forceinline__ void dmbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KBD_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KBD_DONE;\n"
        "bra KBD_WAIT;\n"
        "KBD_DONE:\n"
        "}\n" ::"r"(dsmem_u32(bar)),
        "r"(parity)
        : "memory");
}

// one plain atomic by the calling lane.  Written in PTX because nvcc turns `if (lane == 0) atomicAdd(...)` into its
// warp-aggregated form, whose result broadcast (SHFL) waits for the atomic at once and defeats issuing it a step early.
__device__ __forceinline__ uint32_t atom_add_u32_raw(unsigned int *p, uint32_t v)
{
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v));
    return r;
}

__device__ __forceinline__ bool contains_events(const uint8_t *uk, uint32_t n)  // bytes.Contains(rawKey, "/events/")
{
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        w = (w << 8) | uk[i];
        if (i >= 7 && w == 0x2f6576656e74732full) return true;
    }
    return false;
}

// pipeline stage 0: the tile a sub-tile belongs to
struct TileRef {
    uint32_t valid, sid;
    TileDev t;
};

__device__ __forceinline__ TileRef fetch_tile(const TileDev *__restrict__ tiles, uint32_t sid, uint32_t n_sub)
{
    TileRef r;
    r.valid = sid < n_sub;
    r.sid = sid;
    if (r.valid) {
        const uint4 *p = (const uint4 *)(tiles + (sid >> 5));
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        r.t.req = a.x;
        r.t.rec0 = a.y;
        r.t.n = a.z;
        r.t.flat0 = a.w;
        r.t.lo = b.x;
        r.t.pad = b.y;
        r.t.read_rev = ((uint64_t)b.w << 32) | b.z;
    } else {
        r.t.req = r.t.rec0 = r.t.n = r.t.flat0 = r.t.lo = r.t.pad = 0;
        r.t.read_rev = 0;
    }
    return r;
}

// everything a warp needs to know about one 32-record sub-tile
struct SubDesc {
    uint32_t valid;     // sub-tile exists
    uint32_t sid;       // flat sub-tile id (= flat slot / 32)
    uint32_t r0, nrec;  // first record, records in the sub-tile (0 for padding sub-tiles)
    uint32_t lo;        // first record of the request (the LCP of that record is never used)
    uint64_t read_rev;
    // per lane (stage 1)
    uint32_t ko, kl, vl;  // koff16[r], klen[r], vlen[r]
    uint64_t vo;          // voff16[r]
    uint32_t pko, pkl;    // lane 0 only: koff16[r0-1], klen[r0-1] when r0 > lo
    // stage 2
    uint32_t base16, span;  // staged chunk interval [base16, base16+span)
    uint4 v0;               // first 16 bytes of the value when it has to be inspected
};

// pipeline stage 1: the record directory of the sub-tile (all loads independent of each other)
__device__ __forceinline__ SubDesc load_desc(const StoreDev &st, const TileRef &tr, uint32_t lane)
{
    SubDesc d;
    d.valid = tr.valid;
    d.sid = tr.sid;
    d.r0 = d.nrec = d.lo = 0;
    d.read_rev = 0;
    d.ko = d.kl = d.vl = d.pko = d.pkl = 0;
    d.vo = 0;
    d.base16 = d.span = 0;
    d.v0 = make_uint4(0, 0, 0, 0);
    if (!d.valid) return d;
    const uint32_t sub = tr.sid & 31;
    if (sub * 32 >= tr.t.n) return d;  // padding sub-tile of the request's last tile
    d.r0 = tr.t.rec0 + sub * 32;
    d.nrec = min(32u, tr.t.n - sub * 32);
    d.lo = tr.t.lo;
    d.read_rev = tr.t.read_rev;
    // one 16-byte packed directory entry per record (a single long-latency load per lane, see StoreDev::dir)
    if (lane < d.nrec) {
        const uint4 e = __ldg(st.dir + d.r0 + lane);
        d.ko = e.x;
        d.kl = e.y & 0xffffu;
        d.vl = e.z;
        d.vo = ((uint64_t)(e.y >> 16) << 32) | e.w;
    }
    if (lane == 0 && d.r0 > d.lo) {
        const uint4 e = __ldg(st.dir + d.r0 - 1);
        d.pko = e.x;
        d.pkl = e.y & 0xffffu;
    }
    return d;
}

// pipeline stage 2: start the asynchronous copy of the sub-tile's key bytes (plus the record before it) into `buf`
// and the value probe of the 9-byte values (tombstone literal / deleted-flag revision record)
__device__ __forceinline__ void issue_stage(const StoreDev &st, const ScanMode &mode, SubDesc &d, uint4 *buf,
                                            uint64_t *bar, uint32_t lane)
{
    if (!d.valid || d.nrec == 0) return;
    const bool halo = d.r0 > d.lo;
    d.base16 = __shfl_sync(0xffffffffu, halo ? d.pko : d.ko, 0);
    const uint32_t end16 = __shfl_sync(0xffffffffu, d.ko + ((d.kl + 15) >> 4), d.nrec - 1);
    d.span = end16 - d.base16;
    if (d.span <= KB_WARP_STAGE_CHUNKS) {
        // one bulk (TMA) copy for the whole sub-tile: [base16, base16+span) chunks -> buf.  The buffer was last read
        // (generic proxy) two steps ago and a __syncwarp separates those reads from this point; the proxy fence orders
        // them before the async-proxy write.
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsmem_u32(bar)), "r"(d.span * 16)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dsmem_u32(buf)),
                         "l"(st.kslab + d.base16), "r"(d.span * 16), "r"(dsmem_u32(bar))
                         : "memory");
        }
    }
    // only 9-byte values are ever inspected by the range path; the TTL sweep also reads revision-record values
    if (lane < d.nrec && d.vl >= 8 && (d.vl == 9 || mode.ttl_scan)) d.v0 = st.vslab[d.vo];
}

template <bool STAGED>
__device__ __forceinline__ uint32_t decode_record(const ScanMode &mode, const SubDesc &d, const uint4 *kp,
                                                  const uint4 *pp, uint32_t len, uint32_t plen, bool has_prev)
{
    const uint8_t *kb = (const uint8_t *)kp;
    uint32_t lcp = KB_LCP_INF;
    if (has_prev) {
        const uint32_t m = min(len, plen);
        const uint32_t nch = (m + 15) >> 4;
        lcp = m;
        // four chunks per step; one OR-reduced difference word per chunk, one branch per step.  In the staged
        // path the loads may run up to three chunks past the shorter key: shared memory is always readable and a
        // difference found at or beyond m is clamped to m below.
        for (uint32_t c0 = 0; c0 < nch; c0 += 4) {
            uint4 x[4], y[4];
            uint32_t dw[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t c = STAGED ? c0 + j : min(c0 + j, nch - 1);
                x[j] = kp[c];
                y[j] = pp[c];
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                dw[j] = (x[j].x ^ y[j].x) | (x[j].y ^ y[j].y) | (x[j].z ^ y[j].z) | (x[j].w ^ y[j].w);
            if (dw[0] | dw[1] | dw[2] | dw[3]) {
                const int j = dw[0] ? 0 : dw[1] ? 1 : dw[2] ? 2 : 3;
                const uint4 xa = j == 0 ? x[0] : j == 1 ? x[1] : j == 2 ? x[2] : x[3];
                const uint4 ya = j == 0 ? y[0] : j == 1 ? y[1] : j == 2 ? y[2] : y[3];
                lcp = min(m, (c0 + j) * 16 + (uint32_t)first_diff16(xa, ya));
                break;
            }
        }
    }
    uint32_t flags = 0;
    // coder.Decode (normal.go:58-70); keys shorter than 13 bytes are undecodable (Go would panic)
    bool dec_ok = len >= 13;
    if (dec_ok) dec_ok = (((const uint32_t *)kp)[0] == MAGIC_LE) && (kb[len - 9] == 0x24);
    if (dec_ok) {
        uint64_t rev;
        if (STAGED) {
            // the 8 revision bytes sit at an arbitrary offset: two aligned 64-bit loads + funnel shift + byte swap
            const uint32_t off = len - 8, sh = (off & 7) * 8;
            const uint64_t *w = (const uint64_t *)(kb + (off & ~7u));
            const uint64_t lo = w[0], hi = w[1];
            const uint64_t le = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
            rev = ((uint64_t)bswap32((uint32_t)le) << 32) | bswap32((uint32_t)(le >> 32));
        } else {
            rev = be64_bytes(kb + len - 8);
        }
        flags |= KB_M_DEC_OK;
        if (rev == 0) flags |= KB_M_REV0;
        const uint32_t vl = d.vl;
        const uint4 v0 = d.v0;
        const uint64_t vrev = ((uint64_t)bswap32(v0.x) << 32) | bswap32(v0.y);
        bool expired = false;
        if (mode.ttl_scan && contains_events(kb + 4, len - 13)) {  // compactIfExpired scanner.go:566-591
            if (rev == 0) {
                if (vl >= 8 && vrev <= mode.timeout_rev) {
                    expired = true;
                    flags |= KB_M_TTLREV;
                }
            } else if (rev <= mode.timeout_rev) {
                expired = true;
                flags |= KB_M_TTLOBJ;
            }
        }
        if (!expired && rev <= d.read_rev) {  // scanner.go:451-453
            flags |= KB_M_TRIG;
            if (vl == 9 && v0.x == 0x626d6f74u && v0.y == 0x6e6f7473u && (v0.z & 0xffu) == 0x65u)
                flags |= KB_M_TOMB;  // "tombstone" util.go:28
            bool prevok = true;
            if (mode.compact && rev == 0 && vl == 9) {  // scanner.go:476-491
                if (vrev > d.read_rev)
                    prevok = false;  // `continue` without updating prev (Q5)
                else
                    flags |= KB_M_REVDEL;
            }
            if (prevok) flags |= KB_M_PREVOK;
        }
    }
    return lcp | flags;
}

// pipeline stage 3
__device__ __forceinline__ void process_sub(const StoreDev &st, const ScanMode &mode, const SubDesc &d,
                                            const uint4 *buf, uint32_t lane, uint32_t *__restrict__ meta,
                                            uint2 *__restrict__ sub_agg)
{
    const unsigned FULLM = 0xffffffffu;
    if (d.nrec == 0) {
        if (lane == 0) sub_agg[d.sid] = make_uint2(KB_NONE, KB_LCP_INF);
        return;
    }
    const bool valid = lane < d.nrec;
    const bool staged = d.span <= KB_WARP_STAGE_CHUNKS;
    // previous record's offset / length: lane-1, or the halo record for lane 0
    uint32_t pko = __shfl_up_sync(FULLM, d.ko, 1), pkl = __shfl_up_sync(FULLM, d.kl, 1);
    if (lane == 0) {
        pko = d.pko;
        pkl = d.pkl;
    }
    uint32_t word = KB_LCP_INF;
    if (valid) {
        const uint32_t r = d.r0 + lane;
        const bool has_prev = r > d.lo;
        if (staged) {
            word = decode_record<true>(mode, d, buf + (d.ko - d.base16), buf + (has_prev ? pko - d.base16 : 0), d.kl, pkl,
                                       has_prev);
        } else {
            word = decode_record<false>(mode, d, st.kslab + d.ko, st.kslab + (has_prev ? pko : d.ko), d.kl, pkl,
                                        has_prev);
        }
        meta[d.sid * 32 + lane] = word;
    }
    // sub-tile aggregate: (last PREVOK slot, min LCP of the records after it)
    const unsigned pm = __ballot_sync(FULLM, valid && (word & KB_M_PREVOK));
    uint32_t mval = valid ? (word & KB_M_LCP_MASK) : KB_LCP_INF;
    uint32_t L = KB_NONE;
    if (pm) {
        const uint32_t top = 31 - __clz(pm);
        if (lane <= top) mval = KB_LCP_INF;
        L = d.sid * 32 + top;
    }
    mval = __reduce_min_sync(FULLM, mval);
    if (lane == 0) sub_agg[d.sid] = make_uint2(L, mval);
}

__global__ void __launch_bounds__(DECODE_WARPS * 32, 1)
k_decode_lcp(StoreDev st, const ReqDev *__restrict__ reqs, const TileDev *__restrict__ tiles, uint32_t n_sub,
             ScanMode mode, uint32_t *__restrict__ meta, uint2 *__restrict__ sub_agg, unsigned int *__restrict__ work_ctr)
{
    extern __shared__ uint4 stage[];  // DECODE_WARPS x DECODE_STAGES x KB_WARP_STAGE_CHUNKS (+ 4 chunks of slack)
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *buf0 = stage + (size_t)warp * DECODE_STAGES * KB_WARP_STAGE_CHUNKS;
    // The first three sub-tiles of a warp are assigned statically (they fill the pipeline); the rest are handed out
    // through a global counter (left at zero by k_emit, which follows every decode pass), so a CTA that starts late --
    // its SM was still busy with another stream's kernel -- simply takes fewer.  The atomic is issued one step before
    // its result is used: lane 0 keeps the raw value and the warp picks it up with a shuffle at the next step.
    const uint32_t stride = gridDim.x * DECODE_WARPS;
    const uint32_t dyn_base = 3 * stride;
    uint32_t sid = blockIdx.x * DECODE_WARPS + warp;
    uint32_t raw = 0;

    // one mbarrier per stage buffer; the phase of buffer b flips each time it is filled
    __shared__ uint64_t bars[DECODE_WARPS * DECODE_STAGES];
    uint64_t *bar0 = bars + warp * DECODE_STAGES;
    if (lane == 0) {
        dmbar_init(bar0);
        dmbar_init(bar0 + 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // prologue: fill the pipeline
    TileRef tA = fetch_tile(tiles, sid, n_sub);
    sid += stride;
    SubDesc dB = load_desc(st, tA, lane);
    tA = fetch_tile(tiles, sid, n_sub);
    sid += stride;
    SubDesc dA = load_desc(st, tA, lane);
    tA = fetch_tile(tiles, sid, n_sub);
    if (lane == 0) raw = atom_add_u32_raw(work_ctr, 1u);
    issue_stage(st, mode, dB, buf0, bar0, lane);
    uint32_t fills0 = 0, fills1 = 0;  // completed-phase counters of the two buffers (parity = count & 1)
    // The body is unrolled six times (lcm of the 3 descriptor roles and the 2 stage buffers) so that the role
    // rotation dC <- dB <- dA is pure register renaming inside the body; moves remain only on the back edge.
    bool more = dB.valid != 0;
    while (more) {
#pragma unroll
        for (int u = 0; u < 6; u++) {
            if (more) {
                const SubDesc dC = dB;  // key bytes + value probe in flight since the previous step
                dB = dA;                // directory loaded one step ago
                dA = load_desc(st, tA, lane);
                sid = dyn_base + __shfl_sync(0xffffffffu, raw, 0);
                tA = fetch_tile(tiles, sid, n_sub);
                if (lane == 0) raw = atom_add_u32_raw(work_ctr, 1u);
                issue_stage(st, mode, dB, buf0 + ((u + 1) & 1) * KB_WARP_STAGE_CHUNKS, bar0 + ((u + 1) & 1), lane);
                // wait for dC's bytes (only if a copy was issued for it: real, staged sub-tile)
                if (dC.nrec != 0 && dC.span <= KB_WARP_STAGE_CHUNKS) {
                    if ((u & 1) == 0) {
                        dmbar_wait(bar0, fills0 & 1);
                        fills0++;
                    } else {
                        dmbar_wait(bar0 + 1, fills1 & 1);
                        fills1++;
                    }
                }
                process_sub(st, mode, dC, buf0 + (u & 1) * KB_WARP_STAGE_CHUNKS, lane, meta, sub_agg);
                __syncwarp();
                more = dB.valid != 0;
            }
        }
    }
}

}  // namespace
