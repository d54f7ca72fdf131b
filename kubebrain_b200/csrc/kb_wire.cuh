// kb_wire.cuh -- the scan answer as etcd protobuf elements, written by the device (included by kb_scan.cu).
//
// Replaces the per-kv materialisation of the etcd-compatible server: kvToEtcdKv (pkg/server/etcd/backendshim.go:427-436)
// + the gogo-protobuf Marshal of etcdserverpb.RangeResponse.kvs (List, backendshim.go:269-282) or of
// etcdserverpb.WatchResponse.events (range stream, backendshim.go:349-363; batches cut by receiver.go:119-138).
// A protobuf message is the concatenation of its fields, so the device writes the repeated-field ELEMENTS back to
// back (one per emitted kv, reference order) and the host only prepends / appends the few header and trailer bytes
// (kb_wire_range_head / _tail / kb_wire_watch_head): nothing is re-materialised per kv on the CPU.
//
//   KVS    element:  12 <len(kv)>  kv                       kv = [0a <ul> user key] [18 <rev>] [2a <vl> value]
//   EVENTS element:  5a <len(ev)>  ev,  ev = 12 <len(kv)> kv      (proto3: empty key / value and rev 0 are omitted)
//
// Field numbers: go.etcd.io/etcd/api/v3 v3.5.2 (reference go.mod:28), pinned by tests/test_wire.py.
#pragma once

#include "kb_internal.cuh"

namespace {

__device__ __forceinline__ uint32_t varint_len(uint64_t v) { return v ? (uint32_t)(70 - __clzll((long long)v)) / 7u : 1u; }

struct WireSizes {
    uint32_t ul, vl;
    uint64_t rev;
    uint64_t body, kv, elem;  // mvccpb.KeyValue body, kv field (tag + len + body), whole element
};

__device__ __forceinline__ WireSizes wire_sizes(uint32_t ul, uint32_t vl, uint64_t rev, int wire)
{
    WireSizes s;
    s.ul = ul;
    s.vl = vl;
    s.rev = rev;
    s.body = (ul ? 1 + varint_len(ul) + (uint64_t)ul : 0) + (rev ? 1 + varint_len(rev) : 0) +
             (vl ? 1 + varint_len(vl) + (uint64_t)vl : 0);
    s.kv = 1 + varint_len(s.body) + s.body;
    s.elem = wire == KB_WIRE_KVS_I ? s.kv : 1 + varint_len(s.kv) + s.kv;
    return s;
}

// response bytes of one emitted record: padded [key][value] in the arena modes, element bytes in the wire modes
__device__ __forceinline__ uint64_t kv_resp_bytes(const StoreDev &st, uint32_t rec, int wire)
{
    const uint32_t kl = st.klen[rec], vl = st.vlen[rec];
    if (!wire) return (uint64_t)pad16(kl) + pad16(vl);
    const uint64_t rev = be64_bytes((const uint8_t *)(st.kslab + st.koff16[rec]) + kl - 8);
    return wire_sizes(kl - 13, vl, rev, wire).elem;
}

struct __align__(16) WireJob {
    uint64_t dst;     // first byte of the element in the arena
    uint64_t vsrc16;  // value slab chunk
    uint64_t rev;
    uint32_t ksrc16;  // key slab chunk (internal key: magic | user key | '$' | rev)
    uint32_t ul, vl;
    uint32_t pad;
};

struct WireOut {
    uint32_t *rec_idx;
    uint64_t *rev;
    uint64_t *key_off;
    uint32_t *key_len;
    uint64_t *val_off;
    uint32_t *val_len;
    uint64_t *elem_off;  // n_kvs + 1
};

__device__ __forceinline__ uint32_t put_varint(uint8_t *p, uint64_t v)
{
    uint32_t n = 0;
    while (v >= 0x80) {
        p[n++] = (uint8_t)(v | 0x80);
        v >>= 7;
    }
    p[n++] = (uint8_t)v;
    return n;
}

// the bytes in front of the user key and between the user key and the value
__device__ __forceinline__ void wire_headers(const WireSizes &s, int wire, uint8_t *h1, uint32_t &n1, uint8_t *h2,
                                             uint32_t &n2)
{
    n1 = 0;
    if (wire == KB_WIRE_EVENTS_I) {
        h1[n1++] = 0x5a;
        n1 += put_varint(h1 + n1, s.kv);
    }
    h1[n1++] = 0x12;
    n1 += put_varint(h1 + n1, s.body);
    if (s.ul) {
        h1[n1++] = 0x0a;
        n1 += put_varint(h1 + n1, s.ul);
    }
    n2 = 0;
    if (s.rev) {
        h2[n2++] = 0x18;
        n2 += put_varint(h2 + n2, s.rev);
    }
    if (s.vl) {
        h2[n2++] = 0x2a;
        n2 += put_varint(h2 + n2, s.vl);
    }
}

// thread per emitted kv: where its element goes and what the per-kv view arrays say
__global__ void __launch_bounds__(256)
k_wire_jobs(StoreDev st, const ReqDev *__restrict__ reqs, uint32_t nreq, const uint64_t *__restrict__ job_first,
            const uint64_t *__restrict__ arena_base, const uint32_t *__restrict__ sel,
            const uint64_t *__restrict__ slot, int wire, WireJob *__restrict__ jobs, WireOut out)
{
    const uint64_t n_kvs = job_first[nreq];
    if (n_kvs == 0 && blockIdx.x == 0 && threadIdx.x == 0) out.elem_off[0] = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_kvs;
         k += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = nreq;  // request of kv k: last q with job_first[q] <= k
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (job_first[mid] <= k) lo = mid; else hi = mid;
        }
        const uint32_t q = lo;
        const uint64_t s = reqs[q].sel_base + (k - job_first[q]);
        const uint32_t rec = sel[s];
        const uint64_t E = arena_base[q] + slot[s];
        const uint32_t kl = st.klen[rec], vl = st.vlen[rec];
        const uint32_t ksrc16 = st.koff16[rec];
        const uint64_t rev = be64_bytes((const uint8_t *)(st.kslab + ksrc16) + kl - 8);
        const WireSizes ws = wire_sizes(kl - 13, vl, rev, wire);
        uint8_t h1[32], h2[24];
        uint32_t n1, n2;
        wire_headers(ws, wire, h1, n1, h2, n2);
        WireJob j;
        j.dst = E;
        j.vsrc16 = st.voff16[rec];
        j.rev = rev;
        j.ksrc16 = ksrc16;
        j.ul = ws.ul;
        j.vl = vl;
        j.pad = 0;
        jobs[k] = j;
        out.rec_idx[k] = rec;
        out.rev[k] = rev;
        out.key_off[k] = E + n1;
        out.key_len[k] = ws.ul;
        out.val_off[k] = E + n1 + ws.ul + n2;
        out.val_len[k] = vl;
        out.elem_off[k] = E;
        if (k == n_kvs - 1) out.elem_off[n_kvs] = E + ws.elem;
    }
}

// bytes [lo, hi) of a 16-byte chunk as a mask (0xff per selected byte); lo/hi are clamped to [0,16]
__device__ __forceinline__ uint4 byte_mask16(int lo, int hi)
{
    lo = max(lo, 0);
    hi = min(hi, 16);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int a = max(lo - 4 * k, 0), b = min(hi - 4 * k, 4);  // byte range inside word k
        uint32_t m = 0;
        if (b > a) {
            const uint32_t upto_b = b >= 4 ? 0xffffffffu : ((1u << (8 * b)) - 1u);
            const uint32_t upto_a = a >= 4 ? 0xffffffffu : ((1u << (8 * a)) - 1u);
            m = upto_b & ~upto_a;
        }
        w[k] = m;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 16 bytes starting `s` bytes (0..15) into the 32-byte pair (lo, hi)
__device__ __forceinline__ uint4 funnel16(const uint4 &lo, const uint4 &hi, uint32_t s)
{
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t ws = s >> 2, bs = (s & 3) * 8;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t a = 0, b = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {  // ws is 0..3: select without dynamic register indexing
            if (ws == (uint32_t)t) {
                a = w[t + k];
                b = w[t + k + 1 < 8 ? t + k + 1 : 7];
            }
        }
        o[k] = __funnelshift_r(a, b, bs);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// the part of the element's chunk [e0, e0+16) that comes from a global-memory segment: segment bytes
// [begin, begin+len) of the element are bytes [skip, skip+len) behind the 16-byte aligned pointer `base`
__device__ __forceinline__ void seg_global(uint4 &acc, long long e0, long long begin, uint32_t len,
                                           const uint4 *__restrict__ base, uint32_t skip)
{
    const long long lo = begin - e0, hi = begin + (long long)len - e0;  // segment range in chunk coordinates
    if (hi <= 0 || lo >= 16 || len == 0) return;
    const long long t = (long long)skip + (e0 - begin);  // source byte of chunk byte 0 (may be negative)
    const long long ci = t >> 4;                          // floor
    const uint32_t s = (uint32_t)(t & 15);
    const long long last = ((long long)skip + len - 1) >> 4;  // last chunk holding segment bytes
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (ci >= 0 && ci <= last) a = ldg_stream(base + ci);
    if (s != 0 && ci + 1 >= 0 && ci + 1 <= last) b = ldg_stream(base + ci + 1);
    const uint4 v = funnel16(a, b, s);
    const uint4 m = byte_mask16((int)max(lo, 0ll), (int)min(hi, 16ll));
    acc.x |= v.x & m.x;
    acc.y |= v.y & m.y;
    acc.z |= v.z & m.z;
    acc.w |= v.w & m.w;
}

// ... and the part that comes from a short header held in shared memory
__device__ __forceinline__ void seg_shared(uint4 &acc, long long e0, long long begin, uint32_t len, const uint8_t *h)
{
    const long long lo = begin - e0, hi = begin + (long long)len - e0;
    if (hi <= 0 || lo >= 16 || len == 0) return;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (j >= lo && j < hi) w[j >> 2] |= (uint32_t)h[j - lo] << ((j & 3) * 8);
    }
    acc.x |= w[0];
    acc.y |= w[1];
    acc.z |= w[2];
    acc.w |= w[3];
}

constexpr int WIRE_WARPS = 8;

// warp per element: every lane produces aligned 16-byte chunks of the destination from the four segments
// [header | user key | rev + value header | value]; only the first and last chunk of an element (shared with its
// neighbours, which other warps write) fall back to byte stores.
__global__ void __launch_bounds__(WIRE_WARPS * 32)
k_wire_copy(StoreDev st, const WireJob *__restrict__ jobs, const uint64_t *__restrict__ n_kvs_dev, int wire,
            uint8_t *__restrict__ arena)
{
    __shared__ uint8_t hdr[WIRE_WARPS][64];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_kvs = *n_kvs_dev;
    uint8_t *h1 = hdr[warp], *h2 = hdr[warp] + 32;
    for (uint64_t k = (uint64_t)blockIdx.x * WIRE_WARPS + warp; k < n_kvs; k += (uint64_t)gridDim.x * WIRE_WARPS) {
        const WireJob j = jobs[k];
        const WireSizes ws = wire_sizes(j.ul, j.vl, j.rev, wire);
        uint32_t n1 = 0, n2 = 0;
        __syncwarp();  // the previous element's readers are done with the header scratch
        if (lane == 0) wire_headers(ws, wire, h1, n1, h2, n2);
        __syncwarp();
        n1 = __shfl_sync(0xffffffffu, n1, 0);
        n2 = __shfl_sync(0xffffffffu, n2, 0);
        const long long total = (long long)ws.elem;
        const uint32_t lead = (uint32_t)(j.dst & 15);
        uint4 *dst0 = (uint4 *)(arena + (j.dst - lead));
        const long long nchunks = (lead + total + 15) >> 4;
        const long long b_key = n1, b_h2 = b_key + j.ul, b_val = b_h2 + n2;
        const uint4 *ksrc = st.kslab + j.ksrc16;
        const uint4 *vsrc = st.vslab + j.vsrc16;
        for (long long c = lane; c < nchunks; c += 32) {
            const long long e0 = c * 16 - lead;
            uint4 acc = make_uint4(0, 0, 0, 0);
            if (e0 + 16 > b_val) {
                seg_global(acc, e0, b_val, j.vl, vsrc, 0);
            }
            if (e0 < b_val) {  // one of the first chunks: headers and key
                seg_shared(acc, e0, 0, n1, h1);
                seg_global(acc, e0, b_key, j.ul, ksrc, 4);
                seg_shared(acc, e0, b_h2, n2, h2);
            }
            if (e0 >= 0 && e0 + 16 <= total) {
                stg_stream(dst0 + c, acc);
            } else {
                uint8_t *d = (uint8_t *)(dst0 + c);
#pragma unroll
                for (int b = 0; b < 16; b++)
                    if (e0 + b >= 0 && e0 + b < total) d[b] = (uint8_t)byte_of(acc, b);
            }
        }
    }
}

}  // namespace
