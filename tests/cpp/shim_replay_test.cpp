// shim_replay_test.cpp -- replays, call for call, what the Go shim under go/ does through the C ABI
// (go/pkg/storage/b200/storage.go, go/pkg/backend/scanner/b200/{kb,events}.go), with a std::map standing in for the
// durable engine, and checks every answer against the CPU oracle (oracle/libkboracle.so).  Go cannot be compiled in
// the build image, so this is the executable form of the shim's call sequences:
//   loadSnapshot        one ascending iteration of the engine -> kb_load_sorted           (storage.go loadSnapshot)
//   batch.Commit        engine commit, then kb_apply_batch with expire_unix                (storage.go batch.Commit)
//   store.Del / DelCurrent                                                                (storage.go Del, DelCurrent)
//   expireLoop          kb_expire(now)                                                     (storage.go expireLoop)
//   WatchAdd x n        ids returned in registration order (the round-1 operand-order bug) (kb.go WatchAdd)
//   EventSlab + Fanout  collector batches -> kb_watch_match -> per-batch messages          (events.go)
//   Compact             kb_expire, kb_compact_sweep, victims applied in bulk              (kb.go Compact, storage.go ApplyVictims)
//   RangeResponseWire   limit+1 scan, arena cut at elem_off[limit], More / Count           (kb.go RangeResponseWire)
//   empty inputs        LoadSorted(n = 0), Sweep with an empty bound
// usage: shim_replay_test            (needs a CUDA device)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/kb_b200.h"
#include "../../oracle/kb_oracle.h"

#define CHECK(c)                                                                                         \
    do {                                                                                                 \
        if (!(c)) {                                                                                      \
            std::printf("FAIL %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, ctx ? kb_last_error(ctx) : ""); \
            std::exit(1);                                                                                \
        }                                                                                                \
    } while (0)

typedef std::string Bytes;
static kb_ctx *ctx = nullptr;

static Bytes be64(uint64_t v)
{
    Bytes b(8, '\0');
    for (int i = 0; i < 8; i++) b[i] = (char)(v >> (8 * (7 - i)));
    return b;
}
static Bytes ikey(const Bytes &uk, uint64_t rev) { return Bytes("\x57\xfb\x80\x8b", 4) + uk + "$" + be64(rev); }

// ---- the "durable engine": a sorted map with per-key expiry, iterated the way badger is ----------------------
struct MiniEngine {
    std::map<Bytes, Bytes> kv;
    std::map<Bytes, uint64_t> expire;  // key -> unix second
    void put(const Bytes &k, const Bytes &v, uint64_t exp)
    {
        kv[k] = v;
        if (exp) expire[k] = exp; else expire.erase(k);
    }
    void del(const Bytes &k)
    {
        kv.erase(k);
        expire.erase(k);
    }
    void advance(uint64_t now)  // what a TTL engine stops returning
    {
        for (auto it = expire.begin(); it != expire.end();)
            if (it->second <= now) {
                kv.erase(it->first);
                it = expire.erase(it);
            } else
                ++it;
    }
};

struct Packed {
    std::vector<uint8_t> keys, vals;
    std::vector<uint64_t> koff{0}, voff{0};
    std::vector<Bytes> key_list;
    uint64_t n() const { return koff.size() - 1; }
    ko_store view() const { return ko_store{keys.data(), koff.data(), vals.data(), voff.data(), n()}; }
};

static Packed iterate(const MiniEngine &e)  // storage.Iter over [magic, magic+1): ascending unique keys
{
    Packed p;
    for (auto &it : e.kv) {
        p.keys.insert(p.keys.end(), it.first.begin(), it.first.end());
        p.vals.insert(p.vals.end(), it.second.begin(), it.second.end());
        p.koff.push_back(p.keys.size());
        p.voff.push_back(p.vals.size());
        p.key_list.push_back(it.first);
    }
    if (p.keys.empty()) p.keys.push_back(0);
    if (p.vals.empty()) p.vals.push_back(0);
    return p;
}

// ---- storage.go: the write path ---------------------------------------------------------------------------------
struct WriteOp {
    bool del;
    Bytes key, val;
    uint64_t expire_unix;
};
static void apply_batch(const std::vector<WriteOp> &ops)  // kb.go ApplyBatch
{
    if (ops.empty()) return;
    std::vector<kb_write_op> raw(ops.size());
    for (size_t i = 0; i < ops.size(); i++) {
        memset(&raw[i], 0, sizeof(raw[i]));
        raw[i].type = ops[i].del ? KB_OP_DEL : KB_OP_PUT;
        raw[i].key = ops[i].key.empty() ? nullptr : (const uint8_t *)ops[i].key.data();
        raw[i].key_len = ops[i].key.size();
        raw[i].val = ops[i].val.empty() ? nullptr : (const uint8_t *)ops[i].val.data();
        raw[i].val_len = ops[i].val.size();
        raw[i].expire_unix = ops[i].expire_unix;
    }
    CHECK(kb_apply_batch(ctx, raw.data(), raw.size()) == KB_OK);
}

struct Backend {  // just enough of pkg/backend/txn.go to produce the reference's record formats
    MiniEngine eng;
    uint64_t rev = 1000;
    std::map<Bytes, uint64_t> latest;  // user key -> latest revision (0: deleted / absent)
    std::vector<std::pair<uint64_t, Bytes>> events;

    // Create / Update: batch.CAS(revisionKey, ...) + batch.Put(objectKey, value, ttl); Commit; mirror
    void write(const Bytes &uk, const Bytes &val, uint64_t now, int64_t ttl)
    {
        const uint64_t r = ++rev;
        const uint64_t exp = ttl > 0 ? now + (uint64_t)ttl : 0;
        std::vector<WriteOp> ops = {{false, ikey(uk, 0), be64(r), exp}, {false, ikey(uk, r), val, exp}};
        for (auto &o : ops) eng.put(o.key, o.val, o.expire_unix);  // inner.Commit succeeded
        apply_batch(ops);                                          // ... then the mirror
        latest[uk] = r;
        events.push_back({r, uk});
    }
    // Delete: revision record gets the 9-byte deleted form, the object version a tombstone (txn.go:145-190)
    void remove(const Bytes &uk)
    {
        const uint64_t r = ++rev;
        std::vector<WriteOp> ops = {{false, ikey(uk, 0), be64(r) + Bytes(1, '\0'), 0}, {false, ikey(uk, r), "tombstone", 0}};
        for (auto &o : ops) eng.put(o.key, o.val, 0);
        apply_batch(ops);
        latest[uk] = 0;
        events.push_back({r, uk});
    }
};

// ---- comparisons ---------------------------------------------------------------------------------------------------
// b200Scanner.rangeOnce is two critical sections: kb_range_submit under the engine lock, then kb_range_collect under the
// lock again -- another goroutine's submission may sit between them (two batches in flight on the device)
struct ScanCall {
    Bytes s, t;
    uint64_t read_rev;
    int64_t limit;
    kb_pending *pend;
};

static ScanCall scan_submit(const Bytes &prefix, uint64_t read_rev, int64_t limit)
{
    ScanCall c;
    Bytes end = prefix;
    end.back()++;
    c.s = ikey(prefix, 0);
    c.t = ikey(end, 0);
    c.read_rev = read_rev;
    c.limit = limit;
    c.pend = nullptr;
    kb_range_req rq{(const uint8_t *)c.s.data(), c.s.size(), (const uint8_t *)c.t.data(), c.t.size(), read_rev, limit};
    CHECK(kb_range_submit(ctx, &rq, 1, KB_OUT_HOST, &c.pend) == KB_OK);
    CHECK(c.pend != nullptr);
    return c;  // (the bound keys are only read during the submission: the Go slices are unpinned here)
}

static void scan_collect_check(const MiniEngine &e, ScanCall &c)
{
    const Packed p = iterate(e);
    const ko_store st = p.view();
    ko_result exp;
    ko_result_init(&exp);
    CHECK(ko_range(&st, (const uint8_t *)c.s.data(), c.s.size(), (const uint8_t *)c.t.data(), c.t.size(), c.read_rev, c.limit, 0, 0,
                   &exp) == 0);
    kb_result *res = nullptr;
    CHECK(kb_range_collect(ctx, c.pend, &res) == KB_OK);
    c.pend = nullptr;
    kb_range_view v;
    CHECK(kb_range_view_get(res, &v) == KB_OK);
    CHECK(v.n_kvs == exp.n_emit);
    CHECK(v.req_examined[0] == exp.examined);
    for (uint64_t k = 0; k < v.n_kvs; k++) {
        const uint64_t i = exp.emit[k];
        const Bytes ik((const char *)p.keys.data() + p.koff[i], p.koff[i + 1] - p.koff[i]);
        const Bytes val((const char *)p.vals.data() + p.voff[i], p.voff[i + 1] - p.voff[i]);
        CHECK(Bytes((const char *)v.bytes + v.key_off[k], v.key_len[k]) == ik.substr(4, ik.size() - 13));
        CHECK(Bytes((const char *)v.bytes + v.val_off[k], v.val_len[k]) == val);
        uint64_t rv = 0;
        for (int b = 0; b < 8; b++) rv = (rv << 8) | (uint8_t)ik[ik.size() - 8 + b];
        CHECK(v.rev[k] == rv);
    }
    kb_result_free(ctx, res);
    ko_result_free(&exp);
}

static void check_scan(const MiniEngine &e, const Bytes &prefix, uint64_t read_rev, int64_t limit)
{
    ScanCall c = scan_submit(prefix, read_rev, limit);
    scan_collect_check(e, c);
}

// three goroutines in Range at once: submissions and collections interleave in lock-acquisition order
static void check_concurrent_scans(const MiniEngine &e, uint64_t read_rev)
{
    ScanCall a = scan_submit("/registry/pods/", read_rev, 0);
    ScanCall b = scan_submit("/registry/events/", read_rev, 3);
    ScanCall c = scan_submit("/registry/", read_rev, 0);  // third submission: the first one's rows are read back first
    scan_collect_check(e, b);
    ScanCall d = scan_submit("/registry/pods/", read_rev, 2);
    scan_collect_check(e, a);
    scan_collect_check(e, d);
    scan_collect_check(e, c);
}

static void check_store_equals(const MiniEngine &e)
{
    uint64_t n = 0;
    CHECK(kb_store_info(ctx, &n, nullptr, nullptr) == KB_OK);
    CHECK(n == e.kv.size());
    check_scan(e, "/registry/", ~0ull >> 1, 0);
    check_scan(e, "/registry/", 1005, 0);
    check_scan(e, "/registry/pods/", ~0ull >> 1, 3);
    check_concurrent_scans(e, ~0ull >> 1);
}

int main()
{
    CHECK(kb_abi_version() == KB_ABI_VERSION);
    if (kb_open(0, nullptr, &ctx) != KB_OK) {
        std::printf("no CUDA device: the shim has no CPU fallback\n");
        return 2;
    }
    // ---- empty engine: LoadSorted(n = 0) as kb.go sends it (nil data pointers, one zero offset)
    {
        const uint64_t zero = 0;
        CHECK(kb_load_sorted(ctx, nullptr, &zero, nullptr, &zero, 0) == KB_OK);
        MiniEngine none;
        check_scan(none, "/registry/", 100, 0);
    }
    Backend be;
    const uint64_t now = 1700000000;
    const char *res[] = {"pods", "configmaps", "events"};
    // ---- content before start-up; then loadSnapshot
    for (int i = 0; i < 60; i++) {
        char name[96];
        std::snprintf(name, sizeof(name), "/registry/%s/ns-%02d/obj-%03d", res[i % 3], i % 7, i);
        be.write(name, Bytes(40 + i, 'a' + i % 26), now, i % 3 == 2 ? 30 + i : 0);  // /events/ keys carry a ttl
    }
    for (int i = 0; i < 60; i += 4) {
        char name[96];
        std::snprintf(name, sizeof(name), "/registry/%s/ns-%02d/obj-%03d", res[i % 3], i % 7, i);
        be.write(name, Bytes(17, 'z'), now, i % 3 == 2 ? 100 : 0);
    }
    {
        const Packed p = iterate(be.eng);
        CHECK(kb_load_sorted(ctx, p.keys.data(), p.koff.data(), p.vals.data(), p.voff.data(), p.n()) == KB_OK);
        // the TTLs of the snapshot travel as a batch of re-puts (storage.go: the adaptor re-registers expiry after a load)
        std::vector<WriteOp> ops;
        for (auto &x : be.eng.expire) ops.push_back({false, x.first, be.eng.kv[x.first], x.second});
        apply_batch(ops);
    }
    check_store_equals(be.eng);
    // ---- live writes through batch.Commit, deletes, store.Del / DelCurrent
    for (int i = 0; i < 40; i++) {
        char name[96];
        std::snprintf(name, sizeof(name), "/registry/%s/ns-%02d/obj-%03d", res[i % 3], i % 5, 100 + i);
        be.write(name, Bytes(64, 'A' + i % 26), now + 1, i % 3 == 2 ? 20 : 0);
        if (i % 6 == 0) be.remove(name);
    }
    check_store_equals(be.eng);
    {
        const Bytes victim = be.eng.kv.begin()->first;  // store.Del(ctx, key)
        be.eng.del(victim);
        apply_batch({{true, victim, "", 0}});
        apply_batch({{true, victim, "", 0}});  // deleting an absent key is a no-op
        check_store_equals(be.eng);
    }
    // ---- expireLoop: two ticks
    for (uint64_t t : {now + 25, now + 70}) {
        uint64_t dropped = 0;
        const size_t before = be.eng.kv.size();
        be.eng.advance(t);
        CHECK(kb_expire(ctx, t, &dropped) == KB_OK);
        CHECK(dropped == before - be.eng.kv.size());
        check_store_equals(be.eng);
    }
    // ---- watchers: ids come back in registration order (kb.go WatchAdd reads `id` after the call)
    const char *prefixes[] = {"/registry/pods/", "/registry/pods/ns-01/", "/registry/", "/registry/events/ns-02/", "/nothing/"};
    std::vector<uint64_t> minrev = {0, 1030, 1060, 0, 0};
    for (uint32_t w = 0; w < 5; w++) {
        uint32_t id = 0xdead;
        CHECK(kb_watch_add(ctx, (const uint8_t *)prefixes[w], strlen(prefixes[w]), minrev[w], &id) == KB_OK);
        CHECK(id == w);
    }
    // ---- EventSlab.AppendBatch over collector batches of <= 7 events, then Fanout
    {
        std::vector<uint8_t> keys;
        std::vector<uint64_t> koff{0}, rev, boff{0};
        for (size_t i = 0; i < be.events.size(); i++) {
            keys.insert(keys.end(), be.events[i].second.begin(), be.events[i].second.end());
            koff.push_back(keys.size());
            rev.push_back(be.events[i].first);
            if ((i + 1) % 7 == 0 || i + 1 == be.events.size()) boff.push_back(rev.size());
        }
        kb_events ev{keys.data(), koff.data(), rev.data(), rev.size(), boff.data(), boff.size() - 1};
        kb_result *res2 = nullptr;
        CHECK(kb_watch_match(ctx, &ev, KB_OUT_HOST, &res2) == KB_OK);
        kb_match_view mv;
        CHECK(kb_match_view_get(res2, &mv) == KB_OK);
        std::vector<uint8_t> pre;
        std::vector<uint64_t> poff{0};
        for (auto p : prefixes) {
            pre.insert(pre.end(), p, p + strlen(p));
            poff.push_back(pre.size());
        }
        ko_events oe{keys.data(), koff.data(), rev.data(), rev.size(), boff.data(), boff.size() - 1};
        ko_watchers ow{pre.data(), poff.data(), minrev.data(), 5};
        ko_fanout of;
        CHECK(ko_fanout_run(&oe, &ow, 2, 1, &of) == 0);
        CHECK(mv.n_watchers == 5 && mv.n_deliveries == of.n_deliveries);
        uint64_t messages = 0;
        for (uint32_t w = 0; w < 5; w++) {
            CHECK(mv.start[w] == of.start[w] && mv.start[w + 1] == of.start[w + 1]);
            size_t b = 0;
            long cur = -1;
            for (uint64_t d = mv.start[w]; d < mv.start[w + 1]; d++) {  // events.go Fanout: one message per batch touched
                CHECK(mv.event_idx[d] == of.event_idx[d]);
                while (mv.event_idx[d] >= boff[b + 1]) b++;
                if ((long)b != cur) {
                    messages++;
                    cur = (long)b;
                }
            }
        }
        CHECK(messages == of.n_messages);
        ko_fanout_free(&of);
        kb_result_free(ctx, res2);
    }
    // ---- RangeResponseWire: user limit 5 -> ask 6, cut at elem_off[5], More, Count = 5 + 1
    {
        const Packed p = iterate(be.eng);
        const ko_store st = p.view();
        const Bytes s = ikey("/registry/pods/", 0), t = ikey("/registry/pods0", 0);
        const int64_t limit = 5;
        kb_range_req rq{(const uint8_t *)s.data(), s.size(), (const uint8_t *)t.data(), t.size(), ~0ull >> 1, limit + 1};
        kb_result *r = nullptr;
        CHECK(kb_range_batch(ctx, &rq, 1, KB_OUT_HOST | KB_WIRE_ETCD_KVS, &r) == KB_OK);
        kb_range_view v;
        CHECK(kb_range_view_get(r, &v) == KB_OK);
        CHECK((int64_t)v.n_kvs == limit + 1 && v.elem_off);
        uint8_t head[32], tail[32];
        Bytes got((const char *)head, kb_wire_range_head(77, head));
        got += Bytes((const char *)v.bytes, v.elem_off[limit]);
        got += Bytes((const char *)tail, kb_wire_range_tail(1, limit + 1, tail));
        ko_result exp;
        ko_result_init(&exp);
        CHECK(ko_range(&st, (const uint8_t *)s.data(), s.size(), (const uint8_t *)t.data(), t.size(), ~0ull >> 1, limit + 1, 0, 0, &exp) == 0);
        std::vector<uint8_t> enc(1 << 20);
        std::vector<uint64_t> eoff(limit + 2);
        const uint64_t nb = ko_wire_encode(&st, exp.emit, limit, KO_WIRE_KVS, enc.data(), eoff.data());
        uint8_t oh[32], ot[32];
        Bytes want((const char *)oh, ko_wire_range_head(77, oh));
        want += Bytes((const char *)enc.data(), nb);
        want += Bytes((const char *)ot, ko_wire_range_tail(1, limit + 1, ot));
        CHECK(got == want);
        ko_result_free(&exp);
        kb_result_free(ctx, r);
    }
    // ---- Compact: expire, sweep, victims applied in bulk (engine batch, then the mirror)
    {
        be.eng.advance(now + 200);
        CHECK(kb_expire(ctx, now + 200, nullptr) == KB_OK);
        const Packed p = iterate(be.eng);
        const ko_store st = p.view();
        const Bytes s = ikey("/registry/", 0), t = ikey("/registry0", 0);
        const uint64_t crev = be.rev - 10;
        kb_result *r = nullptr;
        CHECK(kb_compact_sweep(ctx, (const uint8_t *)s.data(), s.size(), (const uint8_t *)t.data(), t.size(), crev, 0, 1,
                               KB_OUT_HOST, &r) == KB_OK);
        kb_compact_view cv;
        CHECK(kb_compact_view_get(r, &cv) == KB_OK);
        std::vector<uint8_t> borders(s.begin(), s.end());
        borders.insert(borders.end(), t.begin(), t.end());
        const uint64_t boff[3] = {0, s.size(), s.size() + t.size()};
        ko_worker_cfg cfg{crev, 0, 1, 0, 1, 0};
        ko_result exp;
        ko_result_init(&exp);
        int total = 0;
        CHECK(ko_scan(&st, borders.data(), boff, 2, &cfg, 0, 0, 1, &exp, &total) == 0);
        CHECK(cv.n_victims == exp.n_victim && (int)cv.count == total);
        std::vector<WriteOp> dels;
        for (uint64_t i = 0; i < cv.n_victims; i++) {
            CHECK(cv.victim_idx[i] == exp.victim[i] && cv.victim_class[i] == exp.vclass[i]);
            dels.push_back({true, p.key_list[cv.victim_idx[i]], "", 0});
        }
        CHECK(cv.n_victims > 10);
        kb_result_free(ctx, r);
        ko_result_free(&exp);
        for (auto &d : dels) be.eng.del(d.key);  // ApplyVictims: one engine batch per chunk ...
        for (size_t i = 0; i < dels.size(); i += 16)  // ... then the same keys leave the mirror
            apply_batch(std::vector<WriteOp>(dels.begin() + i, dels.begin() + std::min(dels.size(), i + 16)));
        CHECK(kb_set_compact_revision(ctx, 0, 0) == KB_OK);
        check_store_equals(be.eng);
        // Sweep with an empty start bound (ptr8 of an empty slice is nil): everything below `end`
        kb_result *r2 = nullptr;
        CHECK(kb_compact_sweep(ctx, nullptr, 0, (const uint8_t *)t.data(), t.size(), crev, 0, 1, KB_OUT_COUNT, &r2) == KB_OK);
        kb_result_free(ctx, r2);
    }
    kb_close(ctx);
    ctx = nullptr;
    std::printf("shim replay ok\n");
    return 0;
}
