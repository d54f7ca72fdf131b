// host_mirror_test.cpp -- the reference's own table tests, written against the C++ host mirror
// (kubebrain_b200/host/kubebrain.hpp).   usage: host_mirror_test cpu | gpu
//   cpu: coder known answer (coder/normal_test.go:23-32), PrefixEnd, compaction borders (compact_test.go:36-81),
//        Ring.FindEvents table (ring_test.go:61-107)
//   gpu: testBackendRange (backend_test.go:740-901) through Backend.List / Count / ListByStream on the B200
#include <cstdio>
#include <cstdlib>
#include <map>

#include "../../kubebrain_b200/host/kubebrain.hpp"

#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);    \
            std::exit(1);                                               \
        }                                                               \
    } while (0)

using kb::Bytes;

struct Ev {
    uint64_t Revision = 0;
};

static void cpu_tests()
{
    kb::NormalCoder c;
    const unsigned char bs[] = {87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 36,
                                0,  0,   0,   0,   0,  0,   0,   0};
    Bytes k((const char *)bs, sizeof(bs)), uk;
    uint64_t rev = 99;
    CHECK(c.Decode(k, &uk, &rev) && uk == "/registry/test" && rev == 0);
    CHECK(c.EncodeObjectKey("/registry/test", 0) == k);
    CHECK(!c.Decode(Bytes("\x00\x00\x00\x00abc$", 8) + Bytes(8, '\0'), &uk, &rev));
    CHECK(!c.Decode(Bytes("\x57\xfb\x80\x8b$", 5), &uk, &rev));
    bool tomb;
    CHECK(kb::ParseRevision(Bytes("\0\0\0\0\0\0\0\x05", 8), &rev, &tomb) && rev == 5 && !tomb);
    CHECK(kb::ParseRevision(Bytes("\0\0\0\0\0\0\0\x05\0", 9), &rev, &tomb) && rev == 5 && tomb);
    CHECK(!kb::ParseRevision("abc", &rev, &tomb));
    CHECK(kb::PrefixEnd("/registry/test/") == "/registry/test0");
    CHECK(kb::PrefixEnd(Bytes("a\xff\xff")) == "b");
    CHECK(kb::PrefixEnd(Bytes("\xff\xff")) == Bytes(1, '\0'));

    kb::Ring<Ev> r(10);
    CHECK(r.FindEvents(5).empty);
    for (uint64_t i = 1; i <= 20; i++) r.Add(Ev{i});
    struct Row {
        uint64_t rev;
        bool high, low;
        uint64_t first;
        size_t n;
    } table[] = {{9, false, true, 0, 0},   {10, false, true, 0, 0},  {11, false, false, 11, 10}, {12, false, false, 12, 9},
                 {15, false, false, 15, 6}, {19, false, false, 19, 2}, {20, false, false, 20, 1},  {21, true, false, 0, 0},
                 {30, true, false, 0, 0}};
    for (auto &t : table) {
        auto ret = r.FindEvents(t.rev);
        CHECK(ret.high == t.high && ret.low == t.low && ret.oldest.Revision == 11 && ret.newest.Revision == 20);
        CHECK(ret.events.size() == t.n);
        for (size_t i = 0; i < t.n; i++) CHECK(ret.events[i].Revision == t.first + i);
    }
    std::printf("cpu ok\n");
}

static void gpu_tests()
{
    kb::Engine eng(0);
    kb::NormalCoder c;
    // compaction borders need no device but live on Backend
    kb::Backend b(eng, "/registry/test", {"/registry/test/pods", "/registry/test/events"});
    auto borders = b.GetCompactBorders();
    const char *exp[] = {"/registry/test/", "/registry/test/events/", "/registry/test/events0",
                         "/registry/test/pods/", "/registry/test/pods0", "/registry/test0"};
    CHECK(borders.size() == 6);
    for (int i = 0; i < 6; i++) CHECK(borders[i] == c.EncodeRevisionKey(exp[i]));

    // testBackendRange: 10 sequential creates (creator/naive.go:53-105 record format), then the 14 read cases
    const int inject = 10;
    uint64_t rev = 1700000000ull;
    std::map<Bytes, Bytes> kv;
    std::vector<kb::KeyValue> kvList;
    auto fmt = [](const Bytes &p, int i) {
        char buf[16];
        std::snprintf(buf, sizeof(buf), "/%05d", i);
        return p + buf;
    };
    const Bytes testKey = "/registry/test/key", endKey = kb::PrefixEnd(testKey);
    for (int i = 0; i < inject; i++) {
        rev++;
        Bytes rb;
        for (int s = 7; s >= 0; s--) rb.push_back((char)((rev >> (8 * s)) & 0xff));
        kv[c.EncodeRevisionKey(fmt(testKey, i))] = rb;
        kv[c.EncodeObjectKey(fmt(testKey, i), rev)] = fmt("val", i);
        kvList.push_back(kb::KeyValue{fmt(testKey, i), fmt("val", i), rev});
    }
    const uint64_t init = rev;
    eng.LoadSorted(std::vector<std::pair<Bytes, Bytes>>(kv.begin(), kv.end()));
    kb::Backend be(eng, "/registry/test");
    be.SetCurrentRevision(init);
    auto sub = [&](int a, int bnd) { return std::vector<kb::KeyValue>(kvList.begin() + a, kvList.begin() + bnd); };
    auto r = be.List(testKey, endKey);
    CHECK(r.Revision == init && r.Kvs == kvList && !r.More);
    r = be.List(testKey, fmt(testKey, inject - 2));
    CHECK(r.Kvs == sub(0, inject - 2) && !r.More);
    r = be.List(testKey, fmt(testKey, inject - 2), 0, inject - 4);
    CHECK(r.Kvs == sub(0, inject - 4) && r.More);
    r = be.List(endKey, fmt(endKey, inject - 2));
    CHECK(r.Kvs.empty() && !r.More);
    bool threw = false;
    try {
        (void)be.List(fmt(endKey, inject - 2), endKey);
    } catch (const kb::Error &e) {
        threw = std::string(e.what()) == "invalid range end";
    }
    CHECK(threw);
    r = be.List(fmt(testKey, 1), fmt(testKey, inject - 1), init - 2, inject - 5);
    CHECK(r.Kvs == sub(1, inject - 4) && r.More);
    r = be.List(testKey, kb::PrefixEnd(testKey), 0, inject - 5);
    CHECK(r.Kvs == sub(0, inject - 5) && r.More);
    // get cases (backend_test.go:800-823)
    kb::KeyValue gkv;
    bool found = false;
    CHECK(be.Get(fmt(testKey, inject - 1), 0, &gkv, &found) == init && found && gkv == kvList[inject - 1]);
    CHECK(be.Get(fmt(testKey, inject - 2), init, &gkv, &found) == init && found && gkv.Revision == init - 1);
    be.Get(fmt(testKey, inject - 1), 1700000000ull, &gkv, &found);
    CHECK(!found);
    be.Get(testKey + "/-0001", 0, &gkv, &found);
    CHECK(!found);
    CHECK(be.Count(testKey, endKey) == (uint64_t)inject);
    CHECK(be.Count(endKey, kb::PrefixEnd(endKey)) == 0);
    std::vector<kb::KeyValue> got;
    for (auto &m : be.ListByStream(c.EncodeObjectKey(testKey, 0), c.EncodeObjectKey(endKey, 0))) {
        CHECK(m.Err.empty());
        got.insert(got.end(), m.Kvs.begin(), m.Kvs.end());
    }
    CHECK(got == kvList);
    // compaction at the current revision: nothing to delete (every object has exactly one live version)
    auto victims = be.Compact(init);
    CHECK(victims.size() == 1 && victims[0].empty());
    // incremental maintenance: update key 3 and delete key 4 through the BatchWrite hook, no reload
    auto be64 = [](uint64_t v) {
        Bytes rb;
        for (int s = 7; s >= 0; s--) rb.push_back((char)((v >> (8 * s)) & 0xff));
        return rb;
    };
    std::vector<kb::Engine::WriteOp> ops;
    ops.push_back({false, c.EncodeRevisionKey(fmt(testKey, 3)), be64(init + 1)});
    ops.push_back({false, c.EncodeObjectKey(fmt(testKey, 3), init + 1), "new"});
    ops.push_back({false, c.EncodeRevisionKey(fmt(testKey, 4)), be64(init + 2) + Bytes(1, '\0')});
    ops.push_back({false, c.EncodeObjectKey(fmt(testKey, 4), init + 2), "tombstone"});
    ops.push_back({true, c.EncodeObjectKey(fmt(testKey, 9), init), ""});  // drop the only version of key 9
    eng.ApplyBatch(ops);
    be.SetCurrentRevision(init + 2);
    r = be.List(testKey, endKey);
    CHECK(r.Kvs.size() == (size_t)inject - 2 && r.Kvs[3].Value == "new" && r.Kvs[3].Revision == init + 1 &&
          r.Kvs[4].Key == fmt(testKey, 5));
    r = be.List(testKey, endKey, init);  // time travel still sees the old versions of 3 and 4 (9 was removed)
    CHECK(r.Kvs.size() == (size_t)inject - 1 && r.Kvs[3].Value == fmt("val", 3) && r.Kvs[4].Value == fmt("val", 4));
    be.Get(fmt(testKey, 4), 0, &gkv, &found);
    CHECK(!found);
    // etcd wire path: one kv, bytes checked by hand against the protobuf encoding
    //   RangeResponse{header{revision: 7}, kvs:[{key:"/registry/test/key/00000", mod_revision: R, value:"val/00000"}], count: 1}
    {
        kb::Scanner sc(eng);
        const uint64_t R = 1700000001ull;
        Bytes wire = sc.RangeResponseWire(c.EncodeObjectKey(fmt(testKey, 0), 0), c.EncodeObjectKey(fmt(testKey, 1), 0),
                                          init + 2, 0, 7, false);
        Bytes kvb;
        kvb += "\x0a\x18" + fmt(testKey, 0);  // key = 1, 24 bytes
        kvb += "\x18";                          // mod_revision = 3
        for (uint64_t v = R; ; v >>= 7) {
            if (v >= 0x80) kvb.push_back((char)(v | 0x80)); else { kvb.push_back((char)v); break; }
        }
        kvb += "\x2a\x09" + fmt("val", 0);      // value = 5, 9 bytes
        Bytes exp = Bytes("\x0a\x02\x18\x07", 4) + "\x12" + Bytes(1, (char)kvb.size()) + kvb + Bytes("\x20\x01", 2);
        CHECK(wire == exp);
        auto msgs = sc.RangeStreamWire(c.EncodeObjectKey(testKey, 0), c.EncodeObjectKey(endKey, 0), init + 2);
        CHECK(msgs.size() == 2);  // one batch (8 live keys) + the cancel message
        CHECK(msgs[0].substr(0, 3) == Bytes("\x0a\x00\x5a", 3));
        CHECK(msgs[1][0] == 0x0a && msgs[1].substr(msgs[1].size() - 2) == Bytes("\x20\x01", 2));
    }
    std::printf("gpu ok\n");
}

int main(int argc, char **argv)
{
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    try {
        if (mode == "cpu") cpu_tests();
        if (mode == "gpu") gpu_tests();
    } catch (const kb::Error &e) {
        std::printf("FAIL kb::Error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
