/*
 * kb_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of KubeBrain's MVCC range-scan / compaction-sweep / watch fan-out
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libkbb200.so) never links, loads or calls it.
 *
 * Every function cites the reference file:line (relative to the kubewharf/kubebrain tree) whose
 * behaviour it restates.  The reference is Go-only and no Go toolchain exists in the build image, so
 * the reference itself cannot be compiled here (no oracle/_ref); parity is pinned instead by the
 * reference's own golden vectors G1..G7, reproduced in tests/test_oracle_golden.py.
 * For BULK inputs the reference holds no fixture: parity beyond the 10-key tables is "unpinned" and is
 * defined as equality with this restatement.
 *
 * Store model: a sorted array of unique internal keys with values -- exactly what
 * storage.Iter yields on a snapshot (pkg/storage/badger/iter.go:39-83, pkg/storage/memkv/iter.go:52-100):
 * byte-lexicographic order (bytes.Compare), half-open [start,end), reverse iff start > end.
 */
#ifndef KB_ORACLE_H
#define KB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- store ------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t  *keys;   /* packed internal keys, record i = keys[koff[i] .. koff[i+1])   */
    const uint64_t *koff;   /* n+1 byte offsets                                               */
    const uint8_t  *vals;   /* packed values                                                  */
    const uint64_t *voff;   /* n+1 byte offsets                                               */
    uint64_t        n;
} ko_store;

/* ---- coder (pkg/backend/coder/normal.go:25-70, rev.go:22-47) ------------------------------ */
#define KO_MAGIC0 0x57
#define KO_MAGIC1 0xfb
#define KO_MAGIC2 0x80
#define KO_MAGIC3 0x8b
#define KO_SPLIT  0x24 /* '$' */

enum {
    KO_OK = 0,
    KO_EDECODE_MAGIC = -1,   /* normal.go:59-61 */
    KO_EDECODE_SPLIT = -2,   /* normal.go:63-65 */
    KO_EDECODE_SHORT = -3,   /* Go would panic (index out of range); the oracle reports it */
    KO_EREVFORMAT    = -4,   /* rev.go:46 ErrInvalidRevFormat */
    KO_EINVAL        = -5,
    KO_ECOMPACTED    = -6,   /* scanner.go:618-624 range revision < compact revision */
    KO_ENOMEM        = -7
};

/* normal.go:42-50; out must hold ulen+13 bytes; returns ulen+13 */
size_t ko_encode_object_key(const uint8_t *ukey, size_t ulen, uint64_t rev, uint8_t *out);
/* normal.go:58-70; on success *uk_off = 4, *uk_len = len-13 */
int ko_decode(const uint8_t *ikey, size_t len, size_t *uk_off, size_t *uk_len, uint64_t *rev);
/* rev.go:32-47 */
int ko_parse_revision(const uint8_t *val, size_t len, uint64_t *rev, int *deleted);
/* pkg/backend/util.go:70-83; out must hold max(len,1) bytes; returns the length written */
size_t ko_prefix_end(const uint8_t *prefix, size_t len, uint8_t *out);

/* ---- iterator helpers (badger/iter.go:39-83) ---------------------------------------------- */
int      ko_bytes_compare(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen);
/* first record index whose key >= key */
uint64_t ko_lower_bound(const ko_store *s, const uint8_t *key, size_t len);

/* ---- worker.run (pkg/backend/scanner/scanner.go:389-516) ----------------------------------- */
enum { /* victim classes, in the order the reference issues the deletes */
    KO_V_SUPERSEDED = 1, /* scanner.go:465-469 older version of the same key: store.Del        */
    KO_V_TOMBSTONE  = 2, /* scanner.go:472-475 tombstone-valued version: store.Del             */
    KO_V_REVRECORD  = 3, /* scanner.go:477-491 deleted-flag revision record: store.DelCurrent  */
    KO_V_TTL_REVREC = 4, /* scanner.go:576-581 expired /events/ revision record: DelCurrent    */
    KO_V_TTL_OBJECT = 5  /* scanner.go:582-585 expired /events/ object version: Del            */
};

typedef struct {
    uint64_t  read_rev;      /* workerConfig.revision                                         */
    int64_t   limit;         /* commonResultReceiver.limit (<=0: unlimited)                   */
    int       compact;       /* workerConfig.compact                                          */
    uint64_t  timeout_rev;   /* workerConfig.timeoutRevision                                  */
    int       support_ttl;   /* store.SupportTTL()                                            */
    int       collect;       /* 1: commonResultReceiver, 0: emptyResultReceiver                */
} ko_worker_cfg;

typedef struct {
    /* emitted kvs as indices of the store record that supplied key/value/revision             */
    uint64_t *emit;   uint64_t n_emit,   cap_emit;
    /* ordered delete calls: record index + class                                              */
    uint64_t *victim; uint8_t *vclass; uint64_t n_victim, cap_victim;
    int       count;        /* worker.run's returned count (0 when the limit stopped the loop) */
    int       limit_stop;   /* 1 iff the loop ended by !needMore() (scanner.go:499-502 quirk)  */
    uint64_t  examined;     /* records pulled from the iterator                                */
    uint64_t  val_size;     /* "storage.scan_worker.size"                                      */
} ko_result;

void ko_result_init(ko_result *r);
void ko_result_free(ko_result *r);

/* One worker over the half-open internal-key interval [start,end). */
int ko_worker_run(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
                  const ko_worker_cfg *cfg, ko_result *out);

/* scanner.go:202-225 adjustPartitionsBorders: borders = P+1 internal keys (packed + offsets), sorted by
 * the caller's partition Start; rewrites interior borders in place into out (packed) / out_off. */
int ko_adjust_partition_borders(const uint8_t *borders, const uint64_t *boff, uint64_t n_borders,
                                uint8_t *out, uint64_t *out_off);

/* scanner.go:227-304 scan(): fan out over partitions, merge in partition order, total count.
 * borders as above (n_borders >= 2).  compact_rev_present/compact_rev: the stored compact_key record for
 * checkCompactRace (scanner.go:594-626).  threads: worker threads (1 = sequential). */
int ko_scan(const ko_store *s, const uint8_t *borders, const uint64_t *boff, uint64_t n_borders,
            const ko_worker_cfg *cfg, int compact_rev_present, uint64_t compact_rev, int threads,
            ko_result *out, int *total_count);

/* scanner.go:83-119 Range (limit>0 => single worker, else scan over the given partition borders) */
int ko_range(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
             uint64_t read_rev, int64_t limit, int compact_rev_present, uint64_t compact_rev,
             ko_result *out);

/* range.go:91-121 getInternalVal + :81-87 get: returns record index or -1 (not found / tombstone -> -2) */
int64_t ko_get(const ko_store *s, const uint8_t *ukey, size_t ulen, uint64_t rev, uint64_t *mod_rev);

/* compact.go:107-127 getCompactBorders. prefixes[0] = Config.Prefix, rest = SkippedPrefixes.
 * Writes 2*n internal keys, sorted, into out/out_off (out_off has 2n+1 entries). */
int ko_compact_borders(const uint8_t *prefixes, const uint64_t *poff, uint64_t n_prefixes,
                       uint8_t *out, uint64_t *out_off);

/* ---- Ring (pkg/backend/ring.go:24-118) ----------------------------------------------------- */
typedef struct ko_ring ko_ring;
ko_ring *ko_ring_new(int64_t capacity);
void     ko_ring_free(ko_ring *r);
void     ko_ring_reset(ko_ring *r);
void     ko_ring_add(ko_ring *r, uint64_t revision, uint64_t payload);
typedef struct {
    int empty, high, low;
    uint64_t newest_rev, oldest_rev;
    uint64_t n_events;   /* events copied into revs/payloads */
} ko_find_ret;
/* FindEvents; revs/payloads must hold capacity entries */
void ko_ring_find(const ko_ring *r, uint64_t revision, ko_find_ret *ret, uint64_t *revs, uint64_t *payloads);

/* ---- watch fan-out (pkg/backend/watch.go:119-159, watcherhub.go:78-92) ---------------------- */
typedef struct {
    const uint8_t  *keys;  const uint64_t *koff;   /* event user keys (Event.Kv.Key) */
    const uint64_t *rev;                            /* Event.Revision                 */
    uint64_t        n;
    const uint64_t *batch_off; uint64_t n_batches; /* batches [batch_off[b], batch_off[b+1]) */
} ko_events;

typedef struct {
    const uint8_t  *prefixes; const uint64_t *poff; /* watcher prefixes */
    const uint64_t *min_rev;                         /* processEvents' revision argument */
    uint64_t        n;
} ko_watchers;

typedef struct {
    uint64_t *start;   /* n_watchers+1 offsets into event_idx                                */
    uint32_t *event_idx;
    uint64_t  n_deliveries;
    uint64_t  n_messages; /* non-empty (watcher,batch) sends: watch.go:128-130 */
} ko_fanout;
void ko_fanout_free(ko_fanout *f);
/* every watcher runs filterByRevision+filterByPrefix over every batch; threads shard the watchers.
 * alloc_per_batch=1 reproduces watch.go:141's make([]*Event,0,len) per watcher per batch. */
int ko_fanout_run(const ko_events *ev, const ko_watchers *w, int threads, int alloc_per_batch, ko_fanout *out);

/* watch.go:37-99 registration decision.  mode: 0 live-only from min_rev=*live_rev, 1 error(empty cache),
 * 2 error(too old), 3 catch-up then live. catch-up event payload indices written to catchup (cap entries). */
typedef struct {
    int      mode;
    uint64_t live_rev;      /* revision handed to processEvents */
    uint64_t n_catchup;
    uint64_t err_rev;       /* revision quoted in the error string */
} ko_watch_reg;
/* keys of cached events are looked up through payload -> event index in ev */
int ko_watch_register(const ko_ring *ring, const ko_events *ev, const uint8_t *prefix, size_t plen,
                      uint64_t revision, uint64_t current_rev, ko_watch_reg *reg,
                      uint64_t *catchup, uint64_t cap);
/* watch.go:102-117 catchUpEvents chunk sizes; returns number of chunks, sizes[] filled (cap entries) */
uint64_t ko_catchup_chunks(uint64_t n_events, uint64_t *sizes, uint64_t cap);

/* ---- etcd wire encoding of a scan answer ---------------------------------------
 * The etcd-compatible server turns every emitted kv into mvccpb.KeyValue{Key, Value, ModRevision}
 * (pkg/server/etcd/backendshim.go:427-436 kvToEtcdKv) inside etcdserverpb.RangeResponse.kvs (List,
 * backendshim.go:269-282) or inside mvccpb.Event{Kv} of etcdserverpb.WatchResponse.events (range stream,
 * backendshim.go:349-363).  The message schemas live in the absent dependency go.etcd.io/etcd/api/v3 v3.5.2
 * (go.mod:28); their published field numbers are restated here and pinned in tests/test_wire.py against the
 * protobuf runtime: KeyValue{key=1 bytes, create_revision=2, mod_revision=3, version=4, value=5 bytes, lease=6},
 * Event{type=1, kv=2, prev_kv=3}, ResponseHeader{cluster_id=1, member_id=2, revision=3, raft_term=4},
 * RangeResponse{header=1, kvs=2, more=3, count=4}, WatchResponse{header=1, watch_id=2, created=3, canceled=4,
 * compact_revision=5, cancel_reason=6, fragment=7, events=11}.  proto3: zero / empty fields are not emitted,
 * fields appear in field-number order. */
enum { KO_WIRE_KVS = 1 /* RangeResponse.kvs element */, KO_WIRE_EVENTS = 2 /* WatchResponse.events element */ };
/* bytes of one repeated-field element (tag + length + message) */
uint64_t ko_wire_elem_size(uint64_t uk_len, uint64_t val_len, uint64_t rev, int mode);
/* encode the records rec[0..n) of `s` (their user key, value and key revision) as consecutive elements;
 * elem_off gets n+1 byte offsets; out may be NULL to size only.  Returns the total byte count. */
uint64_t ko_wire_encode(const ko_store *s, const uint64_t *rec, uint64_t n, int mode, uint8_t *out,
                        uint64_t *elem_off);
/* message framing around the elements: each returns the byte count written to out (out may be NULL) */
uint64_t ko_wire_range_head(uint64_t header_rev, uint8_t *out);               /* RangeResponse.header           */
uint64_t ko_wire_range_tail(int more, int64_t count, uint8_t *out);           /* RangeResponse.more / .count    */
uint64_t ko_wire_watch_head(uint64_t header_rev, int canceled, const uint8_t *reason, uint64_t reason_len,
                            uint8_t *out);                                     /* WatchResponse fields 1..6      */

/* ---- CPU-baseline timing variants (bench.py only) ------------------------------------------ */
/* faithful=1 performs the per-record heap copies the badger iterator performs
 * (iter.go:85-92 KeyCopy/ValueCopy; scanner.go:441,495 call Val() twice). Returns emitted count. */
int64_t ko_bench_scan(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
                      uint64_t read_rev, int64_t limit, int faithful, int threads, uint64_t *examined,
                      uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
