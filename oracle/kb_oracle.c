/*
 * kb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See kb_oracle.h.
 *
 * Restates, function by function, the reference loops of kubewharf/kubebrain's range-scan,
 * compaction-sweep and watch fan-out path.  Citations are file:line in the reference tree.
 * Parity status: pinned by the reference's golden vectors G1..G7 (tests/test_oracle_golden.py);
 * unpinned for bulk inputs (the reference has no bulk fixture and cannot be built here: Go-only,
 * no Go toolchain, un-vendored modules).
 */
#include "kb_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static const uint8_t KO_TOMBSTONE[9] = {'t', 'o', 'm', 'b', 's', 't', 'o', 'n', 'e'}; /* util.go:28 */
static const uint8_t KO_EVENTS[8] = {'/', 'e', 'v', 'e', 'n', 't', 's', '/'};         /* util.go:30 */

/* ------------------------------------------------------------------------------------------ */
/* coder                                                                                       */
/* ------------------------------------------------------------------------------------------ */

static inline uint64_t be64(const uint8_t *p)
{
    return ((uint64_t)p[0] << 56) | ((uint64_t)p[1] << 48) | ((uint64_t)p[2] << 40) | ((uint64_t)p[3] << 32) |
           ((uint64_t)p[4] << 24) | ((uint64_t)p[5] << 16) | ((uint64_t)p[6] << 8) | (uint64_t)p[7];
}

static inline void put_be64(uint8_t *p, uint64_t v)
{
    for (int i = 7; i >= 0; i--) {
        p[i] = (uint8_t)(v & 0xff);
        v >>= 8;
    }
}

/* coder/normal.go:42-50: magic | userKey | '$' | BE64(revision) */
size_t ko_encode_object_key(const uint8_t *ukey, size_t ulen, uint64_t rev, uint8_t *out)
{
    out[0] = KO_MAGIC0;
    out[1] = KO_MAGIC1;
    out[2] = KO_MAGIC2;
    out[3] = KO_MAGIC3;
    if (ulen) memcpy(out + 4, ukey, ulen);
    out[4 + ulen] = KO_SPLIT;
    put_be64(out + 4 + ulen + 1, rev);
    return ulen + 13;
}

/* coder/normal.go:58-70.  Go indexes internalKey[:4] and internalKey[len-9] and slices [4:len-9]
 * without a length check, so a key shorter than 13 bytes with a correct magic panics; the oracle
 * reports KO_EDECODE_SHORT for every key shorter than 13 bytes (callers treat it as "skip"). */
int ko_decode(const uint8_t *ikey, size_t len, size_t *uk_off, size_t *uk_len, uint64_t *rev)
{
    if (len < 13) return KO_EDECODE_SHORT;
    if (ikey[0] != KO_MAGIC0 || ikey[1] != KO_MAGIC1 || ikey[2] != KO_MAGIC2 || ikey[3] != KO_MAGIC3)
        return KO_EDECODE_MAGIC;
    if (ikey[len - 9] != KO_SPLIT) return KO_EDECODE_SPLIT;
    *rev = be64(ikey + len - 8);
    *uk_off = 4;
    *uk_len = len - 13;
    return KO_OK;
}

/* coder/rev.go:32-47 */
int ko_parse_revision(const uint8_t *val, size_t len, uint64_t *rev, int *deleted)
{
    if (len == 8) {
        *rev = be64(val);
        *deleted = 0;
        return KO_OK;
    }
    if (len == 9) {
        *rev = be64(val);
        *deleted = 1;
        return KO_OK;
    }
    *rev = 0;
    *deleted = 0;
    return KO_EREVFORMAT;
}

/* backend/util.go:70-83 */
size_t ko_prefix_end(const uint8_t *prefix, size_t len, uint8_t *out)
{
    if (len) memcpy(out, prefix, len);
    for (size_t i = len; i-- > 0;) {
        if (out[i] < 0xff) {
            out[i] = (uint8_t)(out[i] + 1);
            return i + 1;
        }
    }
    out[0] = 0; /* noPrefixEnd, util.go:29 */
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* iterator over the sorted snapshot (badger/iter.go:39-83)                                     */
/* ------------------------------------------------------------------------------------------ */

int ko_bytes_compare(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen)
{
    size_t m = alen < blen ? alen : blen;
    int c = m ? memcmp(a, b, m) : 0;
    if (c) return c < 0 ? -1 : 1;
    if (alen == blen) return 0;
    return alen < blen ? -1 : 1;
}

static inline const uint8_t *rec_key(const ko_store *s, uint64_t i, size_t *len)
{
    *len = (size_t)(s->koff[i + 1] - s->koff[i]);
    return s->keys + s->koff[i];
}

static inline const uint8_t *rec_val(const ko_store *s, uint64_t i, size_t *len)
{
    *len = (size_t)(s->voff[i + 1] - s->voff[i]);
    return s->vals + s->voff[i];
}

uint64_t ko_lower_bound(const ko_store *s, const uint8_t *key, size_t len)
{
    uint64_t lo = 0, hi = s->n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        size_t kl;
        const uint8_t *k = rec_key(s, mid, &kl);
        if (ko_bytes_compare(k, kl, key, len) < 0)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* first record index whose key > key */
static uint64_t upper_bound(const ko_store *s, const uint8_t *key, size_t len)
{
    uint64_t lo = 0, hi = s->n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        size_t kl;
        const uint8_t *k = rec_key(s, mid, &kl);
        if (ko_bytes_compare(k, kl, key, len) <= 0)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

typedef struct {
    const ko_store *s;
    int reverse;       /* iter.go:44 reverse iff start > end */
    int seeked;
    int64_t pos;       /* current record */
    int64_t first;     /* first record to yield */
    const uint8_t *end;
    size_t elen;
    uint64_t limit, counter;
    int eof;
} ko_iter;

static void iter_init(ko_iter *it, const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end,
                      size_t elen, uint64_t limit)
{
    it->s = s;
    it->reverse = ko_bytes_compare(start, slen, end, elen) > 0;
    it->seeked = 0;
    it->end = end;
    it->elen = elen;
    it->limit = limit;
    it->counter = 0;
    it->eof = 0;
    if (!it->reverse)
        it->first = (int64_t)ko_lower_bound(s, start, slen); /* Seek(start): first key >= start */
    else
        it->first = (int64_t)upper_bound(s, start, slen) - 1; /* reverse Seek: last key <= start */
    it->pos = it->first;
}

/* returns 1 when positioned on a record, 0 at io.EOF (iter.go:65-83, inRange :50-63) */
static int iter_next(ko_iter *it)
{
    if (it->eof) return 0;
    if (it->seeked)
        it->pos += it->reverse ? -1 : 1;
    else
        it->seeked = 1;
    if (it->pos < 0 || it->pos >= (int64_t)it->s->n) {
        it->eof = 1;
        return 0;
    }
    if (it->limit != 0 && it->counter >= it->limit) {
        it->eof = 1;
        return 0;
    }
    it->counter++;
    size_t kl;
    const uint8_t *k = rec_key(it->s, (uint64_t)it->pos, &kl);
    int cmp = ko_bytes_compare(k, kl, it->end, it->elen);
    if (it->reverse ? !(cmp > 0) : !(cmp < 0)) {
        it->eof = 1;
        return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* result receiver (scanner/receiver.go:62-103)                                                 */
/* ------------------------------------------------------------------------------------------ */

void ko_result_init(ko_result *r) { memset(r, 0, sizeof(*r)); }

void ko_result_free(ko_result *r)
{
    free(r->emit);
    free(r->victim);
    free(r->vclass);
    memset(r, 0, sizeof(*r));
}

static int res_append(ko_result *r, uint64_t idx)
{
    if (r->n_emit == r->cap_emit) {
        uint64_t nc = r->cap_emit ? r->cap_emit * 2 : 1024;
        uint64_t *p = (uint64_t *)realloc(r->emit, nc * sizeof(uint64_t));
        if (!p) return KO_ENOMEM;
        r->emit = p;
        r->cap_emit = nc;
    }
    r->emit[r->n_emit++] = idx;
    return KO_OK;
}

static int res_victim(ko_result *r, uint64_t idx, uint8_t cls)
{
    if (r->n_victim == r->cap_victim) {
        uint64_t nc = r->cap_victim ? r->cap_victim * 2 : 1024;
        uint64_t *p = (uint64_t *)realloc(r->victim, nc * sizeof(uint64_t));
        if (!p) return KO_ENOMEM;
        r->victim = p;
        uint8_t *q = (uint8_t *)realloc(r->vclass, nc);
        if (!q) return KO_ENOMEM;
        r->vclass = q;
        r->cap_victim = nc;
    }
    r->victim[r->n_victim] = idx;
    r->vclass[r->n_victim] = cls;
    r->n_victim++;
    return KO_OK;
}

static inline int is_tombstone(const uint8_t *v, size_t len)
{
    return len == 9 && memcmp(v, KO_TOMBSTONE, 9) == 0;
}

static int contains_events(const uint8_t *k, size_t len) /* bytes.Contains(rawKey, "/events/") scanner.go:573 */
{
    if (len < 8) return 0;
    for (size_t i = 0; i + 8 <= len; i++)
        if (k[i] == '/' && memcmp(k + i, KO_EVENTS, 8) == 0) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* worker.run (scanner/scanner.go:389-516)                                                      */
/* ------------------------------------------------------------------------------------------ */

int ko_worker_run(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
                  const ko_worker_cfg *cfg, ko_result *out)
{
    ko_iter it;
    iter_init(&it, s, start, slen, end, elen, 0); /* scanner.go:395 Iter(start,end,tso,0) */

    int count = 0;
    uint64_t val_size = 0;
    /* scanner.go:408-414 loop state; prevUserKey == nil compares equal to an empty key (Q6) */
    const uint8_t *prev_uk = NULL;
    size_t prev_uk_len = 0;
    uint64_t prev_rev = 0;
    const uint8_t *prev_val = NULL;
    size_t prev_val_len = 0;
    uint64_t prev_idx = 0;
    int eof = 0, rc = KO_OK;
    const int limited = cfg->collect && cfg->limit > 0;

    /* receiver.reset() scanner.go:403 */
    out->n_emit = 0;

    for (;;) {
        if (limited && (int64_t)out->n_emit >= cfg->limit) break; /* needMore() receiver.go:82-87 */
        if (!iter_next(&it)) {
            eof = 1;
            break;
        }
        out->examined++;
        uint64_t idx = (uint64_t)it.pos;
        size_t klen, uk_off, uk_len, vlen;
        uint64_t rev;
        const uint8_t *key = rec_key(s, idx, &klen);
        if (ko_decode(key, klen, &uk_off, &uk_len, &rev) != KO_OK) continue; /* scanner.go:436-439 (Q2) */
        const uint8_t *uk = key + uk_off;
        const uint8_t *val = rec_val(s, idx, &vlen);
        val_size += vlen;

        /* compactIfExpired scanner.go:566-591 */
        if (!cfg->support_ttl && cfg->timeout_rev != 0 && contains_events(uk, uk_len)) {
            if (rev == 0) {
                if (vlen < 8) return KO_EINVAL; /* Go: value[:8] panics */
                uint64_t r = be64(val);
                if (r <= cfg->timeout_rev) {
                    if ((rc = res_victim(out, idx, KO_V_TTL_REVREC))) return rc;
                    continue;
                }
            } else if (rev <= cfg->timeout_rev) {
                if ((rc = res_victim(out, idx, KO_V_TTL_OBJECT))) return rc;
                continue;
            }
        }

        if (rev > cfg->read_rev) continue; /* scanner.go:451-453 (Q1) */

        int same = (uk_len == prev_uk_len) && (uk_len == 0 || memcmp(uk, prev_uk, uk_len) == 0);
        if (!same) {
            /* scanner.go:457-462 */
            if (prev_rev > 0 && !is_tombstone(prev_val, prev_val_len)) {
                if (cfg->collect && (rc = res_append(out, prev_idx))) return rc;
                count++;
            }
        } else if (cfg->compact && prev_rev > 0) {
            /* scanner.go:463-469 */
            if ((rc = res_victim(out, prev_idx, KO_V_SUPERSEDED))) return rc;
        }
        /* scanner.go:471-475 */
        if (cfg->compact && is_tombstone(val, vlen)) {
            if ((rc = res_victim(out, idx, KO_V_TOMBSTONE))) return rc;
        }
        /* scanner.go:476-491 */
        if (cfg->compact && rev == 0 && vlen == 9) {
            uint64_t obj_rev = be64(val);
            if (obj_rev > cfg->read_rev) continue; /* (Q5) prev is NOT updated */
            if ((rc = res_victim(out, idx, KO_V_REVRECORD))) return rc;
        }
        /* scanner.go:493-495 */
        prev_rev = rev;
        prev_uk = uk;
        prev_uk_len = uk_len;
        prev_val = val;
        prev_val_len = vlen;
        prev_idx = idx;
    }

    out->val_size += val_size;
    if (!eof) {
        /* scanner.go:499-502: err is nil (not io.EOF) when the limit stopped the loop -> (0, nil) (Q4) */
        out->limit_stop = 1;
        out->count = 0;
        return KO_OK;
    }
    /* scanner.go:503-507 trailing object */
    if (prev_rev > 0 && !is_tombstone(prev_val, prev_val_len) &&
        !(limited && (int64_t)out->n_emit >= cfg->limit)) {
        if (cfg->collect && (rc = res_append(out, prev_idx))) return rc;
        count++;
    }
    out->count = count;
    return KO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* adjustPartitionsBorders (scanner.go:202-225) over contiguous partitions                      */
/* ------------------------------------------------------------------------------------------ */

int ko_adjust_partition_borders(const uint8_t *borders, const uint64_t *boff, uint64_t n_borders, uint8_t *out,
                                uint64_t *out_off)
{
    uint64_t w = 0;
    out_off[0] = 0;
    for (uint64_t i = 0; i < n_borders; i++) {
        const uint8_t *b = borders + boff[i];
        size_t len = (size_t)(boff[i + 1] - boff[i]);
        size_t uk_off, uk_len;
        uint64_t rev;
        int interior = (i != 0) && (i != n_borders - 1);
        if (interior && ko_decode(b, len, &uk_off, &uk_len, &rev) == KO_OK && rev != 0) {
            /* border is an object key: move it forward to the revision key of the same user key */
            w += ko_encode_object_key(b + uk_off, uk_len, 0, out + w);
        } else {
            memcpy(out + w, b, len);
            w += len;
        }
        out_off[i + 1] = w;
    }
    return KO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* scan (scanner.go:227-304) + checkCompactRace (:594-626) + receiver merge (receiver.go:72-80)  */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const ko_store *s;
    const uint8_t *start, *end;
    size_t slen, elen;
    const ko_worker_cfg *cfg;
    ko_result res;
    int rc;
} scan_task;

typedef struct {
    scan_task *tasks;
    uint64_t n;
    uint64_t next;
    pthread_mutex_t mu;
} scan_pool;

static void *scan_thread(void *arg)
{
    scan_pool *p = (scan_pool *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        uint64_t i = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (i >= p->n) break;
        scan_task *t = &p->tasks[i];
        t->rc = ko_worker_run(t->s, t->start, t->slen, t->end, t->elen, t->cfg, &t->res);
    }
    return NULL;
}

int ko_scan(const ko_store *s, const uint8_t *borders, const uint64_t *boff, uint64_t n_borders,
            const ko_worker_cfg *cfg, int compact_rev_present, uint64_t compact_rev, int threads, ko_result *out,
            int *total_count)
{
    if (n_borders < 2) return KO_EINVAL;
    /* checkCompactRace: a range below the stored compact revision is refused */
    if (!cfg->compact && compact_rev_present && compact_rev > cfg->read_rev) return KO_ECOMPACTED;

    uint64_t total = boff[n_borders];
    uint8_t *adj = (uint8_t *)malloc(total + 16 * n_borders + 16);
    uint64_t *aoff = (uint64_t *)malloc((n_borders + 1) * sizeof(uint64_t));
    if (!adj || !aoff) {
        free(adj);
        free(aoff);
        return KO_ENOMEM;
    }
    ko_adjust_partition_borders(borders, boff, n_borders, adj, aoff);

    uint64_t P = n_borders - 1;
    scan_task *tasks = (scan_task *)calloc(P, sizeof(scan_task));
    for (uint64_t i = 0; i < P; i++) {
        tasks[i].s = s;
        tasks[i].start = adj + aoff[i];
        tasks[i].slen = (size_t)(aoff[i + 1] - aoff[i]);
        tasks[i].end = adj + aoff[i + 1];
        tasks[i].elen = (size_t)(aoff[i + 2] - aoff[i + 1]);
        tasks[i].cfg = cfg;
        ko_result_init(&tasks[i].res);
    }
    scan_pool pool = {tasks, P, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads <= 1 || P == 1) {
        scan_thread(&pool);
    } else {
        if ((uint64_t)threads > P) threads = (int)P;
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, scan_thread, &pool);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        free(th);
    }

    int rc = KO_OK, count = 0;
    for (uint64_t i = 0; i < P && rc == KO_OK; i++) rc = tasks[i].rc;
    if (rc == KO_OK) {
        const int limited = cfg->collect && cfg->limit > 0;
        for (uint64_t i = 0; i < P; i++) {
            ko_result *r = &tasks[i].res;
            count += r->count;
            out->examined += r->examined;
            out->val_size += r->val_size;
            /* merge in partition order, truncating to limit (receiver.go:72-80) */
            for (uint64_t k = 0; k < r->n_emit; k++) {
                if (limited && (int64_t)out->n_emit >= cfg->limit) break;
                if ((rc = res_append(out, r->emit[k]))) break;
            }
            for (uint64_t k = 0; k < r->n_victim && rc == KO_OK; k++) rc = res_victim(out, r->victim[k], r->vclass[k]);
        }
    }
    for (uint64_t i = 0; i < P; i++) ko_result_free(&tasks[i].res);
    free(tasks);
    free(adj);
    free(aoff);
    out->count = count;
    if (total_count) *total_count = count;
    return rc;
}

/* scanner.go:83-119 */
int ko_range(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
             uint64_t read_rev, int64_t limit, int compact_rev_present, uint64_t compact_rev, ko_result *out)
{
    ko_worker_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.read_rev = read_rev;
    cfg.limit = limit;
    cfg.collect = 1;
    cfg.support_ttl = 1;
    if (limit > 0) {
        /* rangeWithLimit: one worker over the whole interval */
        if (compact_rev_present && compact_rev > read_rev) return KO_ECOMPACTED;
        return ko_worker_run(s, start, slen, end, elen, &cfg, out);
    }
    /* badger GetPartitions returns the single partition [start,end) (badger.go:52-54) */
    uint8_t *b = (uint8_t *)malloc(slen + elen + 1);
    uint64_t off[3] = {0, slen, slen + elen};
    memcpy(b, start, slen);
    memcpy(b + slen, end, elen);
    int total;
    int rc = ko_scan(s, b, off, 2, &cfg, compact_rev_present, compact_rev, 1, out, &total);
    free(b);
    return rc;
}

/* range.go:91-121 getInternalVal (reverse iterator, limit 1) + :81-87 get */
int64_t ko_get(const ko_store *s, const uint8_t *ukey, size_t ulen, uint64_t rev, uint64_t *mod_rev)
{
    if (rev == 0) rev = UINT64_MAX;
    uint8_t *sk = (uint8_t *)malloc(2 * (ulen + 13));
    uint8_t *ek = sk + ulen + 13;
    ko_encode_object_key(ukey, ulen, rev, sk);
    ko_encode_object_key(ukey, ulen, 0, ek);
    ko_iter it;
    iter_init(&it, s, sk, ulen + 13, ek, ulen + 13, 1);
    int64_t ret = -1;
    *mod_rev = 0;
    if (iter_next(&it)) {
        size_t klen, uk_off = 0, uk_len = 0, vlen;
        uint64_t mrev = 0;
        const uint8_t *key = rec_key(s, (uint64_t)it.pos, &klen);
        int ok = ko_decode(key, klen, &uk_off, &uk_len, &mrev) == KO_OK;
        if (ok && mrev != 0 && uk_len == ulen && (ulen == 0 || memcmp(key + uk_off, ukey, ulen) == 0)) {
            const uint8_t *val = rec_val(s, (uint64_t)it.pos, &vlen);
            *mod_rev = mrev;
            ret = is_tombstone(val, vlen) ? -2 : it.pos;
        }
    }
    free(sk);
    return ret;
}

/* compact.go:107-127 */
typedef struct {
    uint8_t *p;
    size_t len;
} ko_slice;

static int slice_cmp(const void *a, const void *b)
{
    const ko_slice *x = (const ko_slice *)a, *y = (const ko_slice *)b;
    return ko_bytes_compare(x->p, x->len, y->p, y->len);
}

int ko_compact_borders(const uint8_t *prefixes, const uint64_t *poff, uint64_t n_prefixes, uint8_t *out,
                       uint64_t *out_off)
{
    ko_slice *sl = (ko_slice *)calloc(2 * n_prefixes, sizeof(ko_slice));
    if (!sl) return KO_ENOMEM;
    for (uint64_t i = 0; i < n_prefixes; i++) {
        size_t len = (size_t)(poff[i + 1] - poff[i]);
        uint8_t *key = (uint8_t *)malloc(len + 2);
        memcpy(key, prefixes + poff[i], len);
        if (len == 0 || key[len - 1] != '/') key[len++] = '/'; /* strings.HasSuffix(key, "/") */
        uint8_t *pe = (uint8_t *)malloc(len + 1);
        size_t pelen = ko_prefix_end(key, len, pe);
        sl[2 * i].p = (uint8_t *)malloc(len + 13);
        sl[2 * i].len = ko_encode_object_key(key, len, 0, sl[2 * i].p);
        sl[2 * i + 1].p = (uint8_t *)malloc(pelen + 13);
        sl[2 * i + 1].len = ko_encode_object_key(pe, pelen, 0, sl[2 * i + 1].p);
        free(key);
        free(pe);
    }
    qsort(sl, 2 * n_prefixes, sizeof(ko_slice), slice_cmp);
    uint64_t w = 0;
    out_off[0] = 0;
    for (uint64_t i = 0; i < 2 * n_prefixes; i++) {
        memcpy(out + w, sl[i].p, sl[i].len);
        w += sl[i].len;
        out_off[i + 1] = w;
        free(sl[i].p);
    }
    free(sl);
    return KO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Ring (backend/ring.go:24-118)                                                                */
/* ------------------------------------------------------------------------------------------ */

struct ko_ring {
    int64_t s, e;
    int64_t l;
    uint64_t *rev;
    uint64_t *payload;
};

ko_ring *ko_ring_new(int64_t capacity)
{
    ko_ring *r = (ko_ring *)calloc(1, sizeof(ko_ring));
    r->l = capacity;
    r->rev = (uint64_t *)calloc((size_t)capacity, sizeof(uint64_t));
    r->payload = (uint64_t *)calloc((size_t)capacity, sizeof(uint64_t));
    return r;
}

void ko_ring_free(ko_ring *r)
{
    if (!r) return;
    free(r->rev);
    free(r->payload);
    free(r);
}

void ko_ring_reset(ko_ring *r) { r->s = r->e = 0; } /* ring.go:57-61 */

static inline int64_t ring_index(const ko_ring *r, int64_t i) { return i % r->l; } /* ring.go:67-69 */

void ko_ring_add(ko_ring *r, uint64_t revision, uint64_t payload) /* ring.go:38-46 */
{
    int64_t i = ring_index(r, r->e);
    r->rev[i] = revision;
    r->payload[i] = payload;
    if (r->e == r->s + r->l) r->s++;
    r->e++;
}

void ko_ring_find(const ko_ring *r, uint64_t revision, ko_find_ret *ret, uint64_t *revs, uint64_t *payloads)
{
    memset(ret, 0, sizeof(*ret));
    if (r->e == 0) { /* ring.go:89-92 isEmpty */
        ret->empty = 1;
        return;
    }
    ret->newest_rev = r->rev[ring_index(r, r->e - 1)];
    ret->oldest_rev = r->rev[ring_index(r, r->s)];
    if (revision > ret->newest_rev) {
        ret->high = 1;
        return;
    }
    if (revision < ret->oldest_rev) {
        ret->low = 1;
        return;
    }
    /* sort.Search over [0, e-s) for the first Revision >= revision (ring.go:105-107) */
    int64_t n = r->e - r->s, lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (!(r->rev[ring_index(r, r->s + mid)] >= revision))
            lo = mid + 1;
        else
            hi = mid;
    }
    int64_t idx = lo, cnt = n - idx;
    ret->n_events = (uint64_t)cnt;
    /* wrap-aware copy, ring.go:111-117 (restated as the same two-segment copy) */
    int64_t from = ring_index(r, r->s + idx), to = ring_index(r, r->e);
    if (to > from) {
        for (int64_t k = 0; k < to - from && k < cnt; k++) {
            revs[k] = r->rev[from + k];
            payloads[k] = r->payload[from + k];
        }
        return;
    }
    int64_t k = 0;
    for (int64_t j = from; j < r->l && k < cnt; j++, k++) {
        revs[k] = r->rev[j];
        payloads[k] = r->payload[j];
    }
    for (int64_t j = 0; j < to && k < cnt; j++, k++) {
        revs[k] = r->rev[j];
        payloads[k] = r->payload[j];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* watch fan-out (watch.go:119-159; watcherhub.go:78-92 hands every batch to every watcher)      */
/* ------------------------------------------------------------------------------------------ */

static inline int has_prefix(const uint8_t *k, size_t klen, const uint8_t *p, size_t plen)
{
    return klen >= plen && (plen == 0 || memcmp(k, p, plen) == 0); /* bytes.HasPrefix */
}

typedef struct {
    const ko_events *ev;
    const ko_watchers *w;
    uint64_t w_lo, w_hi;
    int alloc_per_batch;
    uint32_t **lists;   /* per watcher delivery list */
    uint64_t *counts, *caps;
    uint64_t messages;
} fan_task;

static void fan_push(fan_task *t, uint64_t wi, uint32_t e)
{
    if (t->counts[wi] == t->caps[wi]) {
        uint64_t nc = t->caps[wi] ? t->caps[wi] * 2 : 16;
        t->lists[wi] = (uint32_t *)realloc(t->lists[wi], nc * sizeof(uint32_t));
        t->caps[wi] = nc;
    }
    t->lists[wi][t->counts[wi]++] = e;
}

static void *fan_thread(void *arg)
{
    fan_task *t = (fan_task *)arg;
    const ko_events *ev = t->ev;
    for (uint64_t wi = t->w_lo; wi < t->w_hi; wi++) {
        const uint8_t *p = t->w->prefixes + t->w->poff[wi];
        size_t plen = (size_t)(t->w->poff[wi + 1] - t->w->poff[wi]);
        uint64_t min_rev = t->w->min_rev[wi];
        for (uint64_t b = 0; b < ev->n_batches; b++) {
            uint64_t lo = ev->batch_off[b], hi = ev->batch_off[b + 1];
            /* filterByRevision watch.go:153-159: strip LEADING events below min_rev only */
            while (lo < hi && ev->rev[lo] < min_rev) lo++;
            /* filterByPrefix watch.go:140-150 (allocates the output list per call, :141) */
            void *scratch = NULL;
            if (t->alloc_per_batch) scratch = malloc((size_t)(hi - lo) * sizeof(void *) + 1);
            uint64_t before = t->counts[wi];
            for (uint64_t e = lo; e < hi; e++) {
                const uint8_t *k = ev->keys + ev->koff[e];
                size_t klen = (size_t)(ev->koff[e + 1] - ev->koff[e]);
                if (has_prefix(k, klen, p, plen)) {
                    if (scratch) ((void **)scratch)[e - lo] = (void *)k;
                    fan_push(t, wi, (uint32_t)e);
                }
            }
            if (t->counts[wi] > before) t->messages++; /* watch.go:128-130 only non-empty lists are sent */
            free(scratch);
        }
    }
    return NULL;
}

void ko_fanout_free(ko_fanout *f)
{
    free(f->start);
    free(f->event_idx);
    memset(f, 0, sizeof(*f));
}

int ko_fanout_run(const ko_events *ev, const ko_watchers *w, int threads, int alloc_per_batch, ko_fanout *out)
{
    if (threads < 1) threads = 1;
    uint64_t W = w->n;
    uint32_t **lists = (uint32_t **)calloc(W ? W : 1, sizeof(uint32_t *));
    uint64_t *counts = (uint64_t *)calloc(W ? W : 1, sizeof(uint64_t));
    uint64_t *caps = (uint64_t *)calloc(W ? W : 1, sizeof(uint64_t));
    fan_task *tasks = (fan_task *)calloc((size_t)threads, sizeof(fan_task));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        tasks[t].ev = ev;
        tasks[t].w = w;
        tasks[t].w_lo = W * (uint64_t)t / (uint64_t)threads;
        tasks[t].w_hi = W * (uint64_t)(t + 1) / (uint64_t)threads;
        tasks[t].alloc_per_batch = alloc_per_batch;
        tasks[t].lists = lists;
        tasks[t].counts = counts;
        tasks[t].caps = caps;
    }
    if (threads == 1) {
        fan_thread(&tasks[0]);
    } else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, fan_thread, &tasks[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    out->start = (uint64_t *)malloc((W + 1) * sizeof(uint64_t));
    uint64_t total = 0;
    for (uint64_t i = 0; i < W; i++) {
        out->start[i] = total;
        total += counts[i];
    }
    out->start[W] = total;
    out->event_idx = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < W; i++) {
        if (counts[i]) memcpy(out->event_idx + out->start[i], lists[i], counts[i] * sizeof(uint32_t));
        free(lists[i]);
    }
    out->n_deliveries = total;
    out->n_messages = 0;
    for (int t = 0; t < threads; t++) out->n_messages += tasks[t].messages;
    free(lists);
    free(counts);
    free(caps);
    free(tasks);
    free(th);
    return KO_OK;
}

/* watch.go:37-99 */
int ko_watch_register(const ko_ring *ring, const ko_events *ev, const uint8_t *prefix, size_t plen,
                      uint64_t revision, uint64_t current_rev, ko_watch_reg *reg, uint64_t *catchup, uint64_t cap)
{
    memset(reg, 0, sizeof(*reg));
    if (revision == 0) { /* watch.go:54-57 */
        reg->mode = 0;
        reg->live_rev = 0;
        return KO_OK;
    }
    uint64_t *revs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ring->l ? ring->l : 1));
    uint64_t *pay = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ring->l ? ring->l : 1));
    ko_find_ret ret;
    ko_ring_find(ring, revision, &ret, revs, pay);
    if (ret.empty) { /* watch.go:61-72 */
        if (revision > current_rev) {
            reg->mode = 0;
            reg->live_rev = revision;
        } else {
            reg->mode = 1;
            reg->err_rev = current_rev;
        }
    } else if (ret.high) { /* watch.go:74-77 */
        reg->mode = 0;
        reg->live_rev = revision;
    } else if (ret.low) { /* watch.go:79-85 */
        reg->mode = 2;
        reg->err_rev = ret.oldest_rev;
    } else { /* watch.go:87-97 */
        uint64_t n = 0;
        for (uint64_t i = 0; i < ret.n_events; i++) {
            uint64_t e = pay[i];
            const uint8_t *k = ev->keys + ev->koff[e];
            size_t klen = (size_t)(ev->koff[e + 1] - ev->koff[e]);
            if (has_prefix(k, klen, prefix, plen)) {
                if (n < cap) catchup[n] = e;
                n++;
            }
        }
        reg->mode = 3;
        reg->n_catchup = n;
        reg->live_rev = n > 0 ? ret.newest_rev + 1 : revision;
    }
    free(revs);
    free(pay);
    return KO_OK;
}

/* watch.go:102-117 (eventBatchSize = 300 backend.go:41, resultChanLength = 100 watch.go:30) */
uint64_t ko_catchup_chunks(uint64_t n_events, uint64_t *sizes, uint64_t cap)
{
    uint64_t batch = 300, n = 0, left = n_events;
    if (n_events > 100 * 300) batch = n_events / (100 - 1);
    for (;;) {
        if (left > batch) {
            if (n < cap) sizes[n] = batch;
            n++;
            left -= batch;
        } else {
            if (n < cap) sizes[n] = left;
            n++;
            break;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* CPU-baseline timing variants (bench.py only)                                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const ko_store *s;
    uint64_t lo, hi;      /* record interval of this partition */
    uint64_t read_rev;
    int64_t limit;
    int faithful;
    int64_t emitted;
    uint64_t examined, checksum;
} bench_task;

typedef struct {
    uint8_t *key;   /* KeyCopy */
    uint8_t *val;   /* ValueCopy */
    size_t klen, vlen;
} held_kv;

/* The same state machine as ko_worker_run (non-compact), operating on a record interval, with the
 * badger iterator's per-record heap copies when faithful (iter.go:85-92: Key()=KeyCopy(nil),
 * Val()=ValueCopy(nil); worker.run calls Val() twice per record, scanner.go:441 and :495). */
static void *bench_thread(void *arg)
{
    bench_task *t = (bench_task *)arg;
    const ko_store *s = t->s;
    const uint8_t *prev_uk = NULL, *prev_val = NULL;
    size_t prev_uk_len = 0, prev_val_len = 0;
    uint64_t prev_rev = 0, prev_idx = 0;
    uint8_t *prev_keybuf = NULL, *prev_valbuf = NULL;
    held_kv *held = NULL;
    uint64_t n_held = 0, cap_held = 0;
    int64_t emitted = 0;
    uint64_t checksum = 0, examined = 0;
    int eof = 1;

    for (uint64_t idx = t->lo; idx < t->hi; idx++) {
        if (t->limit > 0 && emitted >= t->limit) {
            eof = 0;
            break;
        }
        examined++;
        size_t klen, vlen, uk_off, uk_len;
        uint64_t rev;
        const uint8_t *key = rec_key(s, idx, &klen);
        const uint8_t *val = rec_val(s, idx, &vlen);
        uint8_t *keybuf = NULL, *valbuf = NULL;
        if (t->faithful) {
            keybuf = (uint8_t *)malloc(klen ? klen : 1);
            memcpy(keybuf, key, klen);
            key = keybuf;
        }
        if (ko_decode(key, klen, &uk_off, &uk_len, &rev) != KO_OK) {
            free(keybuf);
            continue;
        }
        if (t->faithful) { /* first Val(): scanner.go:441 */
            uint8_t *v1 = (uint8_t *)malloc(vlen ? vlen : 1);
            memcpy(v1, val, vlen);
            checksum += v1[0];
            free(v1);
        }
        if (rev > t->read_rev) {
            free(keybuf);
            continue;
        }
        const uint8_t *uk = key + uk_off;
        int same = (uk_len == prev_uk_len) && (uk_len == 0 || memcmp(uk, prev_uk, uk_len) == 0);
        if (!same) {
            if (prev_rev > 0 && !is_tombstone(prev_val, prev_val_len)) {
                emitted++;
                checksum += prev_idx * 1315423911ull + prev_rev;
                if (t->faithful) { /* result keeps the copies alive */
                    if (n_held == cap_held) {
                        cap_held = cap_held ? cap_held * 2 : 1024;
                        held = (held_kv *)realloc(held, cap_held * sizeof(held_kv));
                    }
                    held[n_held].key = prev_keybuf;
                    held[n_held].val = prev_valbuf;
                    n_held++;
                    prev_keybuf = prev_valbuf = NULL;
                }
            }
        }
        if (t->faithful) { /* second Val(): scanner.go:495 */
            valbuf = (uint8_t *)malloc(vlen ? vlen : 1);
            memcpy(valbuf, val, vlen);
            val = valbuf;
            free(prev_keybuf);
            free(prev_valbuf);
            prev_keybuf = keybuf;
            prev_valbuf = valbuf;
        }
        prev_rev = rev;
        prev_uk = uk;
        prev_uk_len = uk_len;
        prev_val = val;
        prev_val_len = vlen;
        prev_idx = idx;
    }
    if (eof && prev_rev > 0 && !is_tombstone(prev_val, prev_val_len) && !(t->limit > 0 && emitted >= t->limit)) {
        emitted++;
        checksum += prev_idx * 1315423911ull + prev_rev;
    }
    free(prev_keybuf);
    free(prev_valbuf);
    for (uint64_t i = 0; i < n_held; i++) {
        free(held[i].key);
        free(held[i].val);
    }
    free(held);
    t->emitted = emitted;
    t->examined = examined;
    t->checksum = checksum;
    return NULL;
}

int64_t ko_bench_scan(const ko_store *s, const uint8_t *start, size_t slen, const uint8_t *end, size_t elen,
                      uint64_t read_rev, int64_t limit, int faithful, int threads, uint64_t *examined,
                      uint64_t *checksum)
{
    uint64_t lo = ko_lower_bound(s, start, slen), hi = ko_lower_bound(s, end, elen);
    if (hi < lo) hi = lo;
    if (threads < 1 || limit > 0) threads = 1; /* rangeWithLimit is a single worker, scanner.go:96-118 */
    bench_task *tasks = (bench_task *)calloc((size_t)threads, sizeof(bench_task));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    uint64_t cut = lo;
    for (int t = 0; t < threads; t++) {
        uint64_t nxt = (t == threads - 1) ? hi : lo + (hi - lo) * (uint64_t)(t + 1) / (uint64_t)threads;
        /* snap the border like adjustPartitionsBorders: to the revision key of the border's user key */
        if (t != threads - 1 && nxt > cut && nxt < hi) {
            size_t klen, uk_off, uk_len;
            uint64_t rev;
            const uint8_t *key = rec_key(s, nxt, &klen);
            if (ko_decode(key, klen, &uk_off, &uk_len, &rev) == KO_OK && rev != 0) {
                uint8_t *rk = (uint8_t *)malloc(uk_len + 13);
                ko_encode_object_key(key + uk_off, uk_len, 0, rk);
                uint64_t b = ko_lower_bound(s, rk, uk_len + 13);
                free(rk);
                if (b > cut) nxt = b; else nxt = cut;
            }
        }
        if (nxt < cut) nxt = cut;
        tasks[t].s = s;
        tasks[t].lo = cut;
        tasks[t].hi = nxt;
        tasks[t].read_rev = read_rev;
        tasks[t].limit = limit;
        tasks[t].faithful = faithful;
        cut = nxt;
    }
    if (threads == 1) {
        bench_thread(&tasks[0]);
    } else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, bench_thread, &tasks[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    int64_t emitted = 0;
    uint64_t ex = 0, cs = 0;
    for (int t = 0; t < threads; t++) {
        emitted += tasks[t].emitted;
        ex += tasks[t].examined;
        cs += tasks[t].checksum;
    }
    if (examined) *examined = ex;
    if (checksum) *checksum = cs;
    free(tasks);
    free(th);
    return emitted;
}

/* ============================================================================================
 * etcd wire encoding (see kb_oracle.h): protobuf base-128 varints, length-delimited fields
 * ============================================================================================ */
static uint64_t varint_len(uint64_t v)
{
    uint64_t n = 1;
    while (v >= 0x80) {
        v >>= 7;
        n++;
    }
    return n;
}

static uint64_t put_varint(uint8_t *out, uint64_t v)
{
    uint64_t n = 0;
    while (v >= 0x80) {
        if (out) out[n] = (uint8_t)(v | 0x80);
        v >>= 7;
        n++;
    }
    if (out) out[n] = (uint8_t)v;
    return n + 1;
}

/* mvccpb.KeyValue{Key, Value, ModRevision} body (backendshim.go:427-436) */
static uint64_t kv_body_size(uint64_t uk_len, uint64_t val_len, uint64_t rev)
{
    uint64_t n = 0;
    if (uk_len) n += 1 + varint_len(uk_len) + uk_len;   /* field 1, bytes  */
    if (rev) n += 1 + varint_len(rev);                  /* field 3, int64 (two's complement varint) */
    if (val_len) n += 1 + varint_len(val_len) + val_len; /* field 5, bytes */
    return n;
}

uint64_t ko_wire_elem_size(uint64_t uk_len, uint64_t val_len, uint64_t rev, int mode)
{
    const uint64_t body = kv_body_size(uk_len, val_len, rev);
    const uint64_t kv = 1 + varint_len(body) + body; /* RangeResponse.kvs (2) / Event.kv (2): tag 0x12 */
    if (mode == KO_WIRE_KVS) return kv;
    return 1 + varint_len(kv) + kv; /* WatchResponse.events (11): tag 0x5a around Event{kv} */
}

uint64_t ko_wire_encode(const ko_store *s, const uint64_t *rec, uint64_t n, int mode, uint8_t *out,
                        uint64_t *elem_off)
{
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t r = rec[i];
        const uint8_t *k = s->keys + s->koff[r];
        const uint64_t kl = s->koff[r + 1] - s->koff[r], vl = s->voff[r + 1] - s->voff[r];
        const uint8_t *v = s->vals + s->voff[r];
        const uint64_t ul = kl - 13, rev = be64(k + kl - 8);
        const uint64_t body = kv_body_size(ul, vl, rev);
        if (elem_off) elem_off[i] = w;
        if (mode == KO_WIRE_EVENTS) {
            if (out) out[w] = 0x5a;
            w += 1;
            w += put_varint(out ? out + w : NULL, 1 + varint_len(body) + body);
        }
        if (out) out[w] = 0x12;
        w += 1;
        w += put_varint(out ? out + w : NULL, body);
        if (ul) {
            if (out) out[w] = 0x0a;
            w += 1;
            w += put_varint(out ? out + w : NULL, ul);
            if (out) memcpy(out + w, k + 4, ul);
            w += ul;
        }
        if (rev) {
            if (out) out[w] = 0x18;
            w += 1;
            w += put_varint(out ? out + w : NULL, rev);
        }
        if (vl) {
            if (out) out[w] = 0x2a;
            w += 1;
            w += put_varint(out ? out + w : NULL, vl);
            if (out) memcpy(out + w, v, vl);
            w += vl;
        }
    }
    if (elem_off) elem_off[n] = w;
    return w;
}

/* etcdserverpb.ResponseHeader{Revision} as field 1 of the enclosing response (kv.go:253-257 txnHeader); a header
 * with revision 0 is still a non-nil message: tag + zero length */
static uint64_t put_header(uint64_t header_rev, uint8_t *out)
{
    uint64_t w = 0;
    const uint64_t body = header_rev ? 1 + varint_len(header_rev) : 0;
    if (out) out[w] = 0x0a;
    w += 1;
    w += put_varint(out ? out + w : NULL, body);
    if (header_rev) {
        if (out) out[w] = 0x18;
        w += 1;
        w += put_varint(out ? out + w : NULL, header_rev);
    }
    return w;
}

uint64_t ko_wire_range_head(uint64_t header_rev, uint8_t *out) { return put_header(header_rev, out); }

uint64_t ko_wire_range_tail(int more, int64_t count, uint8_t *out)
{
    uint64_t w = 0;
    if (more) {
        if (out) {
            out[w] = 0x18;
            out[w + 1] = 1;
        }
        w += 2;
    }
    if (count) {
        if (out) out[w] = 0x20;
        w += 1;
        w += put_varint(out ? out + w : NULL, (uint64_t)count);
    }
    return w;
}

uint64_t ko_wire_watch_head(uint64_t header_rev, int canceled, const uint8_t *reason, uint64_t reason_len,
                            uint8_t *out)
{
    uint64_t w = put_header(header_rev, out);
    if (canceled) {
        if (out) {
            out[w] = 0x20;
            out[w + 1] = 1;
        }
        w += 2;
    }
    if (reason_len) {
        if (out) out[w] = 0x32;
        w += 1;
        w += put_varint(out ? out + w : NULL, reason_len);
        if (out) memcpy(out + w, reason, reason_len);
        w += reason_len;
    }
    return w;
}
