/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load this.  A plain-C restatement of oracle/evlfu.py: BatchEvLFU -- the batch-granular generalisation of the
 * reference's sequential EvLFU (/root/reference/cache_algo/EvLFU_C1.py:32-166: set() :32-63, update_agg_hit() :65-78,
 * request_to_ev_lfu() :97-166) that DESIGN.md section 3 freezes -- fast enough to follow the CUDA path at the
 * BASELINE sizes (33.76 M rows, 4.39 M cache entries, 53 248 keys per batch), where the Python oracle takes hours.
 * It is pinned to oracle/evlfu.py (which is pinned to the reference itself) by tests/test_oracle_c_batch.py.
 *
 * State, as in EvLFU_C1.py:7-19: vals (key -> agg_hit) = an open-addressing map onto a node pool; lists[0..T] = one
 * doubly linked FIFO per agg_hit bucket (append at the tail, evict from the head); n_perfect; max_perfect =
 * int(cap * 0.95); flush = int(0.3 * cap) + 1.  Keys are (table << 40) | row.
 *
 *   gcc -O2 -shared -fPIC oracle/evlfu_batch.c -o oracle/_ref/libevlfu_batch.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EMPTY 0xFFFFFFFFu
#define NIL 0xFFFFFFFFu

typedef struct {
    uint64_t key;
    uint32_t prev, next;
    int32_t bucket;
} Node;

typedef struct {
    int T;                      /* agg_hit range 0..T */
    int64_t cap, max_perfect, flush_n;
    int64_t n_perfect, size;
    /* key -> node */
    uint32_t *map;
    uint64_t map_mask;
    Node *pool;
    uint32_t pool_cap, free_head;
    uint32_t head[64], tail[64];
    int64_t count[64];
    /* batch scratch */
    uint64_t *bk_key;           /* batch-local map: key -> winner (agg, pos) */
    int64_t *bk_val;
    uint64_t bk_mask;
    uint32_t *bk_used;
    int64_t bk_n;
    int64_t max_keys;
} Cache;

static uint64_t mix(uint64_t k) {
    k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull;
    k ^= k >> 27; k *= 0x94d049bb133111ebull;
    k ^= k >> 31;
    return k;
}

static uint32_t map_find(const Cache *c, uint64_t key) {
    uint64_t i = mix(key) & c->map_mask;
    while (1) {
        uint32_t n = c->map[i];
        if (n == EMPTY) return NIL;
        if (c->pool[n].key == key) return n;
        i = (i + 1) & c->map_mask;
    }
}

static void map_put(Cache *c, uint64_t key, uint32_t node) {
    uint64_t i = mix(key) & c->map_mask;
    while (c->map[i] != EMPTY) i = (i + 1) & c->map_mask;
    c->map[i] = node;
}

/* backward-shift deletion keeps the probe sequences intact without tombstones */
static void map_del(Cache *c, uint64_t key) {
    uint64_t i = mix(key) & c->map_mask;
    while (c->pool[c->map[i]].key != key) i = (i + 1) & c->map_mask;
    uint64_t j = i;
    while (1) {
        j = (j + 1) & c->map_mask;
        uint32_t n = c->map[j];
        if (n == EMPTY) break;
        uint64_t home = mix(c->pool[n].key) & c->map_mask;
        /* can the entry at j move to i?  yes unless its home lies cyclically in (i, j] */
        int stay = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
        if (!stay) {
            c->map[i] = n;
            i = j;
        }
    }
    c->map[i] = EMPTY;
}

static void list_remove(Cache *c, uint32_t n) {
    Node *nd = &c->pool[n];
    int b = nd->bucket;
    if (nd->prev != NIL) c->pool[nd->prev].next = nd->next; else c->head[b] = nd->next;
    if (nd->next != NIL) c->pool[nd->next].prev = nd->prev; else c->tail[b] = nd->prev;
    c->count[b]--;
}

static void list_append(Cache *c, uint32_t n, int b) {
    Node *nd = &c->pool[n];
    nd->bucket = b;
    nd->next = NIL;
    nd->prev = c->tail[b];
    if (c->tail[b] != NIL) c->pool[c->tail[b]].next = n; else c->head[b] = n;
    c->tail[b] = n;
    c->count[b]++;
}

static void drop(Cache *c, uint32_t n) {           /* remove an entry altogether */
    list_remove(c, n);
    map_del(c, c->pool[n].key);
    c->pool[n].next = c->free_head;
    c->free_head = n;
    c->size--;
}

void *evb_create(int64_t capacity, int n_tables, int64_t max_keys_per_batch, double flush_rate, double perfect_item_cap) {
    Cache *c = (Cache *)calloc(1, sizeof(Cache));
    c->T = n_tables;
    c->cap = capacity;
    c->max_perfect = (int64_t)((double)capacity * perfect_item_cap);        /* EvLFU_C1.py:29 */
    c->flush_n = (int64_t)(flush_rate * (double)capacity) + 1;              /* :39 */
    c->max_keys = max_keys_per_batch;
    uint64_t need = (uint64_t)(capacity + max_keys_per_batch) * 2 + 16, m = 1;
    while (m < need) m <<= 1;
    c->map_mask = m - 1;
    c->map = (uint32_t *)malloc(m * sizeof(uint32_t));
    memset(c->map, 0xFF, m * sizeof(uint32_t));
    c->pool_cap = (uint32_t)(capacity + max_keys_per_batch + 8);
    c->pool = (Node *)malloc((size_t)c->pool_cap * sizeof(Node));
    for (uint32_t i = 0; i < c->pool_cap; ++i) c->pool[i].next = (i + 1 < c->pool_cap) ? i + 1 : NIL;
    c->free_head = 0;
    for (int b = 0; b < 64; ++b) c->head[b] = c->tail[b] = NIL;
    need = (uint64_t)max_keys_per_batch * 2 + 16;
    m = 1;
    while (m < need) m <<= 1;
    c->bk_mask = m - 1;
    c->bk_key = (uint64_t *)malloc(m * sizeof(uint64_t));
    c->bk_val = (int64_t *)malloc(m * sizeof(int64_t));
    memset(c->bk_key, 0xFF, m * sizeof(uint64_t));
    c->bk_used = (uint32_t *)malloc((size_t)max_keys_per_batch * sizeof(uint32_t));
    return c;
}

void evb_destroy(void *h) {
    Cache *c = (Cache *)h;
    free(c->map); free(c->pool); free(c->bk_key); free(c->bk_val); free(c->bk_used); free(c);
}

typedef struct { int64_t pos; uint64_t key; int agg; } Winner;
static int by_pos(const void *a, const void *b) {
    int64_t x = ((const Winner *)a)->pos, y = ((const Winner *)b)->pos;
    return (x > y) - (x < y);
}

/* One batch (oracle/evlfu.py: BatchEvLFU.lookup_batch without the approximate substitution).
 *   idx [Tl][B] int64 (the reference's lS_i), table_ids [Tl] global ids, agg_in [B] or NULL
 *   hit [B][Tl] uint8 out, agg_out [B] int32 out
 *   evicted / flushed: key streams of this batch (capacity = max_keys / flush_n + 1), counts in n_out[0..2] = evicted,
 *   flushed, inserted.  Returns 0, or -1 if a scratch array would overflow. */
int evb_lookup_batch(void *h, const int64_t *idx, int Tl, int64_t B, const int32_t *table_ids, const int32_t *agg_in,
                     uint8_t *hit, int32_t *agg_out, int64_t *evicted, int64_t *flushed, int64_t *n_out) {
    Cache *c = (Cache *)h;
    const int T = c->T;
    if (B * Tl > c->max_keys) return -1;
    /* 1. probe against the state at batch start (EvLFU_C1.py:113-120) */
    for (int64_t s = 0; s < B; ++s) {
        int a = 0;
        for (int t = 0; t < Tl; ++t) {
            uint64_t key = ((uint64_t)table_ids[t] << 40) | (uint64_t)idx[(int64_t)t * B + s];
            int hh = map_find(c, key) != NIL;
            hit[s * Tl + t] = (uint8_t)hh;
            a += hh;
        }
        agg_out[s] = agg_in ? agg_in[s] : a;
    }
    /* 2. per key the winning occurrence = max (agg, position); a hit whose bucket is already >= agg asks for nothing */
    c->bk_n = 0;
    int any_perfect = 0;
    for (int64_t s = 0; s < B; ++s) {
        const int a = agg_out[s];
        if (a == T) any_perfect = 1;
        for (int t = 0; t < Tl; ++t) {
            const int64_t p = s * Tl + t;
            uint64_t key = ((uint64_t)table_ids[t] << 40) | (uint64_t)idx[(int64_t)t * B + s];
            if (hit[p]) {
                uint32_t n = map_find(c, key);
                if (c->pool[n].bucket >= a) continue;
            }
            const int64_t val = ((int64_t)a << 40) | p;            /* (agg, pos) ordered lexicographically */
            uint64_t i = mix(key) & c->bk_mask;
            while (c->bk_key[i] != ~0ull && c->bk_key[i] != key) i = (i + 1) & c->bk_mask;
            if (c->bk_key[i] == ~0ull) {
                c->bk_key[i] = key;
                c->bk_val[i] = val;
                c->bk_used[c->bk_n++] = (uint32_t)i;
            } else if (val > c->bk_val[i]) {
                c->bk_val[i] = val;
            }
        }
    }
    /* 3. winners in position order: promote (remove + append) or insert (evlfu.py: sorted(winners, key=pos)) */
    Winner *w = (Winner *)malloc((size_t)(c->bk_n + 1) * sizeof(Winner));
    for (int64_t k = 0; k < c->bk_n; ++k) {
        uint32_t i = c->bk_used[k];
        w[k].key = c->bk_key[i];
        w[k].agg = (int)(c->bk_val[i] >> 40);
        w[k].pos = c->bk_val[i] & ((1ll << 40) - 1);
        c->bk_key[i] = ~0ull;
    }
    qsort(w, (size_t)c->bk_n, sizeof(Winner), by_pos);
    int64_t inserted = 0;
    uint64_t prot = ~0ull;
    int64_t prot_rank = -1;
    for (int64_t k = 0; k < c->bk_n; ++k) {
        uint32_t n = map_find(c, w[k].key);
        if (n != NIL) {
            list_remove(c, n);
        } else {
            if (c->free_head == NIL) { free(w); return -1; }
            n = c->free_head;
            c->free_head = c->pool[n].next;
            c->pool[n].key = w[k].key;
            map_put(c, w[k].key, n);
            c->size++;
            inserted++;
            const int64_t rank = ((int64_t)w[k].agg << 40) | w[k].pos;
            if (rank > prot_rank) { prot_rank = rank; prot = w[k].key; }
        }
        list_append(c, n, w[k].agg);
    }
    free(w);
    /* 4. flush rule (EvLFU_C1.py:36-44): the oldest flush_n keys of bucket T go */
    int64_t n_fl = 0;
    if (inserted > 0 && c->n_perfect >= c->max_perfect) {
        int64_t want = c->flush_n < c->count[T] ? c->flush_n : c->count[T];
        for (int64_t k = 0; k < want; ++k) {
            uint32_t n = c->head[T];
            flushed[n_fl++] = (int64_t)c->pool[n].key;
            drop(c, n);
        }
        c->n_perfect = c->count[T];
    }
    /* 5. evict back down to capacity in (bucket, FIFO) order, never the highest-ranked new key */
    int64_t n_ev = 0;
    int64_t need = c->size - c->cap;
    for (int b = 0; need > 0 && b <= T; ++b) {
        uint32_t n = c->head[b];
        while (n != NIL && need > 0) {
            uint32_t nx = c->pool[n].next;
            if (c->pool[n].key != prot) {
                evicted[n_ev++] = (int64_t)c->pool[n].key;
                drop(c, n);
                need--;
            }
            n = nx;
        }
    }
    /* 6. EvLFU_C1.py:163-165 */
    if (any_perfect) c->n_perfect = c->count[T];
    n_out[0] = n_ev;
    n_out[1] = n_fl;
    n_out[2] = inserted;
    return 0;
}

int64_t evb_size(void *h) { return ((Cache *)h)->size; }
int64_t evb_n_perfect(void *h) { return ((Cache *)h)->n_perfect; }

/* resident keys per bucket in FIFO order; off[T + 2] */
void evb_state(void *h, int64_t *keys, int64_t *off) {
    Cache *c = (Cache *)h;
    int64_t k = 0;
    for (int b = 0; b <= c->T; ++b) {
        off[b] = k;
        for (uint32_t n = c->head[b]; n != NIL; n = c->pool[n].next) keys[k++] = (int64_t)c->pool[n].key;
    }
    off[c->T + 1] = k;
}
