// Trace-driven L1 / L2 hit-rate model of the warp-per-row SpMM: 148 SMs, each an LRU
// cache of `cap` state rows, over one shared LRU L2; warps of an SM advance round-robin, `quad` edges per turn.
// An odd N reads the files written by local_greedy.c (the refined order of the graph of N - 1 rows).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define NSM 148
typedef struct { int *key; int *prev, *next; int *slot_of; int head, tail, size, cap, hsize, tomb; int *htab; } lru_t;
static void lru_init(lru_t *c, int cap) {
    c->cap = cap; c->size = 0; c->tomb = 0; c->head = c->tail = -1; c->hsize = 1; while (c->hsize < cap * 4) c->hsize <<= 1;
    c->key = malloc(sizeof(int) * cap); c->prev = malloc(sizeof(int) * cap); c->next = malloc(sizeof(int) * cap);
    c->htab = malloc(sizeof(int) * c->hsize); memset(c->htab, -1, sizeof(int) * c->hsize);
}
static inline unsigned hsh(int k) { return (unsigned)k * 2654435761u; }
static int h_find(lru_t *c, int k) { unsigned m = c->hsize - 1, i = hsh(k) & m; while (c->htab[i] != -1) { if (c->htab[i] >= 0 && c->key[c->htab[i]] == k) return c->htab[i]; i = (i + 1) & m; } return -1; }
static void h_insert(lru_t *c, int node) { unsigned m = c->hsize - 1, i = hsh(c->key[node]) & m; while (c->htab[i] >= 0) i = (i + 1) & m; c->htab[i] = node; }
static void h_erase(lru_t *c, int node) { unsigned m = c->hsize - 1, i = hsh(c->key[node]) & m; while (c->htab[i] != node) i = (i + 1) & m; c->htab[i] = -2; }
static void unlink_node(lru_t *c, int n) { if (c->prev[n] >= 0) c->next[c->prev[n]] = c->next[n]; else c->head = c->next[n]; if (c->next[n] >= 0) c->prev[c->next[n]] = c->prev[n]; else c->tail = c->prev[n]; }
static void push_front(lru_t *c, int n) { c->prev[n] = -1; c->next[n] = c->head; if (c->head >= 0) c->prev[c->head] = n; c->head = n; if (c->tail < 0) c->tail = n; }

static int lru_access(lru_t *c, int k) {  // 1 = hit
    int n = h_find(c, k);
    if (n >= 0) { unlink_node(c, n); push_front(c, n); return 1; }
    if (c->size < c->cap) n = c->size++;
    else { n = c->tail; unlink_node(c, n); h_erase(c, n); if (++c->tomb > c->hsize / 4) { /* rebuild */ memset(c->htab, -1, sizeof(int) * c->hsize); c->tomb = 0; for (int i = 0; i < c->size; ++i) if (i != n) h_insert(c, i); } }
    c->key[n] = k; h_insert(c, n); push_front(c, n); return 0;
}
int main(int argc, char **argv) {
    int N = atoi(argv[1]); int NF = N; if (N % 2 == 1) N -= 1; int rows_per_cta = atoi(argv[2]); int ctas_per_sm = atoi(argv[3]); int cap = atoi(argv[4]);
    int contiguous = atoi(argv[5]);  // 0: global round-robin CTA queue; 1: each SM owns a contiguous range of rows
    int quad = argc > 6 ? atoi(argv[6]) : 4; int K = argc > 7 ? atoi(argv[7]) : 256; int l2cap = argc > 8 ? atoi(argv[8]) : 157000; lru_t l2; lru_init(&l2, l2cap); long l2hits = 0, l2acc = 0;
    char fn[256]; sprintf(fn, "/tmp/sim/ptr_%d.bin", NF); FILE *f = fopen(fn, "rb"); int *ptr = malloc(sizeof(int) * (N + 1)); fread(ptr, 4, N + 1, f); fclose(f);
    int nnz = ptr[N]; sprintf(fn, "/tmp/sim/idx_%d.bin", NF); f = fopen(fn, "rb"); int *idx = malloc(sizeof(int) * (size_t)nnz); fread(idx, 4, nnz, f); fclose(f);
    int n_cta = (N + rows_per_cta - 1) / rows_per_cta;
    lru_t *cache = malloc(sizeof(lru_t) * NSM); for (int s = 0; s < NSM; ++s) lru_init(&cache[s], cap);
    int nslot = NSM * ctas_per_sm; int *slot_cta = malloc(sizeof(int) * nslot); int *cur = malloc(sizeof(int) * nslot * rows_per_cta);
    int next_cta = 0; int per_sm = (n_cta + NSM - 1) / NSM; int *sm_next = malloc(sizeof(int) * NSM); for (int s = 0; s < NSM; ++s) sm_next[s] = s * per_sm;
    long hits = 0, acc = 0; int live = 0;
    for (int i = 0; i < nslot; ++i) slot_cta[i] = -1;
    if (contiguous == 2) {
        int W = rows_per_cta * ctas_per_sm; long *t_sm = calloc(NSM, sizeof(long)); int *wrow = malloc(sizeof(int) * NSM * W); int *wcur = malloc(sizeof(int) * NSM * W);
        for (int i = 0; i < NSM * W; ++i) wrow[i] = -1;
        for (;;) { int progressed = 0;
            for (int s = 0; s < NSM; ++s) for (int w = 0; w < W; ++w) { int i = s * W + w;
                if (wrow[i] < 0) { long t = t_sm[s]; long blk = t / K, within = t % K; long r = blk * (long)NSM * K + (long)s * K + within; if (r >= N) { if (blk * (long)NSM * K >= N) continue; t_sm[s]++; progressed = 1; continue; } t_sm[s]++; wrow[i] = (int)r; wcur[i] = ptr[r]; }
                int r = wrow[i], e = wcur[i], e1 = ptr[r + 1]; int lim = e + quad < e1 ? e + quad : e1;
                for (; e < lim; ++e) { int h = lru_access(&cache[s], idx[e]); hits += h; ++acc; if (!h) { l2hits += lru_access(&l2, idx[e]); ++l2acc; } }
                wcur[i] = e; if (e >= e1) wrow[i] = -1; progressed = 1; }
            if (!progressed) break; }
        printf("L2 hit rate %.3f dram rows %ld | superblock K %d warps/sm %d cap %d: accesses %ld L1 hit rate %.3f\n", (double)l2hits / l2acc, l2acc - l2hits, K, W, cap, acc, (double)hits / acc);
        return 0;
    }
    for (;;) {
        int progressed = 0;
        for (int s = 0; s < NSM; ++s) for (int k = 0; k < ctas_per_sm; ++k) {
            int sl = s * ctas_per_sm + k;
            if (slot_cta[sl] < 0) {
                int c = -1;
                if (contiguous) { if (sm_next[s] < (s + 1) * per_sm && sm_next[s] < n_cta) c = sm_next[s]++; }
                else if (next_cta < n_cta) c = next_cta++;
                if (c < 0) continue;
                slot_cta[sl] = c; for (int w = 0; w < rows_per_cta; ++w) { int r = c * rows_per_cta + w; cur[sl * rows_per_cta + w] = r < N ? ptr[r] : -1; }
            }
            int c = slot_cta[sl], done = 1;
            for (int w = 0; w < rows_per_cta; ++w) {
                int r = c * rows_per_cta + w; if (r >= N) continue; int e = cur[sl * rows_per_cta + w]; int e1 = ptr[r + 1];
                if (e >= e1) continue;
                int lim = e + quad < e1 ? e + quad : e1;
                for (; e < lim; ++e) { int h = lru_access(&cache[s], idx[e]); hits += h; ++acc; if (!h) { l2hits += lru_access(&l2, idx[e]); ++l2acc; } }
                cur[sl * rows_per_cta + w] = e; if (e < e1) done = 0; progressed = 1;
            }
            if (done) slot_cta[sl] = -1;
        }
        if (!progressed) break;
    }
    printf("L2 hit rate %.3f  dram rows %ld | ", (double)l2hits / l2acc, l2acc - l2hits); printf("rows/cta %d ctas/sm %d cap %d contiguous %d: accesses %ld hit rate %.3f\n", rows_per_cta, ctas_per_sm, cap, contiguous, acc, (double)hits / acc);
    return 0;
}
