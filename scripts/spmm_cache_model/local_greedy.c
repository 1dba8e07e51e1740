// Local refinement of a banded (Cuthill-McKee) cell order: inside every block of B consecutive rows the rows
// are re-ordered greedily so that the next row is the unplaced row with the most edges into the last W
// placed rows.  Blocks keep their position (the L2 band structure is untouched); rows keep their edge order.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char **argv) {
    int N = atoi(argv[1]), B = atoi(argv[2]), W = atoi(argv[3]);
    char fn[256]; sprintf(fn, "/tmp/sim/ptr_%d.bin", N); FILE *f = fopen(fn, "rb"); int *ptr = malloc(4 * (size_t)(N + 1)); fread(ptr, 4, N + 1, f); fclose(f);
    int nnz = ptr[N]; sprintf(fn, "/tmp/sim/idx_%d.bin", N); f = fopen(fn, "rb"); int *idx = malloc(4 * (size_t)nnz); fread(idx, 4, nnz, f); fclose(f);
    int *order = malloc(4 * (size_t)N), *score = calloc(N, 4); char *placed = calloc(N, 1);
    int pos = 0;
    for (int b0 = 0; b0 < N; b0 += B) {
        int b1 = b0 + B < N ? b0 + B : N, nextfree = b0, start = pos;
        while (pos - start < b1 - b0) {
            int best = -1, bs = 0;
            int w0 = pos - W > start ? pos - W : start;
            for (int p = w0; p < pos; ++p) { int u = order[p]; for (int e = ptr[u]; e < ptr[u + 1]; ++e) { int v = idx[e]; if (v >= b0 && v < b1 && !placed[v] && (score[v] > bs || (score[v] == bs && best >= 0 && v < best))) { best = v; bs = score[v]; } } }
            if (best < 0) { while (placed[nextfree]) ++nextfree; best = nextfree; }
            placed[best] = 1; order[pos++] = best;
            for (int e = ptr[best]; e < ptr[best + 1]; ++e) { int v = idx[e]; if (v >= b0 && v < b1) score[v]++; }
            if (pos - start > W) { int u = order[pos - W - 1]; for (int e = ptr[u]; e < ptr[u + 1]; ++e) { int v = idx[e]; if (v >= b0 && v < b1) score[v]--; } }
        }
        for (int p = start; p < pos; ++p) { int u = order[p]; for (int e = ptr[u]; e < ptr[u + 1]; ++e) { int v = idx[e]; if (v >= b0 && v < b1) score[v] = 0; } }
    }
    int *inv = malloc(4 * (size_t)N); for (int i = 0; i < N; ++i) inv[order[i]] = i;
    int *nptr = malloc(4 * (size_t)(N + 1)), *nidx = malloc(4 * (size_t)nnz); nptr[0] = 0;
    for (int i = 0; i < N; ++i) { int u = order[i]; int d = ptr[u + 1] - ptr[u]; for (int k = 0; k < d; ++k) nidx[nptr[i] + k] = inv[idx[ptr[u] + k]]; nptr[i + 1] = nptr[i] + d; }
    int N2 = N + 1;  // written under the pseudo-size N+1 so that lru_model can read it
    sprintf(fn, "/tmp/sim/ptr_%d.bin", N2); f = fopen(fn, "wb"); fwrite(nptr, 4, N + 1, f); fclose(f);
    sprintf(fn, "/tmp/sim/idx_%d.bin", N2); f = fopen(fn, "wb"); fwrite(nidx, 4, nnz, f); fclose(f);
    // overlap of a row's neighbour set with the union of the previous 1 / 7 rows
    double o1 = 0, o7 = 0; int cnt = 0; char *mark = calloc(N, 1);
    for (int r = 8; r < N; r += 97) {
        for (int back = 1; back <= 7; ++back) for (int e = nptr[r - back]; e < nptr[r - back + 1]; ++e) mark[nidx[e]] |= back == 1 ? 3 : 2;
        int h1 = 0, h7 = 0, d = nptr[r + 1] - nptr[r]; for (int e = nptr[r]; e < nptr[r + 1]; ++e) { h1 += mark[nidx[e]] & 1; h7 += (mark[nidx[e]] >> 1) & 1; }
        for (int back = 1; back <= 7; ++back) for (int e = nptr[r - back]; e < nptr[r - back + 1]; ++e) mark[nidx[e]] = 0;
        if (d) { o1 += (double)h1 / d; o7 += (double)h7 / d; ++cnt; }
    }
    printf("B %d W %d: overlap with previous row %.3f, with previous 7 rows %.3f\n", B, W, o1 / cnt, o7 / cnt);
    return 0;
}
