/* Oracle matcher (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).
 *
 * Restates what the reference obtains from the third-party, un-vendored
 * py-bsds500 `correspond_pixels.correspond_pixels(bmap1, bmap2, max_dist)`
 * (call sites: eval_depth_edges.py:50-52 and :130-132): a one-to-one
 * assignment between boundary pixels of two maps, pairs allowed iff their
 * Euclidean distance is <= max_dist * diagonal, of MAXIMUM CARDINALITY (the
 * BSDS min-cost formulation with outlier cost 100x the radius).  The callers
 * only count matched pixels (eval_depth_edges.py:133-143), and that count is
 * unique for any maximum matching, so Hopcroft-Karp is a sufficient oracle.
 *
 * PARITY UNPINNED: py-bsds500 is absent from /root/reference and this image;
 * cross-checked against scipy.sparse.csgraph.maximum_bipartite_matching.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int h, w, noff;
    int *offy, *offx;
    const uint8_t *a, *b;   /* a: side-1 map (pred), b: side-2 map (gt) */
    int32_t *idx_b;         /* pixel -> compact index on side 2, -1 if none */
    int32_t *pix_a, *pix_b; /* compact index -> pixel */
    int na, nb;
    int32_t *mate_a, *mate_b, *dist, *queue, *iter;
} hk_t;

static int hk_bfs(hk_t *g) {
    int head = 0, tail = 0, found = 0;
    for (int i = 0; i < g->na; i++) {
        if (g->mate_a[i] < 0) { g->dist[i] = 0; g->queue[tail++] = i; }
        else g->dist[i] = -1;
    }
    while (head < tail) {
        int u = g->queue[head++];
        int py = g->pix_a[u] / g->w, px = g->pix_a[u] % g->w;
        for (int k = 0; k < g->noff; k++) {
            int y = py + g->offy[k], x = px + g->offx[k];
            if (y < 0 || y >= g->h || x < 0 || x >= g->w) continue;
            int v = g->idx_b[y * g->w + x];
            if (v < 0) continue;
            int u2 = g->mate_b[v];
            if (u2 < 0) found = 1;
            else if (g->dist[u2] < 0) { g->dist[u2] = g->dist[u] + 1; g->queue[tail++] = u2; }
        }
    }
    return found;
}

/* iterative DFS along the BFS layering */
static int hk_dfs(hk_t *g, int root, int32_t *stack_u, int32_t *stack_v) {
    int sp = 0;
    stack_u[0] = root;
    while (sp >= 0) {
        int u = stack_u[sp];
        int py = g->pix_a[u] / g->w, px = g->pix_a[u] % g->w;
        int advanced = 0;
        while (g->iter[u] < g->noff) {
            int k = g->iter[u]++;
            int y = py + g->offy[k], x = px + g->offx[k];
            if (y < 0 || y >= g->h || x < 0 || x >= g->w) continue;
            int v = g->idx_b[y * g->w + x];
            if (v < 0) continue;
            int u2 = g->mate_b[v];
            if (u2 < 0) {
                /* augment along the stack */
                stack_v[sp] = v;
                for (int s = sp; s >= 0; s--) {
                    g->mate_a[stack_u[s]] = stack_v[s];
                    g->mate_b[stack_v[s]] = stack_u[s];
                }
                return 1;
            }
            if (g->dist[u2] == g->dist[u] + 1) {
                stack_v[sp] = v;
                stack_u[++sp] = u2;
                advanced = 1;
                break;
            }
        }
        if (!advanced) { g->dist[u] = -1; sp--; }
    }
    return 0;
}

/* Returns the matching size; match_a / match_b (may be NULL) get 1 at matched
 * pixels of each map.  radius is in pixels (already multiplied by the diagonal). */
int64_t oracle_correspond_pixels(const uint8_t *a, const uint8_t *b, int h, int w,
                                 double radius, uint8_t *match_a, uint8_t *match_b) {
    hk_t g; memset(&g, 0, sizeof g);
    g.h = h; g.w = w; g.a = a; g.b = b;
    int r = (int)radius; if (r < 0) r = 0;
    int side = 2 * r + 1;
    g.offy = malloc(sizeof(int) * side * side); g.offx = malloc(sizeof(int) * side * side);
    double r2 = radius * radius;
    /* nearest offsets first so the greedy start is good */
    for (int d2 = 0; d2 <= 2 * r * r; d2++)
        for (int dy = -r; dy <= r; dy++) for (int dx = -r; dx <= r; dx++)
            if (dy * dy + dx * dx == d2 && (double)d2 <= r2) { g.offy[g.noff] = dy; g.offx[g.noff] = dx; g.noff++; }
    size_t n = (size_t)h * w;
    g.idx_b = malloc(sizeof(int32_t) * (n ? n : 1));
    g.pix_a = malloc(sizeof(int32_t) * (n ? n : 1)); g.pix_b = malloc(sizeof(int32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        g.idx_b[i] = -1;
        if (a[i]) g.pix_a[g.na++] = (int32_t)i;
        if (b[i]) { g.idx_b[i] = g.nb; g.pix_b[g.nb++] = (int32_t)i; }
    }
    int na = g.na ? g.na : 1, nb = g.nb ? g.nb : 1;
    g.mate_a = malloc(sizeof(int32_t) * na); g.mate_b = malloc(sizeof(int32_t) * nb);
    g.dist = malloc(sizeof(int32_t) * na); g.queue = malloc(sizeof(int32_t) * na);
    g.iter = malloc(sizeof(int32_t) * na);
    int32_t *su = malloc(sizeof(int32_t) * na), *sv = malloc(sizeof(int32_t) * na);
    memset(g.mate_a, 0xff, sizeof(int32_t) * na); memset(g.mate_b, 0xff, sizeof(int32_t) * nb);
    int64_t size = 0;
    while (hk_bfs(&g)) {
        memset(g.iter, 0, sizeof(int32_t) * na);
        int progressed = 0;
        for (int u = 0; u < g.na; u++)
            if (g.mate_a[u] < 0 && hk_dfs(&g, u, su, sv)) { size++; progressed = 1; }
        if (!progressed) break;
    }
    if (match_a) { memset(match_a, 0, n); for (int u = 0; u < g.na; u++) if (g.mate_a[u] >= 0) match_a[g.pix_a[u]] = 1; }
    if (match_b) { memset(match_b, 0, n); for (int v = 0; v < g.nb; v++) if (g.mate_b[v] >= 0) match_b[g.pix_b[v]] = 1; }
    free(g.offy); free(g.offx); free(g.idx_b); free(g.pix_a); free(g.pix_b);
    free(g.mate_a); free(g.mate_b); free(g.dist); free(g.queue); free(g.iter); free(su); free(sv);
    return size;
}
