/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h; parity unpinned).
 *
 * Exact k-nearest-neighbour search standing in for pcl::KdTreeFLANN::nearestKSearch as used by
 * the LOAM fork.  Contract frozen by BASELINE.json north_star:
 * neighbours ordered by (squared distance, index) lexicographically, i.e. ties -> lowest index.
 * d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in float32 without contraction (FLANN L2_Simple order).
 * Two implementations: brute force (the definition) and a kd-tree (what the CPU baseline times).
 */
#include "vlo_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

static inline float d2f(const orc_pt *a, const orc_pt *b)
{
    float dx = a->x - b->x, dy = a->y - b->y, dz = a->z - b->z;
    return (dx * dx + dy * dy) + dz * dz;
}

/* insert (d,i) into ascending list of length k (lexicographic) */
static inline void topk_insert(float *bd, int *bi, int k, float d, int i)
{
    if (!(d < bd[k - 1] || (d == bd[k - 1] && i < bi[k - 1]))) return;
    int j = k - 1;
    while (j > 0 && (d < bd[j - 1] || (d == bd[j - 1] && i < bi[j - 1]))) {
        bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; j--;
    }
    bd[j] = d; bi[j] = i;
}

void orc_knn_brute(const orc_pt *cloud, int n, const orc_pt *q, int nq, int k, int *idx, float *d2)
{
    for (int qi = 0; qi < nq; qi++) {
        float *bd = d2 + (size_t)qi * k; int *bi = idx + (size_t)qi * k;
        for (int j = 0; j < k; j++) { bd[j] = INFINITY; bi[j] = -1; }
        for (int i = 0; i < n; i++) topk_insert(bd, bi, k, d2f(&cloud[i], &q[qi]), i);
    }
}

/* ---------------- kd-tree ---------------- */
typedef struct {
    float lo[3], hi[3];
    int left, right;     /* children, -1 for leaf */
    int begin, end;      /* range in perm */
} kd_node;

struct orc_kdtree {
    const orc_pt *cloud;
    int n;
    int *perm;
    kd_node *nodes;
    int n_nodes, cap;
};

#define KD_LEAF 10

static float coord(const orc_pt *p, int a) { return a == 0 ? p->x : (a == 1 ? p->y : p->z); }

static int kd_new_node(orc_kdtree *t)
{
    if (t->n_nodes == t->cap) { t->cap *= 2; t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap); }
    return t->n_nodes++;
}

/* nth_element on perm[b,e) by axis a */
static void kd_select(orc_kdtree *t, int b, int e, int nth, int a)
{
    int *p = t->perm;
    while (e - b > 1) {
        float pv = coord(&t->cloud[p[(b + e) / 2]], a);
        int i = b, j = e - 1;
        while (i <= j) {
            while (coord(&t->cloud[p[i]], a) < pv) i++;
            while (coord(&t->cloud[p[j]], a) > pv) j--;
            if (i <= j) { int tmp = p[i]; p[i] = p[j]; p[j] = tmp; i++; j--; }
        }
        if (nth <= j) e = j + 1;
        else if (nth >= i) b = i;
        else return;
    }
}

static int kd_build_rec(orc_kdtree *t, int b, int e)
{
    int id = kd_new_node(t);
    kd_node nd;
    for (int a = 0; a < 3; a++) { nd.lo[a] = FLT_MAX; nd.hi[a] = -FLT_MAX; }
    for (int i = b; i < e; i++)
        for (int a = 0; a < 3; a++) {
            float v = coord(&t->cloud[t->perm[i]], a);
            if (v < nd.lo[a]) nd.lo[a] = v;
            if (v > nd.hi[a]) nd.hi[a] = v;
        }
    nd.begin = b; nd.end = e; nd.left = nd.right = -1;
    if (e - b > KD_LEAF) {
        int ax = 0; float best = nd.hi[0] - nd.lo[0];
        for (int a = 1; a < 3; a++) if (nd.hi[a] - nd.lo[a] > best) { best = nd.hi[a] - nd.lo[a]; ax = a; }
        if (best > 0.0f) {
            int mid = (b + e) / 2;
            kd_select(t, b, e, mid, ax);
            int l = kd_build_rec(t, b, mid);
            int r = kd_build_rec(t, mid, e);
            nd.left = l; nd.right = r;
        }
    }
    t->nodes[id] = nd;
    return id;
}

orc_kdtree *orc_kdtree_build(const orc_pt *cloud, int n)
{
    orc_kdtree *t = (orc_kdtree *)calloc(1, sizeof(orc_kdtree));
    t->cloud = cloud; t->n = n;
    t->perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) t->perm[i] = i;
    t->cap = 64; t->nodes = (kd_node *)malloc(sizeof(kd_node) * (size_t)t->cap);
    if (n > 0) kd_build_rec(t, 0, n);
    return t;
}

void orc_kdtree_free(orc_kdtree *t)
{
    if (!t) return;
    free(t->perm); free(t->nodes); free(t);
}

static inline float box_d2(const kd_node *nd, const orc_pt *q)
{
    float s = 0.0f;
    float v[3] = { q->x, q->y, q->z };
    for (int a = 0; a < 3; a++) {
        float g = 0.0f;
        if (v[a] < nd->lo[a]) g = nd->lo[a] - v[a];
        else if (v[a] > nd->hi[a]) g = v[a] - nd->hi[a];
        s += g * g;
    }
    return s;
}

static void kd_search(const orc_kdtree *t, int id, const orc_pt *q, int k, float *bd, int *bi)
{
    const kd_node *nd = &t->nodes[id];
    /* conservative prune: a rounding slack keeps equal-distance candidates reachable */
    if (box_d2(nd, q) > bd[k - 1] * 1.00001f + 1e-12f) return;
    if (nd->left < 0) {
        for (int i = nd->begin; i < nd->end; i++) {
            int pi = t->perm[i];
            topk_insert(bd, bi, k, d2f(&t->cloud[pi], q), pi);
        }
        return;
    }
    float dl = box_d2(&t->nodes[nd->left], q), dr = box_d2(&t->nodes[nd->right], q);
    if (dl <= dr) { kd_search(t, nd->left, q, k, bd, bi); kd_search(t, nd->right, q, k, bd, bi); }
    else          { kd_search(t, nd->right, q, k, bd, bi); kd_search(t, nd->left, q, k, bd, bi); }
}

void orc_kdtree_knn(const orc_kdtree *t, const orc_pt *q, int nq, int k, int *idx, float *d2)
{
    for (int qi = 0; qi < nq; qi++) {
        float *bd = d2 + (size_t)qi * k; int *bi = idx + (size_t)qi * k;
        for (int j = 0; j < k; j++) { bd[j] = INFINITY; bi[j] = -1; }
        if (t->n > 0) kd_search(t, 0, &q[qi], k, bd, bi);
    }
}
