/* TEST INFRASTRUCTURE -- CPU oracle: the whole per-scan path chained the way the reference's CPU
 * processes run it (multiScanRegistration -> laserMapping / laserOdometry -> D-opt filter), used as the
 * reported CPU baseline (bench.py cpu_baseline / --impl reference) and by parity tests.  The map's
 * kd-trees are built once and reused, which favours the CPU side (upstream rebuilds the sub-map tree on
 * every mapping call).  Frames are independent, so the multi-threaded variant simply runs one frame per
 * thread (pthreads).
 */
#include "vlo_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void orc_mapping_register_trees(const orc_config *c,
                                const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                                const orc_pt *corner_map, int n_cm, const orc_kdtree *kc,
                                const orc_pt *surf_map, int n_sm, const orc_kdtree *ks,
                                const float *seed, orc_reg_result *res);

typedef struct {
    orc_config cfg;
    const orc_pt *corner_map, *surf_map; int n_cm, n_sm;
    orc_kdtree *kc, *ks;
} orc_map;

orc_map *orc_map_create(const orc_config *c, const orc_pt *corner_map, int n_cm, const orc_pt *surf_map, int n_sm)
{
    orc_map *m = (orc_map *)calloc(1, sizeof(orc_map));
    m->cfg = *c; m->corner_map = corner_map; m->surf_map = surf_map; m->n_cm = n_cm; m->n_sm = n_sm;
    m->kc = orc_kdtree_build(corner_map, n_cm);
    m->ks = orc_kdtree_build(surf_map, n_sm);
    return m;
}

void orc_map_free(orc_map *m)
{
    if (!m) return;
    orc_kdtree_free(m->kc); orc_kdtree_free(m->ks); free(m);
}

/* one frame: organise + extract + scan-to-map registration (+ D-opt gate inside the result) */
void orc_frame_scan_to_map(const orc_map *m, const float *raw, int n, int stride, const float *seed, orc_reg_result *res,
                           orc_feature_counts *counts)
{
    const orc_config *c = &m->cfg;
    int R = c->n_rings;
    int cap = n > 0 ? n : 1;
    orc_pt *cloud = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)cap);
    orc_pt *lflat = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)cap);
    orc_pt *lsharp = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)cap);
    int8_t *label = (int8_t *)malloc((size_t)cap);
    float *curv = (float *)malloc(sizeof(float) * (size_t)cap);
    int *sharp = (int *)malloc(sizeof(int) * (size_t)cap), *ls = (int *)malloc(sizeof(int) * (size_t)cap), *fl = (int *)malloc(sizeof(int) * (size_t)cap);
    int *rs = (int *)malloc(sizeof(int) * (size_t)(R + 1)), *lsr = (int *)malloc(sizeof(int) * (size_t)(R + 1)), *lfr = (int *)malloc(sizeof(int) * (size_t)(R + 1));
    orc_feature_counts fc;
    orc_organise(c, raw, n, stride, cloud, rs, NULL);
    orc_extract(c, cloud, rs, label, curv, NULL, sharp, ls, fl, lflat, lsr, lfr, &fc);
    for (int i = 0; i < fc.n_less_sharp; i++) lsharp[i] = cloud[ls[i]];
    /* the stacks LaserMapping optimises with are VoxelGrid-filtered (cornerFilterSize / surfaceFilterSize); cloud[] is
     * free to be reused as the output buffers now */
    orc_pt *cds = cloud, *sds = cloud + fc.n_less_sharp;
    int ncd = orc_voxel_downsample(lsharp, fc.n_less_sharp, c->corner_filter_size, cds);
    int nsd = orc_voxel_downsample(lflat, fc.n_less_flat, c->surface_filter_size, sds);
    orc_mapping_register_trees(c, cds, ncd, sds, nsd, m->corner_map, m->n_cm, m->kc,
                               m->surf_map, m->n_sm, m->ks, seed, res);
    if (counts) *counts = fc;
    free(cloud); free(lflat); free(lsharp); free(label); free(curv); free(sharp); free(ls); free(fl); free(rs); free(lsr); free(lfr);
}

typedef struct {
    const orc_map *m; const float *raw; const int *offsets; int stride; const float *seeds; orc_reg_result *res;
    int first, step, n;
} frame_job;

static void *frame_worker(void *arg)
{
    frame_job *j = (frame_job *)arg;
    for (int k = j->first; k < j->n; k += j->step)
        orc_frame_scan_to_map(j->m, j->raw + (size_t)j->offsets[k] * j->stride, j->offsets[k + 1] - j->offsets[k], j->stride,
                              j->seeds + 6 * k, &j->res[k], NULL);
    return NULL;
}

/* raw: concatenated clouds; offsets[n+1] in points; frames are interleaved over n_threads threads */
void orc_batch_scan_to_map(const orc_map *m, const float *raw, const int *offsets, int n, int stride, const float *seeds,
                           orc_reg_result *res, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; frame_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (frame_job){ m, raw, offsets, stride, seeds, res, t, n_threads, n };
        if (n_threads == 1) frame_worker(&jobs[t]); else pthread_create(&th[t], NULL, frame_worker, &jobs[t]);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}
