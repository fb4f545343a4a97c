/* ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's CPU 1D NMS extension
 *   libs/nms/src/nms_cpu.cpp:20-63   (nms_1d_cpu:     greedy hard NMS)
 *   libs/nms/src/nms_cpu.cpp:72-172  (softnms_1d_cpu: selection-sort style soft-NMS)
 * plus the decode step of Evaluator._collect_segments (libs/worker_v2.py:1131-1187) and
 * segment voting (libs/nms/nms.py:64-103) for timing the CPU baseline without Python
 * loops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load this.
 *
 * Parity status: PINNED against the compiled, unmodified reference extension
 * (oracle/_ref/nms_1d_cpu_vg.so, built by oracle/build_ref.py) in
 * tests/test_nms_oracle.py, and against tests/golden/ fixtures produced by the reference.
 *
 * Float semantics: everything is IEEE float32 with no FMA contraction (build with
 * -ffp-contract=off); exp is glibc expf, the same function std::exp(float) resolves to in
 * the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* soft-NMS.  segs (n,2), scores (n) are not modified.  dets (n,3) and inds (n) are written
 * for the first `return value` rows.  max_iters < 0: run all outer steps (the reference);
 * max_iters >= 0: stop after that many outer steps (only the first max_num_segs rows are
 * ever consumed by libs/nms/nms.py:54-59).  method: 0 hard, 1 linear, 2 gaussian. */
int64_t oracle_softnms_1d(const float *segs, const float *scores, int64_t n_in,
                          float *dets, int64_t *inds_out, float iou_thresh, float sigma,
                          float min_score, int method, int64_t max_iters) {
    if (n_in <= 0) return 0;
    int64_t nsegs = n_in;
    float *x1 = (float *)malloc(sizeof(float) * n_in * 4);
    float *x2 = x1 + n_in, *sc = x2 + n_in, *areas = sc + n_in;
    int64_t *inds = (int64_t *)malloc(sizeof(int64_t) * n_in);
    for (int64_t i = 0; i < n_in; i++) {
        x1[i] = segs[2 * i];
        x2[i] = segs[2 * i + 1];
        sc[i] = scores[i];
        /* Tensor areas_t = x2_t - x1_t + 1e-6  (float tensor + double scalar -> float add) */
        areas[i] = (x2[i] - x1[i]) + 1e-6f;
        inds[i] = i;
    }
    int64_t i = 0;
    for (; i < nsegs; i++) {
        if (max_iters >= 0 && i >= max_iters) break;
        float max_score = sc[i];
        int64_t max_pos = i;
        for (int64_t pos = i + 1; pos < nsegs; pos++) {
            if (max_score < sc[pos]) { max_score = sc[pos]; max_pos = pos; }
        }
        float ix1 = dets[i * 3 + 0] = x1[max_pos];
        float ix2 = dets[i * 3 + 1] = x2[max_pos];
        float iscore = dets[i * 3 + 2] = sc[max_pos];
        float iarea = areas[max_pos];
        int64_t iind = inds[max_pos];
        x1[max_pos] = x1[i]; x2[max_pos] = x2[i]; sc[max_pos] = sc[i];
        areas[max_pos] = areas[i]; inds[max_pos] = inds[i];
        x1[i] = ix1; x2[i] = ix2; sc[i] = iscore; areas[i] = iarea; inds[i] = iind;

        int64_t pos = i + 1;
        while (pos < nsegs) {
            float xx1 = ix1 > x1[pos] ? ix1 : x1[pos];
            float xx2 = ix2 < x2[pos] ? ix2 : x2[pos];
            float d = xx2 - xx1;
            float inter = d > 0.f ? d : 0.f;
            float ovr = inter / (iarea + areas[pos] - inter);
            float weight = 1.f;
            if (method == 0) {
                if (ovr >= iou_thresh) weight = 0.f;
            } else if (method == 1) {
                if (ovr >= iou_thresh) weight = 1.f - ovr;
            } else if (method == 2) {
                weight = expf(-(ovr * ovr) / sigma);
            }
            sc[pos] *= weight;
            if (sc[pos] < min_score) {
                x1[pos] = x1[nsegs - 1]; x2[pos] = x2[nsegs - 1]; sc[pos] = sc[nsegs - 1];
                areas[pos] = areas[nsegs - 1]; inds[pos] = inds[nsegs - 1];
                nsegs--;
                pos--;
            }
            pos++;
        }
    }
    int64_t n_out = i < nsegs ? i : nsegs;
    if (max_iters < 0) n_out = nsegs;
    memcpy(inds_out, inds, sizeof(int64_t) * n_out);
    free(x1);
    free(inds);
    return n_out;
}

typedef struct { float s; int64_t i; } kv_t;
static int cmp_desc_stable(const void *a, const void *b) {
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* hard NMS: stable descending order (the reference's sort is unstable; ties are excluded
 * from reference comparisons).  Returns the number of kept indices written to keep. */
int64_t oracle_nms_1d(const float *segs, const float *scores, int64_t n, int64_t *keep,
                      float iou_thresh) {
    if (n <= 0) return 0;
    kv_t *ord = (kv_t *)malloc(sizeof(kv_t) * n);
    float *areas = (float *)malloc(sizeof(float) * n);
    char *sel = (char *)malloc(n);
    for (int64_t i = 0; i < n; i++) {
        ord[i].s = scores[i]; ord[i].i = i; sel[i] = 1;
        areas[i] = (segs[2 * i + 1] - segs[2 * i]) + 1e-6f;
    }
    qsort(ord, n, sizeof(kv_t), cmp_desc_stable);
    for (int64_t _i = 0; _i < n; _i++) {
        if (!sel[_i]) continue;
        int64_t i = ord[_i].i;
        float ix1 = segs[2 * i], ix2 = segs[2 * i + 1], iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; _j++) {
            if (!sel[_j]) continue;
            int64_t j = ord[_j].i;
            float xx1 = ix1 > segs[2 * j] ? ix1 : segs[2 * j];
            float xx2 = ix2 < segs[2 * j + 1] ? ix2 : segs[2 * j + 1];
            float d = xx2 - xx1;
            float inter = d > 0.f ? d : 0.f;
            float ovr = inter / (iarea + areas[j] - inter);
            if (ovr >= iou_thresh) sel[_j] = 0;
        }
    }
    int64_t k = 0;
    for (int64_t _i = 0; _i < n; _i++) if (sel[_i]) keep[k++] = ord[_i].i;
    free(ord); free(areas); free(sel);
    return k;
}

/* decode for one query over the level-major flat point list (libs/worker_v2.py:1131-1187):
 * scores (p) already = sigmoid(logit) * mask; offsets (p,2); coord/stride (p).
 * Writes up to topk (seg(2), score, flat idx) rows in descending-score order (ties: ascending
 * flat index) after the length filter.  Returns the row count. */
int64_t oracle_decode(const float *scores, const float *offsets, const float *coord,
                      const float *stride, int64_t p, float pre_nms_thresh, int64_t topk,
                      float seg_len_thresh, float *segs_out, float *scores_out,
                      int64_t *idx_out) {
    kv_t *c = (kv_t *)malloc(sizeof(kv_t) * (p > 0 ? p : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < p; i++)
        if (scores[i] > pre_nms_thresh) { c[m].s = scores[i]; c[m].i = i; m++; }
    qsort(c, m, sizeof(kv_t), cmp_desc_stable);
    if (m > topk) m = topk;
    int64_t k = 0;
    for (int64_t r = 0; r < m; r++) {
        int64_t i = c[r].i;
        float left = coord[i] - offsets[2 * i] * stride[i];
        float right = coord[i] + offsets[2 * i + 1] * stride[i];
        if (right - left > seg_len_thresh) {
            segs_out[2 * k] = left; segs_out[2 * k + 1] = right;
            scores_out[k] = c[r].s; idx_out[k] = i; k++;
        }
    }
    free(c);
    return k;
}
