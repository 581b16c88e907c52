// monopsr_b200/csrc/kitti_eval.cu -- KITTI object-detection AP evaluator (host code only; lives in a .cu so that the one
// build recipe covers it).  C ABI: include/monopsr_b200_eval.h.
//
// Restates the evaluator the reference shells out to after inference
// (scripts/offline_eval/kitti_native_eval/evaluate_object_3d_offline.cpp; line numbers below refer to it):
//   overlaps        imageBoxOverlap :225-262, toPolygon :267-290, groundBoxOverlap :292-313, box3DOverlap :315-345
//   clean_data      cleanData :383-459           (which ground truth / detections count, are ignored, or are "other")
//   statistics      computeStatistics :461-642   (greedy matching, TP / FP / FN, DontCare areas, orientation similarity)
//   thresholds      getThresholds :347-381       (scores at 41 equally spaced recall positions)
//   eval_class      eval_class :648-744
// boost::geometry is replaced by Sutherland-Hodgman clipping of one oriented rectangle by the other (both convex).
// Checked against the reference's own evaluator compiled from its source (oracle/build_ref.sh, oracle/boost_shim):
// tests/test_kitti_eval.py, tests/golden/kitti_eval_golden.json.
#include "../../include/monopsr_b200_eval.h"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <functional>
#include <vector>

#if defined(__GNUC__)
#define MPB_EVAL_API extern "C" __attribute__((visibility("default")))
#else
#define MPB_EVAL_API extern "C"
#endif

namespace mpb_eval {

enum { TYPE = 0, TRUNC = 1, OCC = 2, ALPHA = 3, X1 = 4, Y1 = 5, X2 = 6, Y2 = 7, H = 8, W = 9, L = 10, T1 = 11, T2 = 12,
       T3 = 13, RY = 14, SCORE = 15 };
const int kMinHeight[3] = {40, 25, 25};
const int kMaxOcclusion[3] = {0, 1, 2};
const double kMaxTruncation[3] = {0.15, 0.3, 0.5};
const double kNoDetection = -10000000;

struct Pt { double x, y; };

// bird's-eye-view rectangle of a box: corners (+-l/2, +-w/2) rotated by ry, moved to (t1, t3)
static void bev_corners(const double* b, Pt out[4]) {
    const double c = cos(b[RY]), s = sin(b[RY]);
    const double cx[4] = {b[L] / 2, b[L] / 2, -b[L] / 2, -b[L] / 2};
    const double cz[4] = {b[W] / 2, -b[W] / 2, -b[W] / 2, b[W] / 2};
    for (int i = 0; i < 4; i++) {
        out[i].x = c * cx[i] + s * cz[i] + b[T1];
        out[i].y = -s * cx[i] + c * cz[i] + b[T3];
    }
}
static double ring_area2(const Pt* r, int n) {       // twice the signed area
    double a = 0;
    for (int i = 0; i < n; i++) a += r[i].x * r[(i + 1) % n].y - r[(i + 1) % n].x * r[i].y;
    return a;
}
// area of the intersection of two convex quadrilaterals
static double convex_intersection_area(const Pt a[4], const Pt b[4]) {
    Pt cur[16], nxt[16];
    int n = 4;
    for (int i = 0; i < 4; i++) cur[i] = a[i];
    const double orient = ring_area2(b, 4) >= 0 ? 1.0 : -1.0;
    for (int e = 0; e < 4 && n > 0; e++) {
        const Pt c0 = b[e], c1 = b[(e + 1) % 4];
        const double ex = c1.x - c0.x, ey = c1.y - c0.y;
        int m = 0;
        for (int j = 0; j < n; j++) {
            const Pt p = cur[j], q = cur[(j + 1) % n];
            const double dp = orient * (ex * (p.y - c0.y) - ey * (p.x - c0.x));
            const double dq = orient * (ex * (q.y - c0.y) - ey * (q.x - c0.x));
            if (dp >= 0) nxt[m++] = p;
            if ((dp >= 0) != (dq >= 0)) {
                const double t = dp / (dp - dq);
                nxt[m].x = p.x + t * (q.x - p.x);
                nxt[m].y = p.y + t * (q.y - p.y);
                m++;
            }
        }
        n = m;
        for (int j = 0; j < n; j++) cur[j] = nxt[j];
    }
    return n >= 3 ? fabs(ring_area2(cur, n)) * 0.5 : 0.0;
}

static double image_overlap(const double* a, const double* b, int criterion) {
    const double w = std::min(a[X2], b[X2]) - std::max(a[X1], b[X1]);
    const double h = std::min(a[Y2], b[Y2]) - std::max(a[Y1], b[Y1]);
    if (w <= 0 || h <= 0) return 0;
    const double inter = w * h;
    const double aa = (a[X2] - a[X1]) * (a[Y2] - a[Y1]), ba = (b[X2] - b[X1]) * (b[Y2] - b[Y1]);
    if (criterion == -1) return inter / (aa + ba - inter);
    if (criterion == 0) return inter / aa;
    if (criterion == 1) return inter / ba;
    return -1;
}
static double ground_overlap(const double* d, const double* g, int criterion) {
    Pt dp[4], gp[4];
    bev_corners(d, dp);
    bev_corners(g, gp);
    const double inter = convex_intersection_area(gp, dp);
    const double da = fabs(ring_area2(dp, 4)) * 0.5, ga = fabs(ring_area2(gp, 4)) * 0.5;
    if (criterion == -1) return inter / (da + ga - inter);
    if (criterion == 0) return inter / da;
    return inter / ga;
}
static double box3d_overlap(const double* d, const double* g, int criterion) {
    Pt dp[4], gp[4];
    bev_corners(d, dp);
    bev_corners(g, gp);
    const double ymax = std::min(d[T2], g[T2]);
    const double ymin = std::max(d[T2] - d[H], g[T2] - g[H]);
    const double inter_vol = convex_intersection_area(gp, dp) * std::max(0.0, ymax - ymin);
    const double det_vol = d[H] * d[L] * d[W], gt_vol = g[H] * g[L] * g[W];
    if (criterion == -1) return inter_vol / (det_vol + gt_vol - inter_vol);
    if (criterion == 0) return inter_vol / det_vol;
    return inter_vol / gt_vol;
}
static double overlap(const double* d, const double* g, int metric, int criterion) {
    if (metric == MPB_KITTI_IMAGE) return image_overlap(d, g, criterion);
    if (metric == MPB_KITTI_GROUND) return ground_overlap(d, g, criterion);
    return box3d_overlap(d, g, criterion);
}

struct Image {
    const double* gt; int ngt;
    const double* det; int ndet;
    std::vector<int> ignored_gt, ignored_det;     // 0 = evaluated, 1 = ignored, -1 = other class
    std::vector<int> dontcare;                    // rows of gt
};

static void clean_data(Image& im, int cls, int difficulty, int& n_gt) {
    for (int i = 0; i < im.ngt; i++) {
        const double* g = im.gt + (size_t)i * MPB_KITTI_GT_COLS;
        const int type = (int)g[TYPE];
        const double height = g[Y2] - g[Y1];
        int valid_class;
        if (type == cls) valid_class = 1;
        else if (cls == MPB_KITTI_PEDESTRIAN && type == MPB_KITTI_PERSON_SITTING) valid_class = 0;
        else if (cls == MPB_KITTI_CAR && type == MPB_KITTI_VAN) valid_class = 0;
        else valid_class = -1;
        const bool ignore = (int)g[OCC] > kMaxOcclusion[difficulty] || g[TRUNC] > kMaxTruncation[difficulty] ||
                            height <= kMinHeight[difficulty];
        if (valid_class == 1 && !ignore) { im.ignored_gt.push_back(0); n_gt++; }
        else if (valid_class == 0 || (ignore && valid_class == 1)) im.ignored_gt.push_back(1);
        else im.ignored_gt.push_back(-1);
    }
    for (int i = 0; i < im.ngt; i++)
        if ((int)im.gt[(size_t)i * MPB_KITTI_GT_COLS + TYPE] == MPB_KITTI_DONTCARE) im.dontcare.push_back(i);
    for (int j = 0; j < im.ndet; j++) {
        const double* d = im.det + (size_t)j * MPB_KITTI_DET_COLS;
        const int height = (int)fabs(d[Y1] - d[Y2]);            // (sic) truncated to an integer
        if (height < kMinHeight[difficulty]) im.ignored_det.push_back(1);
        else if ((int)d[TYPE] == cls) im.ignored_det.push_back(0);
        else im.ignored_det.push_back(-1);
    }
}

struct Stat {
    std::vector<double> v;
    double similarity = 0, similarity_ground = 0;
    int tp = 0, fp = 0, fn = 0;
};

static Stat statistics(const Image& im, bool compute_fp, int metric, double min_overlap, bool compute_aos,
                       bool compute_aos_ground, double thresh) {
    Stat stat;
    std::vector<double> delta, delta_ground;
    std::vector<char> assigned(im.ndet, 0), below(im.ndet, 0);
    auto D = [&](int j) { return im.det + (size_t)j * MPB_KITTI_DET_COLS; };
    auto G = [&](int i) { return im.gt + (size_t)i * MPB_KITTI_GT_COLS; };
    if (compute_fp)
        for (int j = 0; j < im.ndet; j++)
            if (D(j)[SCORE] < thresh) below[j] = 1;
    for (int i = 0; i < im.ngt; i++) {
        if (im.ignored_gt[i] == -1) continue;
        int det_idx = -1;
        double valid_detection = kNoDetection, max_overlap = 0;
        bool assigned_ignored_det = false;
        for (int j = 0; j < im.ndet; j++) {
            if (im.ignored_det[j] == -1 || assigned[j] || below[j]) continue;
            const double o = overlap(D(j), G(i), metric, -1);
            if (!compute_fp && o > min_overlap && D(j)[SCORE] > valid_detection) {
                det_idx = j;                        // recall thresholds: the best-scoring candidate
                valid_detection = D(j)[SCORE];
            } else if (compute_fp && o > min_overlap && (o > max_overlap || assigned_ignored_det) && im.ignored_det[j] == 0) {
                max_overlap = o;                    // precision: the candidate that overlaps most ...
                det_idx = j;
                valid_detection = 1;
                assigned_ignored_det = false;
            } else if (compute_fp && o > min_overlap && valid_detection == kNoDetection && im.ignored_det[j] == 1) {
                det_idx = j;                        // ... or, failing that, a too-small detection
                valid_detection = 1;
                assigned_ignored_det = true;
            }
        }
        if (valid_detection == kNoDetection && im.ignored_gt[i] == 0) {
            stat.fn++;
        } else if (valid_detection != kNoDetection && (im.ignored_gt[i] == 1 || im.ignored_det[det_idx] == 1)) {
            assigned[det_idx] = 1;
        } else if (valid_detection != kNoDetection) {
            stat.tp++;
            stat.v.push_back(D(det_idx)[SCORE]);
            if (compute_aos) delta.push_back(G(i)[ALPHA] - D(det_idx)[ALPHA]);
            if (compute_aos_ground) delta_ground.push_back(fabs(G(i)[RY] - D(det_idx)[RY]));
            assigned[det_idx] = 1;
        }
    }
    if (compute_fp) {
        for (int j = 0; j < im.ndet; j++)
            if (!(assigned[j] || im.ignored_det[j] == -1 || im.ignored_det[j] == 1 || below[j])) stat.fp++;
        int nstuff = 0;                 // detections inside DontCare areas are not false positives
        for (int dc : im.dontcare)
            for (int j = 0; j < im.ndet; j++) {
                if (assigned[j] || im.ignored_det[j] == -1 || im.ignored_det[j] == 1 || below[j]) continue;
                if (overlap(D(j), G(dc), metric, 0) > min_overlap) {
                    assigned[j] = 1;
                    nstuff++;
                }
            }
        stat.fp -= nstuff;
        auto similarity = [&](const std::vector<double>& dl) {
            if (!(stat.tp > 0 || stat.fp > 0)) return -1.0;      // neither TP nor FP: the image does not count
            double s = 0;                                        // false positives contribute 0
            for (double x : dl) s += (1.0 + cos(x)) / 2.0;
            return s;
        };
        if (compute_aos) stat.similarity = similarity(delta);
        if (compute_aos_ground) stat.similarity_ground = similarity(delta_ground);
    }
    return stat;
}

static std::vector<double> thresholds(std::vector<double>& v, double n_groundtruth) {
    std::vector<double> t;
    std::sort(v.begin(), v.end(), std::greater<double>());
    double current_recall = 0;
    for (size_t i = 0; i < v.size(); i++) {
        const double l_recall = (double)(i + 1) / n_groundtruth;
        const double r_recall = i < v.size() - 1 ? (double)(i + 2) / n_groundtruth : l_recall;
        if ((r_recall - current_recall) < (current_recall - l_recall) && i < v.size() - 1) continue;
        t.push_back(v[i]);
        current_recall += 1.0 / (MPB_KITTI_SAMPLE_PTS - 1.0);
    }
    return t;
}

}  // namespace mpb_eval

MPB_EVAL_API double mpb_kitti_overlap(const double* det_row, const double* gt_row, int metric, int criterion) {
    if (!det_row || !gt_row || metric < 0 || metric > 2 || criterion < -1 || criterion > 1) return -1;
    return mpb_eval::overlap(det_row, gt_row, metric, criterion);
}

MPB_EVAL_API int mpb_kitti_eval_class(int n_images, const int* gt_off, const double* gt, const int* det_off,
                                      const double* det, int cls, int difficulty, int metric, double min_overlap,
                                      int compute_aos, int compute_aos_ground, double* precision, double* aos,
                                      double* aos_ground, int* n_thresholds, int* n_gt_out) {
    using namespace mpb_eval;
    if (n_images < 0 || !gt_off || !det_off || !precision || cls < 0 || cls > 2 || difficulty < 0 || difficulty > 2 ||
        metric < 0 || metric > 2 || (compute_aos && !aos) || (compute_aos_ground && !aos_ground))
        return -1;
    if ((gt_off[n_images] > 0 && !gt) || (det_off[n_images] > 0 && !det)) return -1;
    const int N = MPB_KITTI_SAMPLE_PTS;
    std::vector<Image> images(n_images);
    std::vector<double> v;
    int n_gt = 0;
    for (int i = 0; i < n_images; i++) {
        Image& im = images[i];
        if (gt_off[i + 1] < gt_off[i] || det_off[i + 1] < det_off[i]) return -1;
        im.gt = gt + (size_t)gt_off[i] * MPB_KITTI_GT_COLS; im.ngt = gt_off[i + 1] - gt_off[i];
        im.det = det + (size_t)det_off[i] * MPB_KITTI_DET_COLS; im.ndet = det_off[i + 1] - det_off[i];
        clean_data(im, cls, difficulty, n_gt);
        const Stat s = statistics(im, false, metric, min_overlap, false, false, 0);
        v.insert(v.end(), s.v.begin(), s.v.end());
    }
    const std::vector<double> th = thresholds(v, n_gt);
    const int T = (int)th.size();
    std::vector<Stat> pr(T);
    for (int i = 0; i < n_images; i++)
        for (int t = 0; t < T; t++) {
            const Stat s = statistics(images[i], true, metric, min_overlap, compute_aos != 0, compute_aos_ground != 0, th[t]);
            pr[t].tp += s.tp; pr[t].fp += s.fp; pr[t].fn += s.fn;
            if (s.similarity != -1) pr[t].similarity += s.similarity;
            if (s.similarity_ground != -1) pr[t].similarity_ground += s.similarity_ground;
        }
    for (int i = 0; i < N; i++) {
        precision[i] = 0;
        if (compute_aos) aos[i] = 0;
        if (compute_aos_ground) aos_ground[i] = 0;
    }
    for (int i = 0; i < T && i < N; i++) {
        precision[i] = pr[i].tp / (double)(pr[i].tp + pr[i].fp);
        if (compute_aos) aos[i] = pr[i].similarity / (double)(pr[i].tp + pr[i].fp);
        if (compute_aos_ground) aos_ground[i] = pr[i].similarity_ground / (double)(pr[i].tp + pr[i].fp);
    }
    for (int i = 0; i < T && i < N; i++) {          // monotone envelope: max over everything to the right
        precision[i] = *std::max_element(precision + i, precision + N);
        if (compute_aos) aos[i] = *std::max_element(aos + i, aos + N);
        if (compute_aos_ground) aos_ground[i] = *std::max_element(aos_ground + i, aos_ground + N);
    }
    if (n_thresholds) *n_thresholds = T;
    if (n_gt_out) *n_gt_out = n_gt;
    return 0;
}
