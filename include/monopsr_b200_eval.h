/* monopsr_b200_eval.h -- C ABI of the KITTI object-detection AP evaluator (host code, no GPU).
 *
 * Replaces the native evaluator the reference shells out to after inference
 * (scripts/offline_eval/kitti_native_eval/evaluate_object_3d_offline.cpp, run through run_eval.sh:20 by
 * src/monopsr/core/evaluator_utils.py:512-535; the `_low_iou` twin differs only in MIN_OVERLAP, :55).
 * One call evaluates one (class, difficulty, metric) triple over all images, i.e. the reference's
 * eval_class (:648-744) with cleanData (:383-459), computeStatistics (:461-642) and getThresholds (:347-381).
 * The host side (monopsr_b200/core/kitti_eval.py) parses the label / result files, decides which classes and
 * metrics are evaluated (loadDetections, :130-175) and prints the reference's "AP" lines (:746-755).
 */
#ifndef MONOPSR_B200_EVAL_H
#define MONOPSR_B200_EVAL_H

#ifdef __cplusplus
extern "C" {
#endif

/* object type codes (anything that is not one of the named types: MPB_KITTI_OTHER) */
enum { MPB_KITTI_CAR = 0, MPB_KITTI_PEDESTRIAN = 1, MPB_KITTI_CYCLIST = 2, MPB_KITTI_VAN = 3,
       MPB_KITTI_PERSON_SITTING = 4, MPB_KITTI_DONTCARE = 5, MPB_KITTI_OTHER = 6 };
enum { MPB_KITTI_EASY = 0, MPB_KITTI_MODERATE = 1, MPB_KITTI_HARD = 2 };
enum { MPB_KITTI_IMAGE = 0, MPB_KITTI_GROUND = 1, MPB_KITTI_BOX3D = 2 };
#define MPB_KITTI_SAMPLE_PTS 41
#define MPB_KITTI_GT_COLS 15   /* type, truncation, occlusion, alpha, x1, y1, x2, y2, h, w, l, t1, t2, t3, ry */
#define MPB_KITTI_DET_COLS 16  /* the same (truncation / occlusion unused) + score */

/* gt  : rows of MPB_KITTI_GT_COLS doubles, image i owns rows gt_off[i] .. gt_off[i+1]-1 (gt_off has n_images+1 entries)
 * det : rows of MPB_KITTI_DET_COLS doubles, likewise with det_off
 * cls : MPB_KITTI_CAR / PEDESTRIAN / CYCLIST;  min_overlap: 0.7 / 0.5 / 0.5 (0.5 / 0.25 / 0.25 for "low IoU")
 * precision / aos / aos_ground : MPB_KITTI_SAMPLE_PTS doubles each (aos / aos_ground may be NULL when not computed);
 *   filled as the reference does: entries past the number of score thresholds stay 0, then every entry becomes the
 *   maximum of itself and everything to its right
 * n_thresholds, n_gt : optional outputs (number of recall sample points reached, number of valid ground-truth objects)
 * returns 0, or -1 on invalid arguments. */
int mpb_kitti_eval_class(int n_images, const int* gt_off, const double* gt, const int* det_off, const double* det,
                         int cls, int difficulty, int metric, double min_overlap, int compute_aos,
                         int compute_aos_ground, double* precision, double* aos, double* aos_ground,
                         int* n_thresholds, int* n_gt);

/* overlap of two boxes given as rows in the layouts above: criterion -1 = intersection over union,
 * 0 = over the detection's area / volume, 1 = over the ground truth's (imageBoxOverlap :225-262,
 * groundBoxOverlap :292-313, box3DOverlap :315-345) */
double mpb_kitti_overlap(const double* det_row, const double* gt_row, int metric, int criterion);

#ifdef __cplusplus
}
#endif
#endif
