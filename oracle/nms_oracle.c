/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
 * may use it, and only as the checker or the timed CPU baseline.
 *
 * Plain-C restatement of the greedy NMS the reference calls at model/_base.py:203
 * (`torchvision.ops.nms(boxes, scores, 0.5)`).  The algorithm lives in a third-party
 * dependency (torchvision, pinned 0.19.1 in requirements.txt:177; 0.26.0 in this image),
 * torchvision/csrc/ops/cpu/nms_kernel.cpp: stable descending sort of the scores, then for each
 * surviving box i suppress every later box j with
 *     inter / (area_i + area_j - inter) > iou_threshold
 * where inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1)), everything in fp32.
 * Parity pin: tests/test_oracle.py checks this file against torchvision.ops.nms itself
 * (random, tie-heavy, degenerate and NaN inputs) and against tests/golden/nms_*.npz.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile) — contraction must
 * stay off so no FMA changes a rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float fmax_std(float a, float b) { return (a < b) ? b : a; } /* std::max */
static float fmin_std(float a, float b) { return (b < a) ? b : a; } /* std::min */

typedef struct { float key; int64_t idx; } item_t;

/* torch.sort(descending=True, stable=True): NaN is the largest value. */
static int greater(float a, float b) {
  int an = a != a, bn = b != b;
  if (an || bn) return an && !bn;
  return a > b;
}

static void merge_sort(item_t* a, item_t* tmp, int64_t n) {
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort(a, tmp, h);
  merge_sort(a + h, tmp, n - h);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) {
    if (greater(a[j].key, a[i].key)) tmp[k++] = a[j++]; /* strict: ties keep the left (lower index) */
    else tmp[k++] = a[i++];
  }
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(item_t));
}

/* boxes (n,4) xyxy fp32, scores (n) fp32 -> keep (<= n) int64; returns the number kept. */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold, int64_t* keep) {
  if (n <= 0) return 0;
  item_t* order = (item_t*)malloc((size_t)n * sizeof(item_t));
  item_t* tmp = (item_t*)malloc((size_t)n * sizeof(item_t));
  float* areas = (float*)malloc((size_t)n * sizeof(float));
  uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) {
    order[i].key = scores[i];
    order[i].idx = i;
    float w = boxes[4 * i + 2] - boxes[4 * i + 0];
    float h = boxes[4 * i + 3] - boxes[4 * i + 1];
    areas[i] = w * h;
  }
  merge_sort(order, tmp, n);
  int64_t kept = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i].idx;
    if (suppressed[i]) continue;
    keep[kept++] = i;
    float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j].idx;
      if (suppressed[j]) continue;
      float xx1 = fmax_std(ix1, boxes[4 * j]);
      float yy1 = fmax_std(iy1, boxes[4 * j + 1]);
      float xx2 = fmin_std(ix2, boxes[4 * j + 2]);
      float yy2 = fmin_std(iy2, boxes[4 * j + 3]);
      float w = fmax_std(0.0f, xx2 - xx1);
      float h = fmax_std(0.0f, yy2 - yy1);
      float inter = w * h;
      float uni = iarea + areas[j];
      uni = uni - inter;
      float ovr = inter / uni;
      if ((double)ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  free(order); free(tmp); free(areas); free(suppressed);
  return kept;
}
