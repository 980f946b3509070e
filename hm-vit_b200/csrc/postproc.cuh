// Detection post-processing on the GPU: replaces VoxelPostprocessor.post_process for the
// intermediate-fusion case (one 'ego' entry, batch size 1)
//   opencood/data_utils/post_processor/voxel_postprocessor.py:232-343 (post_process), :345-397 (delta_to_boxes3d)
//   opencood/utils/box_utils.py:139-184 (boxes_to_corners_3d), :258-296 (project_box3d), :326-357 (range mask),
//                               :575-620 (nms_rotated), :722-772 (remove_large_pred_bbx, remove_bbx_abnormal_z)
//   opencood/utils/common_utils.py:120-158 (compute_iou / convert_format: shapely polygons of the first four corners)
//
//   post_decode_kernel    thread == anchor (h, w, a): sigmoid score, threshold, anchor decoding, 8 corners, projection,
//                         the two sanity filters -> keep flag, score, corners (dense arrays in anchor order)
//   post_compact_kernel   one CTA: ORDER-PRESERVING compaction of the kept anchors (the reference's masked_select order)
//   post_sort_kernel      one CTA: bitonic sort of the candidates by (score descending, candidate index ascending) in shared
//                         memory; the first `top` (1000) go on, like `scores.argsort()[::-1][:top]`
//   post_iou_kernel       rotated IoU of every pair of the top candidates in fp64 (convex polygon clipping of the quadrilaterals
//                         spanned by corners 0..3) -> suppression bit matrix
//   post_nms_kernel       one warp: greedy pass in score order over the bit matrix, range mask, outputs in pick order
// The reference's quirks are kept: `remove_large_pred_bbx` measures its "z length" on the y coordinates and ANDs the
// length itself (box_utils.py:744-750), i.e. keep = x_len <= 6 and y_len <= 6 and y_len != 0.
#pragma once
#include "common.cuh"

namespace hmvit {

constexpr int kPostTop = 1000;           // nms_rotated keeps the 1000 best-scored candidates (box_utils.py:600)
constexpr int kPostMaxCand = 16384;      // candidates the single-CTA sort handles
constexpr int kPostMaskWords = (kPostTop + 63) / 64;

struct PostParams {
  int H, W, A;                 // feature map and anchors per cell
  const float* psm;            // (1, A, H, W)
  const float* rm;             // (1, 7A, H, W)
  const float* anchors;        // (H, W, A, 7): x, y, z, h, w, l, r
  const float* tmat;           // 4x4 row-major projection to the ego frame, or null ('no_post_projection')
  int order_hwl;               // 1: 'hwl' (boxes are x, y, z, h, w, l, yaw), 0: 'lwh'
  float score_thr, nms_thr;
  float range[4];              // x_min, y_min, x_max, y_max (GT_RANGE[0:2], GT_RANGE[3:5])
  // workspace
  uint8_t* keep;               // [HWA]
  float* score;                // [HWA]
  float* corners;              // [HWA][24]
  int* cand;                   // [kPostMaxCand] anchor index of candidate c (anchor order)
  int* n_cand;                 // [1]
  int* order;                  // [kPostTop] candidate index by rank
  int* n_top;                  // [1]
  unsigned long long* mask;    // [kPostTop][kPostMaskWords] bit j of row i: IoU(i, j) > thr, j > i (ranks)
  // outputs
  float* out_boxes;            // [kPostTop][8][3]
  float* out_scores;           // [kPostTop]
  int* out_count;              // [1]
  int* status;                 // [1] 0 ok, 1 more than kPostMaxCand candidates
};

__global__ void __launch_bounds__(256) post_decode_kernel(const PostParams p) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int HW = p.H * p.W, total = HW * p.A;
  if (n >= total) return;
  const int a = n % p.A, pix = n / p.A;
  // classification probability (voxel_postprocessor.py:270-272): sigmoid of psm permuted to (H, W, A)
  const float logit = p.psm[static_cast<size_t>(a) * HW + pix];
  const float prob = 1.0f / (1.0f + expf(-logit));
  bool keep = prob > p.score_thr;
  float c[8][3];
  if (keep) {
    const float* an = p.anchors + static_cast<size_t>(n) * 7;
    float d[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) d[k] = p.rm[static_cast<size_t>(a * 7 + k) * HW + pix];
    // delta_to_boxes3d (:376-395)
    const float ad = sqrtf(an[4] * an[4] + an[5] * an[5]);
    float box[7];
    box[0] = d[0] * ad + an[0];
    box[1] = d[1] * ad + an[1];
    box[2] = d[2] * an[3] + an[2];
    box[3] = expf(d[3]) * an[3];
    box[4] = expf(d[4]) * an[4];
    box[5] = expf(d[5]) * an[5];
    box[6] = d[6] + an[6];
    // boxes_to_corners_3d (box_utils.py:167-184): 'hwl' boxes are reordered to l, w, h first
    float dx = box[3], dy = box[4], dz = box[5];
    if (p.order_hwl) { dx = box[5]; dz = box[3]; }
    const float cosa = cosf(box[6]), sina = sinf(box[6]);
    const float tx[8] = {1, 1, -1, -1, 1, 1, -1, -1}, ty[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, tz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    float zmin = INFINITY, zmax = -INFINITY, xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float lx = dx * (tx[k] * 0.5f), ly = dy * (ty[k] * 0.5f), lz = dz * (tz[k] * 0.5f);
      // points @ [[cos, sin, 0], [-sin, cos, 0], [0, 0, 1]] (common_utils.py:44-49), then + centre
      float x = lx * cosa + ly * (-sina) + box[0];
      float y = lx * sina + ly * cosa + box[1];
      float z = lz + box[2];
      if (p.tmat != nullptr) {            // project_box3d: T @ [x, y, z, 1]
        const float* T = p.tmat;
        const float px = T[0] * x + T[1] * y + T[2] * z + T[3];
        const float py = T[4] * x + T[5] * y + T[6] * z + T[7];
        const float pz = T[8] * x + T[9] * y + T[10] * z + T[11];
        x = px; y = py; z = pz;
      }
      c[k][0] = x; c[k][1] = y; c[k][2] = z;
      xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
      zmin = fminf(zmin, z); zmax = fmaxf(zmax, z);
    }
    const float xl = xmax - xmin, yl = ymax - ymin;
    const bool keep1 = xl <= 6.0f && yl <= 6.0f && yl != 0.0f;          // remove_large_pred_bbx (with its quirk)
    const bool keep2 = zmin >= -3.0f && zmax <= 1.0f;                    // remove_bbx_abnormal_z
    keep = keep1 && keep2;
  }
  p.keep[n] = keep ? 1 : 0;
  if (keep) {
    p.score[n] = prob;
    float* o = p.corners + static_cast<size_t>(n) * 24;
#pragma unroll
    for (int k = 0; k < 8; ++k) { o[k * 3] = c[k][0]; o[k * 3 + 1] = c[k][1]; o[k * 3 + 2] = c[k][2]; }
  }
}

// one CTA of 1024 threads: stable compaction (anchor order) of the kept anchors
__global__ void __launch_bounds__(1024) post_compact_kernel(const PostParams p) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int total = p.H * p.W * p.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int n0 = 0; n0 < total; n0 += 1024) {
    const int n = n0 + threadIdx.x;
    const bool k = n < total && p.keep[n] != 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, sum = 0;
    for (int w = 0; w < 32; ++w) { const int cnt = s_warp[w]; if (w < warp) before += cnt; sum += cnt; }
    const int pos = s_base + before + __popc(bal & ((1u << lane) - 1u));
    if (k && pos < kPostMaxCand) p.cand[pos] = n;
    __syncthreads();
    if (threadIdx.x == 0) s_base += sum;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *p.status = s_base > kPostMaxCand ? 1 : 0;
    *p.n_cand = min(s_base, kPostMaxCand);
  }
}

// one CTA of 1024 threads: bitonic sort of (score descending, candidate index ascending); dynamic smem = 8 bytes x padded size
__global__ void __launch_bounds__(1024) post_sort_kernel(const PostParams p) {
  extern __shared__ unsigned long long s_key[];
  const int n = *p.n_cand;
  int m = 1;
  while (m < n) m <<= 1;
  if (m < 2) m = 2;
  // key = (score bits (positive floats order like unsigned), ~candidate index): a DESCENDING sort of the key puts the higher
  // score first and, among equal scores, the lower candidate index first
  for (int i = threadIdx.x; i < m; i += 1024) {
    unsigned long long key = 0ull;
    if (i < n) key = (static_cast<unsigned long long>(__float_as_uint(p.score[p.cand[i]])) << 32) | (0xffffffffu - static_cast<uint32_t>(i));
    s_key[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = s_key[i], y = s_key[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { s_key[i] = y; s_key[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  const int top = min(n, kPostTop);
  for (int i = threadIdx.x; i < top; i += 1024) p.order[i] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(s_key[i] & 0xffffffffull));
  if (threadIdx.x == 0) *p.n_top = top;
}

// area of the intersection of two convex quadrilaterals (fp64, Sutherland-Hodgman; both made counter-clockwise first)
HMVIT_DEVINL double quad_signed_area(const double (&q)[4][2]) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const int j = (i + 1) & 3; s += q[i][0] * q[j][1] - q[j][0] * q[i][1]; }
  return 0.5 * s;
}
HMVIT_DEVINL double quad_intersection_area(const double (&a_in)[4][2], const double (&b_in)[4][2]) {
  double a[4][2], b[4][2];
  const bool fa = quad_signed_area(a_in) < 0.0, fb = quad_signed_area(b_in) < 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i][0] = a_in[fa ? 3 - i : i][0]; a[i][1] = a_in[fa ? 3 - i : i][1];
    b[i][0] = b_in[fb ? 3 - i : i][0]; b[i][1] = b_in[fb ? 3 - i : i][1];
  }
  double poly[16][2], tmp[16][2];
  int np = 4;
  for (int i = 0; i < 4; ++i) { poly[i][0] = a[i][0]; poly[i][1] = a[i][1]; }
  for (int e = 0; e < 4 && np > 0; ++e) {
    const double x1 = b[e][0], y1 = b[e][1], x2 = b[(e + 1) & 3][0], y2 = b[(e + 1) & 3][1];
    int nt = 0;
    for (int i = 0; i < np; ++i) {
      const double px = poly[i][0], py = poly[i][1], qx = poly[(i + 1) % np][0], qy = poly[(i + 1) % np][1];
      const double sp = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1);      // >= 0: inside (left of the edge)
      const double sq = (x2 - x1) * (qy - y1) - (y2 - y1) * (qx - x1);
      if (sp >= 0.0) { tmp[nt][0] = px; tmp[nt][1] = py; ++nt; }
      if ((sp >= 0.0) != (sq >= 0.0)) {
        const double t = sp / (sp - sq);
        tmp[nt][0] = px + t * (qx - px); tmp[nt][1] = py + t * (qy - py); ++nt;
      }
    }
    np = nt;
    for (int i = 0; i < np; ++i) { poly[i][0] = tmp[i][0]; poly[i][1] = tmp[i][1]; }
  }
  double s = 0.0;
  for (int i = 0; i < np; ++i) { const int j = (i + 1) % np; s += poly[i][0] * poly[j][1] - poly[j][0] * poly[i][1]; }
  return np >= 3 ? fabs(0.5 * s) : 0.0;
}

// grid (ceil(T / 64), T): block (jb, i) handles ranks j = 64 jb .. 64 jb + 63 against rank i
__global__ void __launch_bounds__(64) post_iou_kernel(const PostParams p) {
  const int T = *p.n_top;
  const int i = blockIdx.y, j = blockIdx.x * 64 + threadIdx.x;
  if (i >= T) return;
  bool sup = false;
  if (j < T && j > i) {
    const float* ci = p.corners + static_cast<size_t>(p.cand[p.order[i]]) * 24;
    const float* cj = p.corners + static_cast<size_t>(p.cand[p.order[j]]) * 24;
    double a[4][2], b[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k][0] = ci[k * 3]; a[k][1] = ci[k * 3 + 1]; b[k][0] = cj[k * 3]; b[k][1] = cj[k * 3 + 1]; }
    const double inter = quad_intersection_area(a, b);
    const double uni = fabs(quad_signed_area(a)) + fabs(quad_signed_area(b)) - inter;
    // compute_iou returns float32 (common_utils.py:139-140); `iou > threshold` compares it with the python float
    const float iou = static_cast<float>(inter / uni);
    sup = static_cast<double>(iou) > static_cast<double>(p.nms_thr);
  }
  const unsigned long long bits = __ballot_sync(0xffffffffu, sup);
  __shared__ unsigned int s_hi;
  if (threadIdx.x == 32) s_hi = static_cast<unsigned int>(bits);
  __syncthreads();
  if (threadIdx.x == 0)
    p.mask[static_cast<size_t>(i) * kPostMaskWords + blockIdx.x] = (bits & 0xffffffffull) | (static_cast<unsigned long long>(s_hi) << 32);
}

// one warp: greedy NMS over the ranks, then the range mask; outputs in pick order
__global__ void __launch_bounds__(32) post_nms_kernel(const PostParams p) {
  __shared__ unsigned long long removed[kPostMaskWords];
  const int T = *p.n_top;
  const int lane = threadIdx.x;
  for (int w = lane; w < kPostMaskWords; w += 32) removed[w] = 0ull;
  __syncwarp();
  int count = 0;
  for (int i = 0; i < T; ++i) {
    const bool gone = (removed[i >> 6] >> (i & 63)) & 1ull;     // (uniform)
    if (gone) continue;
    for (int w = lane; w < kPostMaskWords; w += 32) removed[w] |= p.mask[static_cast<size_t>(i) * kPostMaskWords + w];
    __syncwarp();
    // get_mask_for_boxes_within_range_torch: all 8 corners inside [x_min, x_max] x [y_min, y_max]
    const float* c = p.corners + static_cast<size_t>(p.cand[p.order[i]]) * 24;
    bool in = true;
    if (lane < 8) {
      const float x = c[lane * 3], y = c[lane * 3 + 1];
      in = x >= p.range[0] && y >= p.range[1] && x <= p.range[2] && y <= p.range[3];
    }
    if (__all_sync(0xffffffffu, in)) {
      if (lane < 24) p.out_boxes[static_cast<size_t>(count) * 24 + lane] = c[lane];
      if (lane == 0) p.out_scores[count] = p.score[p.cand[p.order[i]]];
      ++count;
    }
  }
  if (lane == 0) *p.out_count = count;
}

}  // namespace hmvit
