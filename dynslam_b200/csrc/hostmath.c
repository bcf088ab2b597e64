/* hostmath.c — host-side helpers of the Python mirror (dynslam_b200/engine.py), built into libb200host.so.
 * NOT part of libb200fusion / the C-ABI: the C++ shim uses ORUtils' own Matrix4f::inv(). The Python host has no ORUtils,
 * and b200_view.invM_d must carry exactly the bits the reference host would compute (the inverse pose feeds the ray set-up of
 * the allocation and raycast kernels), so the 4x4 inverse is evaluated here in the reference's operation order:
 * cofactor expansion over the transposed matrix (ORUtils/Matrix.h:162-224), each cofactor
 *   (p[a]*t[i] + p[b]*t[j] + p[c]*t[k]) - (p[d]*t[i] + p[e]*t[j] + p[f]*t[k])
 * with the 2x2 sub-determinant products p[] shared between cofactors. Written as index tables; compiled without
 * contraction (-ffp-contract=off). tests/test_cpu_abi.py compares it with the reference's own inv() bit for bit. */
#include <string.h>

/* products of pairs of the transposed matrix t[]: p[n] = t[PA[n]] * t[PB[n]]  (first table: rows 2,3; second: rows 0,1) */
static const unsigned char PA1[12] = {10, 11, 9, 11, 9, 10, 8, 11, 8, 10, 8, 9}, PB1[12] = {15, 14, 15, 13, 14, 13, 15, 12, 14, 12, 13, 12};
static const unsigned char PA2[12] = {2, 3, 1, 3, 1, 2, 0, 3, 0, 2, 0, 1}, PB2[12] = {7, 6, 7, 5, 6, 5, 7, 4, 6, 4, 5, 4};
/* cofactor n = (p[P[n][0]]*t[T[n][0]] + p[P[n][1]]*t[T[n][1]] + p[P[n][2]]*t[T[n][2]])
 *            - (p[P[n][3]]*t[T[n][0..2] in the order of Q]) */
struct cof { unsigned char pp[3], tp[3], pm[3], tm[3]; };
static const struct cof C1[8] = {
  {{0, 3, 4}, {5, 6, 7}, {1, 2, 5}, {5, 6, 7}},   {{1, 6, 9}, {4, 6, 7}, {0, 7, 8}, {4, 6, 7}},
  {{2, 7, 10}, {4, 5, 7}, {3, 6, 11}, {4, 5, 7}}, {{5, 8, 11}, {4, 5, 6}, {4, 9, 10}, {4, 5, 6}},
  {{1, 2, 5}, {1, 2, 3}, {0, 3, 4}, {1, 2, 3}},   {{0, 7, 8}, {0, 2, 3}, {1, 6, 9}, {0, 2, 3}},
  {{3, 6, 11}, {0, 1, 3}, {2, 7, 10}, {0, 1, 3}}, {{4, 9, 10}, {0, 1, 2}, {5, 8, 11}, {0, 1, 2}}};
static const struct cof C2[8] = {
  {{0, 3, 4}, {13, 14, 15}, {1, 2, 5}, {13, 14, 15}},   {{1, 6, 9}, {12, 14, 15}, {0, 7, 8}, {12, 14, 15}},
  {{2, 7, 10}, {12, 13, 15}, {3, 6, 11}, {12, 13, 15}}, {{5, 8, 11}, {12, 13, 14}, {4, 9, 10}, {12, 13, 14}},
  {{2, 5, 1}, {10, 11, 9}, {4, 0, 3}, {11, 9, 10}},     {{8, 0, 7}, {11, 8, 10}, {6, 9, 1}, {10, 11, 8}},
  {{6, 11, 3}, {9, 11, 8}, {10, 2, 7}, {11, 8, 9}},     {{10, 4, 9}, {10, 8, 9}, {8, 11, 5}, {9, 10, 8}}};

static float cofactor(const struct cof *c, const float *p, const float *t) {
  const float plus = p[c->pp[0]] * t[c->tp[0]] + p[c->pp[1]] * t[c->tp[1]] + p[c->pp[2]] * t[c->tp[2]];
  const float minus = p[c->pm[0]] * t[c->tm[0]] + p[c->pm[1]] * t[c->tm[1]] + p[c->pm[2]] * t[c->tm[2]];
  return plus - minus;
}

int b200h_mat4_inv(const float *m, float *dst) {
  float t[16], p[12], out[16];
  int n;
  for (n = 0; n < 16; ++n) t[n] = m[(n & 3) * 4 + (n >> 2)];          /* transpose */
  for (n = 0; n < 12; ++n) p[n] = t[PA1[n]] * t[PB1[n]];
  for (n = 0; n < 4; ++n) out[n] = cofactor(&C1[n], p, t);
  {
    const float det = t[0] * out[0] + t[1] * out[1] + t[2] * out[2] + t[3] * out[3];
    float s;
    if (det == 0.0f) return 0;
    for (n = 4; n < 8; ++n) out[n] = cofactor(&C1[n], p, t);
    for (n = 0; n < 12; ++n) p[n] = t[PA2[n]] * t[PB2[n]];
    for (n = 0; n < 8; ++n) out[8 + n] = cofactor(&C2[n], p, t);
    s = 1 / det;
    for (n = 0; n < 16; ++n) dst[n] = out[n] * s;
  }
  return 1;
}

/* lhs * rhs, column-major m[col*4+row], accumulated over k ascending from 0.0f (ORUtils/Matrix.h:102-108) */
void b200h_mat4_mul(const float *lhs, const float *rhs, float *out) {
  float r[16];
  int c, row, k;
  for (c = 0; c < 4; ++c) for (row = 0; row < 4; ++row) {
    float acc = 0.0f;
    for (k = 0; k < 4; ++k) acc += lhs[k * 4 + row] * rhs[c * 4 + k];
    r[c * 4 + row] = acc;
  }
  memcpy(out, r, sizeof(r));
}
