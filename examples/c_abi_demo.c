/* Plain-C use of the drop-in boundary (include/smplk.h): no CUDA headers, no Python, host buffers only.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo -L3d-human-body-reconstruction_b200 -lsmplk -lm \
 *       -Wl,-rpath,$PWD/3d-human-body-reconstruction_b200
 *   /tmp/c_abi_demo            (needs an sm_100 GPU; exits 1 with the library's error string otherwise)
 *
 * Builds a small random body model (a 24-joint chain, V vertices, <= 2 weights per vertex) in float64 as the
 * reference's pickles hold it (models/smpl_np.py:124-133), evaluates B bodies through smplk_forward_host --
 * what SMPLModel.set_params(pose, beta, trans) does per body (models/smpl_np.py:158-206) -- and checks the
 * result against the same arithmetic written out in double precision below (Rodrigues, chain, rest-pose
 * removal, blend shapes, skinning).  Prints the largest vertex error; exits 0 when it is <= 1e-5. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smplk.h"

#define V 1500
#define J 24
#define NB 10
#define P (9 * (J - 1))
#define B 7

static double urand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xffffff) / 16777216.0; }
static double nrand(unsigned* s) { double a = 0; for (int i = 0; i < 6; ++i) a += urand(s); return (a - 3.0) * 1.4142; }

static void rodrigues(const double* r, double* R) {      /* models/smpl_np.py:208-228 */
  double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 1e-300) th = 1e-300;
  const double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = cos(th), s = sin(th), k = 1 - c;
  R[0] = c + k * x * x;     R[1] = k * x * y - s * z; R[2] = k * x * z + s * y;
  R[3] = k * x * y + s * z; R[4] = c + k * y * y;     R[5] = k * y * z - s * x;
  R[6] = k * x * z - s * y; R[7] = k * y * z + s * x; R[8] = c + k * z * z;
}

int main(void) {
  unsigned seed = 12345u;
  double* vt = malloc(sizeof(double) * V * 3);
  double* sd = malloc(sizeof(double) * V * 3 * NB);
  double* pd = malloc(sizeof(double) * V * 3 * P);
  double* Jr = calloc((size_t)J * V, sizeof(double));
  double* W = calloc((size_t)V * J, sizeof(double));
  int32_t parents[J];
  for (int i = 0; i < V * 3; ++i) vt[i] = nrand(&seed) * 0.3;
  for (int i = 0; i < V * 3 * NB; ++i) sd[i] = nrand(&seed) * 0.01;
  for (int i = 0; i < V * 3 * P; ++i) pd[i] = nrand(&seed) * 0.001;
  for (int j = 0; j < J; ++j) {
    parents[j] = j == 0 ? -1 : (int)(urand(&seed) * j);
    double sum = 0;
    for (int k = 0; k < 8; ++k) { int v = (int)(urand(&seed) * V); double w = urand(&seed) + 0.1; Jr[(size_t)j * V + v] += w; sum += w; }
    for (int v = 0; v < V; ++v) Jr[(size_t)j * V + v] /= sum;
  }
  for (int v = 0; v < V; ++v) {
    const int j0 = v * J / V, j1 = parents[j0] < 0 ? j0 : parents[j0];
    const double w = 0.5 + 0.5 * urand(&seed);
    W[(size_t)v * J + j0] += w;
    W[(size_t)v * J + j1] += 1.0 - w;
  }
  smplk_model_desc d;
  memset(&d, 0, sizeof(d));
  d.num_verts = V; d.num_joints = J; d.num_betas = NB;
  d.v_template = vt; d.shapedirs = sd; d.posedirs = pd; d.J_regressor = Jr; d.weights = W; d.parents = parents;
  smplk_model* model = NULL;
  if (smplk_model_create(&d, 0, &model) != SMPLK_OK) {
    fprintf(stderr, "smplk_model_create: %s\n", smplk_last_error_string());
    return 1;
  }
  float pose[B][J * 3], beta[B][NB], trans[B][3];
  static float verts[B][V][3];
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < J * 3; ++i) pose[b][i] = (float)(nrand(&seed) * (b == 0 ? 0.0 : 0.4));
    for (int i = 0; i < NB; ++i) beta[b][i] = (float)nrand(&seed);
    for (int i = 0; i < 3; ++i) trans[b][i] = (float)nrand(&seed);
  }
  if (smplk_forward_host(model, B, 0, &beta[0][0], B, &pose[0][0], &trans[0][0], &verts[0][0][0], NULL, NULL) != SMPLK_OK) {
    fprintf(stderr, "smplk_forward_host: %s\n", smplk_last_error_string());
    return 1;
  }
  /* the same arithmetic in double precision */
  double worst = 0;
  double* vs = malloc(sizeof(double) * V * 3);
  double* vp = malloc(sizeof(double) * V * 3);
  for (int b = 0; b < B; ++b) {
    double R[J][9], Jj[J][3], G[J][12], A[J][12], feat[P];
    for (int n = 0; n < V * 3; ++n) { double a = vt[n]; for (int i = 0; i < NB; ++i) a += sd[(size_t)n * NB + i] * beta[b][i]; vs[n] = a; }
    for (int j = 0; j < J; ++j)
      for (int c = 0; c < 3; ++c) { double a = 0; for (int v = 0; v < V; ++v) a += Jr[(size_t)j * V + v] * vs[3 * v + c]; Jj[j][c] = a; }
    for (int j = 0; j < J; ++j) { double r[3] = {pose[b][3 * j], pose[b][3 * j + 1], pose[b][3 * j + 2]}; rodrigues(r, R[j]); }
    for (int j = 1; j < J; ++j) for (int i = 0; i < 9; ++i) feat[9 * (j - 1) + i] = R[j][i] - (i % 4 == 0 ? 1.0 : 0.0);
    for (int n = 0; n < V * 3; ++n) { double a = vs[n]; for (int k = 0; k < P; ++k) a += pd[(size_t)n * P + k] * feat[k]; vp[n] = a; }
    for (int j = 0; j < J; ++j) {
      double t[3] = {Jj[j][0], Jj[j][1], Jj[j][2]};
      if (j == 0) { for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) G[0][4 * r + c] = R[0][3 * r + c]; G[0][4 * r + 3] = t[r]; } continue; }
      const int p = parents[j];
      for (int r = 0; r < 3; ++r) t[r] -= Jj[p][r];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) G[j][4 * r + c] = G[p][4 * r] * R[j][c] + G[p][4 * r + 1] * R[j][3 + c] + G[p][4 * r + 2] * R[j][6 + c];
        G[j][4 * r + 3] = G[p][4 * r] * t[0] + G[p][4 * r + 1] * t[1] + G[p][4 * r + 2] * t[2] + G[p][4 * r + 3];
      }
    }
    for (int j = 0; j < J; ++j)
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[j][4 * r + c] = G[j][4 * r + c];
        A[j][4 * r + 3] = G[j][4 * r + 3] - (G[j][4 * r] * Jj[j][0] + G[j][4 * r + 1] * Jj[j][1] + G[j][4 * r + 2] * Jj[j][2]);
      }
    for (int v = 0; v < V; ++v)
      for (int r = 0; r < 3; ++r) {
        double a = 0;
        for (int j = 0; j < J; ++j) {
          const double w = W[(size_t)v * J + j];
          if (w != 0) a += w * (A[j][4 * r] * vp[3 * v] + A[j][4 * r + 1] * vp[3 * v + 1] + A[j][4 * r + 2] * vp[3 * v + 2] + A[j][4 * r + 3]);
        }
        const double e = fabs(a + trans[b][r] - verts[b][v][r]);
        if (e > worst) worst = e;
      }
  }
  printf("smplk C ABI v%d: %d bodies x %d vertices through smplk_forward_host, max |error| vs double = %.3g m, %llu kernels launched\n",
         smplk_version(), B, V, worst, (unsigned long long)smplk_launch_count());
  smplk_model_destroy(model);
  return worst <= 1e-5 ? 0 : 2;
}
