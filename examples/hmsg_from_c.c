/* The C-ABI of libhmsg_b200.so from plain C (what a cgo / JNI / ctypes binding calls):
 * two synthetic depth frames -> voxel table (A2) -> radius filter (A3) -> node table, then a small kNN (A11).
 *
 *   gcc -std=c99 -Iinclude examples/hmsg_from_c.c -o /tmp/hmsg_from_c -Lholoagent_b200 -lhmsg_b200 \
 *       -Wl,-rpath,$PWD/holoagent_b200
 *
 * tests/test_abi.py compiles and links this file in the CPU suite; it needs a B200 to run. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmsg_b200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int32_t rc_ = (call);                                                           \
    if (rc_ != HMSG_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, hmsg_last_error(ctx));     \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int main(void) {
  enum { H = 120, W = 160, F = 2, D = 128, N = 1000, K = 5 };
  hmsg_ctx* ctx = NULL;
  if (hmsg_ctx_create(0, &ctx) != HMSG_OK) {
    fprintf(stderr, "hmsg_ctx_create: %s\n", hmsg_last_error(NULL));   /* no CUDA device: a hard error, there is no CPU path */
    return 2;
  }
  /* dataset side (memory/hmsg/dataloader/generic.py:19-33): intrinsics, depth scale; cfg.pipeline.voxel_size */
  const double Kmat[9] = {80.0, 0.0, 80.0, 0.0, 80.0, 60.0, 0.0, 0.0, 1.0};
  CHECK(hmsg_scene_begin(ctx, H, W, Kmat, 1000.0f, 0.05, F));
  uint16_t* depth = (uint16_t*)malloc(sizeof(uint16_t) * F * H * W);
  uint8_t* rgb = (uint8_t*)malloc((size_t)F * H * W * 3);
  double poses[F * 16];
  for (int f = 0; f < F; f++) {
    for (int i = 0; i < H * W; i++) {
      depth[f * H * W + i] = (uint16_t)(1500 + 5 * (i % W) + 40 * f);   /* a slanted wall, millimetres */
      rgb[(f * H * W + i) * 3 + 0] = (uint8_t)(i & 255);
      rgb[(f * H * W + i) * 3 + 1] = (uint8_t)((i >> 3) & 255);
      rgb[(f * H * W + i) * 3 + 2] = (uint8_t)(f * 100);
    }
    memset(poses + 16 * f, 0, sizeof(double) * 16);
    poses[16 * f + 0] = poses[16 * f + 5] = poses[16 * f + 10] = poses[16 * f + 15] = 1.0;
    poses[16 * f + 3] = 0.1 * f;                                        /* camera-to-world: shifted along x */
  }
  CHECK(hmsg_scene_add_frames(ctx, depth, rgb, poses, F, 0));
  int64_t n_voxels = 0, n_nodes = 0;
  double min_bound[3];
  CHECK(hmsg_voxel_build(ctx, &n_voxels, min_bound));                   /* graph.py:344-348 */
  CHECK(hmsg_radius_filter(ctx, 10, 0.5, &n_nodes));                    /* graph.py:355-358 (nb_points 1000, radius 1.0 there) */
  double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n_nodes > 0 ? n_nodes : 1));
  CHECK(hmsg_nodes_read(ctx, xyz, NULL, NULL, NULL));
  printf("voxels %lld, nodes %lld, min bound (%.3f %.3f %.3f), first node (%.3f %.3f %.3f)\n", (long long)n_voxels, (long long)n_nodes,
         min_bound[0], min_bound[1], min_bound[2], n_nodes ? xyz[0] : 0.0, n_nodes ? xyz[1] : 0.0, n_nodes ? xyz[2] : 0.0);

  /* retrieval core (graph.py:3126-3133): np.dot(q, E.T), argsort descending, top k */
  float* E = (float*)malloc(sizeof(float) * N * D);
  float q[D];
  uint32_t lcg = 12345u;
  for (int i = 0; i < N * D; i++) {                                    /* uniform in [-1, 1) */
    lcg = lcg * 1664525u + 1013904223u;
    E[i] = (float)(lcg >> 8) * (2.0f / 16777216.0f) - 1.0f;
  }
  for (int j = 0; j < D; j++) q[j] = E[123 * D + j];                    /* row 123 must come back first */
  int64_t ids[K];
  float scores[K];
  CHECK(hmsg_index_set(ctx, E, N, D, 0));
  CHECK(hmsg_query_topk(ctx, q, 1, K, NULL, ids, scores, 0));
  printf("top-%d: ids %lld %lld %lld ..., best score %.4f, launches %lld\n", K, (long long)ids[0], (long long)ids[1], (long long)ids[2],
         scores[0], (long long)hmsg_launch_count(ctx));
  int ok = ids[0] == 123 && n_nodes > 0 && n_nodes <= n_voxels;
  free(E); free(xyz); free(depth); free(rgb);
  CHECK(hmsg_ctx_destroy(ctx));
  return ok ? 0 : 3;
}
