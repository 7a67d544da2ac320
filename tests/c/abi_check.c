/* C99 consumer of include/rfdnet_b200.h: proves the boundary is a plain C ABI (no C++/torch types) and that the
 * argument validation of every entry point runs before any CUDA call (so it works on a machine without a GPU).
 * When a CUDA device is present it also does one real round trip with buffers from the plain CUDA runtime C API
 * (no torch anywhere in the process): FPS + ball query + grouping on a 64-point cloud, checked on the host. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rfdnet_b200.h"

#define EXPECT(cond)                                     \
  do {                                                   \
    if (!(cond)) {                                       \
      fprintf(stderr, "abi_check failed: %s\n", #cond);  \
      return 1;                                          \
    }                                                    \
  } while (0)

int main(void) {
  EXPECT(rfd_abi_version() == RFD_ABI_VERSION);
  EXPECT(strcmp(rfd_status_string(RFD_OK), "ok") == 0);
  EXPECT(rfd_status_string(RFD_ERR_CUDA) != NULL);
  EXPECT(rfd_launch_count() == 0);
  /* invalid arguments */
  EXPECT(rfd_furthest_point_sampling(NULL, 1, 16, 4, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_furthest_point_sampling_xyz(NULL, 1, 0, 4, NULL, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_ball_query(NULL, NULL, 1, 8, 8, 0.1f, 4, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_query_and_group(NULL, NULL, NULL, 1, 8, 8, 0, 0.1f, 4, 0, 0, NULL, NULL, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_three_nn(NULL, NULL, 1, 8, 8, NULL, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_pointwise_mlp_f32(NULL, NULL, NULL, NULL, NULL, 1, 1, 1, 4, 4, 16, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_onet_decode(NULL, 0, 1, 128, NULL, NULL, 1, NULL, NULL, 0.f, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_mlp_chain(RFD_MLP_MODE_F16X3, NULL, 1, 4, 128, NULL, 64, 64, 128, 1, 16, NULL, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  /* round-2 entry points: surface extraction, STN_Group, the extended chain calls */
  EXPECT(rfd_extract_mesh(NULL, 1, 32, 0.0, 1.1, NULL, 0, NULL, 16, 16, NULL, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_extract_mesh((const float *)16, 1, 40, 0.0, 1.1, (void *)16, 0, (int *)16, 16, 16, (int *)16, (unsigned long long *)16, NULL) ==
         RFD_ERR_UNSUPPORTED_SIZE); /* the padded lattice must fit in shared memory: R <= 32 */
  EXPECT(rfd_extract_mesh(NULL, 0, 32, 0.0, 1.1, NULL, 0, NULL, 0, 0, NULL, NULL, NULL) == RFD_OK);
  EXPECT(rfd_query_and_group_rotated((const float *)16, (const float *)16, NULL, NULL, 1, 8, 8, 0, 0.1f, 4, 1, 0, (float *)16, NULL, NULL,
                                     NULL) == RFD_ERR_INVALID_ARGUMENT); /* heading missing */
  EXPECT(rfd_stn_apply(NULL, NULL, 1, 4, 64, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_stn_apply(NULL, NULL, 0, 4, 64, NULL, NULL) == RFD_OK);
  EXPECT(rfd_mlp_chain_ex(RFD_MLP_MODE_F16X3, (const float *)16, 1, 64, 256, (const void *)16, 64, 0, 0, 1, 1, (float *)16, NULL, 0,
                          (const float *)16, 100, NULL, 0, NULL) == RFD_ERR_UNSUPPORTED_SIZE); /* group of 100 rows: not a tile multiple */
  EXPECT(rfd_mlp_chain_rows(RFD_MLP_MODE_F16X3, (const float *)16, 66, 1, 64, 256, (const void *)16, 64, 0, 0, 1, (float *)16, 64, 0, 0,
                            NULL, 0, NULL, 0, NULL) == RFD_ERR_INVALID_ARGUMENT); /* row stride must be a multiple of 4 floats */
  EXPECT(rfd_mlp_chain_rows(RFD_MLP_MODE_F16X3, (const float *)16, 64, 1, 64, 256, (const void *)16, 64, 0, 0, 1, (float *)16, 32, 0, 0,
                            NULL, 0, NULL, 0, NULL) == RFD_ERR_INVALID_ARGUMENT); /* output row narrower than the layer */
  /* empty work is a no-op success */
  EXPECT(rfd_furthest_point_sampling(NULL, 0, 16, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_ball_query(NULL, NULL, 0, 8, 8, 0.1f, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_group_points(NULL, NULL, 2, 3, 8, 0, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_onet_decode(NULL, 0, 0, 128, NULL, NULL, 1, NULL, NULL, 0.f, NULL, NULL) == RFD_OK);
  /* sizes */
  EXPECT(rfd_onet_packed_bytes(1) == (size_t)10 * 4 * 256 * 128);
  EXPECT(rfd_onet_packed_bytes(RFD_ONET_MODE_F16) == (size_t)10 * 4 * 256 * 128);
  EXPECT(rfd_onet_packed_bytes(RFD_ONET_MODE_F16X3) == (size_t)2 * 10 * 4 * 256 * 128);
  EXPECT(rfd_onet_packed_bytes(0) == 0 && rfd_onet_packed_bytes(4) == 0);
  EXPECT(rfd_onet_decode((const float *)1, 0, 1, 128, NULL, NULL, 7, NULL, NULL, 0.f, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  EXPECT(rfd_onet_decode_set_cluster(3) == RFD_ERR_INVALID_ARGUMENT);
  /* weight stages (4 + 2 + 2 K panels) + xyz table (256 x 16 B) + scale / shift tables (2 x 4 KB) */
  EXPECT(rfd_mlp_chain_packed_bytes(RFD_MLP_MODE_F16, 256, 3, 128, 128, 256) == (size_t)4 * 128 * 128 + 2 * 128 * 128 + 2 * 256 * 128 + 4096 + 8192);
  EXPECT(rfd_mlp_chain_packed_bytes(RFD_MLP_MODE_F16X3, 256, 3, 128, 128, 256) == (size_t)2 * (4 * 128 * 128 + 2 * 128 * 128 + 2 * 256 * 128) + 4096 + 8192);
  EXPECT(rfd_mlp_chain_packed_bytes(RFD_MLP_MODE_F16, 256, 0, 300, 128, 256) == 0); /* hidden width > 256 */
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
      /* device round trip: 64 points on a 4x4x4 lattice of pitch 1 (offset so that no point is inside the FPS skip ball) */
      enum { N = 64, M = 8, S = 4 };
      float h_xyz[N * 3], h_new[M * 3], h_grp[3 * M * S];
      int h_idx[M], h_bq[M * S], i, j, k;
      float *d_xyz = NULL, *d_new = NULL, *d_grp = NULL;
      int *d_idx = NULL, *d_bq = NULL;
      for (i = 0; i < N; ++i) {
        h_xyz[3 * i] = 1.f + (float)(i & 3);
        h_xyz[3 * i + 1] = 1.f + (float)((i >> 2) & 3);
        h_xyz[3 * i + 2] = 1.f + (float)(i >> 4);
      }
      EXPECT(cudaMalloc((void **)&d_xyz, sizeof(h_xyz)) == cudaSuccess);
      EXPECT(cudaMalloc((void **)&d_new, sizeof(h_new)) == cudaSuccess);
      EXPECT(cudaMalloc((void **)&d_grp, sizeof(h_grp)) == cudaSuccess);
      EXPECT(cudaMalloc((void **)&d_idx, sizeof(h_idx)) == cudaSuccess);
      EXPECT(cudaMalloc((void **)&d_bq, sizeof(h_bq)) == cudaSuccess);
      EXPECT(cudaMemcpy(d_xyz, h_xyz, sizeof(h_xyz), cudaMemcpyHostToDevice) == cudaSuccess);
      EXPECT(rfd_furthest_point_sampling_xyz(d_xyz, 1, N, M, d_idx, d_new, NULL) == RFD_OK);
      EXPECT(rfd_query_and_group(d_xyz, d_new, NULL, 1, N, M, 0, 1.5f, S, 1, 0, d_grp, NULL, d_bq, NULL) == RFD_OK);
      EXPECT(cudaDeviceSynchronize() == cudaSuccess);
      EXPECT(cudaMemcpy(h_idx, d_idx, sizeof(h_idx), cudaMemcpyDeviceToHost) == cudaSuccess);
      EXPECT(cudaMemcpy(h_new, d_new, sizeof(h_new), cudaMemcpyDeviceToHost) == cudaSuccess);
      EXPECT(cudaMemcpy(h_bq, d_bq, sizeof(h_bq), cudaMemcpyDeviceToHost) == cudaSuccess);
      EXPECT(cudaMemcpy(h_grp, d_grp, sizeof(h_grp), cudaMemcpyDeviceToHost) == cudaSuccess);
      EXPECT(h_idx[0] == 0);  /* the reference always starts at point 0 (sampling_gpu.cu:85-86) */
      EXPECT(h_idx[1] == 63); /* the opposite corner is the unique farthest point */
      for (j = 0; j < M; ++j) {
        EXPECT(h_idx[j] >= 0 && h_idx[j] < N);
        for (i = 0; i < j; ++i) EXPECT(h_idx[i] != h_idx[j]);
        for (k = 0; k < 3; ++k) EXPECT(h_new[3 * j + k] == h_xyz[3 * h_idx[j] + k]);
        /* every neighbour is within the radius, ascending, first = lowest index in the ball; grouped = p - centre */
        for (k = 0; k < S; ++k) {
          const int n = h_bq[j * S + k];
          float d2 = 0.f;
          int a;
          EXPECT(n >= 0 && n < N);
          for (a = 0; a < 3; ++a) {
            const float d = h_xyz[3 * n + a] - h_new[3 * j + a];
            d2 += d * d;
            EXPECT(h_grp[(a * M + j) * S + k] == d);
          }
          EXPECT(d2 < 1.5f * 1.5f);
          if (k) EXPECT(n > h_bq[j * S + k - 1]);  /* a lattice point has >= 4 neighbours within 1.5 incl. itself */
        }
      }
      EXPECT(rfd_launch_count() >= 2);
      cudaFree(d_xyz); cudaFree(d_new); cudaFree(d_grp); cudaFree(d_idx); cudaFree(d_bq);
      printf("abi_check device round trip ok\n");
    } else {
      printf("abi_check: no CUDA device, argument validation only\n");
    }
  }
  printf("abi_check ok\n");
  return 0;
}
