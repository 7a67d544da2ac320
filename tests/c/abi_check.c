/* C99 consumer of include/rfdnet_b200.h: proves the boundary is a plain C ABI (no C++/torch types) and that the
 * argument validation of every entry point runs before any CUDA call (so it works on a machine without a GPU). */
#include <stdio.h>
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
  EXPECT(rfd_sa_mlp_tc(NULL, 1, 4, 8, 16, NULL, NULL, 64, 64, 128, NULL, NULL) == RFD_ERR_INVALID_ARGUMENT);
  /* empty work is a no-op success */
  EXPECT(rfd_furthest_point_sampling(NULL, 0, 16, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_ball_query(NULL, NULL, 0, 8, 8, 0.1f, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_group_points(NULL, NULL, 2, 3, 8, 0, 4, NULL, NULL) == RFD_OK);
  EXPECT(rfd_onet_decode(NULL, 0, 0, 128, NULL, NULL, 1, NULL, NULL, 0.f, NULL, NULL) == RFD_OK);
  /* sizes */
  EXPECT(rfd_onet_packed_bytes(1) == (size_t)10 * 4 * 256 * 128);
  EXPECT(rfd_onet_packed_bytes(3) == 0);
  EXPECT(rfd_sa_mlp_tc_packed_bytes(259, 128, 128, 256) == (size_t)5 * 128 * 128 + 2 * 128 * 128 + 2 * 256 * 128);
  EXPECT(rfd_sa_mlp_tc_packed_bytes(400, 128, 128, 256) == 0);
  printf("abi_check ok\n");
  return 0;
}
