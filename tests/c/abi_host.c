/* A C99 host for libfnnu.so: compiles include/fnnu.h as plain C, links every entry point the header declares and
 * exercises the calls that need no GPU (version, error reporting, argument validation, sizes of a two-operator program).
 * This is the shape of the host the reference's withheld engine/fast_nnunet.cpp plays: no Python, no torch.
 *   gcc -std=c99 -Wall -Wextra -Werror -Iinclude tests/c/abi_host.c -Lfast_nnunet_b200 -lfnnu -o abi_host */
#include <stdio.h>
#include <string.h>

#include "fnnu.h"

#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      fprintf(stderr, "abi_host: check failed: %s (line %d)\n", #cond, __LINE__); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(void) {
  /* every declared entry point, taken by address: an undefined or renamed symbol fails the LINK */
  void* entry_points[] = {
      (void*)fnnu_abi_version, (void*)fnnu_last_error, (void*)fnnu_device_ok, (void*)fnnu_gather_tiles,
      (void*)fnnu_accumulate_tiles, (void*)fnnu_weight_sum, (void*)fnnu_finalize, (void*)fnnu_mem_launches,
      (void*)fnnu_scale_inplace_f32, (void*)fnnu_add_inplace_f32, (void*)fnnu_export_labels,
      (void*)fnnu_pre_nonzero_bbox, (void*)fnnu_pre_filled_mask, (void*)fnnu_pre_channel_stats,
      (void*)fnnu_pre_crop_normalize, (void*)fnnu_pre_resample_workspace_bytes, (void*)fnnu_pre_resample_channel,
      (void*)fnnu_engine_sizes, (void*)fnnu_engine_create, (void*)fnnu_engine_destroy, (void*)fnnu_engine_forward,
      (void*)fnnu_engine_set_backend, (void*)fnnu_engine_launch_counts, (void*)fnnu_engine_profile_op,
      (void*)fnnu_engine_profile_ms, (void*)fnnu_engine_buffer, (void*)fnnu_engine_stats,
  };
  size_t n = sizeof(entry_points) / sizeof(entry_points[0]);
  for (size_t i = 0; i < n; ++i) CHECK(entry_points[i] != NULL);
  CHECK(fnnu_abi_version() == 1);

  /* errors are return codes + a message, never exceptions across the ABI */
  int vol[3] = {8, 8, 8}, patch[3] = {4, 4, 4};
  unsigned char flips[1] = {0};
  int rc = fnnu_gather_tiles(NULL, 1, vol, NULL, 1, patch, flips, 1, NULL, 1, NULL);
  CHECK(rc < 0);
  CHECK(strstr(fnnu_last_error(), "null") != NULL);
  CHECK(fnnu_engine_forward(NULL, 1, NULL) < 0);

  /* a two-operator program (3x3x3 conv with InstanceNorm, 1x1x1 head): the library tells the caller how much device
   * memory to bring; it allocates none itself */
  fnnu_buffer_desc bufs[3];
  memset(bufs, 0, sizeof(bufs));
  for (int b = 0; b < 3; ++b) {
    bufs[b].dims[0] = bufs[b].dims[1] = bufs[b].dims[2] = 16;
  }
  bufs[0].channels = 1;
  bufs[1].channels = 16;
  bufs[2].channels = 2;
  static float w0[16 * 1 * 27], g0[16], b0[16], w1[2 * 16], bias1[2];
  fnnu_op_desc ops[2];
  memset(ops, 0, sizeof(ops));
  ops[0].op = FNNU_OP_CONV;
  ops[0].src = 0; ops[0].dst = 1; ops[0].src2 = -1;
  ops[0].cin = 1; ops[0].cout = 16;
  for (int a = 0; a < 3; ++a) { ops[0].kernel[a] = 3; ops[0].stride[a] = 1; ops[1].kernel[a] = 1; ops[1].stride[a] = 1; }
  ops[0].has_norm = 1; ops[0].norm_eps = 1e-5f; ops[0].act_slope = 0.01f;
  ops[0].weight = w0; ops[0].gamma = g0; ops[0].beta = b0;
  ops[1].op = FNNU_OP_CONV;
  ops[1].src = 1; ops[1].dst = 2; ops[1].src2 = -1;
  ops[1].cin = 16; ops[1].cout = 2;
  ops[1].has_bias = 1; ops[1].act_slope = 1.0f;
  ops[1].weight = w1; ops[1].bias = bias1;
  size_t param_bytes = 0, workspace_bytes = 0;
  rc = fnnu_engine_sizes(bufs, 3, ops, 2, 4, &param_bytes, &workspace_bytes);
  if (rc != 0) fprintf(stderr, "abi_host: fnnu_engine_sizes: %s\n", fnnu_last_error());
  CHECK(rc == 0);
  CHECK(param_bytes > 0 && workspace_bytes >= (size_t)4 * 16 * 16 * 16 * (1 + 16 + 2) * 2);
  /* a malformed program is refused with a message */
  ops[1].src = 7;
  CHECK(fnnu_engine_sizes(bufs, 3, ops, 2, 4, &param_bytes, &workspace_bytes) < 0);
  CHECK(strlen(fnnu_last_error()) > 0);
  printf("abi_host: %zu entry points linked, ABI version %d, engine sizes ok\n", n, fnnu_abi_version());
  return 0;
}
