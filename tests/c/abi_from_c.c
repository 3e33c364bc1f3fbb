/* tests/c/abi_from_c.c -- include/lfb200.h consumed from strict C99 (gcc -std=c99 -pedantic): the boundary is a C ABI, not a C++
 * one.  Built and run by tests/test_host_api.py::test_header_is_plain_c; host-only entry points, then lfb_create. */
#include "lfb200.h"
#include <stdio.h>
int main(void) {
  lfb_lens lens;
  lfb_params p = {0};
  double rays, inter; int jobs;
  if (lfb_builtin_lens(&lens, 3, 550.f) != LFB_OK) return 1;
  p.mode = LFB_MODE_EXACT_GRID; p.width = 1920; p.height = 1080; p.grid_n = 256; p.pair_set = LFB_PAIRS_ALL; p.include_direct = 1;
  if (lfb_count_work(&lens, &p, 1, &rays, &inter, &jobs) != LFB_OK) return 2;
  printf("abi %d jobs %d rays %.0f interactions %.0f sizeof(lens)=%zu light=%zu params=%zu hit=%zu\n", lfb_abi_version(), jobs, rays, inter,
         sizeof(lfb_lens), sizeof(lfb_light), sizeof(lfb_params), sizeof(lfb_ray_hit));
  lfb_engine* e = NULL;
  int rc = lfb_create(&e, 0);
  printf("lfb_create -> %d (%s)\n", rc, rc ? lfb_last_error() : "ok");
  return 0;
}
