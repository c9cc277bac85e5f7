/* A host that is neither Python nor torch, bound to the C ABI of include/pcb200.h (what INTEGRATION.md §2 describes).
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_host.c \
 *       -Lpytorch_connectomics_b200/csrc -lpcb200 -Wl,-rpath,$PWD/pytorch_connectomics_b200/csrc -o /tmp/c_host
 *
 * It plans the sliding-window grid of BASELINE configs[4] (2048^3 volume, 160^3 tiles, 50 % overlap — reference
 * connectomics/inference/window.py:57-134) with the library's host-side integer logic, then, if a B200 is present, builds the
 * blending map on the device.  Without a GPU it stops after the plan (the library has no CPU path for device work and says so).
 * tests/test_c_header.py compiles it as strict C99 and runs it in the CPU suite. */
#include <stdio.h>
#include <stdlib.h>

#include "pcb200.h"

int main(void) {
  const int64_t image[3] = {2048, 2048, 2048}, roi[3] = {160, 160, 160};
  const double overlap[3] = {0.5, 0.5, 0.5};
  int64_t interval[3], count = 0;
  if (pcb_sw_scan_interval(image, roi, overlap, interval) != PCB_OK) {
    fprintf(stderr, "scan interval: %s\n", pcb_last_error());
    return 1;
  }
  if (pcb_sw_plan(PCB_GRID_EAGER, image, roi, overlap, NULL, NULL, 0, &count) != PCB_OK) { /* capacity 0: count only */
    fprintf(stderr, "plan: %s\n", pcb_last_error());
    return 1;
  }
  int64_t* starts = (int64_t*)malloc((size_t)count * 3 * sizeof(int64_t));
  if (!starts || pcb_sw_plan(PCB_GRID_EAGER, image, roi, overlap, NULL, starts, count, &count) != PCB_OK) {
    fprintf(stderr, "plan: %s\n", pcb_last_error());
    return 1;
  }
  int64_t lazy = 0;
  if (pcb_sw_plan(PCB_GRID_LAZY, image, roi, overlap, NULL, NULL, 0, &lazy) != PCB_OK) return 1;
  printf("pcb200 v%d  interval %lld %lld %lld  eager windows %lld (last start %lld %lld %lld)  lazy windows %lld\n",
         pcb_version(), (long long)interval[0], (long long)interval[1], (long long)interval[2], (long long)count,
         (long long)starts[3 * (count - 1)], (long long)starts[3 * (count - 1) + 1], (long long)starts[3 * (count - 1) + 2],
         (long long)lazy);
  free(starts);
  /* a bad argument is PCB_ERR_INVALID with a message, never a crash */
  const int64_t bad_roi[3] = {0, 160, 160};
  if (pcb_sw_scan_interval(image, bad_roi, overlap, interval) != PCB_ERR_INVALID) return 2;
  printf("invalid roi -> \"%s\"\n", pcb_last_error());
  if (!pcb_device_ok()) {
    printf("no B200 here: %s\n", pcb_last_error());
    return 0;
  }
  printf("B200 present: device work (pcb_sw_importance_map, pcb_net_forward, pcb_sw_run, pcb_grad_allreduce) can be enqueued\n");
  return 0;
}
