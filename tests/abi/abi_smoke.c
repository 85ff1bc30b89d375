/* Plain-C consumer of the C ABI: proves include/cdm_b200.h is valid C (no C++, no torch types) and that a
 * program can link libcdm_b200.so directly.  Without a GPU every compute entry point must fail loudly
 * (negative status + message), never fall back.  Built and run by tests/test_cpu_host.py. */
#include <stdio.h>
#include <string.h>

#include "cdm_b200.h"

int main(void) {
  printf("version %d\n", cdm_version());
  int dev = cdm_device_ok();
  printf("device_ok %d (%s)\n", dev, cdm_last_error());
  /* argument checking happens before any device work: a NULL struct is CDM_ERR_ARG on every box */
  int rc = cdm_conv3x3((const cdm_conv3x3_args*)0, (void*)0);
  printf("conv3x3(NULL) -> %d (%s)\n", rc, cdm_last_error());
  if (rc != CDM_ERR_ARG) return 2;
  cdm_conv_out_args co;
  memset(&co, 0, sizeof(co));
  rc = cdm_conv_out(&co, (void*)0);
  if (rc != CDM_ERR_ARG) return 3;
  float dummy[4] = {0};
  rc = cdm_xrank_sum(dummy, 1, 4, dummy, (const cdm_xrank*)0, (void*)0);
  printf("xrank_sum on host pointers -> %d (%s)\n", rc, cdm_last_error());
  if (dev != CDM_OK && rc == CDM_OK) return 4; /* no usable GPU: must not pretend to have computed anything */
  return 0;
}
