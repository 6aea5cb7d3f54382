// Library probes of the C ABI (no compute).
#include "../../include/ergm_b200.h"
#include "common.cuh"

extern "C" int ergm_abi_version(void) { return 1; }
extern "C" int ergm_device_sm_count(void) { return ergm::num_sms(); }
