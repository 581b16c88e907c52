// monopsr_b200/csrc/c_abi.cu -- library identification and launch accounting.
#include "common.cuh"
#include "../../include/monopsr_b200_tfops.h"

namespace mpb {
unsigned long long g_launch_count = 0;
}

MPB_API const char* mpb_version(void) { return "monopsr_b200 0.1 sm_100a"; }

MPB_API unsigned long long mpb_launch_count(void) {
    return __atomic_load_n(&mpb::g_launch_count, __ATOMIC_RELAXED);
}
