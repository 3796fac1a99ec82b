// contact_addon_b200.cu -- single translation unit of libcontact_addon_b200.so (sm_100a).
// Build: see contact_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared).
#include "../../include/contact_addon_b200.h"
#include "api_cntc.cuh"
