// tc_ksplit_fwd.cu -- one translation unit of the tcgen05 kernels (see tc_api.h / tc_gemm.cuh): K-split forward launches
#define PINNK_TC_TU_KS_FWD 1
#include "tc_gemm.cuh"
