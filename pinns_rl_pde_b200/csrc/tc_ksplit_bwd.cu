// tc_ksplit_bwd.cu -- one translation unit of the tcgen05 kernels (see tc_api.h / tc_gemm.cuh): K-split dgrad launches
#define PINNK_TC_TU_KS_BWD 1
#include "tc_gemm.cuh"
