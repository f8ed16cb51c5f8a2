// tc_wgrad.cu -- one translation unit of the tcgen05 kernels (see tc_api.h / tc_gemm.cuh)
#define PINNK_TC_TU_WGRAD 1
#include "tc_gemm.cuh"
