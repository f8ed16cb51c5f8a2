// tc_rows_bwd_y.cu -- one translation unit of the tcgen05 kernels (see tc_api.h / tc_gemm.cuh)
#define PINNK_TC_TU_BWD_Y 1
#include "tc_gemm.cuh"
