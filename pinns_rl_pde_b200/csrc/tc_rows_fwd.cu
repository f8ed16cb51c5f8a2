// tc_rows_fwd.cu -- one translation unit of the tcgen05 kernels (see tc_api.h / tc_gemm.cuh)
#define PINNK_TC_TU_FWD 1
#include "tc_gemm.cuh"
