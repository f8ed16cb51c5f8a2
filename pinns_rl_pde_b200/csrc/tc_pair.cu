// translation unit of the paired reverse kernel (dgrad + tanh adjoint | wgrad on the two CTAs of a cluster)
#define PINNK_TC_TU_PAIR
#include "tc_gemm.cuh"
