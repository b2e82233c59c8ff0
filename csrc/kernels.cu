// Unity translation unit for all device code (one copy of the sf::g_sf_error_code symbol).
#include "gemm_sm100.cu"
#include "elementwise.cu"
#include "optim_push.cu"
#include "conv.cu"
