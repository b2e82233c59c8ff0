// Unity translation unit for all device code (one copy of the sf::g_sf_error_code symbol).
#include "gemm_sm100.cu"
#include "gemm_pair_sm100.cu"
#include "elementwise.cu"
#include "optim_push.cu"
#include "conv.cu"

// Force the (lazy) load of EVERY kernel of this library up front: CUDA's lazy module loading needs a context-wide
// synchronisation the first time a kernel is used, which must never happen in the middle of a training run.
#define SF_PRELOAD(k)                                                   \
  do {                                                                  \
    cudaFuncAttributes fa;                                              \
    cudaError_t e_ = cudaFuncGetAttributes(&fa, k);                     \
    if (e_ != cudaSuccess) return static_cast<int>(e_);                 \
  } while (0)
#define SF_PRELOAD_OPT(K, SYS)                                                                          \
  SF_PRELOAD((sf::K<SF_OPT_SGD, SYS>)); SF_PRELOAD((sf::K<SF_OPT_MOMENTUM, SYS>)); SF_PRELOAD((sf::K<SF_OPT_ADAM, SYS>)); \
  SF_PRELOAD((sf::K<SF_OPT_RMSPROP, SYS>)); SF_PRELOAD((sf::K<SF_OPT_ADAGRAD, SYS>)); SF_PRELOAD((sf::K<SF_OPT_ADADELTA, SYS>)); \
  SF_PRELOAD((sf::K<SF_OPT_ADAGRAD_DA, SYS>)); SF_PRELOAD((sf::K<SF_OPT_FTRL, SYS>)); SF_PRELOAD((sf::K<SF_OPT_PROXIMAL_ADAGRAD, SYS>)); \
  SF_PRELOAD((sf::K<SF_OPT_PROXIMAL_SGD, SYS>))

extern "C" int sf_preload_kernels() {
  SF_PRELOAD(sf::sf_gemm_kernel<32>); SF_PRELOAD(sf::sf_gemm_kernel<64>); SF_PRELOAD(sf::sf_gemm_kernel<128>); SF_PRELOAD(sf::sf_gemm_kernel<256>); SF_PRELOAD(sf::sf_gemm_pair_kernel);
  SF_PRELOAD(sf::cast_transpose_kernel<false>); SF_PRELOAD(sf::cast_transpose_kernel<true>);
  SF_PRELOAD(sf::softmax_xent_kernel); SF_PRELOAD(sf::mse_kernel); SF_PRELOAD(sf::argmax_rows_kernel);
  SF_PRELOAD(sf::im2col_kernel); SF_PRELOAD(sf::col2im_kernel); SF_PRELOAD(sf::maxpool_fwd_kernel); SF_PRELOAD(sf::maxpool_bwd_kernel);
  SF_PRELOAD(sf::pull_kernel<true>); SF_PRELOAD(sf::pull_kernel<false>); SF_PRELOAD(sf::post_kernel); SF_PRELOAD(sf::lock_test_kernel);
  SF_PRELOAD_OPT(push_kernel, true); SF_PRELOAD_OPT(push_kernel, false);
  SF_PRELOAD_OPT(applier_kernel, true); SF_PRELOAD_OPT(applier_kernel, false);
  return 0;
}
