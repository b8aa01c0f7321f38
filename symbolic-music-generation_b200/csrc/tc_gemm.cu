#include "common.cuh"
int txl_gemm_tc(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, int, const TxlEpilogue*, void*, int* handled) { *handled = 0; return 0; }
int txl_relattn_fwd_tc(const void*, const void*, const void*, const void*, const void*, const void*, const float*, const float*, void*, float*, const TxlAttnDims*, void*, int* handled) { *handled = 0; return 0; }
