// tcgen05/TMEM + TMA contractions (3xTF32).  Placeholder until the kernels land: reports the
// path as unsupported so that NTN_GEMM_AUTO uses the SIMT kernels and NTN_GEMM_TCGEN05 fails loudly.
#include "ntn_internal.h"

int ntn_tc_supported(const ntn_handle*) { return 0; }
int ntn_tc_prepare_batch(ntn_handle*) { return NTN_EUNSUP; }
void ntn_tc_release(ntn_handle*) {}
int ntn_tc_w_prep(ntn_handle*, const float*) { return NTN_EUNSUP; }
int ntn_tc_gemm_fwd(ntn_handle*) { return NTN_EUNSUP; }
int ntn_tc_gemm_dw(ntn_handle*) { return NTN_EUNSUP; }
int ntn_tc_gemm_ge(ntn_handle*) { return NTN_EUNSUP; }
