/* hulc_b200.h — C ABI of libhulc_b200.so: the sm_100a kernels behind the HULC training hot path.
 *
 * The reference (lukashermann/hulc) is pure Python: its "FFI" for this path is torch's operator set called from the
 * nn.Module.forward / loss methods cited next to each entry point below (paths relative to the reference root).  A
 * maintainer binds these symbols with ctypes (see INTEGRATION.md and hulc_b200/_lib.py) and calls them from
 * torch.autograd.Function bodies in place of those torch ops.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless the parameter type says otherwise; row-major, leading dimensions
 *     in elements; no torch types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and never synchronises;
 *   - return value: 0 on success, otherwise the cudaError_t of the failing launch / argument check;
 *   - dropout arguments: `drop_p` = 0 disables; with `drop_keep` != NULL it is an injected uint8 keep-mask laid out like
 *     the tensor it applies to, otherwise keep decisions are Philox4x32-10(seed, site, element index) and the backward
 *     entry points regenerate them from the same (seed, site);
 *   - `workspace`: caller-owned scratch (16-byte aligned, zero-initialised once, then reusable by calls on the SAME stream).
 */
#ifndef HULC_B200_H
#define HULC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- dense layers ---------------------------------------------------------------------------------------------------
 * hulc_gemm: C[M,N] = epi(alpha * op(A)[M,K] * op(B)[K,N]); transA=1: A stored KxM; transB=1: B stored NxK (torch Linear
 * weight).  epi(v) = dropout(gate(act(v + bias[n] + addend[(add_mod ? m % add_mod : m), n] + beta*C[m,n]))).
 * Replaces torch.nn.Linear forward / backward everywhere on the path (e.g. plan_encoders/plan_proposal_net.py:26-47,
 * encoders/goal_encoders.py:20-36, decoders/logistic_decoder_rnn.py:278-283) and the per-step matmuls of torch.nn.RNN
 * (decoders/utils/rnn.py:5-14). */
int hulc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
              float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act,
              const float* gate, int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site,
              const unsigned char* drop_keep, float* workspace, size_t workspace_bytes, void* stream);

/* out[c] = beta*out[c] + sum_r X[r*ldx + c]  (bias gradients). */
int hulc_colsum(const float* X, int rows, int cols, int ldx, float* out, float beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HULC_B200_H */
