// fp32-accurate GEMM on the 5th-gen tensor cores (3xTF32 split), the building block of K10 and K13.
#pragma once
#include "common.cuh"

namespace vlmc {

// C[M,N] = beta * C + alpha * A[M,K] * op(B)
//   A row-major [M,K];  b_nk: B row-major [N,K] (C = A B^T), else B row-major [K,N]
//   tri = 1: only the tiles that touch the lower triangle (column block start <= last row of the tile) are computed
//   kc: K elements accumulated inside the tensor core between round-to-nearest flushes into registers (0 = 128)
//   kmode: structural zeros whose k-slices are skipped.  1: B is [K,N] with B[k][j] == 0 for k < j (lower
//          triangular); 2: A[i][k] == 0 for k > i (lower triangular).  0: dense
//   a_km: A is given as a row-major [K,M] array (C = A^T op(B)); needs b_nk = false, kmode = 0; K is then arbitrary
//   max_ctas: cap of the (persistent) grid, 0 = one CTA per SM; a caller that overlaps this GEMM with a kernel on another
//          stream leaves SMs free this way
// Requirements: K, N, lda, ldb, ldc multiples of 4; A, B, C 16-byte aligned.  A or B may alias C only when every
// output tile reads exactly the operand rows it overwrites (N <= tile width: the panel solve of the Cholesky).
int gemm3x(bool b_nk, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
           float beta, float* C, int64_t ldc, int tri, int kc, cudaStream_t st, int kmode = 0, bool a_km = false,
           int max_ctas = 0);

// Side stream + fork/join events of the look-ahead schedules (blocked Cholesky, OBS far updates), one set per
// (device, caller stream, user): chains that run concurrently on different streams must not share a side stream.
struct ChainSide { cudaStream_t stream; cudaEvent_t solved, updated; };
ChainSide* chain_side_for(cudaStream_t st, int user);
bool chain_lookahead_enabled();       // vlmc_chol_set_lookahead / VLMC_CHOL_LOOKAHEAD

}  // namespace vlmc
