// tx_r16.h — register-resident split-radix passes for the power-of-two float FFT / inverse MDCT (tx_r16.cu); used by tx.cu
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct TxR16;
// mode 0: complex FFT of n points (forward or inverse: the difference is in `gather`), mode 1: inverse MDCT with n = len / 2.
// gather[i]: z[i] = src[gather[i]] (ff_tx_gen_ptwo_revtab); exp_nat: the n MDCT twiddles in natural order (mode 1), else NULL.
// Returns NULL when n is not one of the sizes the kernel is built for (512 ... 4096) — the caller keeps its level-by-level kernels.
TxR16 *tx_r16_create(int mode, int n, const int *gather, const float2 *exp_nat, int sm_count);
void tx_r16_destroy(TxR16 *p);
// true when the buffers fit the kernel's bulk-copy loads (16-byte aligned input and input step, 8-byte aligned output and step)
bool tx_r16_accepts(const TxR16 *p, const void *out, const void *in, long long out_step, long long in_step);
int tx_r16_launch(TxR16 *p, cudaStream_t st, void *out, const void *in, long long out_step, long long in_step, long long count);
