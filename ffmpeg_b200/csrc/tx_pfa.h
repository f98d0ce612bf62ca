// tx_pfa.h — compound 15 x M float MDCT (tx_pfa.cu), used by tx.cu for the lengths av_tx_init() gives to mdct_pfa_15xM
#pragma once
#include "common.h"

struct TxPfa;
// len = MDCT length (2 * N * 2^k with N = 15, 5 or 3 and k >= 1), scale as for av_tx_init; returns nullptr (and sets the error string) on failure
TxPfa *tx_pfa_create(int inv, int len, float scale);
// compound complex FFT of len = N * 2^k points (N = 15, 9, 7, 5 or 3; 2^k = 2 ... 512): ff_tx_fft_pfa over fftN_ns and the split-radix transform
TxPfa *tx_pfa_create_fft(int inv, int len);
bool   tx_pfa_fft_length_ok(int len);
void   tx_pfa_free(TxPfa *p);
bool   tx_pfa_length_ok(int len);
int    tx_pfa_launch(TxPfa *p, cudaStream_t st, void *out, const void *in, ptrdiff_t stride, int64_t count, ptrdiff_t out_step, ptrdiff_t in_step);
